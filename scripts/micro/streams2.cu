// Micro-benchmark 2: achieved HBM bandwidth of a streaming kernel that reads M separate arrays (sizes as the level-1
// fields of 512x512x64, not powers of two; pointers passed by value) and writes one, with 8-byte or 16-byte loads per lane,
// at several occupancies.  Question: do many concurrent read streams or narrow loads cap the bandwidth near 4.5 TB/s?
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
template <int M> struct Ptrs { const double* a[M]; };
template <int M, int W>   // W = doubles per lane per load (1 or 2)
__global__ void __launch_bounds__(256) k_sep(Ptrs<M> P, double* __restrict__ out, size_t n) {
  size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * W, stride = (size_t)gridDim.x * blockDim.x * W;
  for (; e + W <= n; e += stride) {
    if (W == 1) {
      double v[M];
#pragma unroll
      for (int m = 0; m < M; ++m) v[m] = __ldg(P.a[m] + e);
      double s = 0;
#pragma unroll
      for (int m = 0; m < M; ++m) s += v[m];
      out[e] = s;
    } else {
      double2 v[M];
#pragma unroll
      for (int m = 0; m < M; ++m) v[m] = __ldg(reinterpret_cast<const double2*>(P.a[m] + e));
      double2 s = make_double2(0, 0);
#pragma unroll
      for (int m = 0; m < M; ++m) { s.x += v[m].x; s.y += v[m].y; }
      *reinterpret_cast<double2*>(out + e) = s;
    }
  }
}
template <int M, int W> void run(size_t n, int ctas_per_sm) {
  Ptrs<M> P; std::vector<double*> h(M); double* out; std::vector<void*> pads;
  for (int m = 0; m < M; ++m) { cudaMalloc(&h[m], n * 8); cudaMemset(h[m], 0, n * 8); P.a[m] = h[m];
    void* pad; cudaMalloc(&pad, (size_t)(3 + 2 * m) << 20); pads.push_back(pad); }     // irregular spacing of the bases
  cudaMalloc(&out, n * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * ctas_per_sm; float ms;
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(a); for (int r = 0; r < 10; ++r) k_sep<M, W><<<grid, 256>>>(P, out, n); cudaEventRecord(b); cudaEventSynchronize(b); }
  cudaEventElapsedTime(&ms, a, b);
  printf("M=%2d  %2d B/lane  %d CTA/SM : %7.1f GB/s  (%.1f us)\n", M, 8 * W, ctas_per_sm, (M + 1) * n * 8.0 / (ms / 10 * 1e-3) / 1e9, ms / 10 * 1e3);
  for (int m = 0; m < M; ++m) cudaFree(h[m]);
  for (void* p : pads) cudaFree(p);
  cudaFree(out);
}
int main() {
  const size_t n = (size_t)64 * 514 * 514;
  for (int c : {2, 4, 8}) { run<1, 1>(n, c); run<1, 2>(n, c); run<4, 1>(n, c); run<4, 2>(n, c); run<17, 1>(n, c); run<17, 2>(n, c); run<31, 1>(n, c); run<31, 2>(n, c); }
  cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
  return 0;
}

// Micro-benchmark: achieved HBM bandwidth of a streaming kernel that reads M arrays and writes one,
// (a) M separate allocations (SoA planes), (b) one allocation interleaved in chunks of CH doubles per array.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
template <int M>
__global__ void __launch_bounds__(256) k_sep(const double* const* __restrict__ a, double* __restrict__ out, size_t n) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; e < n; e += stride) {
    double v[M];
#pragma unroll
    for (int m = 0; m < M; ++m) v[m] = a[m][e];
    double s = 0;
#pragma unroll
    for (int m = 0; m < M; ++m) s += v[m];
    out[e] = s;
  }
}
template <int M>
__global__ void __launch_bounds__(256) k_int(const double* __restrict__ a, double* __restrict__ out, size_t n, int ch) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; e < n; e += stride) {
    const size_t blk = e / ch, off = e % ch;
    const double* base = a + blk * (size_t)ch * M + off;
    double v[M];
#pragma unroll
    for (int m = 0; m < M; ++m) v[m] = base[(size_t)m * ch];
    double s = 0;
#pragma unroll
    for (int m = 0; m < M; ++m) s += v[m];
    out[e] = s;
  }
}
template <int M> void run(size_t n) {
  std::vector<double*> h(M); double** d; double* out; double* inter;
  for (int m = 0; m < M; ++m) { cudaMalloc(&h[m], n * 8); cudaMemset(h[m], 0, n * 8); }
  cudaMalloc(&d, M * sizeof(double*)); cudaMemcpy(d, h.data(), M * sizeof(double*), cudaMemcpyHostToDevice);
  cudaMalloc(&out, n * 8); cudaMalloc(&inter, n * 8 * M); cudaMemset(inter, 0, n * 8 * M);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * 8; float ms;
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(a); for (int r = 0; r < 10; ++r) k_sep<M><<<grid, 256>>>(d, out, n); cudaEventRecord(b); cudaEventSynchronize(b); }
  cudaEventElapsedTime(&ms, a, b);
  printf("M=%2d separate   : %7.1f GB/s\n", M, (M + 1) * n * 8.0 / (ms / 10 * 1e-3) / 1e9);
  for (int ch : {64, 512, 4096}) {
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(a); for (int r = 0; r < 10; ++r) k_int<M><<<grid, 256>>>(inter, out, n, ch); cudaEventRecord(b); cudaEventSynchronize(b); }
    cudaEventElapsedTime(&ms, a, b);
    printf("M=%2d interleaved ch=%4d doubles: %7.1f GB/s\n", M, ch, (M + 1) * n * 8.0 / (ms / 10 * 1e-3) / 1e9);
  }
  for (int m = 0; m < M; ++m) cudaFree(h[m]);
  cudaFree(d); cudaFree(out); cudaFree(inter);
}
int main() {
  const size_t n = (size_t)1 << 24;   // 16.8M doubles per array = 134 MB (as level 1 of 512x512x64)
  run<1>(n); run<2>(n); run<4>(n); run<10>(n); run<17>(n);
  cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
  return 0;
}

// Micro-benchmark: cycles per column of the shared-memory Thomas solve (one lane per column), as the Thomas warp of the
// relaxation kernels runs it.  Variants: (0) the round-1 loop (no restrict), (1) restrict pointers + unroll 8,
// (2) register-blocked: operands of 8 levels preloaded before the dependent chain, (3) as 2 with two columns per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o thomas thomas.cu && ./thomas
#include <cstdio>
#include <cuda_runtime.h>

template <int NZ, int VAR>
__global__ void k(double* out, long long* cyc, int lanes, int reps) {
  extern __shared__ double sm[];
  constexpr int pitch = NZ + 1;
  const int lane = threadIdx.x;
  const int ncol = (VAR == 3) ? 2 * lanes : lanes;
  for (int e = threadIdx.x; e < ncol * pitch * 3; e += blockDim.x) sm[e] = 1.0 + 1e-3 * (e % 97);
  __syncthreads();
  double* s_rhs = sm; double* s_bet = sm + ncol * pitch; double* s_ud = s_bet + ncol * pitch;
  for (int e = threadIdx.x; e < ncol * pitch; e += blockDim.x) { s_bet[e] = 0.4 + 1e-4 * (e % 13); s_ud[e] = -0.3 + 1e-4 * (e % 7); }
  __syncthreads();
  long long t0 = clock64();
  double acc = 0.0;
  for (int r = 0; r < reps; ++r) {
    if (lane < lanes) {
      if (VAR == 0) {
        double* x = s_rhs + lane * pitch; const double* bt = s_bet + lane * pitch; const double* ud = s_ud + lane * pitch;
        double xm = x[0] * bt[0]; x[0] = xm;
        for (int kk = 1; kk < NZ; ++kk) { xm = (x[kk] - ud[kk - 1] * xm) * bt[kk]; x[kk] = xm; }
        for (int kk = NZ - 2; kk >= 0; --kk) { xm = x[kk] - (ud[kk] * bt[kk]) * xm; x[kk] = xm; }
        acc += xm;
      } else if (VAR == 1) {
        double* __restrict__ x = s_rhs + lane * pitch; const double* __restrict__ bt = s_bet + lane * pitch; const double* __restrict__ ud = s_ud + lane * pitch;
        double xm = x[0] * bt[0]; x[0] = xm;
#pragma unroll 8
        for (int kk = 1; kk < NZ; ++kk) { xm = (x[kk] - ud[kk - 1] * xm) * bt[kk]; x[kk] = xm; }
#pragma unroll 8
        for (int kk = NZ - 2; kk >= 0; --kk) { xm = x[kk] - (ud[kk] * bt[kk]) * xm; x[kk] = xm; }
        acc += xm;
      } else if (VAR == 2) {
        double* __restrict__ x = s_rhs + lane * pitch; const double* __restrict__ bt = s_bet + lane * pitch; const double* __restrict__ ud = s_ud + lane * pitch;
        double xm = x[0] * bt[0]; x[0] = xm;
        // forward, blocks of 8 levels: all operands in registers before the chain starts
        for (int k0 = 1; k0 < NZ; k0 += 8) {
          double r[8], u[8], b[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) if (k0 + q < NZ) { r[q] = x[k0 + q]; u[q] = ud[k0 + q - 1]; b[q] = bt[k0 + q]; }
#pragma unroll
          for (int q = 0; q < 8; ++q) if (k0 + q < NZ) { xm = (r[q] - u[q] * xm) * b[q]; x[k0 + q] = xm; }
        }
        for (int k0 = NZ - 2; k0 >= 0; k0 -= 8) {
          double r[8], w[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) if (k0 - q >= 0) { r[q] = x[k0 - q]; w[q] = ud[k0 - q] * bt[k0 - q]; }
#pragma unroll
          for (int q = 0; q < 8; ++q) if (k0 - q >= 0) { xm = r[q] - w[q] * xm; x[k0 - q] = xm; }
        }
        acc += xm;
      } else {
        double* __restrict__ x = s_rhs + lane * pitch; const double* __restrict__ bt = s_bet + lane * pitch; const double* __restrict__ ud = s_ud + lane * pitch;
        double* __restrict__ y = s_rhs + (lanes + lane) * pitch; const double* __restrict__ bt2 = s_bet + (lanes + lane) * pitch; const double* __restrict__ ud2 = s_ud + (lanes + lane) * pitch;
        double xm = x[0] * bt[0]; x[0] = xm; double ym = y[0] * bt2[0]; y[0] = ym;
        for (int k0 = 1; k0 < NZ; k0 += 4) {
          double r[4], u[4], b[4], r2[4], u2[4], b2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) if (k0 + q < NZ) { r[q] = x[k0 + q]; u[q] = ud[k0 + q - 1]; b[q] = bt[k0 + q]; r2[q] = y[k0 + q]; u2[q] = ud2[k0 + q - 1]; b2[q] = bt2[k0 + q]; }
#pragma unroll
          for (int q = 0; q < 4; ++q) if (k0 + q < NZ) { xm = (r[q] - u[q] * xm) * b[q]; x[k0 + q] = xm; ym = (r2[q] - u2[q] * ym) * b2[q]; y[k0 + q] = ym; }
        }
        for (int k0 = NZ - 2; k0 >= 0; k0 -= 4) {
          double r[4], w[4], r2[4], w2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) if (k0 - q >= 0) { r[q] = x[k0 - q]; w[q] = ud[k0 - q] * bt[k0 - q]; r2[q] = y[k0 - q]; w2[q] = ud2[k0 - q] * bt2[k0 - q]; }
#pragma unroll
          for (int q = 0; q < 4; ++q) if (k0 - q >= 0) { xm = r[q] - w[q] * xm; x[k0 - q] = xm; ym = r2[q] - w2[q] * ym; y[k0 - q] = ym; }
        }
        acc += xm + ym;
      }
    }
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) *cyc = (t1 - t0) / reps;
  out[threadIdx.x] = acc;
}

template <int NZ, int VAR> void run(int lanes) {
  double* out; long long* cyc; cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
  const size_t smem = (size_t)3 * 64 * (NZ + 1) * 8;
  cudaFuncSetAttribute(k<NZ, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<NZ, VAR><<<1, 32, smem>>>(out, cyc, lanes, 50);
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const int cols = (VAR == 3) ? 2 * lanes : lanes;
  printf("nz=%3d variant %d lanes %2d cols %2d: %6lld cycles per solve = %.1f cycles per level per pass-pair, err=%s\n", NZ, VAR, lanes, cols, h, (double)h / NZ, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<64, 0>(16); run<64, 1>(16); run<64, 2>(16); run<64, 3>(16);
  run<128, 0>(8); run<128, 1>(8); run<128, 2>(8); run<128, 3>(8);
  run<128, 0>(16); run<128, 2>(16); run<32, 0>(32); run<32, 2>(32);
  return 0;
}

// device_utils.cuh -- helpers shared by the kernel translation units (kernels_solve.cu, kernels_relax.cu).
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace mgb {

#define IDX3(k, j, i, nzl, ny2) ((size_t)(k) + (size_t)(nzl) * ((size_t)(j) + (size_t)(ny2) * (size_t)(i)))

static inline unsigned blocks_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

// one launch path for every kernel (counts the launches for mgb_launch_count; the first failing launch is remembered
// and surfaced by the next API boundary, see note_launch_error)
template <typename... KArgs, typename... Args>
static void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  note_launch_error(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
  ++g_launches;
}

// Store a freshly relaxed value and, on a level without lateral neighbour ranks, its homogeneous-Neumann mirror(s):
// the halo cell next to an edge column (and the corner next to a corner column) takes the same value, which is what
// fill_halo would copy afterwards (src/mg_mpi_exchange.f90:504-598).  Saves the separate halo launch per colour.
template <typename T>
__device__ __forceinline__ void store_mirrored_t(T* __restrict__ a, size_t idx, T v, int j, int i, int ny, int nx, size_t sj, size_t si, bool mirror) {
  a[idx] = v;
  if (!mirror) return;
  const bool w = (i == 1), e = (i == nx), s = (j == 1), n = (j == ny);
  if (w) a[idx - si] = v;
  if (e) a[idx + si] = v;
  if (s) a[idx - sj] = v;
  if (n) a[idx + sj] = v;
  if (w && s) a[idx - si - sj] = v;
  if (w && n) a[idx - si + sj] = v;
  if (e && s) a[idx + si - sj] = v;
  if (e && n) a[idx + si + sj] = v;
}
__device__ __forceinline__ void store_mirrored(double* __restrict__ a, size_t idx, double v, int j, int i, int ny, int nx, size_t sj, size_t si, bool mirror) {
  store_mirrored_t<double>(a, idx, v, j, i, ny, nx, sj, si, mirror);
}

// first column of row i in the set: red-black parity, fixed (four-colour), or the single column j = w - 2 i of the
// Gauss-Seidel wavefront w = jfix (jrule 2)
__device__ __forceinline__ int colour_jstart(const ColourSet& cs, int i) {
  return cs.jrule == 0 ? 1 + ((i + cs.rb) & 1) : (cs.jrule == 1 ? cs.jfix : cs.jfix - 2 * i);
}

// Thomas solve of one column held in shared memory (tridiag, src/mg_relax.f90:308-334, with the reciprocal pivots `bet`
// precomputed): x(k) in place, bu(k) = {bet(k), ud(k)} with ud(k) = cA(2,k+1) (0 at the top).  Same operations, in the same
// order, as the plain loops
//     x0 *= bet0 ; k = 1..nz-1: x(k) = (x(k) - ud(k-1) x(k-1)) bet(k) ; k = nz-2..0: x(k) = x(k) - (ud(k) bet(k)) x(k+1)
// but the shared-memory operands of the NEXT four levels are loaded before the dependent chain of the current four runs:
// under the load/store pressure of the producer warps an LDS costs 100+ cycles, and on the chain it made the solve 3-5x
// slower than its FP64 latency bound (~55 cycles per level), which was what limited the relaxation kernels (profiles/summary_r02.md).
// Returns x(0).
template <int NZ>
__device__ __forceinline__ double thomas_smem(double* __restrict__ x, const double2* __restrict__ bu) {
  static_assert(NZ >= 2, "at least two levels");
  const double2 c0 = bu[0];
  double xm = x[0] * c0.x;
  x[0] = xm;
  double udp = c0.y;                                           // ud(k-1)
  double r[4]; double2 c[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) if (1 + q < NZ) { r[q] = x[1 + q]; c[q] = bu[1 + q]; }
#pragma unroll 1
  for (int k0 = 1; k0 < NZ; k0 += 4) {
    double rn[4]; double2 cn[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) if (k0 + 4 + q < NZ) { rn[q] = x[k0 + 4 + q]; cn[q] = bu[k0 + 4 + q]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) if (k0 + q < NZ) { xm = (r[q] - udp * xm) * c[q].x; x[k0 + q] = xm; udp = c[q].y; }
#pragma unroll
    for (int q = 0; q < 4; ++q) { r[q] = rn[q]; c[q] = cn[q]; }
  }
  // back substitution, descending from k = nz-2
  double w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) if (NZ - 2 - q >= 0) { r[q] = x[NZ - 2 - q]; const double2 t = bu[NZ - 2 - q]; w[q] = t.y * t.x; }
#pragma unroll 1
  for (int k0 = NZ - 2; k0 >= 0; k0 -= 4) {
    double rn[4], wn[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) if (k0 - 4 - q >= 0) { rn[q] = x[k0 - 4 - q]; const double2 t = bu[k0 - 4 - q]; wn[q] = t.y * t.x; }
#pragma unroll
    for (int q = 0; q < 4; ++q) if (k0 - q >= 0) { xm = r[q] - w[q] * xm; x[k0 - q] = xm; }
#pragma unroll
    for (int q = 0; q < 4; ++q) { r[q] = rn[q]; w[q] = wn[q]; }
  }
  return xm;
}

// the plain loops on the same {pivot, off-diagonal} layout (any nz)
__device__ __forceinline__ double thomas_smem_plain(double* x, const double2* bu, int nz) {
  double2 c = bu[0];
  double xm = x[0] * c.x;
  x[0] = xm;
  for (int k = 1; k < nz; ++k) { const double udp = c.y; c = bu[k]; xm = (x[k] - udp * xm) * c.x; x[k] = xm; }
  for (int k = nz - 2; k >= 0; --k) { c = bu[k]; xm = x[k] - (c.y * c.x) * xm; x[k] = xm; }
  return xm;
}

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

}  // namespace mgb

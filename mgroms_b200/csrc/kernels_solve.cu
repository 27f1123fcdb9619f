// kernels_solve.cu -- FP64 kernels of the per-solve hot path for sm_100a:
// halo fill, red-black line relaxation (rank-1 corrected), residual + norm, restriction, prolongation.
// All 3-D fields are (k,j,i) with k fastest and a one-cell halo in j and i.
// Compiled with -fmad=false so that every expression rounds exactly like the CPU oracle's.
#include "kernels.cuh"
#include "device_utils.cuh"

#include <cstdio>
#include <cstdlib>

namespace mgb {

long long g_launches = 0;
static cudaError_t g_launch_err = cudaSuccess;
void note_launch_error(cudaError_t e) { if (e != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = e; }
cudaError_t take_launch_error() { const cudaError_t e = g_launch_err; g_launch_err = cudaSuccess; return e; }

// =====================================================================================================
// generic helpers
// =====================================================================================================
__global__ void k_fill(double* __restrict__ a, double v, size_t n) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; e < n; e += stride) a[e] = v;
}
__global__ void k_copy(double* __restrict__ d, const double* __restrict__ s, size_t n) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; e < n; e += stride) d[e] = s[e];
}
__global__ void k_axpy_all(double* __restrict__ y, const double* __restrict__ x, size_t n) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; e < n; e += stride) y[e] = y[e] + x[e];
}
static unsigned capped_blocks(size_t n, unsigned threads) {
  size_t b = (n + threads - 1) / threads;
  const size_t cap = (size_t)device_sm_count() * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}
void launch_fill(cudaStream_t st, double* a, double v, size_t n) { if (!n) return; launch_k(k_fill, capped_blocks(n, 256), 256, 0, st, a, v, n); }
void launch_copy(cudaStream_t st, double* d, const double* s, size_t n) { if (!n) return; launch_k(k_copy, capped_blocks(n, 256), 256, 0, st, d, s, n); }
void launch_axpy_all(cudaStream_t st, double* y, const double* x, size_t n) { if (!n) return; launch_k(k_axpy_all, capped_blocks(n, 256), 256, 0, st, y, x, n); }

// p += r on the halo ring (j in {0,ny+1} or i in {0,nx+1})
__global__ void k_add_halo_ring(double* __restrict__ p, const double* __restrict__ r, int nz, int ny, int nx) {
  const int ny2 = ny + 2;
  const size_t n_ew = (size_t)2 * ny2 * nz;          // i = 0 and i = nx+1, all j
  const size_t n_sn = (size_t)2 * nx * nz;           // j = 0 and j = ny+1, i = 1..nx
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_ew) {
    int side = (int)(e / ((size_t)ny2 * nz)); size_t rem = e % ((size_t)ny2 * nz);
    size_t idx = IDX3(0, 0, side ? nx + 1 : 0, nz, ny2) + rem;
    p[idx] = p[idx] + r[idx];
  } else if (e < n_ew + n_sn) {
    e -= n_ew;
    int side = (int)(e / ((size_t)nx * nz)); size_t rem = e % ((size_t)nx * nz);
    int i = 1 + (int)(rem / nz), k = (int)(rem % nz);
    size_t idx = IDX3(k, side ? ny + 1 : 0, i, nz, ny2);
    p[idx] = p[idx] + r[idx];
  }
}
void launch_add_halo_ring(cudaStream_t st, double* p, const double* r, int nz, int ny, int nx) {
  size_t n = (size_t)2 * (ny + 2) * nz + (size_t)2 * nx * nz;
  launch_k(k_add_halo_ring, blocks_for(n, 256), 256, 0, st, p, r, nz, ny, nx);
}

// =====================================================================================================
// halo, nh = 1   (reference: src/mg_mpi_exchange.f90:504-598 for the physical-edge rules)
// =====================================================================================================
__global__ void k_halo1_phys(double* __restrict__ a, int nzl, int ny, int nx, Halo1Flags f) {
  const int ny2 = ny + 2;
  const size_t n_ew = (size_t)ny * nzl, n_sn = (size_t)nx * nzl;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 2 * n_ew) {                       // west (side 0) / east (side 1): j = 1..ny contiguous in memory
    int side = (int)(e / n_ew); size_t rem = e % n_ew;
    if (side == 0 && !f.west) a[IDX3(0, 1, 0, nzl, ny2) + rem] = a[IDX3(0, 1, 1, nzl, ny2) + rem];
    if (side == 1 && !f.east) a[IDX3(0, 1, nx + 1, nzl, ny2) + rem] = a[IDX3(0, 1, nx, nzl, ny2) + rem];
    return;
  }
  e -= 2 * n_ew;
  if (e < 2 * n_sn) {                       // south (0) / north (1): i = 1..nx
    int side = (int)(e / n_sn); size_t rem = e % n_sn;
    int i = 1 + (int)(rem / nzl), k = (int)(rem % nzl);
    if (side == 0 && !f.south) a[IDX3(k, 0, i, nzl, ny2)] = a[IDX3(k, 1, i, nzl, ny2)];
    if (side == 1 && !f.north) a[IDX3(k, ny + 1, i, nzl, ny2)] = a[IDX3(k, ny, i, nzl, ny2)];
    return;
  }
  e -= 2 * n_sn;
  if (e < (size_t)4 * nzl) {                // corners with no neighbour at all on either side: mirror of the interior corner
    int c = (int)(e / nzl), k = (int)(e % nzl);
    if (c == 0 && !f.sw && !f.south && !f.west) a[IDX3(k, 0, 0, nzl, ny2)] = a[IDX3(k, 1, 1, nzl, ny2)];
    if (c == 1 && !f.se && !f.south && !f.east) a[IDX3(k, 0, nx + 1, nzl, ny2)] = a[IDX3(k, 1, nx, nzl, ny2)];
    if (c == 2 && !f.ne && !f.north && !f.east) a[IDX3(k, ny + 1, nx + 1, nzl, ny2)] = a[IDX3(k, ny, nx, nzl, ny2)];
    if (c == 3 && !f.nw && !f.north && !f.west) a[IDX3(k, ny + 1, 0, nzl, ny2)] = a[IDX3(k, ny, 1, nzl, ny2)];
  }
}
void launch_halo1_phys(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f) {
  size_t n = (size_t)2 * ny * nzl + (size_t)2 * nx * nzl + (size_t)4 * nzl;
  launch_k(k_halo1_phys, blocks_for(n, 256), 256, 0, st, a, nzl, ny, nx, f);
}
// corners next to a physical wall but with one lateral neighbour: copy from the halo row/column just received
__global__ void k_halo1_flagcorners(double* __restrict__ a, int nzl, int ny, int nx, Halo1Flags f) {
  const int ny2 = ny + 2;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 4 * nzl) return;
  int c = e / nzl, k = e % nzl;
  if (c == 0 && !f.sw) { if (f.south) a[IDX3(k, 0, 0, nzl, ny2)] = a[IDX3(k, 0, 1, nzl, ny2)]; else if (f.west) a[IDX3(k, 0, 0, nzl, ny2)] = a[IDX3(k, 1, 0, nzl, ny2)]; }
  if (c == 1 && !f.se) { if (f.south) a[IDX3(k, 0, nx + 1, nzl, ny2)] = a[IDX3(k, 0, nx, nzl, ny2)]; else if (f.east) a[IDX3(k, 0, nx + 1, nzl, ny2)] = a[IDX3(k, 1, nx + 1, nzl, ny2)]; }
  if (c == 2 && !f.ne) { if (f.north) a[IDX3(k, ny + 1, nx + 1, nzl, ny2)] = a[IDX3(k, ny + 1, nx, nzl, ny2)]; else if (f.east) a[IDX3(k, ny + 1, nx + 1, nzl, ny2)] = a[IDX3(k, ny, nx + 1, nzl, ny2)]; }
  if (c == 3 && !f.nw) { if (f.north) a[IDX3(k, ny + 1, 0, nzl, ny2)] = a[IDX3(k, ny + 1, 1, nzl, ny2)]; else if (f.west) a[IDX3(k, ny + 1, 0, nzl, ny2)] = a[IDX3(k, ny, 0, nzl, ny2)]; }
}
void launch_halo1_flagcorners(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f) {
  launch_k(k_halo1_flagcorners, blocks_for((size_t)4 * nzl, 128), 128, 0, st, a, nzl, ny, nx, f);
}

// Multi-GPU halo of an nh=1 field in two launches around the NCCL group:
//   pre : physical-edge mirrors + physical corners (as k_halo1_phys) and packing of the strided S / N rows to send
//   post: unpacking of the received S / N rows, the "flag" corners (taken from the received rows / columns) and, when
//         x1 != nullptr, the upkeep of the persistent k=1 plane x1 = a(1,:,:) on the whole halo ring
// E/W faces and corner columns are contiguous in the k-fastest layout and travel in place.
__global__ void k_halo1_pre(double* __restrict__ a, int nzl, int ny, int nx, Halo1Flags f, double* __restrict__ packS, double* __restrict__ packN) {
  const int ny2 = ny + 2;
  const size_t n_ew = (size_t)ny * nzl, n_sn = (size_t)nx * nzl;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 2 * n_ew) {
    int side = (int)(e / n_ew); size_t rem = e % n_ew;
    if (side == 0 && !f.west) a[IDX3(0, 1, 0, nzl, ny2) + rem] = a[IDX3(0, 1, 1, nzl, ny2) + rem];
    if (side == 1 && !f.east) a[IDX3(0, 1, nx + 1, nzl, ny2) + rem] = a[IDX3(0, 1, nx, nzl, ny2) + rem];
    return;
  }
  e -= 2 * n_ew;
  if (e < 2 * n_sn) {
    int side = (int)(e / n_sn); size_t rem = e % n_sn;
    int i = 1 + (int)(rem / nzl), k = (int)(rem % nzl);
    if (side == 0) { const double v = a[IDX3(k, 1, i, nzl, ny2)]; if (f.south) packS[rem] = v; else a[IDX3(k, 0, i, nzl, ny2)] = v; }
    else { const double v = a[IDX3(k, ny, i, nzl, ny2)]; if (f.north) packN[rem] = v; else a[IDX3(k, ny + 1, i, nzl, ny2)] = v; }
    return;
  }
  e -= 2 * n_sn;
  if (e < (size_t)4 * nzl) {
    int c = (int)(e / nzl), k = (int)(e % nzl);
    if (c == 0 && !f.sw && !f.south && !f.west) a[IDX3(k, 0, 0, nzl, ny2)] = a[IDX3(k, 1, 1, nzl, ny2)];
    if (c == 1 && !f.se && !f.south && !f.east) a[IDX3(k, 0, nx + 1, nzl, ny2)] = a[IDX3(k, 1, nx, nzl, ny2)];
    if (c == 2 && !f.ne && !f.north && !f.east) a[IDX3(k, ny + 1, nx + 1, nzl, ny2)] = a[IDX3(k, ny, nx, nzl, ny2)];
    if (c == 3 && !f.nw && !f.north && !f.west) a[IDX3(k, ny + 1, 0, nzl, ny2)] = a[IDX3(k, ny, 1, nzl, ny2)];
  }
}
__global__ void k_halo1_post(double* __restrict__ a, int nzl, int ny, int nx, Halo1Flags f, const double* __restrict__ recvS, const double* __restrict__ recvN,
                             double* __restrict__ x1) {
  const int ny2 = ny + 2;
  const size_t n_sn = (size_t)nx * nzl;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 2 * n_sn) {                        // received S / N rows -> halo rows j = 0 / ny+1
    int side = (int)(e / n_sn); size_t rem = e % n_sn;
    int i = 1 + (int)(rem / nzl), k = (int)(rem % nzl);
    if (side == 0) { if (f.south) a[IDX3(k, 0, i, nzl, ny2)] = recvS[rem]; if (x1 && k == 0) x1[(size_t)0 + (size_t)ny2 * i] = f.south ? recvS[rem] : a[IDX3(0, 0, i, nzl, ny2)]; }
    else { if (f.north) a[IDX3(k, ny + 1, i, nzl, ny2)] = recvN[rem]; if (x1 && k == 0) x1[(size_t)(ny + 1) + (size_t)ny2 * i] = f.north ? recvN[rem] : a[IDX3(0, ny + 1, i, nzl, ny2)]; }
    return;
  }
  e -= 2 * n_sn;
  if (e < (size_t)4 * nzl) {                 // corners: received in place, or derived from a received row / column next to a wall
    int c = (int)(e / nzl), k = (int)(e % nzl);
    const size_t lastS = (size_t)k + (size_t)nzl * (nx - 1), firstS = (size_t)k;          // index into recvS / recvN for i = nx / i = 1
    double v; size_t idx; size_t c2d;
    if (c == 0) { idx = IDX3(k, 0, 0, nzl, ny2); c2d = 0;
      if (!f.sw) { if (f.south) a[idx] = recvS[firstS]; else if (f.west) a[idx] = a[IDX3(k, 1, 0, nzl, ny2)]; } }
    else if (c == 1) { idx = IDX3(k, 0, nx + 1, nzl, ny2); c2d = (size_t)ny2 * (nx + 1);
      if (!f.se) { if (f.south) a[idx] = recvS[lastS]; else if (f.east) a[idx] = a[IDX3(k, 1, nx + 1, nzl, ny2)]; } }
    else if (c == 2) { idx = IDX3(k, ny + 1, nx + 1, nzl, ny2); c2d = (size_t)(ny + 1) + (size_t)ny2 * (nx + 1);
      if (!f.ne) { if (f.north) a[idx] = recvN[lastS]; else if (f.east) a[idx] = a[IDX3(k, ny, nx + 1, nzl, ny2)]; } }
    else { idx = IDX3(k, ny + 1, 0, nzl, ny2); c2d = (size_t)(ny + 1);
      if (!f.nw) { if (f.north) a[idx] = recvN[firstS]; else if (f.west) a[idx] = a[IDX3(k, ny, 0, nzl, ny2)]; } }
    v = a[idx];
    if (x1 && k == 0) x1[c2d] = v;
    return;
  }
  e -= (size_t)4 * nzl;
  if (x1 && e < (size_t)2 * ny) {            // E / W halo columns of the k=1 plane
    int side = (int)(e / ny), j = 1 + (int)(e % ny);
    const int i = side ? nx + 1 : 0;
    x1[(size_t)j + (size_t)ny2 * i] = a[IDX3(0, j, i, nzl, ny2)];
  }
}
// ---- the same exchange in ONE launch over peer memory (tagged NVLink stores), see kernels.cuh -------------------------
__device__ __forceinline__ void ll_store(uint4* dst, double v, unsigned tag) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"((unsigned)u), "r"(tag), "r"((unsigned)(u >> 32)), "r"(tag) : "memory");
}
// Poll a mailbox word until both tags carry the number of this exchange.  A neighbour that never writes must not hang the
// GPU: after B.timeout_ms of wall clock (globaltimer) the wait gives up, raises the mapped host word (the API call fails at
// its end) and the device-side abort flag (every later wait of this and of the following launches returns at once).
// `ok` turns false: the caller leaves the halo untouched instead of storing the stale word.
__device__ __forceinline__ double ll_wait(const uint4* src, unsigned tag, const P2PBox& B, bool& ok) {
  unsigned lo, t0, hi, t1;
  unsigned long long start = 0; int spins = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(src) : "memory");
    if (t0 == tag && t1 == tag) break;
    if (!ok) break;
    if (++spins == 64) { start = globaltimer_ns(); if (*(volatile int*)B.abort_flag) { ok = false; break; } }
    if (spins > 64) {
      __nanosleep(64);
      if ((spins & 1023) == 0) {
        if (*(volatile int*)B.abort_flag) { ok = false; break; }
        if (globaltimer_ns() - start > (unsigned long long)B.timeout_ms * 1000000ull) { *B.err = 1; *(volatile int*)B.abort_flag = 1; ok = false; break; }
      }
    }
  }
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
// One exchange (number s) of an nh = 1 field by `nthr` cooperating threads (this one is number tid0): usable as a kernel of its own
// (k_halo1_p2p) and from inside the fused relaxation kernels of small / mid-size levels, whose CTAs are co-resident.
// CG: values of `a` written by other CTAs of the same launch are read with ld.global.cg (the L1s are not coherent).
template <bool CG>
__device__ __forceinline__ void halo1_p2p_body(double* a, int nzl, int ny, int nx, const Halo1Flags& f, const P2PBox& B, double* x1,
                                               unsigned long long s, size_t tid0, size_t nthr) {
  auto LD = [](const double* q) { return CG ? __ldcg(q) : *q; };
  const int ny2 = ny + 2;
  const size_t n_ew = (size_t)ny * nzl, n_sn = (size_t)nx * nzl;
  const size_t row = (size_t)B.nzmax * nx, col = (size_t)B.nzmax * ny;
  const size_t offS = 0, offE = row, offN = row + col, offW = 2 * row + col, offC = 2 * row + 2 * col;
  const size_t stride = offC + 4 * (size_t)B.nzmax;
  const unsigned tag = (unsigned)s;
  const size_t par = (size_t)(s & 1ull) * stride;
  const size_t n1 = 2 * n_ew + 2 * n_sn + 4 * (size_t)nzl;
  // (1) physical mirrors in place, boundary faces into the neighbours' mailboxes
  for (size_t e = tid0; e < n1; e += nthr) {
    if (e < 2 * n_ew) {
      const int side = (int)(e / n_ew); const size_t rem = e % n_ew;
      if (side == 0) { const double v = LD(a + IDX3(0, 1, 1, nzl, ny2) + rem);
        if (f.west) ll_store(B.peer[3] + par + offE + rem, v, tag);
        else { a[IDX3(0, 1, 0, nzl, ny2) + rem] = v; if (x1 && rem % nzl == 0) x1[(size_t)(1 + rem / nzl)] = v; } }
      else { const double v = LD(a + IDX3(0, 1, nx, nzl, ny2) + rem);
        if (f.east) ll_store(B.peer[1] + par + offW + rem, v, tag);
        else { a[IDX3(0, 1, nx + 1, nzl, ny2) + rem] = v; if (x1 && rem % nzl == 0) x1[(size_t)(1 + rem / nzl) + (size_t)ny2 * (nx + 1)] = v; } }
      continue;
    }
    size_t e2 = e - 2 * n_ew;
    if (e2 < 2 * n_sn) {
      const int side = (int)(e2 / n_sn); const size_t rem = e2 % n_sn;
      const int i = 1 + (int)(rem / nzl), k = (int)(rem % nzl);
      if (side == 0) { const double v = LD(a + IDX3(k, 1, i, nzl, ny2));
        if (f.south) ll_store(B.peer[0] + par + offN + rem, v, tag);
        else { a[IDX3(k, 0, i, nzl, ny2)] = v; if (x1 && k == 0) x1[(size_t)ny2 * i] = v; } }
      else { const double v = LD(a + IDX3(k, ny, i, nzl, ny2));
        if (f.north) ll_store(B.peer[2] + par + offS + rem, v, tag);
        else { a[IDX3(k, ny + 1, i, nzl, ny2)] = v; if (x1 && k == 0) x1[(size_t)(ny + 1) + (size_t)ny2 * i] = v; } }
      continue;
    }
    e2 -= 2 * n_sn;
    const int c = (int)(e2 / nzl), k = (int)(e2 % nzl);
    const int jc = (c < 2) ? 1 : ny, ic = (c == 1 || c == 2) ? nx : 1;            // interior corner cell of corner c (SW,SE,NE,NW)
    const int jh = (c < 2) ? 0 : ny + 1, ih = (c == 1 || c == 2) ? nx + 1 : 0;    // its halo corner
    const double v = LD(a + IDX3(k, jc, ic, nzl, ny2));
    const int diag = c == 0 ? f.sw : c == 1 ? f.se : c == 2 ? f.ne : f.nw;
    const int sn = (c < 2) ? f.south : f.north, ew = (c == 1 || c == 2) ? f.east : f.west;
    if (diag) ll_store(B.peer[4 + c] + par + offC + (size_t)((c + 2) & 3) * B.nzmax + k, v, tag);   // lands in the opposite corner slot
    else if (!sn && !ew) { a[IDX3(k, jh, ih, nzl, ny2)] = v; if (x1 && k == 0) x1[(size_t)jh + (size_t)ny2 * ih] = v; }
  }
  // (2) poll my mailbox for the words of exchange s and unpack them
  const uint4* __restrict__ M = B.mine + par;
  bool ok = true;                                          // false after a timed-out wait: nothing more is stored
  for (size_t e = tid0; e < n1; e += nthr) {
    if (e < 2 * n_ew) {
      const int side = (int)(e / n_ew); const size_t rem = e % n_ew;
      if (side == 0 && f.west) { const double v = ll_wait(M + offW + rem, tag, B, ok); if (ok) { a[IDX3(0, 1, 0, nzl, ny2) + rem] = v; if (x1 && rem % nzl == 0) x1[(size_t)(1 + rem / nzl)] = v; } }
      if (side == 1 && f.east) { const double v = ll_wait(M + offE + rem, tag, B, ok); if (ok) { a[IDX3(0, 1, nx + 1, nzl, ny2) + rem] = v; if (x1 && rem % nzl == 0) x1[(size_t)(1 + rem / nzl) + (size_t)ny2 * (nx + 1)] = v; } }
      continue;
    }
    size_t e2 = e - 2 * n_ew;
    if (e2 < 2 * n_sn) {
      const int side = (int)(e2 / n_sn); const size_t rem = e2 % n_sn;
      const int i = 1 + (int)(rem / nzl), k = (int)(rem % nzl);
      if (side == 0 && f.south) { const double v = ll_wait(M + offS + rem, tag, B, ok); if (ok) { a[IDX3(k, 0, i, nzl, ny2)] = v; if (x1 && k == 0) x1[(size_t)ny2 * i] = v; } }
      if (side == 1 && f.north) { const double v = ll_wait(M + offN + rem, tag, B, ok); if (ok) { a[IDX3(k, ny + 1, i, nzl, ny2)] = v; if (x1 && k == 0) x1[(size_t)(ny + 1) + (size_t)ny2 * i] = v; } }
      continue;
    }
    e2 -= 2 * n_sn;
    // corners (src/mg_mpi_exchange.f90:540-598,720-743): diagonal neighbour > S/N neighbour (end of the received row) >
    // W/E neighbour (end of the received column) > physical mirror (done in step 1)
    const int c = (int)(e2 / nzl), k = (int)(e2 % nzl);
    const int jh = (c < 2) ? 0 : ny + 1, ih = (c == 1 || c == 2) ? nx + 1 : 0;
    const int diag = c == 0 ? f.sw : c == 1 ? f.se : c == 2 ? f.ne : f.nw;
    const int sn = (c < 2) ? f.south : f.north, ew = (c == 1 || c == 2) ? f.east : f.west;
    const size_t rowend = (size_t)k + (size_t)nzl * ((c == 1 || c == 2) ? nx - 1 : 0);   // i = nx or i = 1 of a received row
    const size_t colend = (size_t)k + (size_t)nzl * ((c < 2) ? 0 : ny - 1);              // j = 1 or j = ny of a received column
    double v;
    if (diag) v = ll_wait(M + offC + (size_t)c * B.nzmax + k, tag, B, ok);
    else if (sn) v = ll_wait(M + ((c < 2) ? offS : offN) + rowend, tag, B, ok);
    else if (ew) v = ll_wait(M + ((c == 1 || c == 2) ? offE : offW) + colend, tag, B, ok);
    else continue;
    if (!ok) continue;
    a[IDX3(k, jh, ih, nzl, ny2)] = v;
    if (x1 && k == 0) x1[(size_t)jh + (size_t)ny2 * ih] = v;
  }
}
__global__ void __launch_bounds__(256) k_halo1_p2p(double* __restrict__ a, int nzl, int ny, int nx, Halo1Flags f, P2PBox B, double* __restrict__ x1) {
  const unsigned long long s = *B.seq + 1;                 // number of this exchange; *B.seq is advanced by the last block to leave
  halo1_p2p_body<false>(a, nzl, ny, nx, f, B, x1, s, (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
  // every block has read *B.seq by now; the last one to leave advances it for the next launch
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(B.done, 1u) == gridDim.x - 1) { *B.done = 0; *B.seq = s; }
}
void launch_halo1_p2p(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f, const P2PBox& B, double* x1) {
  const size_t n = (size_t)2 * ny * nzl + (size_t)2 * nx * nzl + (size_t)4 * nzl;
  static const unsigned cap = []() {                                 // all blocks must be co-resident (the kernel spins on its peers)
    int per = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_halo1_p2p, 256, 0);
    if (per > 4) per = 4;
    return (unsigned)(per > 0 ? per : 1) * (unsigned)device_sm_count();
  }();
  unsigned nb = blocks_for(n, 256); if (nb > cap) nb = cap;
  launch_k(k_halo1_p2p, nb, 256, 0, st, a, nzl, ny, nx, f, B, x1);
}
void launch_halo1_pre(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f, double* packS, double* packN) {
  size_t n = (size_t)2 * ny * nzl + (size_t)2 * nx * nzl + (size_t)4 * nzl;
  launch_k(k_halo1_pre, blocks_for(n, 256), 256, 0, st, a, nzl, ny, nx, f, packS, packN);
}
void launch_halo1_post(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f, const double* recvS, const double* recvN, double* x1) {
  size_t n = (size_t)2 * nx * nzl + (size_t)4 * nzl + (size_t)2 * ny;
  launch_k(k_halo1_post, blocks_for(n, 256), 256, 0, st, a, nzl, ny, nx, f, recvS, recvN, x1);
}

// =====================================================================================================
// halo, nh = 2 (zr, zw)   (reference: src/mg_mpi_exchange.f90:947-1101)
// array a(nzl, -1:ny+2, -1:nx+2): index k + nzl*((j+1) + (ny+4)*(i+1))
// =====================================================================================================
#define IDXH2(k, j, i, nzl, ny4) ((size_t)(k) + (size_t)(nzl) * ((size_t)((j) + 1) + (size_t)(ny4) * (size_t)((i) + 1)))
__global__ void k_halo2_phys(double* __restrict__ a, int nzl, int ny, int nx, Halo1Flags f) {
  const int ny4 = ny + 4;
  const size_t n_ew = (size_t)ny * nzl, n_sn = (size_t)nx * nzl;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 2 * n_ew) {
    int side = (int)(e / n_ew); size_t rem = e % n_ew; int j = 1 + (int)(rem / nzl), k = (int)(rem % nzl);
    if (side == 0 && !f.west) { double A = a[IDXH2(k, j, 1, nzl, ny4)], B = a[IDXH2(k, j, 2, nzl, ny4)];
      a[IDXH2(k, j, 0, nzl, ny4)] = A; a[IDXH2(k, j, -1, nzl, ny4)] = 2.0 * A - B; }
    if (side == 1 && !f.east) { double A = a[IDXH2(k, j, nx, nzl, ny4)], B = a[IDXH2(k, j, nx - 1, nzl, ny4)];
      a[IDXH2(k, j, nx + 1, nzl, ny4)] = A; a[IDXH2(k, j, nx + 2, nzl, ny4)] = 2.0 * A - B; }
    return;
  }
  e -= 2 * n_ew;
  if (e < 2 * n_sn) {
    int side = (int)(e / n_sn); size_t rem = e % n_sn; int i = 1 + (int)(rem / nzl), k = (int)(rem % nzl);
    if (side == 0 && !f.south) { double A = a[IDXH2(k, 1, i, nzl, ny4)], B = a[IDXH2(k, 2, i, nzl, ny4)];
      a[IDXH2(k, 0, i, nzl, ny4)] = A; a[IDXH2(k, -1, i, nzl, ny4)] = 2.0 * A - B; }
    if (side == 1 && !f.north) { double A = a[IDXH2(k, ny, i, nzl, ny4)], B = a[IDXH2(k, ny - 1, i, nzl, ny4)];
      a[IDXH2(k, ny + 1, i, nzl, ny4)] = A; a[IDXH2(k, ny + 2, i, nzl, ny4)] = 2.0 * A - B; }
    return;
  }
  e -= 2 * n_sn;
  if (e < (size_t)4 * nzl) {
    int c = (int)(e / nzl), k = (int)(e % nzl);
    for (int aa = 1; aa <= 2; ++aa) for (int cc = 1; cc <= 2; ++cc) {
      if (c == 0 && !f.sw && !f.south && !f.west) a[IDXH2(k, 1 - aa, 1 - cc, nzl, ny4)] = a[IDXH2(k, aa, cc, nzl, ny4)];
      if (c == 1 && !f.se && !f.south && !f.east) a[IDXH2(k, 1 - aa, nx + cc, nzl, ny4)] = a[IDXH2(k, aa, nx + 1 - cc, nzl, ny4)];
      if (c == 2 && !f.ne && !f.north && !f.east) a[IDXH2(k, ny + aa, nx + cc, nzl, ny4)] = a[IDXH2(k, ny + 1 - aa, nx + 1 - cc, nzl, ny4)];
      if (c == 3 && !f.nw && !f.north && !f.west) a[IDXH2(k, ny + aa, 1 - cc, nzl, ny4)] = a[IDXH2(k, ny + 1 - aa, cc, nzl, ny4)];
    }
  }
}
void launch_halo2_phys(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f) {
  size_t n = (size_t)2 * ny * nzl + (size_t)2 * nx * nzl + (size_t)4 * nzl;
  launch_k(k_halo2_phys, blocks_for(n, 256), 256, 0, st, a, nzl, ny, nx, f);
}
__global__ void k_halo2_flagcorners(double* __restrict__ a, int nzl, int ny, int nx, Halo1Flags f) {
  const int ny4 = ny + 4;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 4 * nzl) return;
  int c = e / nzl, k = e % nzl;
  for (int aa = 1; aa <= 2; ++aa) for (int cc = 1; cc <= 2; ++cc) {
    if (c == 0 && !f.sw) { if (f.south) a[IDXH2(k, 1 - aa, 1 - cc, nzl, ny4)] = a[IDXH2(k, 1 - aa, cc, nzl, ny4)]; else if (f.west) a[IDXH2(k, 1 - aa, 1 - cc, nzl, ny4)] = a[IDXH2(k, aa, 1 - cc, nzl, ny4)]; }
    if (c == 1 && !f.se) { if (f.south) a[IDXH2(k, 1 - aa, nx + cc, nzl, ny4)] = a[IDXH2(k, 1 - aa, nx + 1 - cc, nzl, ny4)]; else if (f.east) a[IDXH2(k, 1 - aa, nx + cc, nzl, ny4)] = a[IDXH2(k, aa, nx + cc, nzl, ny4)]; }
    if (c == 2 && !f.ne) { if (f.north) a[IDXH2(k, ny + aa, nx + cc, nzl, ny4)] = a[IDXH2(k, ny + aa, nx + 1 - cc, nzl, ny4)]; else if (f.east) a[IDXH2(k, ny + aa, nx + cc, nzl, ny4)] = a[IDXH2(k, ny + 1 - aa, nx + cc, nzl, ny4)]; }
    if (c == 3 && !f.nw) { if (f.north) a[IDXH2(k, ny + aa, 1 - cc, nzl, ny4)] = a[IDXH2(k, ny + aa, cc, nzl, ny4)]; else if (f.west) a[IDXH2(k, ny + aa, 1 - cc, nzl, ny4)] = a[IDXH2(k, ny + 1 - aa, 1 - cc, nzl, ny4)]; }
  }
}
void launch_halo2_flagcorners(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f) {
  launch_k(k_halo2_flagcorners, blocks_for((size_t)4 * nzl, 128), 128, 0, st, a, nzl, ny, nx, f);
}

// pack a rectangular (j0..j0+nrows-1) x (i0..i0+ncols-1) patch of an array with halo width nh into a dense buffer
// buf(k, jr, ic), k fastest  -- used for S/N faces and corners of the inter-GPU exchange
__global__ void k_pack_rows(const double* __restrict__ a, double* __restrict__ buf, int nzl, int nyh, int nh, int j0, int nrows, int i0, int ncols) {
  size_t n = (size_t)nzl * nrows * ncols;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int k = (int)(e % nzl); size_t t = e / nzl; int jr = (int)(t % nrows), ic = (int)(t / nrows);
  buf[e] = a[(size_t)k + (size_t)nzl * ((size_t)(j0 + jr - 1 + nh) + (size_t)nyh * (size_t)(i0 + ic - 1 + nh))];
}
__global__ void k_unpack_rows(double* __restrict__ a, const double* __restrict__ buf, int nzl, int nyh, int nh, int j0, int nrows, int i0, int ncols) {
  size_t n = (size_t)nzl * nrows * ncols;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int k = (int)(e % nzl); size_t t = e / nzl; int jr = (int)(t % nrows), ic = (int)(t / nrows);
  a[(size_t)k + (size_t)nzl * ((size_t)(j0 + jr - 1 + nh) + (size_t)nyh * (size_t)(i0 + ic - 1 + nh))] = buf[e];
}
// j0,i0 are Fortran indices (may be <= 0 for halo cells); nyh = ny + 2*nh
void launch_pack_rows(cudaStream_t st, const double* a, double* buf, int nzl, int ny, int nx, int nh, int j0, int nrows, int i0, int ncols) {
  size_t n = (size_t)nzl * nrows * ncols; (void)nx;
  launch_k(k_pack_rows, blocks_for(n, 256), 256, 0, st, a, buf, nzl, ny + 2 * nh, nh, j0, nrows, i0, ncols);
}
void launch_unpack_rows(cudaStream_t st, double* a, const double* buf, int nzl, int ny, int nx, int nh, int j0, int nrows, int i0, int ncols) {
  size_t n = (size_t)nzl * nrows * ncols; (void)nx;
  launch_k(k_unpack_rows, blocks_for(n, 256), 256, 0, st, a, buf, nzl, ny + 2 * nh, nh, j0, nrows, i0, ncols);
}

// =====================================================================================================
// relaxation
// =====================================================================================================
// x1(j,i) = p(1,j,i) for every column incl. halo: snapshot of the bottom plane taken before a colour is updated
__global__ void k_snapshot_x1(const double* __restrict__ p, double* __restrict__ x1, int nz, size_t ncols) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncols) x1[c] = p[c * nz];
}
void launch_snapshot_x1(cudaStream_t st, const LevelView& L) {
  size_t ncols = (size_t)(L.ny + 2) * (L.nx + 2);
  launch_k(k_snapshot_x1, blocks_for(ncols, 256), 256, 0, st, L.p, L.x1, L.nz, ncols);
}

static void relax_tiling(const LevelView& L, const ColourSet& cs, int& TI, int& TJ, int& nti, int& ntj, size_t& smem) {
  const int nyc = cs.jrule == 2 ? 1 : (L.ny + 1) / 2; // columns of the set per visited i (upper bound)
  const int nvi = (L.nx - cs.istart) / cs.istep + 1; // visited i
  static const int tile_pts = []() { const char* e = std::getenv("MGB_A4_TILE"); const int v = e ? std::atoi(e) : 1024; return v < 256 ? 256 : v; }();
  int tc_max = tile_pts / L.nz;                      // columns per tile: ~1024 points (two 25 KB buffers per CTA leave most of the L1)
  static const int tc_floor = []() { const char* e = std::getenv("MGB_A4_TCMIN"); const int v = e ? std::atoi(e) : 8; return v < 4 ? 4 : v; }();
  if (tc_max < tc_floor) tc_max = tc_floor;          // nz = 128: 8 columns (3 CTAs / SM); 16 would need 99 KB per CTA (2 CTAs / SM)
  if (tc_max > 1024) tc_max = 1024;
  while (tc_max > 1 && (size_t)tc_max * (L.nz + 1) * 3 * sizeof(double) > 100 * 1024) tc_max /= 2;
  while (tc_max * L.nz > 256 && ((size_t)nyc * nvi + tc_max - 1) / tc_max < 296) tc_max /= 2;   // small level: more, smaller tiles
  TJ = nyc < tc_max ? nyc : tc_max;
  TI = tc_max / TJ; if (TI < 1) TI = 1; if (TI > nvi) TI = nvi;
  ntj = (nyc + TJ - 1) / TJ; nti = (nvi + TI - 1) / TI;
  smem = ((((size_t)TI * TJ * (L.nz + 1) + 1) & ~(size_t)1) + (size_t)2 * TI * TJ * (L.nz + 1)) * sizeof(double);
}
// Kernel A, warp-specialised persistent variant.  A CTA of 8 warps loops over tiles with two shared-memory tile
// buffers: warps 0..6 ("producers") run phase 1 (27 loads per lane in flight, rhs/bet/ud -> smem buffer) and the coalesced
// write-back of the tile solved two steps earlier, warp 7 ("solver") only runs Thomas (one lane per column).  Named
// barriers (bar.sync / bar.arrive) hand the buffers over, so producers never idle behind the serial Thomas phase and
// the memory pipe is fed continuously.  Same arithmetic, operation by operation, as k_relax_A.

// NZC > 0: nz known at compile time (power of two) so that every k/j offset of the 27 operand addresses is an immediate
// and e -> (k, column) is a shift; 0: generic.  tj_shift >= 0 when TJ is a power of two.  All element indices are 32-bit
// (a level holds < 2^31 points per GPU): one IMAD.WIDE per operand array instead of 64-bit add chains.
template <int MINB, int NZC>
__global__ void __launch_bounds__(256, MINB) k_relax_A4(LevelView L, ColourSet cs, int TI, int TJ, int ntj, int ntiles, int tj_shift, int thv) {
  extern __shared__ __align__(16) double sm[];
  const int nz = NZC ? NZC : L.nz, ny2 = L.ny + 2, pitch = nz + 1;
  const int TC = TI * TJ;
  const int rhs_sz = (TC * pitch + 1) & ~1;                 // the pair array behind it must start on a 16-byte boundary
  const int bufsz = rhs_sz + 2 * TC * pitch;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sj = nz; const int si = nz * ny2;
  const int nelem = TC * nz;
  constexpr int NPROD = 7 * 32;
  // element e of a tile -> (k, column c) -> (ci, cj) -> (i, j); returns false for padding columns
  auto locate = [&](int e, int tile_i, int tile_j, int& k, int& c, int& i, int& j) -> bool {
    if (NZC) { k = e & (NZC - 1); c = e / (NZC ? NZC : 1); } else { k = e % nz; c = e / nz; }
    int ci, cj;
    if (tj_shift >= 0) { ci = c >> tj_shift; cj = c & (TJ - 1); } else { ci = c / TJ; cj = c - ci * TJ; }
    i = cs.istart + (tile_i * TI + ci) * cs.istep;
    if (i > L.nx) return false;
    j = colour_jstart(cs, i) + 2 * (tile_j * TJ + cj);
    return j <= L.ny && j >= 1;
  };
  // barrier ids: 1,2 = buffer b full ; 3,4 = buffer b empty
  if (warp < 7) {
    const double* __restrict__ p = L.p; const double* __restrict__ b = L.b;
    const double* __restrict__ c2 = L.cA[1]; const double* __restrict__ c3 = L.cA[2]; const double* __restrict__ c4 = L.cA[3];
    const double* __restrict__ c5 = L.cA[4]; const double* __restrict__ c6 = L.cA[5]; const double* __restrict__ c7 = L.cA[6];
    const double* __restrict__ c8 = L.cA[7]; const double* __restrict__ bet = L.bet;
    double* __restrict__ pw = L.p;
    const bool mirror = cs.fuse_halo == 1 && cs.bottom_mode != 1;
    auto writeback = [&](const double* s_x, int t) {         // coalesced along k; final values also go to the Neumann halo
      const int t_j = t % ntj, t_i = t / ntj;
      for (int e = threadIdx.x; e < nelem; e += NPROD) {
        int k, c, i, j;
        if (!locate(e, t_i, t_j, k, c, i, j)) continue;
        const int idx = k + nz * (j + ny2 * i);
        const double v = s_x[c * pitch + k];
        if (!mirror) pw[idx] = v;
        else store_mirrored(pw, (size_t)idx, v, j, i, L.ny, L.nx, (size_t)sj, (size_t)si, true);
      }
    };
    int n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
      const int bsel = n & 1;
      double* s_rhs = sm + bsel * bufsz; double2* s_bu = reinterpret_cast<double2*>(s_rhs + rhs_sz);   // {pivot, off-diagonal} pairs
      if (n >= 2) {                                         // tile n-2 sits solved in this buffer: write it back, then reuse the buffer
        named_bar_sync(3 + bsel, 256);
        writeback(s_rhs, tile - 2 * (int)gridDim.x);
        named_bar_sync(5, NPROD);
      }
      const int tile_j = tile % ntj, tile_i = tile / ntj;
      for (int e = threadIdx.x; e < nelem; e += NPROD) {
        int k, c, i, j;
        if (!locate(e, tile_i, tile_j, k, c, i, j)) continue;
        const int idx = k + nz * (j + ny2 * i);
        const int im = idx - si, ip = idx + si;             // same k, j in rows i-1 / i+1
        const double bb = b[idx], bt = bet[idx], udv = c2[idx + 1];
        const double a3 = c3[idx],           q3 = p[idx - sj + 1];
        const double a3n = c3[idx + sj - 1], q3n = p[idx + sj - 1];
        const double a4 = c4[idx],           q4 = p[idx - sj];
        const double a4n = c4[idx + sj],     q4n = p[idx + sj];
        const double a5 = c5[idx],           q5 = p[idx - sj - 1];
        const double a5n = c5[idx + sj + 1], q5n = p[idx + sj + 1];
        const double a6 = c6[idx],           q6 = p[im + 1];
        const double a6n = c6[ip - 1],       q6n = p[ip - 1];
        const double a7 = c7[idx],           q7 = p[im];
        const double a7n = c7[ip],           q7n = p[ip];
        const double a8 = c8[idx],           q8 = p[im - 1];
        const double a8n = c8[ip + 1],       q8n = p[ip + 1];
        const bool lo = (k > 0), hi = (k < nz - 1);
        double acc = bb;
        if (hi) acc = acc - a3 * q3;
        if (lo) acc = acc - a3n * q3n;
        acc = acc - a4 * q4;
        acc = acc - a4n * q4n;
        if (lo) acc = acc - a5 * q5;
        if (hi) acc = acc - a5n * q5n;
        if (hi) acc = acc - a6 * q6;
        if (lo) acc = acc - a6n * q6n;
        acc = acc - a7 * q7;
        acc = acc - a7n * q7n;
        if (lo) acc = acc - a8 * q8;
        if (hi) acc = acc - a8n * q8n;
        if (k == 0) {
          if (cs.bottom_mode == 1) {
            const int c2d = j + ny2 * i;
            acc = acc - L.c5b[c2d - 1 + ny2] * L.x1[c2d - 1 + ny2];
            acc = acc - L.c8b[c2d + 1 + ny2] * L.x1[c2d + 1 + ny2];
          } else if (cs.bottom_mode == 2) {
            acc = acc - a5 * p[im + sj];
            acc = acc - c5[ip - sj] * p[ip - sj];
            acc = acc - a8 * p[im - sj];
            acc = acc - c8[ip + sj] * p[ip + sj];
          }
        }
        const int so = c * pitch + k;
        s_rhs[so] = acc;
        s_bu[so] = make_double2(bt, hi ? udv : 0.0);
      }
      __threadfence_block();
      named_bar_arrive(1 + bsel, 256);                      // buffer full
    }
    // drain: the last (up to two) tiles of this CTA
    for (int m = (n >= 2 ? n - 2 : 0); m < n; ++m) {
      const int bsel = m & 1;
      named_bar_sync(3 + bsel, 256);
      writeback(sm + bsel * bufsz, (int)blockIdx.x + m * (int)gridDim.x);
    }
  } else {
    int n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
      const int bsel = n & 1;
      double* s_rhs = sm + bsel * bufsz; const double2* s_bu = reinterpret_cast<const double2*>(s_rhs + rhs_sz);
      named_bar_sync(1 + bsel, 256);                        // wait for the producers
      const int tile_j = tile % ntj, tile_i = tile / ntj;
      for (int c = lane; c < TC; c += 32) {
        const int ci = c / TJ, cj = c % TJ;
        const int i = cs.istart + (tile_i * TI + ci) * cs.istep;
        if (i > L.nx) continue;
        const int j = colour_jstart(cs, i) + 2 * (tile_j * TJ + cj);
        if (j > L.ny || j < 1) continue;
        double* x = s_rhs + c * pitch; const double2* bu = s_bu + c * pitch;
        const double xm = (NZC && thv) ? thomas_smem<NZC ? NZC : 2>(x, bu) : thomas_smem_plain(x, bu, nz);
        if (cs.bottom_mode == 1) L.q1[(size_t)j + (size_t)ny2 * i] = xm;
      }
      __syncwarp();
      __threadfence_block();
      named_bar_arrive(3 + bsel, 256);                      // solved: the producers write it back
    }
  }
}
// 0: plain Thomas loops (shared-memory loads on the dependent chain) ; 1: operands of the next four levels prefetched.
// Measured inside the kernels (profiles/summary_r02.md): the prefetching form issues its loads in bursts that the in-order
// warp cannot overlap with its own FP64 chain, and is slower where the Thomas warp is critical (nz = 128) -> default 0.
int thomas_variant() { static const int v = []() { const char* e = std::getenv("MGB_THOMAS"); return e ? std::atoi(e) : 0; }(); return v; }
template <int MINB>
static void launch_A4_nz(cudaStream_t st, const LevelView& L, const ColourSet& cs, int grid, size_t smem, int TI, int TJ, int ntj, int ntiles, int tj_shift) {
  switch (L.nz) {
    case 128: launch_k(k_relax_A4<MINB, 128>, grid, 256, smem, st, L, cs, TI, TJ, ntj, ntiles, tj_shift, thomas_variant()); break;
    case 64: launch_k(k_relax_A4<MINB, 64>, grid, 256, smem, st, L, cs, TI, TJ, ntj, ntiles, tj_shift, thomas_variant()); break;
    case 32: launch_k(k_relax_A4<MINB, 32>, grid, 256, smem, st, L, cs, TI, TJ, ntj, ntiles, tj_shift, thomas_variant()); break;
    case 16: launch_k(k_relax_A4<MINB, 16>, grid, 256, smem, st, L, cs, TI, TJ, ntj, ntiles, tj_shift, thomas_variant()); break;
    default: launch_k(k_relax_A4<MINB, 0>, grid, 256, smem, st, L, cs, TI, TJ, ntj, ntiles, tj_shift, thomas_variant()); break;
  }
}
template <int MINB> static void a4_attrs() {
  cudaFuncSetAttribute(k_relax_A4<MINB, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_relax_A4<MINB, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_relax_A4<MINB, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_relax_A4<MINB, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_relax_A4<MINB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
void launch_relax_A(cudaStream_t st, const LevelView& L, const ColourSet& cs) {
  int TI, TJ, nti, ntj; size_t smem1;
  relax_tiling(L, cs, TI, TJ, nti, ntj, smem1);
  const int ntiles = nti * ntj;
  static const int minb = []() { const char* e = std::getenv("MGB_A4_CTAS"); const int v = e ? std::atoi(e) : 3; return v < 2 ? 2 : (v > 3 ? 3 : v); }();
  const int cap = device_sm_count() * minb;
  const int grid = ntiles < cap ? ntiles : cap;        // persistent: minb CTAs per SM
  int tj_shift = -1; for (int q = 0; q < 12; ++q) if ((1 << q) == TJ) tj_shift = q;
  if (minb == 2) launch_A4_nz<2>(st, L, cs, grid, 2 * smem1, TI, TJ, ntj, ntiles, tj_shift);
  else launch_A4_nz<3>(st, L, cs, grid, 2 * smem1, TI, TJ, ntj, ntiles, tj_shift);
}

// Kernel B: the sequential part of the red-black bottom coupling, on the 2-D k=1 plane only.
// For i ascending: s(j,i) = -(c5b(j,i) x1(j+1,i-1) + c8b(j,i) x1(j-1,i-1)) ; x1(j,i) = q1(j,i) + s(j,i) g1(j,i)
// One CTA marches over i.  x1(:,i-1) lives in shared memory (two ping-pong rows); the five operand rows of the
// coming steps (c5b, c8b, q1, g1 and the old x1 for the halo entries) are streamed into a D-deep shared-memory
// ring with cp.async (LDGSTS) so that no global-memory latency sits on the dependency chain.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int D, bool ONE_WARP>
__global__ void __launch_bounds__(1024) k_relax_B(LevelView L, ColourSet cs) {
  extern __shared__ __align__(16) double sm[];
  const int ny = L.ny, nx = L.nx, ny2 = ny + 2, nyc = (ny + 1) / 2;
  // this CTA commits s(:, c0..c1); it starts W columns earlier from the OLD x1 (contraction: error ~ rho^W)
  const int c0 = 1 + blockIdx.x * L.rb_chunk, c1 = min(nx, c0 + L.rb_chunk - 1), i0 = max(1, c0 - L.rb_W);
  const int nsteps = c1 - i0 + 1;
  double* row0 = sm;                 // x1(:, i-1)
  double* row1 = sm + ny2;           // x1(:, i)
  double* ring = sm + 2 * ny2;       // [D][5][ny2]
  const double* planes[5] = {L.c5b, L.c8b, L.q1, L.g1, L.x1};
  const int chunks_per_row = ny2 / 2;              // 16-byte chunks (ny2 is even)
  auto issue = [&](int step) {                     // step = 0.. ; column i0 + step
    if (step < nsteps) {
      double* slot = ring + (size_t)(step % D) * 5 * ny2;
      const size_t col = (size_t)ny2 * (i0 + step);
#pragma unroll
      for (int a = 0; a < 5; ++a)
        for (int q = threadIdx.x; q < chunks_per_row; q += blockDim.x) cp_async16(slot + a * ny2 + 2 * q, planes[a] + col + 2 * q);
    }
    cp_async_commit();
  };
  auto barrier = [&]() { if (ONE_WARP) __syncwarp(); else __syncthreads(); };
  for (int j = threadIdx.x; j < ny2; j += blockDim.x) row0[j] = L.x1[(size_t)ny2 * (i0 - 1) + j];
  for (int d = 0; d <= D - 2; ++d) issue(d);
  double* prev = row0; double* cur = row1;
  for (int st = 0; st < nsteps; ++st) {
    const int i = i0 + st;
    cp_async_wait<D - 2>();
    barrier();
    issue(st + D - 1);
    const double* slot = ring + (size_t)(st % D) * 5 * ny2;
    const double *r5 = slot, *r8 = slot + ny2, *rq = slot + 2 * ny2, *rg = slot + 3 * ny2, *rx = slot + 4 * ny2;
    const int js = 1 + ((i + cs.rb) & 1);
    const size_t col = (size_t)ny2 * i;
    for (int jc = threadIdx.x; jc < nyc; jc += blockDim.x) {
      const int j = js + 2 * jc;
      if (j <= ny) {
        const double sv = -(r5[j] * prev[j + 1] + r8[j] * prev[j - 1]);
        cur[j] = rq[j] + sv * rg[j];
        if (i >= c0) L.s[col + j] = sv;
      }
    }
    if (threadIdx.x == 0) { cur[0] = rx[0]; cur[ny + 1] = rx[ny + 1]; }
    double* t = prev; prev = cur; cur = t;
  }
  cp_async_wait<0>();
}
void launch_relax_B(cudaStream_t st, const LevelView& L, const ColourSet& cs) {
  const int ny2 = L.ny + 2, nyc = (L.ny + 1) / 2;
  int threads = 32; while (threads < nyc && threads < 1024) threads *= 2;
  // ring depth by shared-memory budget (<= ~200 KB): rows are ny2 doubles, 5 rows per step
  int D = 8;
  while (D > 2 && (size_t)(2 + 5 * D) * ny2 * sizeof(double) > 200 * 1024) D /= 2;
  const size_t smem = (size_t)(2 + 5 * D) * ny2 * sizeof(double);
  const int nchunks = (L.nx + L.rb_chunk - 1) / L.rb_chunk;
  if (threads == 32 && D == 8) launch_k(k_relax_B<8, true>, nchunks, 32, smem, st, L, cs);
  else if (D == 8) launch_k(k_relax_B<8, false>, nchunks, threads, smem, st, L, cs);
  else if (D == 4) launch_k(k_relax_B<4, false>, nchunks, threads, smem, st, L, cs);
  else launch_k(k_relax_B<2, false>, nchunks, threads, smem, st, L, cs);
}

// Kernel C: p(:,j,i) += s(j,i) * g(:,j,i) on the columns of the colour
__global__ void __launch_bounds__(256) k_relax_C(LevelView L, ColourSet cs, int nyc) {
  const int nz = L.nz, ny2 = L.ny + 2;
  // one i per blockIdx.y ; a block covers 4*256 consecutive (k, colour column) elements, four per thread (loads first)
  const int i = cs.istart + blockIdx.y * cs.istep;
  const int js = colour_jstart(cs, i);
  const int n = nyc * nz;
  const int e0 = blockIdx.x * 1024 + threadIdx.x;
  int idx[4], jj[4]; double pv[4], gv[4], sv[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = e0 + 256 * u;
    const int jc = e / nz, k = e - jc * nz;
    jj[u] = js + 2 * jc;
    const bool ok = e < n && jj[u] <= L.ny;
    idx[u] = ok ? k + nz * (jj[u] + ny2 * i) : -1;
    if (ok) { pv[u] = L.p[idx[u]]; gv[u] = L.g[idx[u]]; sv[u] = L.s[jj[u] + ny2 * i]; }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (idx[u] < 0) continue;
    const int j = jj[u];
    const double v = pv[u] + sv[u] * gv[u];
    const bool k0 = (idx[u] - nz * (j + ny2 * i)) == 0;
    const bool edge = (cs.fuse_halo == 1) && (i == 1 || i == L.nx || j <= 2 || j >= L.ny - 1);
    if (!edge) {
      L.p[idx[u]] = v;
      if (k0 && cs.fuse_halo) L.x1[j + ny2 * i] = v;   // fuse_halo == 2: keep x1 current, halo done by the exchange
    } else {
      store_mirrored(L.p, (size_t)idx[u], v, j, i, L.ny, L.nx, (size_t)nz, (size_t)nz * ny2, true);
      if (k0) store_mirrored(L.x1, (size_t)j + (size_t)ny2 * i, v, j, i, L.ny, L.nx, 1, (size_t)ny2, true);
    }
  }
}
// Kernels B and C fused (neighbour-free levels): a CTA owns one chunk of rows (c0..c1) and a block of 2*TJ columns.
// It first evaluates the chunked bottom recurrence on the trapezoid of the k=1 plane that its block depends on (warm-up
// rows i0..c0-1 and E = c1-i0 extra columns on each side, so the values it commits are bit-identical to kernel B's),
// entirely in shared memory, then streams p += s g over its rows with the Neumann mirrors.  The k=1 plane is double
// buffered (x1old is read, x1new is written completely) because other CTAs still need the old values.
// Saves the B launch and its serial latency on every colour.
// CG: read everything another CTA of the same launch may have written (x1old, q1, p) with ld.global.cg -- used by the
// cooperative mid-level kernel, where the L1s are not coherent; the stand-alone kernel reads them the plain way.
template <bool CG> __device__ __forceinline__ double ld_shared_data(const double* q) { return CG ? __ldcg(q) : *q; }
template <bool CG>
__device__ __forceinline__ void bc_block(const LevelView& L, int rb, const double* __restrict__ x1old, double* __restrict__ x1new, int TJ,
                                         int bx, int by, double* sm) {
  const int nx = L.nx, ny = L.ny, ny2 = ny + 2, nz = L.nz;
  const int c0 = 1 + by * L.rb_chunk, c1 = min(nx, c0 + L.rb_chunk - 1), i0 = max(1, c0 - L.rb_W);
  const int E = c1 - i0;
  const int jA = 2 * bx * TJ + 1, jB = min(ny, jA + 2 * TJ - 1);
  const int wlo = max(0, jA - 1 - E), whi = min(ny + 1, jB + 1 + E);
  const int width = whi - wlo + 1, pitch = 2 * TJ + 2 * (L.rb_chunk + L.rb_W) + 4;
  double* rowA = sm + 1; double* rowB = sm + pitch + 1;      // index [j - wlo], one pad cell on each side
  double* sblk = sm + 2 * pitch;                             // [row - c0][j - jA]
  // operands of all trapezoid rows, [i - i0][j - wlo]: fetched in one L2 round trip instead of one per (serial) row
  double* sq = sblk + L.rb_chunk * 2 * TJ; double* s5 = sq + (E + 1) * pitch; double* s8 = s5 + (E + 1) * pitch; double* sg = s8 + (E + 1) * pitch;
  for (int t = threadIdx.x; t < (E + 1) * width; t += blockDim.x) {
    const int r = t / width, jj = t - r * width;
    const size_t c = (size_t)(wlo + jj) + (size_t)ny2 * (i0 + r);
    sq[r * pitch + jj] = ld_shared_data<CG>(L.q1 + c); s5[r * pitch + jj] = L.c5b[c]; s8[r * pitch + jj] = L.c8b[c]; sg[r * pitch + jj] = L.g1[c];
  }
  double* sh0 = sg + (E + 1) * pitch; double* shN = sh0 + (E + 1);       // x1old(0, i) and x1old(ny+1, i) of the trapezoid rows
  for (int t = threadIdx.x; t < 2 * (E + 1); t += blockDim.x) {
    const int r = t >> 1;
    if (t & 1) { if (whi == ny + 1) shN[r] = ld_shared_data<CG>(x1old + (size_t)(ny + 1) + (size_t)ny2 * (i0 + r)); }
    else if (wlo == 0) sh0[r] = ld_shared_data<CG>(x1old + (size_t)ny2 * (i0 + r));
  }
  for (int t = threadIdx.x; t < width; t += blockDim.x) rowA[t] = ld_shared_data<CG>(x1old + (size_t)(wlo + t) + (size_t)ny2 * (i0 - 1));
  if (threadIdx.x == 0) { rowA[-1] = 0.0; rowA[width] = 0.0; rowB[-1] = 0.0; rowB[width] = 0.0; }
  double* prev = rowA; double* cur = rowB;
  // this thread's entry of a step: the t-th column of the step's parity inside [max(1,wlo+1), min(ny,whi-1)]
  const int lo = max(1, wlo + 1), hi = min(ny, whi - 1);
  auto my_j = [&](int i) { const int js = 1 + ((i + rb) & 1); int j = lo + (((lo ^ js) & 1) ? 1 : 0) + 2 * (int)threadIdx.x; return j; };
  __syncthreads();
  for (int i = i0; i <= c1; ++i) {
    const int j = my_j(i);
    if (j <= hi) {
      const int o = (i - i0) * pitch + (j - wlo);
      const double sv = -(s5[o] * prev[j + 1 - wlo] + s8[o] * prev[j - 1 - wlo]);
      cur[j - wlo] = sq[o] + sv * sg[o];
      if (i >= c0 && j >= jA && j <= jB) sblk[(i - c0) * 2 * TJ + (j - jA)] = sv;
    }
    if (threadIdx.x == 0) {
      if (wlo == 0) cur[0] = sh0[i - i0];
      if (whi == ny + 1) cur[ny + 1 - wlo] = shN[i - i0];
    }
    __syncthreads();
    double* tmp = prev; prev = cur; cur = tmp;
  }
  // ---- C: stream the block's rows
  const int nrows = c1 - c0 + 1, ncols = (jB - jA + 2) / 2;                 // colour columns per row in the block
  const size_t sj = nz, si = (size_t)nz * ny2;
  const int nelem = nrows * ncols * nz;
#pragma unroll 4
  for (int e = threadIdx.x; e < nelem; e += blockDim.x) {
    const int k = e % nz; int t = e / nz; const int cj = t % ncols, r = t / ncols;
    const int i = c0 + r;
    const int js = 1 + ((i + rb) & 1);
    const int j = jA + (js - 1) + 2 * cj;
    if (j > jB) continue;
    const double sv = sblk[r * 2 * TJ + (j - jA)];
    const size_t idx = IDX3(k, j, i, nz, ny2);
    const double v = ld_shared_data<CG>(L.p + idx) + sv * L.g[idx];
    store_mirrored(L.p, idx, v, j, i, ny, nx, sj, si, true);
    if (k == 0) {
      const size_t c2d = (size_t)j + (size_t)ny2 * i;
      store_mirrored(x1new, c2d, v, j, i, ny, nx, 1, (size_t)ny2, true);
      const int jp = (js == 1) ? j + 1 : j - 1;                               // the other-colour cell of the pair keeps its value
      if (jp >= 1 && jp <= ny) store_mirrored(x1new, (size_t)jp + (size_t)ny2 * i, ld_shared_data<CG>(x1old + (size_t)jp + (size_t)ny2 * i), jp, i, ny, nx, 1, (size_t)ny2, true);
    }
  }
}
__global__ void __launch_bounds__(256) k_relax_BC(LevelView L, ColourSet cs, const double* __restrict__ x1old, double* __restrict__ x1new, int TJ) {
  extern __shared__ __align__(16) double sm[];
  bc_block<false>(L, cs.rb, x1old, x1new, TJ, blockIdx.x, blockIdx.y, sm);
}
void launch_relax_BC(cudaStream_t st, const LevelView& L, const ColourSet& cs, const double* x1old, double* x1new) {
  const int nyc = (L.ny + 1) / 2;
  const int TJ = nyc < 8 ? nyc : 8;          // narrow blocks: many CTAs, the trapezoid overhead is 2-D only
  const int ntj = (nyc + TJ - 1) / TJ, nchunks = (L.nx + L.rb_chunk - 1) / L.rb_chunk;
  const int pitch = 2 * TJ + 2 * (L.rb_chunk + L.rb_W) + 4;
  const size_t smem = ((size_t)2 * pitch + (size_t)L.rb_chunk * 2 * TJ + (size_t)(L.rb_chunk + L.rb_W) * (4 * pitch + 2)) * sizeof(double);
  dim3 grid(ntj, nchunks);
  launch_k(k_relax_BC, grid, 256, smem, st, L, cs, x1old, x1new, TJ);
}
void launch_relax_C(cudaStream_t st, const LevelView& L, const ColourSet& cs) {
  const int nyc = (L.ny + 1) / 2;
  const int nvi = (L.nx - cs.istart) / cs.istep + 1;
  dim3 grid(blocks_for((size_t)nyc * L.nz, 1024), nvi);
  launch_k(k_relax_C, grid, 256, 0, st, L, cs, nyc);
}

// Fused relaxation of a SMALL level: all `nsweeps` red-black sweeps (both colours, bottom-coupling recurrence,
// correction and the Neumann halo of p) in ONE launch of ONE CTA.  Coarse levels are launch/latency bound: this
// replaces 10*nsweeps launches by one.  Only valid when the level has no lateral neighbour rank (single GPU, or
// a level gathered onto one rank).  Same arithmetic, operation by operation, as kernels A/B/C + halo.
//   shared memory: six 2-D planes (x1, q1, c5b, c8b, g1, s) of (ny+2)*(nx+2) doubles + per-warp x1 rows
// ASM (compile time, so that the staged operands are addressed with LDS/STS instead of generic loads):
//   0: 3-D operands in global memory (L2) ; 1: p, cA2..8, bet staged in shared memory (b and g stay in L2) ; 2: b and g as well
// NB (levels WITH neighbour ranks, peer-memory transport): after every colour the CTA runs the halo exchange itself
// (halo1_p2p_body: stores into the neighbours' mailboxes, polls its own); the Neumann mirrors written on sides that have a
// neighbour are overwritten by it.  `nsweeps` sweeps = 2 nsweeps exchanges, numbered from *B.seq + 1.
template <int NZ, int ASM>
__global__ void __launch_bounds__(1024) k_relax_small(LevelView L, int nsweeps, int rank1, int hasnb, Halo1Flags hf, P2PBox B) {
  constexpr int all_smem = ASM;
  extern __shared__ __align__(16) double sm[];
  const int nx = L.nx, ny = L.ny, ny2 = ny + 2, nyc = ny / 2;
  const int n2 = ny2 * (nx + 2);
  double *x1 = sm, *q1 = sm + n2, *c5b = sm + 2 * n2, *c8b = sm + 3 * n2, *g1 = sm + 4 * n2, *ss = sm + 5 * n2, *rows = sm + 6 * n2;   // rows: [8 warps][2][ny2]
  const int sj = NZ; const int si = NZ * ny2;
  // the 3-D operands stay in global memory (L2) or, when the whole level fits (all_smem), are staged in shared memory
  const int n3 = NZ * n2;
  double* const base = rows + 16 * ny2;
  if constexpr (ASM > 0) {
    const double* src[11] = {L.p, L.cA[1], L.cA[2], L.cA[3], L.cA[4], L.cA[5], L.cA[6], L.cA[7], L.bet, L.b, L.g};
    constexpr int na = ASM == 2 ? 11 : 9;
    for (int a = 0; a < na; ++a) for (int e = threadIdx.x; e < n3; e += blockDim.x) base[a * n3 + e] = src[a][e];
  }
  double* const p = ASM > 0 ? base : L.p;
  const double* const c2 = ASM > 0 ? base + n3 : L.cA[1];     const double* const c3 = ASM > 0 ? base + 2 * n3 : L.cA[2];
  const double* const c4 = ASM > 0 ? base + 3 * n3 : L.cA[3]; const double* const c5 = ASM > 0 ? base + 4 * n3 : L.cA[4];
  const double* const c6 = ASM > 0 ? base + 5 * n3 : L.cA[5]; const double* const c7 = ASM > 0 ? base + 6 * n3 : L.cA[6];
  const double* const c8 = ASM > 0 ? base + 7 * n3 : L.cA[7]; const double* const bet = ASM > 0 ? base + 8 * n3 : L.bet;
  const double* const b = ASM == 2 ? base + 9 * n3 : L.b;     const double* const g = ASM == 2 ? base + 10 * n3 : L.g;
  if (rank1) for (int e = threadIdx.x; e < n2; e += blockDim.x) { c5b[e] = L.c5b[e]; c8b[e] = L.c8b[e]; g1[e] = L.g1[e]; }
  __syncthreads();
  // x1 = p(1,:,:) once; afterwards the final stores of each colour keep it (and the Neumann halo of p) current
  if (rank1) { for (int e = threadIdx.x; e < n2; e += blockDim.x) x1[e] = p[(size_t)e * NZ]; __syncthreads(); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const unsigned long long seq0 = hasnb ? *B.seq : 0ull; int nex = 0;
#ifdef MGB_SMALL_TIMING
  long long tA = 0, tB = 0, tC = 0, t0 = clock64();
#define SM_T(acc) { const long long t1 = clock64(); acc += t1 - t0; t0 = t1; }
#else
#define SM_T(acc)
#endif
  for (int it = 0; it < nsweeps; ++it) for (int rb = 1; rb <= 2; ++rb) {
    // ---- A: thread per column
    for (int c = threadIdx.x; c < nx * nyc; c += blockDim.x) {
      const int i = 1 + c / nyc, jc = c % nyc;
      const int j = 1 + ((i + rb) & 1) + 2 * jc;
      const int idx0 = NZ * (j + ny2 * i);
      double x[NZ];
#pragma unroll
      for (int k = 0; k < NZ; ++k) {
        const int idx = idx0 + k;
        double acc = b[idx];
        if (k < NZ - 1) acc = acc - c3[idx] * p[idx - sj + 1];
        if (k > 0)      acc = acc - c3[idx + sj - 1] * p[idx + sj - 1];
        acc = acc - c4[idx] * p[idx - sj];
        acc = acc - c4[idx + sj] * p[idx + sj];
        if (k > 0)      acc = acc - c5[idx] * p[idx - sj - 1];
        if (k < NZ - 1) acc = acc - c5[idx + sj + 1] * p[idx + sj + 1];
        if (k < NZ - 1) acc = acc - c6[idx] * p[idx - si + 1];
        if (k > 0)      acc = acc - c6[idx + si - 1] * p[idx + si - 1];
        acc = acc - c7[idx] * p[idx - si];
        acc = acc - c7[idx + si] * p[idx + si];
        if (k > 0)      acc = acc - c8[idx] * p[idx - si - 1];
        if (k < NZ - 1) acc = acc - c8[idx + si + 1] * p[idx + si + 1];
        if (k == 0 && rank1) {
          const int c2d = j + ny2 * i;
          acc = acc - c5b[c2d - 1 + ny2] * x1[c2d - 1 + ny2];
          acc = acc - c8b[c2d + 1 + ny2] * x1[c2d + 1 + ny2];
        }
        x[k] = acc;
      }
      double xm = x[0] * bet[idx0];
      x[0] = xm;
#pragma unroll
      for (int k = 1; k < NZ; ++k) { xm = (x[k] - c2[idx0 + k] * xm) * bet[idx0 + k]; x[k] = xm; }
#pragma unroll
      for (int k = NZ - 2; k >= 0; --k) { xm = x[k] - (c2[idx0 + k + 1] * bet[idx0 + k]) * xm; x[k] = xm; }
      if (rank1) {            // not final yet: q only
#pragma unroll
        for (int k = 0; k < NZ; ++k) p[idx0 + k] = x[k];
        q1[j + ny2 * i] = xm;
      } else {                // final value of this colour: also its Neumann mirrors
#pragma unroll
        for (int k = 0; k < NZ; ++k) store_mirrored(p, (size_t)(idx0 + k), x[k], j, i, ny, nx, (size_t)sj, (size_t)si, true);
      }
    }
    __syncthreads();
    SM_T(tA)
    if (rank1) {
      // ---- B: chunked bottom recurrence, one warp per chunk (same chunks and warm-up as kernel B).  When the colour's
      //      columns of a row fit one warp (ny <= 64) the previous row lives in registers and travels by shuffles.
      const int wpar = nwarps < 8 ? nwarps : 8;
      const int nchunks = (nx + L.rb_chunk - 1) / L.rb_chunk;
      if (warp < wpar) {
        for (int c = warp; c < nchunks; c += wpar) {
          const int c0 = 1 + c * L.rb_chunk, c1 = min(nx, c0 + L.rb_chunk - 1), i0 = max(1, c0 - L.rb_W);
          if (nyc <= 32) {
            const bool act = lane < nyc;
            const int l = act ? lane : 0;
            int js = 1 + ((i0 + rb) & 1);
            double xv = x1[(3 - js) + 2 * l + ny2 * (i0 - 1)];               // x1 at the colour's positions of row i0-1
            for (int i = i0; i <= c1; ++i, js = 3 - js) {
              const double up = __shfl_up_sync(0xffffffffu, xv, 1), dn = __shfl_down_sync(0xffffffffu, xv, 1);
              const double h0 = x1[ny2 * (i - 1)], hN = x1[ny + 1 + ny2 * (i - 1)];
              const int j = js + 2 * l, c2d = j + ny2 * i;
              double xp, xmn;                                                // x1(j+1,i-1), x1(j-1,i-1)
              if (js == 1) { xp = xv; xmn = (l == 0) ? h0 : up; } else { xmn = xv; xp = (l == nyc - 1) ? hN : dn; }
              const double sv = -(c5b[c2d] * xp + c8b[c2d] * xmn);
              xv = q1[c2d] + sv * g1[c2d];
              if (act && i >= c0) ss[c2d] = sv;
            }
          } else {
            double* prev = rows + warp * 2 * ny2; double* cur = prev + ny2;
            for (int j = lane; j < ny2; j += 32) prev[j] = x1[j + ny2 * (i0 - 1)];
            __syncwarp();
            for (int i = i0; i <= c1; ++i) {
              const int js = 1 + ((i + rb) & 1);
              for (int jc = lane; jc < nyc; jc += 32) {
                const int j = js + 2 * jc, c2d = j + ny2 * i;
                const double sv = -(c5b[c2d] * prev[j + 1] + c8b[c2d] * prev[j - 1]);
                cur[j] = q1[c2d] + sv * g1[c2d];
                if (i >= c0) ss[c2d] = sv;
              }
              if (lane == 0) { cur[0] = x1[ny2 * i]; cur[ny + 1] = x1[ny + 1 + ny2 * i]; }
              __syncwarp();
              double* t = prev; prev = cur; cur = t;
            }
          }
        }
      }
      __syncthreads();
      SM_T(tB)
      // ---- C: p += s g on the colour's columns; final values -> Neumann mirrors and the x1 plane
      for (int c = threadIdx.x; c < nx * nyc; c += blockDim.x) {
        const int i = 1 + c / nyc, jc = c % nyc;
        const int j = 1 + ((i + rb) & 1) + 2 * jc;
        const int idx0 = NZ * (j + ny2 * i);
        const double sv = ss[j + ny2 * i];
#pragma unroll
        for (int k = 0; k < NZ; ++k) {
          const double v = p[idx0 + k] + sv * g[idx0 + k];
          store_mirrored(p, (size_t)(idx0 + k), v, j, i, ny, nx, (size_t)sj, (size_t)si, true);
          if (k == 0) store_mirrored(x1, (size_t)(j + ny2 * i), v, j, i, ny, nx, 1, (size_t)ny2, true);
        }
      }
      __syncthreads();
      SM_T(tC)
    }
    if (hasnb) {
      halo1_p2p_body<false>(p, NZ, ny, nx, hf, B, rank1 ? x1 : nullptr, seq0 + (unsigned long long)(++nex), threadIdx.x, blockDim.x);
      __syncthreads();
    }
  }
  if (hasnb && threadIdx.x == 0) *B.seq = seq0 + (unsigned long long)nex;
#ifdef MGB_SMALL_TIMING
  if (threadIdx.x == 0 && nsweeps >= 20) printf("small nz=%d nx=%d ny=%d threads=%d W=%d all_smem=%d: per colour cycles A=%lld B=%lld C=%lld\n", NZ, nx, ny, (int)blockDim.x, L.rb_W, all_smem,
         tA / (2 * nsweeps), tB / (2 * nsweeps), tC / (2 * nsweeps));
#endif
  if (all_smem) for (int e = threadIdx.x; e < n3; e += blockDim.x) L.p[e] = p[e];
}
// Fused relaxation of a MID-SIZE level without lateral neighbour ranks: all `nsweeps` red-black sweeps in ONE cooperative
// launch of G co-resident CTAs that meet at software grid barriers (an atomic counter in L2) between the phases
// A (column solves) -> B (chunked bottom recurrence) -> C (rank-1 update + Neumann mirrors) of every colour.
// These levels (10^4 .. 10^6 points) are launch-latency bound: a colour costs three ~1 us barriers here instead of two or
// three dependent launches of ~6 us.  Same arithmetic, operation by operation, as kernels A4 / B / C.
// Values written by other CTAs during the launch (p, q1, s, x1) are always read with ld.global.cg (L2): the L1s are not
// coherent inside a launch.  Coefficients are read-only and stay L1-resident across sweeps (a CTA keeps its columns).
//   bar[0]: barrier counter, bar[1]: exit counter (the last CTA to leave zeroes both for the next launch)
// NB (levels WITH neighbour ranks, peer-memory transport): every colour ends with the halo exchange run by all CTAs of the
// launch (halo1_p2p_body) and one more grid barrier; see k_relax_small.
template <int NZ>
__global__ void __launch_bounds__(512, 1) k_relax_mid(LevelView L, int nsweeps, int rank1, int TC, unsigned* __restrict__ bar, int hasnb, Halo1Flags hf, P2PBox B) {
  extern __shared__ __align__(16) double sm[];
  const int nx = L.nx, ny = L.ny, ny2 = ny + 2, nyc = ny / 2, pitch = NZ + 1;
  constexpr int sj = NZ; const int si = NZ * ny2;
  const int G = gridDim.x;
  const int ncol = nx * nyc;                                  // columns of one colour; column c -> row i = 1 + c / nyc
  const int ntiles = (ncol + TC - 1) / TC;
  double* s_rhs = sm; double2* s_bu = reinterpret_cast<double2*>(sm + TC * pitch);     // {pivot, off-diagonal} pairs
  unsigned target = 0;
  auto gbar = [&]() {
    target += G;
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");   // release: the CTA's stores (ordered by bar.sync) are out
      while (ld_acquire_gpu_u32(bar) < target) { }
    }
    __syncthreads();
  };
  const double* __restrict__ b = L.b; const double* __restrict__ bet = L.bet;
  const double* __restrict__ c2 = L.cA[1]; const double* __restrict__ c3 = L.cA[2]; const double* __restrict__ c4 = L.cA[3];
  const double* __restrict__ c5 = L.cA[4]; const double* __restrict__ c6 = L.cA[5]; const double* __restrict__ c7 = L.cA[6];
  const double* __restrict__ c8 = L.cA[7];
  double* p = L.p;
  double* x1cur = L.x1; double* x1alt = L.x1b;
  const unsigned long long seq0 = hasnb ? *B.seq : 0ull; int nex = 0;
  if (rank1) {   // x1 = p(1,:,:) on the whole plane incl. the halo; afterwards every colour writes the next k=1 plane completely
    const int n2 = ny2 * (nx + 2);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += G * blockDim.x) x1cur[e] = __ldcg(p + (size_t)e * NZ);
    gbar();
  }
#ifdef MGB_MID_TIMING
  long long tA = 0, tB1 = 0, tBC = 0, tB2 = 0, t0;
#define MID_T(acc) { const long long t1 = clock64(); acc += t1 - t0; t0 = t1; }
  t0 = clock64();
#else
#define MID_T(acc)
#endif
  for (int it = 0; it < nsweeps; ++it) for (int rb = 1; rb <= 2; ++rb) {
    // ---- A: tiles of TC columns; rhs by (k, column) threads, Thomas by one thread per column, coalesced write-back
    for (int tile = blockIdx.x; tile < ntiles; tile += G) {
      const int cbase = tile * TC, tcn = min(TC, ncol - cbase);
      for (int e = threadIdx.x; e < tcn * NZ; e += blockDim.x) {
        const int k = e & (NZ - 1), cl = e / NZ, c = cbase + cl;
        const int i = 1 + c / nyc, jc = c - (i - 1) * nyc;
        const int j = 1 + ((i + rb) & 1) + 2 * jc;
        const int idx = k + NZ * (j + ny2 * i);
        const int im = idx - si, ip = idx + si;
        const double bb = b[idx], bt = bet[idx], udv = c2[idx + 1];
        const double a3 = c3[idx],           q3 = __ldcg(p + idx - sj + 1);
        const double a3n = c3[idx + sj - 1], q3n = __ldcg(p + idx + sj - 1);
        const double a4 = c4[idx],           q4 = __ldcg(p + idx - sj);
        const double a4n = c4[idx + sj],     q4n = __ldcg(p + idx + sj);
        const double a5 = c5[idx],           q5 = __ldcg(p + idx - sj - 1);
        const double a5n = c5[idx + sj + 1], q5n = __ldcg(p + idx + sj + 1);
        const double a6 = c6[idx],           q6 = __ldcg(p + im + 1);
        const double a6n = c6[ip - 1],       q6n = __ldcg(p + ip - 1);
        const double a7 = c7[idx],           q7 = __ldcg(p + im);
        const double a7n = c7[ip],           q7n = __ldcg(p + ip);
        const double a8 = c8[idx],           q8 = __ldcg(p + im - 1);
        const double a8n = c8[ip + 1],       q8n = __ldcg(p + ip + 1);
        const bool lo = (k > 0), hi = (k < NZ - 1);
        double acc = bb;
        if (hi) acc = acc - a3 * q3;
        if (lo) acc = acc - a3n * q3n;
        acc = acc - a4 * q4;
        acc = acc - a4n * q4n;
        if (lo) acc = acc - a5 * q5;
        if (hi) acc = acc - a5n * q5n;
        if (hi) acc = acc - a6 * q6;
        if (lo) acc = acc - a6n * q6n;
        acc = acc - a7 * q7;
        acc = acc - a7n * q7n;
        if (lo) acc = acc - a8 * q8;
        if (hi) acc = acc - a8n * q8n;
        if (k == 0 && rank1) {
          const int c2d = j + ny2 * i;
          acc = acc - L.c5b[c2d - 1 + ny2] * __ldcg(x1cur + c2d - 1 + ny2);
          acc = acc - L.c8b[c2d + 1 + ny2] * __ldcg(x1cur + c2d + 1 + ny2);
        }
        const int so = cl * pitch + k;
        s_rhs[so] = acc; s_bu[so] = make_double2(bt, hi ? udv : 0.0);
      }
      __syncthreads();
      for (int cl = threadIdx.x; cl < tcn; cl += blockDim.x) {
        const double xm = thomas_smem_plain(s_rhs + cl * pitch, s_bu + cl * pitch, NZ);
        if (rank1) { const int c = cbase + cl; const int i = 1 + c / nyc, jc = c - (i - 1) * nyc; L.q1[1 + ((i + rb) & 1) + 2 * jc + ny2 * i] = xm; }
      }
      __syncthreads();
      for (int e = threadIdx.x; e < tcn * NZ; e += blockDim.x) {
        const int k = e & (NZ - 1), cl = e / NZ, c = cbase + cl;
        const int i = 1 + c / nyc, jc = c - (i - 1) * nyc;
        const int j = 1 + ((i + rb) & 1) + 2 * jc;
        const int idx = k + NZ * (j + ny2 * i);
        const double v = s_rhs[cl * pitch + k];
        if (rank1) p[idx] = v;                               // q only; final after C
        else store_mirrored(p, (size_t)idx, v, j, i, ny, nx, (size_t)sj, (size_t)si, true);
      }
      __syncthreads();
    }
    MID_T(tA)
    gbar();
    MID_T(tB1)
    if (rank1) {
      // ---- B + C fused per (chunk, block of 2*TJ columns) task, exactly as k_relax_BC: trapezoid recurrence in shared memory,
      //      then p += s g with the Neumann mirrors; the k=1 plane is double buffered (other tasks still read the old one)
      const int TJ = nyc < 8 ? nyc : 8, ntj = (nyc + TJ - 1) / TJ, nchunks = (nx + L.rb_chunk - 1) / L.rb_chunk;
      for (int task = blockIdx.x; task < ntj * nchunks; task += G) {
        bc_block<true>(L, rb, x1cur, x1alt, TJ, task % ntj, task / ntj, sm);
        __syncthreads();
      }
      { double* t = x1cur; x1cur = x1alt; x1alt = t; }
      MID_T(tBC)
      gbar();
      MID_T(tB2)
    }
    if (hasnb) {
      halo1_p2p_body<true>(p, NZ, ny, nx, hf, B, rank1 ? x1cur : nullptr, seq0 + (unsigned long long)(++nex),
                           (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)G * blockDim.x);
      gbar();
    }
  }
#ifdef MGB_MID_TIMING
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == G - 1))
    printf("mid nz=%d nx=%d G=%d cta=%d W=%d: per colour cycles A=%lld bar1=%lld BC=%lld bar2=%lld\n", NZ, nx, G, (int)blockIdx.x, L.rb_W,
           tA / (2 * nsweeps), tB1 / (2 * nsweeps), tBC / (2 * nsweeps), tB2 / (2 * nsweeps));
#endif
  // leave the counters zeroed for the next launch: the last CTA out does it
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); if (atomicAdd(bar + 1, 1u) == (unsigned)G - 1) { bar[0] = 0; bar[1] = 0; if (hasnb) *B.seq = seq0 + (unsigned long long)nex; } }
}
static int mid_tc(const LevelView& L) { (void)L; return 64; }
static size_t mid_limit() { static const size_t v = []() { const char* e = std::getenv("MGB_MID_LIMIT"); return e ? (size_t)std::atoll(e) : (size_t)1500000; }(); return v; }
bool relax_mid_supported(const LevelView& L) {
  const bool nz_ok = (L.nz == 2 || L.nz == 4 || L.nz == 8 || L.nz == 16 || L.nz == 32);
  const size_t n3 = (size_t)L.nz * (L.ny + 2) * (L.nx + 2);
  return nz_ok && (L.ny % 2 == 0) && n3 <= mid_limit() && L.rb_W <= RB_WMAX;   // chunked recurrence only
}
void launch_relax_mid(cudaStream_t st, const LevelView& L, int nsweeps, int rank1, unsigned* bar, int hasnb, const Halo1Flags& hf, const P2PBox& B) {
  const int TC = mid_tc(L);
  const int ncol = L.nx * (L.ny / 2), ntiles = (ncol + TC - 1) / TC;
  const int sms = device_sm_count();                                 // one CTA per SM (512 threads, launch bound 1): co-resident by construction,
  const int G = ntiles < sms ? ntiles : sms;                         // and the cooperative launch itself refuses a grid that is not
  const int nyc = L.ny / 2, TJ = nyc < 8 ? nyc : 8;
  const size_t bc_pitch = (size_t)2 * TJ + 2 * (L.rb_chunk + L.rb_W) + 4;
  const size_t bc_doubles = 2 * bc_pitch + (size_t)L.rb_chunk * 2 * TJ + (size_t)(L.rb_chunk + L.rb_W) * (4 * bc_pitch + 2);
  const size_t a_doubles = (size_t)3 * TC * (L.nz + 1);
  const size_t smem = (a_doubles > bc_doubles ? a_doubles : bc_doubles) * sizeof(double);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(G); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;   // co-residency is required (grid barriers)
  cfg.attrs = at; cfg.numAttrs = 1;
  switch (L.nz) {
    case 2: note_launch_error(cudaLaunchKernelEx(&cfg, k_relax_mid<2>, L, nsweeps, rank1, TC, bar, hasnb, hf, B)); break;
    case 4: note_launch_error(cudaLaunchKernelEx(&cfg, k_relax_mid<4>, L, nsweeps, rank1, TC, bar, hasnb, hf, B)); break;
    case 8: note_launch_error(cudaLaunchKernelEx(&cfg, k_relax_mid<8>, L, nsweeps, rank1, TC, bar, hasnb, hf, B)); break;
    case 16: note_launch_error(cudaLaunchKernelEx(&cfg, k_relax_mid<16>, L, nsweeps, rank1, TC, bar, hasnb, hf, B)); break;
    default: note_launch_error(cudaLaunchKernelEx(&cfg, k_relax_mid<32>, L, nsweeps, rank1, TC, bar, hasnb, hf, B)); break;
  }
  ++g_launches;
}
int device_sm_count() {
  static const int n = []() { int dev = 0, v = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); return v > 0 ? v : 148; }();
  return n;
}
void kernels_init() {
  relax_fused_init();
  prolong_bulk_init();
  a4_attrs<2>(); a4_attrs<3>();
  cudaFuncSetAttribute(k_relax_BC, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_relax_mid<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_relax_mid<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_relax_mid<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_relax_mid<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_relax_mid<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_relax_B<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_relax_B<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_relax_B<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_relax_B<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_relax_small<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_relax_small<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
}
bool relax_small_supported(const LevelView& L) {
  const size_t n2 = (size_t)(L.ny + 2) * (L.nx + 2);
  const bool nz_ok = (L.nz == 2 || L.nz == 4 || L.nz == 8 || L.nz == 16);
  // single-CTA relaxation pays off on tiny levels only; a 2-level (nz = 2) grid may be a little wider (coarsest level of a multi-GPU job)
  static const bool no_mid = []() { const char* e = std::getenv("MGB_NO_MID_KERNEL"); return e && e[0] == '1'; }();
  if (no_mid) return nz_ok && (L.ny % 2 == 0) && (n2 <= (size_t)34 * 34 || (L.nz == 2 && n2 <= (size_t)66 * 34));
  // larger levels go to the cooperative multi-CTA kernel (~7 us per colour floor); a single CTA beats that on 16x16 and on 2-level grids up to 32x32
  return nz_ok && (L.ny % 2 == 0) && (n2 <= (size_t)18 * 18 || (L.nz == 2 && n2 <= (size_t)34 * 34));
}
void launch_relax_small(cudaStream_t st, const LevelView& L, int nsweeps, int rank1, int hasnb, const Halo1Flags& hf, const P2PBox& B) {
  const size_t n2 = (size_t)(L.ny + 2) * (L.nx + 2);
  size_t smem = (n2 * 6 + (size_t)16 * (L.ny + 2)) * sizeof(double);
  const size_t smem11 = smem + (size_t)11 * L.nz * n2 * sizeof(double), smem9 = smem + (size_t)9 * L.nz * n2 * sizeof(double);
  const int all_smem = smem11 <= 226 * 1024 ? 2 : (smem9 <= 226 * 1024 ? 1 : 0);
  if (all_smem == 2) smem = smem11; else if (all_smem == 1) smem = smem9;
  int threads = 64; const int ncol = L.nx * (L.ny / 2);
  while (threads < ncol && threads < 1024) threads *= 2;
#define SMALL_CASE(NZv) \
  if (all_smem == 2) launch_k(k_relax_small<NZv, 2>, 1, threads, smem, st, L, nsweeps, rank1, hasnb, hf, B); \
  else if (all_smem == 1) launch_k(k_relax_small<NZv, 1>, 1, threads, smem, st, L, nsweeps, rank1, hasnb, hf, B); \
  else launch_k(k_relax_small<NZv, 0>, 1, threads, smem, st, L, nsweeps, rank1, hasnb, hf, B);
  switch (L.nz) {
    case 2: SMALL_CASE(2) break;
    case 4: SMALL_CASE(4) break;
    case 8: SMALL_CASE(8) break;
    default: SMALL_CASE(16) break;
  }
#undef SMALL_CASE
}

// =====================================================================================================
// residual + norm   (src/mg_relax.f90:421-515)
// =====================================================================================================
__device__ __forceinline__ double block_sum_256(double v, double* s_red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  return t;   // valid in thread 0
}

// one block handles a chunk of one i-slab: elements (k, j) flattened, j = 1..ny
template <int NZC>   // NZC > 0: nz at compile time (power of two): immediates for the k/j offsets, shifts for e -> (k, j)
__global__ void __launch_bounds__(256, 3) k_residual(LevelView L, int real_matrix, int chunks_per_i, double* __restrict__ partials) {
  __shared__ double s_red[8];
  const int nz = NZC ? NZC : L.nz, ny = L.ny, ny2 = ny + 2;
  const int i = 1 + blockIdx.x / chunks_per_i, chunk = blockIdx.x % chunks_per_i;
  const int sj = nz; const int si = nz * ny2;
  const int base_i = nz * (1 + ny2 * i);                 // element index of (k=0, j=1, i): the slab is contiguous in (k, j)
  const double* __restrict__ p = L.p; const double* __restrict__ b = L.b;
  const double* __restrict__ c1 = L.cA[0]; const double* __restrict__ c2 = L.cA[1]; const double* __restrict__ c3 = L.cA[2];
  const double* __restrict__ c4 = L.cA[3]; const double* __restrict__ c5 = L.cA[4]; const double* __restrict__ c6 = L.cA[5];
  const double* __restrict__ c7 = L.cA[6]; const double* __restrict__ c8 = L.cA[7];
  double* __restrict__ r = L.r;
  const int nelem = ny * nz;
  const int per_chunk = (nelem + chunks_per_i - 1) / chunks_per_i;
  const int e0 = chunk * per_chunk, e1 = min(nelem, e0 + per_chunk);
  double sum = 0.0;
  for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const int k = NZC ? (e & (NZC - 1)) : e % nz;
    const int idx = base_i + e;
    const double bb = b[idx];
    const double a1 = c1[idx],           q1v = p[idx];
    const double a2 = c2[idx],           q2 = p[idx - 1];
    const double a2n = c2[idx + 1],      q2n = p[idx + 1];
    const double a3 = c3[idx],           q3 = p[idx - sj + 1];
    const double a3n = c3[idx + sj - 1], q3n = p[idx + sj - 1];
    const double a4 = c4[idx],           q4 = p[idx - sj];
    const double a4n = c4[idx + sj],     q4n = p[idx + sj];
    const double a5 = c5[idx],           q5 = p[idx - sj - 1];
    const double a5n = c5[idx + sj + 1], q5n = p[idx + sj + 1];
    const double a6 = c6[idx],           q6 = p[idx - si + 1];
    const double a6n = c6[idx + si - 1], q6n = p[idx + si - 1];
    const double a7 = c7[idx],           q7 = p[idx - si];
    const double a7n = c7[idx + si],     q7n = p[idx + si];
    const double a8 = c8[idx],           q8 = p[idx - si - 1];
    const double a8n = c8[idx + si + 1], q8n = p[idx + si + 1];
    const bool lo = (k > 0), hi = (k < nz - 1);
    double acc = bb;
    acc = acc - a1 * q1v;
    if (lo) acc = acc - a2 * q2;
    if (hi) acc = acc - a2n * q2n;
    if (hi) acc = acc - a3 * q3;
    if (lo) acc = acc - a3n * q3n;
    acc = acc - a4 * q4;
    acc = acc - a4n * q4n;
    if (lo) acc = acc - a5 * q5;
    if (hi) acc = acc - a5n * q5n;
    if (hi) acc = acc - a6 * q6;
    if (lo) acc = acc - a6n * q6n;
    acc = acc - a7 * q7;
    acc = acc - a7n * q7n;
    if (lo) acc = acc - a8 * q8;
    if (hi) acc = acc - a8n * q8n;
    if (k == 0 && real_matrix) {
      acc = acc - a5 * p[idx + sj - si];
      acc = acc - c5[idx - sj + si] * p[idx - sj + si];
      acc = acc - a8 * p[idx - sj - si];
      acc = acc - c8[idx + sj + si] * p[idx + sj + si];
    }
    r[idx] = acc;
    sum += acc * acc;
  }
  const double t = block_sum_256(sum, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}
// fixed-order final reduction (deterministic): one block
__global__ void __launch_bounds__(256) k_reduce_final(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double s_red[8];
  double v = 0.0;
  for (int e = threadIdx.x; e < n; e += blockDim.x) v += partials[e];
  const double t = block_sum_256(v, s_red);
  if (threadIdx.x == 0) *out = t;
}
static int residual_chunks(const LevelView& L) {
  const int nelem = L.ny * L.nz;
  int chunks = (nelem + 4095) / 4096;        // <= 16 elements per thread
  return chunks < 1 ? 1 : chunks;
}
int residual_num_partials(const LevelView& L) { return L.nx * residual_chunks(L); }
void launch_residual(cudaStream_t st, const LevelView& L, int real_matrix, double* partials, double* sum_out) {
  const int chunks = residual_chunks(L);
  switch (L.nz) {
    case 128: launch_k(k_residual<128>, L.nx * chunks, 256, 0, st, L, real_matrix, chunks, partials); break;
    case 64: launch_k(k_residual<64>, L.nx * chunks, 256, 0, st, L, real_matrix, chunks, partials); break;
    case 32: launch_k(k_residual<32>, L.nx * chunks, 256, 0, st, L, real_matrix, chunks, partials); break;
    case 16: launch_k(k_residual<16>, L.nx * chunks, 256, 0, st, L, real_matrix, chunks, partials); break;
    default: launch_k(k_residual<0>, L.nx * chunks, 256, 0, st, L, real_matrix, chunks, partials); break;
  }
  launch_k(k_reduce_final, 1, 256, 0, st, partials, L.nx * chunks, sum_out);
}
__global__ void __launch_bounds__(256) k_sumsq(const double* __restrict__ a, int nz, int ny, int chunks_per_i, double* __restrict__ partials) {
  __shared__ double s_red[8];
  const int ny2 = ny + 2;
  const int i = 1 + blockIdx.x / chunks_per_i, chunk = blockIdx.x % chunks_per_i;
  const int nelem = ny * nz;
  const int per_chunk = (nelem + chunks_per_i - 1) / chunks_per_i;
  const int e0 = chunk * per_chunk, e1 = min(nelem, e0 + per_chunk);
  double sum = 0.0;
  for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) { const double v = a[IDX3(0, 1, i, nz, ny2) + e]; sum += v * v; }
  const double t = block_sum_256(sum, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}
void launch_sumsq(cudaStream_t st, const LevelView& L, const double* a, double* partials, double* sum_out) {
  const int chunks = residual_chunks(L);
  launch_k(k_sumsq, L.nx * chunks, 256, 0, st, a, L.nz, L.ny, chunks, partials);
  launch_k(k_reduce_final, 1, 256, 0, st, partials, L.nx * chunks, sum_out);
}

// =====================================================================================================
// intergrid transfers
// =====================================================================================================
// restriction: y(k2,j2,i2) = sum of the 8 fine cells (src/mg_intergrids.f90:149-160), same summation order
__global__ void __launch_bounds__(256) k_restrict8(const double* __restrict__ xf, int nzf, int nyf, double* __restrict__ yc, int nzc, int nyc, int nxc) {
  const int n = nzc * nyc * nxc;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int t = e / nzc, k2 = e - t * nzc;
  const int i2 = 1 + t / nyc, j2 = 1 + (t - (i2 - 1) * nyc);
  const int k = 2 * k2, j = 2 * j2 - 1, i = 2 * i2 - 1;       // k 0-based
  const int nyf2 = nyf + 2;
  const int sj = nzf, si = nzf * nyf2;
  const int f00 = k + nzf * (j + nyf2 * i);
  const double2 a = *reinterpret_cast<const double2*>(xf + f00);             // (k, k+1) at (j, i)
  const double2 b = *reinterpret_cast<const double2*>(xf + f00 + si);        // (j, i+1)
  const double2 c = *reinterpret_cast<const double2*>(xf + f00 + sj);        // (j+1, i)
  const double2 d = *reinterpret_cast<const double2*>(xf + f00 + sj + si);   // (j+1, i+1)
  double z = a.x + b.x + c.x + d.x + a.y + b.y + c.y + d.y;
  yc[k2 + nzc * (j2 + (nyc + 2) * i2)] = z;
}
void launch_restrict8(cudaStream_t st, const double* xf, int nzf, int nyf, int nxf, double* yc, int nzc, int nyc, int nxc) {
  (void)nxf;
  size_t n = (size_t)nzc * nyc * nxc;
  launch_k(k_restrict8, blocks_for(n, 256), 256, 0, st, xf, nzf, nyf, yc, nzc, nyc, nxc);
}

// fine2coarse of a neighbour-free coarse level in ONE launch: b_c = restriction with its Neumann halo, p_c = 0 everywhere and
// (F-cycle descent) r_c = b_c -- what restrict + fill_halo + fill + copy do in four (src/mg_intergrids.f90:16-72, src/mg_solvers.f90:113-118)
__global__ void __launch_bounds__(256) k_restrict8_fused(const double* __restrict__ xf, int nzf, int nyf, double* __restrict__ yc, double* __restrict__ pc,
                                                         double* __restrict__ rc, int nzc, int nyc, int nxc) {
  const int n = nzc * nyc * nxc;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int t = e / nzc, k2 = e - t * nzc;
  const int i2 = 1 + t / nyc, j2 = 1 + (t - (i2 - 1) * nyc);
  const int k = 2 * k2, j = 2 * j2 - 1, i = 2 * i2 - 1;       // k 0-based
  const int nyf2 = nyf + 2;
  const int sj = nzf, si = nzf * nyf2;
  const int f00 = k + nzf * (j + nyf2 * i);
  const double2 a = *reinterpret_cast<const double2*>(xf + f00);             // (k, k+1) at (j, i)
  const double2 b = *reinterpret_cast<const double2*>(xf + f00 + si);        // (j, i+1)
  const double2 c = *reinterpret_cast<const double2*>(xf + f00 + sj);        // (j+1, i)
  const double2 d = *reinterpret_cast<const double2*>(xf + f00 + sj + si);   // (j+1, i+1)
  const double z = a.x + b.x + c.x + d.x + a.y + b.y + c.y + d.y;
  const size_t o = (size_t)(k2 + nzc * (j2 + (nyc + 2) * i2));
  const size_t csj = nzc, csi = (size_t)nzc * (nyc + 2);
  store_mirrored(yc, o, z, j2, i2, nyc, nxc, csj, csi, true);
  store_mirrored(pc, o, 0.0, j2, i2, nyc, nxc, csj, csi, true);
  if (rc) store_mirrored(rc, o, z, j2, i2, nyc, nxc, csj, csi, true);
}
void launch_restrict8_fused(cudaStream_t st, const double* xf, int nzf, int nyf, int nxf, double* yc, double* pc, double* rc, int nzc, int nyc, int nxc) {
  (void)nxf;
  size_t n = (size_t)nzc * nyc * nxc;
  launch_k(k_restrict8_fused, blocks_for(n, 256), 256, 0, st, xf, nzf, nyf, yc, pc, rc, nzc, nyc, nxc);
}

// prolongation (src/mg_intergrids.f90:366-450 linear, :336-363 nearest); thread per fine interior point
__global__ void __launch_bounds__(256) k_prolong(const double* __restrict__ xc, int nzc, int nyc, int nxc,
                                                double* __restrict__ rf, double* __restrict__ pf, int linear, int mode) {
  const int nzf = 2 * nzc, nyf = 2 * nyc, nxf = 2 * nxc;
  const size_t n = (size_t)nzf * nyf * nxf;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int kf = (int)(e % nzf); size_t t = e / nzf; const int jf = 1 + (int)(t % nyf), i_f = 1 + (int)(t / nyf);
  const int K = kf + 1;                                   // Fortran k
  const int j2 = (jf + 1) / 2, i2 = (i_f + 1) / 2;
  const int dj = (jf & 1) ? -1 : +1, di = (i_f & 1) ? -1 : +1;   // odd fine index -> neighbour below, even -> above
  const int nyc2 = nyc + 2;
  double v;
  if (!linear) {
    v = xc[IDX3((K + 1) / 2 - 1, j2, i2, nzc, nyc2)];
  } else {
    const double a = 9.0 / 16.0, b = 3.0 / 16.0, c = 1.0 / 16.0, d = 27.0 / 64.0, ee = 9.0 / 64.0, f = 3.0 / 64.0, g = 1.0 / 64.0;
    if (K == 1 || K == nzf) {
      const int k2 = (K == 1) ? 0 : nzc - 1;
      const size_t o = IDX3(k2, j2, i2, nzc, nyc2);
      const double x00 = xc[o], xdd = xc[IDX3(k2, j2 + dj, i2 + di, nzc, nyc2)], xd0 = xc[IDX3(k2, j2 + dj, i2, nzc, nyc2)], x0d = xc[IDX3(k2, j2, i2 + di, nzc, nyc2)];
      v = a * x00 + c * xdd + b * xd0 + b * x0d;
      if (K == nzf) v = 0.5 * v;
    } else {
      const int k2 = (K + 1) / 2, kp = k2 - ((K % 2) * 2 - 1);      // Fortran indices
      const double x00 = xc[IDX3(k2 - 1, j2, i2, nzc, nyc2)], xdd = xc[IDX3(k2 - 1, j2 + dj, i2 + di, nzc, nyc2)];
      const double xd0 = xc[IDX3(k2 - 1, j2 + dj, i2, nzc, nyc2)], x0d = xc[IDX3(k2 - 1, j2, i2 + di, nzc, nyc2)];
      const double y00 = xc[IDX3(kp - 1, j2, i2, nzc, nyc2)], ydd = xc[IDX3(kp - 1, j2 + dj, i2 + di, nzc, nyc2)];
      const double yd0 = xc[IDX3(kp - 1, j2 + dj, i2, nzc, nyc2)], y0d = xc[IDX3(kp - 1, j2, i2 + di, nzc, nyc2)];
      v = d * x00 + f * xdd + ee * xd0 + ee * x0d + ee * y00 + g * ydd + f * yd0 + f * y0d;
    }
  }
  const size_t of = IDX3(kf, jf, i_f, nzf, nyf + 2);
  if (mode == 0) rf[of] = v; else pf[of] = pf[of] + v;
}
// thread per COARSE interior point: computes its 8 fine children from the 27 surrounding coarse values (each child with
// exactly the expression of the reference, see k_prolong) and writes them as (k,k+1) double2 pairs: 27 cached loads
// per 8 fine points instead of 64, and 16-byte coalesced read-modify-write of p.
// mode 2 = mode 1 (pf += interp) on a neighbour-free fine level: the Neumann halo of pf is written too (no fill_halo launch)
template <int mode>
__global__ void __launch_bounds__(256, 3) k_prolong8(const double* __restrict__ xc, int nzc, int nyc, int nxc,
                                                 double* __restrict__ rf, double* __restrict__ pf) {
  const int n = nzc * nyc * nxc;                                              // 32-bit element indices throughout
  const int e_raw = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e_raw < n;
  const int e = valid ? e_raw : n - 1;                                        // idle lanes still take part in the shuffles
  const int lane = threadIdx.x & 31;
  const int t = e / nzc, k2 = 1 + (e - t * nzc);
  const int i2 = 1 + t / nyc, j2 = 1 + (t - (i2 - 1) * nyc);                  // Fortran indices
  const int nyc2 = nyc + 2, nzf = 2 * nzc, nyf2 = 2 * nyc + 2;
  const double a = 9.0 / 16.0, b = 3.0 / 16.0, c = 1.0 / 16.0, d = 27.0 / 64.0, ee = 9.0 / 64.0, f = 3.0 / 64.0, g = 1.0 / 64.0;
  // the four fine (k,k+1) pairs of p are the only DRAM reads: issue them first so they fly during the coarse loads
  int of[4]; double2 pv[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int jf = 2 * j2 - 1 + (q >> 1), i_f = 2 * i2 - 1 + (q & 1);
    of[q] = 2 * (k2 - 1) + nzf * (jf + nyf2 * i_f);
    pv[q] = (mode == 0) ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(pf + of[q]);   // modes 1, 2
  }
  // coarse value at (k2, j2+dj, i2+di) and its vertical neighbours: the levels k2-1 / k2+1 of the same column sit in the
  // adjacent lanes (k2 is the fastest index), so they travel by shuffle; only a warp-boundary lane loads them
  const bool need_lo = (lane == 0) && (k2 > 1), need_hi = (lane == 31) && (k2 < nzc);
  auto ld3 = [&](int dj, int di, double& xm, double& x0, double& xp) {
    const int o = nzc * ((j2 + dj) + nyc2 * (i2 + di)) + k2 - 1;
    x0 = xc[o];
    xm = __shfl_up_sync(0xffffffffu, x0, 1); xp = __shfl_down_sync(0xffffffffu, x0, 1);
    if (need_lo) xm = xc[o - 1];
    if (need_hi) xp = xc[o + 1];
  };
  double Cm[3], C0[3], Cp[3];                                                  // centre row (dj = 0), di = -1, 0, +1
#pragma unroll
  for (int di = -1; di <= 1; ++di) ld3(0, di, Cm[di + 1], C0[di + 1], Cp[di + 1]);
#pragma unroll
  for (int cj = 0; cj < 2; ++cj) {
    const int dj = cj ? +1 : -1;
    double Sm[3], S0[3], Sp[3];                                                // side row (dj), di = -1, 0, +1
#pragma unroll
    for (int di = -1; di <= 1; ++di) ld3(dj, di, Sm[di + 1], S0[di + 1], Sp[di + 1]);
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
      const int di = ci ? +1 : -1, q = cj * 2 + ci;
      const double x00 = C0[1], xdd = S0[di + 1], xd0 = S0[1], x0d = C0[di + 1];
      double vlo, vhi;
      if (k2 == 1) vlo = a * x00 + c * xdd + b * xd0 + b * x0d;                                  // fine K = 1
      else { const double y00 = Cm[1], ydd = Sm[di + 1], yd0 = Sm[1], y0d = Cm[di + 1];
             vlo = d * x00 + f * xdd + ee * xd0 + ee * x0d + ee * y00 + g * ydd + f * yd0 + f * y0d; }
      if (k2 == nzc) { vhi = a * x00 + c * xdd + b * xd0 + b * x0d; vhi = 0.5 * vhi; }            // fine K = nzf
      else { const double y00 = Cp[1], ydd = Sp[di + 1], yd0 = Sp[1], y0d = Cp[di + 1];
             vhi = d * x00 + f * xdd + ee * xd0 + ee * x0d + ee * y00 + g * ydd + f * yd0 + f * y0d; }
      if (!valid) continue;
      if (mode == 0) *reinterpret_cast<double2*>(rf + of[q]) = make_double2(vlo, vhi);
      else {
        const double2 v = make_double2(pv[q].x + vlo, pv[q].y + vhi);
        *reinterpret_cast<double2*>(pf + of[q]) = v;
        if (mode == 2) {      // mirrors of an edge / corner fine column (16-byte aligned: nzf is even)
          const int jf = 2 * j2 - 1 + cj, i_f = 2 * i2 - 1 + ci, nyf = 2 * nyc, nxf = 2 * nxc;
          const int sj = nzf, si = nzf * nyf2;
          const bool w = (i_f == 1), ea = (i_f == nxf), so = (jf == 1), no = (jf == nyf);
          if (w) *reinterpret_cast<double2*>(pf + of[q] - si) = v;
          if (ea) *reinterpret_cast<double2*>(pf + of[q] + si) = v;
          if (so) *reinterpret_cast<double2*>(pf + of[q] - sj) = v;
          if (no) *reinterpret_cast<double2*>(pf + of[q] + sj) = v;
          if (w && so) *reinterpret_cast<double2*>(pf + of[q] - si - sj) = v;
          if (w && no) *reinterpret_cast<double2*>(pf + of[q] - si + sj) = v;
          if (ea && so) *reinterpret_cast<double2*>(pf + of[q] + si - sj) = v;
          if (ea && no) *reinterpret_cast<double2*>(pf + of[q] + si + sj) = v;
        }
      }
    }
  }
}
void launch_prolong(cudaStream_t st, const double* xc, int nzc, int nyc, int nxc, double* rf, double* pf, int linear, int mode) {
  if (linear) {
    size_t n = (size_t)nzc * nyc * nxc;
    if (mode == 0) launch_k(k_prolong8<0>, blocks_for(n, 256), 256, 0, st, xc, nzc, nyc, nxc, rf, pf);
    else if (mode == 1) launch_k(k_prolong8<1>, blocks_for(n, 256), 256, 0, st, xc, nzc, nyc, nxc, rf, pf);
    else launch_k(k_prolong8<2>, blocks_for(n, 256), 256, 0, st, xc, nzc, nyc, nxc, rf, pf);
    return;
  }
  size_t n = (size_t)8 * nzc * nyc * nxc;
  launch_k(k_prolong, blocks_for(n, 256), 256, 0, st, xc, nzc, nyc, nxc, rf, pf, linear, mode);
}

// =====================================================================================================
// gather / split
// =====================================================================================================
__global__ void k_gather_place(const double* __restrict__ x, int nzl, int nys, int nxs, double* __restrict__ y, int ny, int l, int m) {
  size_t n = (size_t)nzl * nys * nxs;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int k = (int)(e % nzl); size_t t = e / nzl; int j = 1 + (int)(t % nys), i = 1 + (int)(t / nys);
  y[IDX3(k, j + m * nys, i + l * nxs, nzl, ny + 2)] = x[IDX3(k, j, i, nzl, nys + 2)];
}
void launch_gather_place(cudaStream_t st, const double* x, int nzl, int nys, int nxs, double* y, int ny, int nx, int l, int m) {
  (void)nx; size_t n = (size_t)nzl * nys * nxs;
  launch_k(k_gather_place, blocks_for(n, 256), 256, 0, st, x, nzl, nys, nxs, y, ny, l, m);
}
__global__ void k_split(const double* __restrict__ x, int nzl, int ny, double* __restrict__ y, int nys, int nxs, int l, int m) {
  size_t n = (size_t)nzl * (nys + 2) * (nxs + 2);
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int k = (int)(e % nzl); size_t t = e / nzl; int j = (int)(t % (nys + 2)), i = (int)(t / (nys + 2));
  y[e] = x[IDX3(k, j + m * nys, i + l * nxs, nzl, ny + 2)];
}
void launch_split(cudaStream_t st, const double* x, int nzl, int ny, int nx, double* y, int nys, int nxs, int l, int m) {
  (void)nx; size_t n = (size_t)nzl * (nys + 2) * (nxs + 2);
  launch_k(k_split, blocks_for(n, 256), 256, 0, st, x, nzl, ny, y, nys, nxs, l, m);
}

// =====================================================================================================
// cA layout conversion for the accessors (AoS of the reference <-> device SoA planes)
// =====================================================================================================
__global__ void k_soa_to_aos8(double* const* __restrict__ planes, double* __restrict__ aos, size_t npts) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; e < npts * 8; e += stride) aos[e] = planes[e & 7][e >> 3];
}
__global__ void k_aos8_to_soa(const double* __restrict__ aos, double* const* __restrict__ planes, size_t npts) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; e < npts * 8; e += stride) planes[e & 7][e >> 3] = aos[e];
}
void launch_soa_to_aos8(cudaStream_t st, double* const* planes_dev, double* aos, size_t npts) {
  launch_k(k_soa_to_aos8, capped_blocks(npts * 8, 256), 256, 0, st, planes_dev, aos, npts);
}
void launch_aos8_to_soa(cudaStream_t st, const double* aos, double* const* planes_dev, size_t npts) {
  launch_k(k_aos8_to_soa, capped_blocks(npts * 8, 256), 256, 0, st, aos, planes_dev, npts);
}

}  // namespace mgb

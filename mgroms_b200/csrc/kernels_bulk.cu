// kernels_bulk.cu -- coarse2fine on big levels with the fine field staged through shared memory by the bulk-copy engine
// (cp.async.bulk + mbarrier, the 1-D form of TMA; SASS: UBLKCP).
//
// Reference: coarse2fine_3D_linear (src/mg_intergrids.f90:366-450) fused with p = p + r (:226), as k_prolong8<1>.
// The register version (kernels_solve.cu: k_prolong8) keeps the four 16-byte loads of p in registers while it interpolates: at
// three CTAs per SM that is ~48 KB in flight per SM, short of what the HBM latency-bandwidth product needs (0.59-0.65 of the
// copy roofline).  Here a persistent CTA owns tiles of one coarse row x 8 coarse columns x all levels = two contiguous
// chunks of the fine field (rows 2 i2 - 1 and 2 i2, 16 fine columns each).  One thread asks the copy engine for the next
// tile's chunks (8-32 KB in flight per CTA, no registers), everybody adds the interpolated values to the tile that has
// arrived -- in shared memory -- and the copy engine writes it back.  Same arithmetic, expression by expression.
#include "kernels.cuh"
#include "device_utils.cuh"

#include <cstdlib>

namespace mgb {

namespace {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes),
               "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int PB_TJ = 8;         // coarse columns per tile
constexpr int PB_THREADS = 256;

// pf += P xc on the interior (the caller refreshes the halo of pf).  NZC = nzc (power of two, >= 8).
template <int NZC>
__global__ void __launch_bounds__(PB_THREADS) k_prolong_bulk(const double* __restrict__ xc, int nyc, int nxc, double* __restrict__ pf, int ntj, int ntiles) {
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) unsigned long long full[2];
  constexpr int nzf = 2 * NZC;
  constexpr int CHUNK = 2 * PB_TJ * nzf;                     // doubles of one fine row of a tile (16 fine columns)
  const int nyc2 = nyc + 2, nyf2 = 2 * nyc + 2;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const double a = 9.0 / 16.0, b = 3.0 / 16.0, c = 1.0 / 16.0, d = 27.0 / 64.0, ee = 9.0 / 64.0, f = 3.0 / 64.0, g = 1.0 / 64.0;
  // tile t -> coarse row i2 = 1 + t / ntj, coarse columns j2a .. j2a + tj - 1
  auto geom = [&](int t, int& i2, int& j2a, int& tj) { i2 = 1 + t / ntj; j2a = 1 + (t - (i2 - 1) * ntj) * PB_TJ; tj = min(PB_TJ, nyc - j2a + 1); };
  auto issue = [&](int t, int stage) {                         // thread 0: both fine rows of tile t -> stage
    int i2, j2a, tj; geom(t, i2, j2a, tj);
    const unsigned bytes = (unsigned)(2 * tj * nzf * sizeof(double));
    mbar_expect_tx(&full[stage], 2 * bytes);
#pragma unroll
    for (int r = 0; r < 2; ++r)
      bulk_load(sm + (stage * 2 + r) * CHUNK, pf + (size_t)nzf * ((size_t)(2 * j2a - 1) + (size_t)nyf2 * (size_t)(2 * i2 - 1 + r)), bytes, &full[stage]);
  };
  int it = 0;
  if (threadIdx.x == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    if (threadIdx.x == 0 && tile + (int)gridDim.x < ntiles) {
      bulk_wait_read0();                                       // the write-back that last read the other stage has left shared memory
      issue(tile + gridDim.x, stage ^ 1);
    }
    int i2, j2a, tj; geom(tile, i2, j2a, tj);
    double* s0 = sm + (stage * 2) * CHUNK;                     // fine row 2 i2 - 1 ; + CHUNK: fine row 2 i2
    // interpolated values first (the coarse field comes from L1 / L2), then the tile, which has been flying in meanwhile
    for (int e0 = 0; e0 < tj * NZC; e0 += PB_THREADS) {
      const int e_raw = e0 + threadIdx.x;
      const bool valid = e_raw < tj * NZC;
      const int e = valid ? e_raw : tj * NZC - 1;              // idle lanes still take part in the shuffles
      const int jl = e / NZC, k2 = 1 + (e - jl * NZC);
      const int j2 = j2a + jl;
      const bool need_lo = (lane == 0) && (k2 > 1), need_hi = (lane == 31) && (k2 < NZC);
      auto ld3 = [&](int dj, int di, double& xm, double& x0, double& xp) {
        const int o = NZC * ((j2 + dj) + nyc2 * (i2 + di)) + k2 - 1;
        x0 = xc[o];
        xm = __shfl_up_sync(0xffffffffu, x0, 1); xp = __shfl_down_sync(0xffffffffu, x0, 1);
        if (need_lo) xm = xc[o - 1];
        if (need_hi) xp = xc[o + 1];
      };
      double Cm[3], C0[3], Cp[3];
#pragma unroll
      for (int di = -1; di <= 1; ++di) ld3(0, di, Cm[di + 1], C0[di + 1], Cp[di + 1]);
      double vlo[4], vhi[4];
#pragma unroll
      for (int cj = 0; cj < 2; ++cj) {
        const int dj = cj ? +1 : -1;
        double Sm[3], S0[3], Sp[3];
#pragma unroll
        for (int di = -1; di <= 1; ++di) ld3(dj, di, Sm[di + 1], S0[di + 1], Sp[di + 1]);
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int di = ci ? +1 : -1, q = cj * 2 + ci;
          const double x00 = C0[1], xdd = S0[di + 1], xd0 = S0[1], x0d = C0[di + 1];
          if (k2 == 1) vlo[q] = a * x00 + c * xdd + b * xd0 + b * x0d;                                  // fine K = 1
          else { const double y00 = Cm[1], ydd = Sm[di + 1], yd0 = Sm[1], y0d = Cm[di + 1];
                 vlo[q] = d * x00 + f * xdd + ee * xd0 + ee * x0d + ee * y00 + g * ydd + f * yd0 + f * y0d; }
          if (k2 == NZC) { double t = a * x00 + c * xdd + b * xd0 + b * x0d; vhi[q] = 0.5 * t; }          // fine K = nzf
          else { const double y00 = Cp[1], ydd = Sp[di + 1], yd0 = Sp[1], y0d = Cp[di + 1];
                 vhi[q] = d * x00 + f * xdd + ee * xd0 + ee * x0d + ee * y00 + g * ydd + f * yd0 + f * y0d; }
        }
      }
      if (e0 == 0) mbar_wait(&full[stage], (unsigned)((it >> 1) & 1));
      if (valid) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {                          // q = 2 cj + ci: fine column 2 jl + cj of fine row ci
          double2* w = reinterpret_cast<double2*>(s0 + (q & 1) * CHUNK + nzf * (2 * jl + (q >> 1)) + 2 * (k2 - 1));
          const double2 pv = *w;
          *w = make_double2(pv.x + vlo[q], pv.y + vhi[q]);
        }
      }
    }
    __syncthreads();                                           // the tile is complete in shared memory
    if (threadIdx.x == 0) {
      fence_async_smem();
      const unsigned bytes = (unsigned)(2 * tj * nzf * sizeof(double));
#pragma unroll
      for (int r = 0; r < 2; ++r)
        bulk_store(pf + (size_t)nzf * ((size_t)(2 * j2a - 1) + (size_t)nyf2 * (size_t)(2 * i2 - 1 + r)), sm + (stage * 2 + r) * CHUNK, bytes);
      bulk_commit();
    }
  }
  if (threadIdx.x == 0) bulk_wait_all();
}

template <int NZC> constexpr size_t pb_smem() { return (size_t)4 * 2 * PB_TJ * 2 * NZC * sizeof(double); }
template <int NZC>
int pb_cap() {
  static const int cap = []() {
    int per = 0;
    cudaFuncSetAttribute(k_prolong_bulk<NZC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb_smem<NZC>());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_prolong_bulk<NZC>, PB_THREADS, pb_smem<NZC>());
    static const int want = []() { const char* e = std::getenv("MGB_BULK_PROLONG_CTAS"); return e ? std::atoi(e) : 6; }();
    if (per > want) per = want;
    return (per > 0 ? per : 1) * device_sm_count();
  }();
  return cap;
}
template <int NZC>
void launch_pb(cudaStream_t st, const double* xc, int nyc, int nxc, double* pf) {
  constexpr size_t smem = pb_smem<NZC>();
  const int cap = pb_cap<NZC>();
  const int ntj = (nyc + PB_TJ - 1) / PB_TJ, ntiles = ntj * nxc;
  launch_k(k_prolong_bulk<NZC>, ntiles < cap ? ntiles : cap, PB_THREADS, smem, st, xc, nyc, nxc, pf, ntj, ntiles);
}

}  // namespace

bool prolong_bulk_supported(int nzc, int nyc, int nxc) {
  static const bool off = []() { const char* e = std::getenv("MGB_NO_BULK_PROLONG"); return e && e[0] == '1'; }();
  static const size_t minpts = []() { const char* e = std::getenv("MGB_BULK_PROLONG_MIN"); return e ? (size_t)std::atoll(e) : (size_t)200000; }();
  if (off) return false;
  if (!(nzc == 8 || nzc == 16 || nzc == 32 || nzc == 64)) return false;
  return (size_t)nzc * nyc * nxc >= minpts;                    // big levels; the small ones are latency bound and keep the fused-halo kernel
}
void prolong_bulk_init() {
  // occupancy queries and shared-memory attributes outside any stream capture
  (void)pb_cap<8>(); (void)pb_cap<16>(); (void)pb_cap<32>(); (void)pb_cap<64>();
}
void launch_prolong_bulk(cudaStream_t st, const double* xc, int nzc, int nyc, int nxc, double* pf) {
  switch (nzc) {
    case 64: launch_pb<64>(st, xc, nyc, nxc, pf); break;
    case 32: launch_pb<32>(st, xc, nyc, nxc, pf); break;
    case 16: launch_pb<16>(st, xc, nyc, nxc, pf); break;
    default: launch_pb<8>(st, xc, nyc, nxc, pf); break;
  }
}

}  // namespace mgb

// kernels_relax.cu -- one colour of the red-black line relaxation in ONE persistent launch (kernels A + B + C fused).
//
// Reference: relax_3D_8_RB + relax_3D_8_heart + tridiag (src/mg_relax.f90:151-334) in the exact rank-1 form of
// DESIGN.md section 3:  q = T^-1 rhs0 (A),  s from the chunked k=1 recurrence (B),  p = q + s g (C).
// The three-launch path (kernels_solve.cu) writes q to HBM, runs B while the GPU idles, and re-reads q in C.  Here a tile
// stays in shared memory until its s is known, so p is written once: 68 + 4 B/pt (the read of g) instead of 80.
//
// A CTA is NPW producer warps + one Thomas warp + one recurrence warp and walks tiles of 8 rows (one recurrence chunk)
// x TJ columns of the colour:
//   producers   P1(n): 27 operands per point -> rhs / pivots / off-diagonal into shared memory (slot n%3 / n%2), and in the
//               same pass the write-back WB(n-2): p = q + s g of the tile solved two steps earlier (g is the only load)
//   Thomas warp T(n):  one lane per column, in place; publishes q1 = q(k=1) to the 2-D plane and raises the tile's flag
//   recurrence  S(n):  waits for the flags of the tiles its dependency cone touches (same chunk +-d tiles, the chunks of
//               the warm-up rows), stages their q1 / c5b / c8b / g1 in shared memory (the tile's dead pivot slot) and runs
//               the chunk recurrence on the cone with warp shuffles -> s of the tile's columns
// Named barriers hand the slots over; flags are epoch-stamped (no reset between launches); the k=1 plane is double
// buffered (x1old read, x1new written completely) because tiles of other CTAs still read the old values.
// Same arithmetic, operation by operation, as k_relax_A4 / k_relax_B / k_relax_C.
#include "kernels.cuh"
#include "device_utils.cuh"

#include <cstdlib>

namespace mgb {

namespace {

template <int NZ> struct FGeom {
  static constexpr int CH = RB_CHUNK;          // rows per tile = one chunk of the bottom recurrence
  static constexpr int TJ = 128 / NZ;          // colour columns per row of a tile (NZ 128 -> 1, 64 -> 2, 32 -> 4, 16 -> 8)
  static constexpr int TC = CH * TJ;           // columns per tile (TC * NZ = 1024 points)
  static constexpr int PITCH = NZ + 1;
  static constexpr int SLOT = TC * PITCH;      // doubles per array slot
  static constexpr int SMEM_DOUBLES = 7 * SLOT + 3 * TC;   // 3 rhs slots, 2 x (bet, ud) slots, 3 x s per column
};

// A flag that never rises (a lost CTA, a bug) must not hang the GPU: after L.timeout_ms of wall clock the wait gives up and
// raises the mapped host word (the API call fails at its end) and the device abort flag (later waits return at once).
__device__ __forceinline__ bool wait_flag(const unsigned* f, unsigned ep, const LevelView& L) {
  if (ld_acquire_gpu_u32(f) == ep) return true;
  const unsigned long long t0 = globaltimer_ns();
  for (int spins = 1;; ++spins) {
    __nanosleep(40);
    if (ld_acquire_gpu_u32(f) == ep) return true;
    if ((spins & 255) == 0) {
      if (*(volatile int*)L.abort_flag) return false;
      if (globaltimer_ns() - t0 > (unsigned long long)L.timeout_ms * 1000000ull) { *L.err = 2; *(volatile int*)L.abort_flag = 1; return false; }
    }
  }
}

// recurrence warp: s of the columns of `tile` (see the header); stg = staging area (2 * SLOT doubles), s_out[TC]
template <int NZ>
__device__ __forceinline__ void fused_S(const LevelView& L, int rb, const double* __restrict__ x1old, int tile, int ntj, unsigned ep,
                                        double* __restrict__ stg, double* __restrict__ s_out) {
  using G = FGeom<NZ>;
  const int lane = threadIdx.x & 31;
  const int nx = L.nx, ny = L.ny, ny2 = ny + 2, nyc = ny >> 1;
  const int chunk = tile / ntj, t = tile - chunk * ntj;
  const int c0 = 1 + chunk * G::CH, c1 = min(nx, c0 + G::CH - 1), i0 = max(1, c0 - L.rb_W);
  const int R = c1 - i0 + 1;
  const int mA = t * G::TJ, mB = min(nyc - 1, mA + G::TJ - 1);
  const int Mw = (R + 1) >> 1, TP = G::TJ + 2 * Mw, m0 = mA - Mw;   // window of colour-pair indices m0 .. m0 + TP - 1 (TP <= 32)
  // 1. the tiles whose q1 the cone reads: chunks of rows i0..c1, tile columns of the window
  {
    const int clo = (i0 - 1) / G::CH;
    const int tlo = max(0, m0) / G::TJ, thi = min(nyc - 1, mB + Mw) / G::TJ, ntc = thi - tlo + 1;
    const int nf = (chunk - clo + 1) * ntc;
    for (int f = lane; f < nf; f += 32) {
      const int cc = clo + f / ntc, tt = tlo + f % ntc;
      const int other = cc * ntj + tt;
      if (other != tile) wait_flag(L.tflags + other, ep, L);
    }
    __syncwarp();
    __threadfence();
  }
  // 2. operands of the cone: [row][plane][lane] ; halo values x1old(0, i-1), x1old(ny+1, i-1) per row
  double* hal = stg + 4 * R * TP;
  for (int x = lane; x < R * TP; x += 32) {
    const int rr = x / TP, ll = x - rr * TP;
    const int i = i0 + rr, m = m0 + ll;
    if (m >= 0 && m < nyc) {
      const int c2d = 1 + ((i + rb) & 1) + 2 * m + ny2 * i;
      double* o = stg + (size_t)(rr * 4) * TP + ll;
      o[0] = __ldcg(L.q1 + c2d); o[TP] = L.c5b[c2d]; o[2 * TP] = L.c8b[c2d]; o[3 * TP] = L.g1[c2d];
    }
  }
  for (int x = lane; x < 2 * R; x += 32) hal[x] = x1old[(size_t)((x & 1) ? ny + 1 : 0) + (size_t)ny2 * (i0 - 1 + (x >> 1))];
  __syncwarp();
  // 3. the recurrence, one lane per colour pair index; the previous row travels by shuffles
  const int m = m0 + lane;
  const bool valid = lane < TP && m >= 0 && m < nyc;
  int js = 1 + ((i0 - 1 + rb) & 1);
  double xv = valid ? x1old[(size_t)(js + 2 * m) + (size_t)ny2 * (i0 - 1)] : 0.0;
  for (int i = i0; i <= c1; ++i) {
    js = 3 - js;
    const int rr = i - i0;
    const double up = __shfl_up_sync(0xffffffffu, xv, 1), dn = __shfl_down_sync(0xffffffffu, xv, 1);
    double xp, xmn;                                            // x1(j+1,i-1), x1(j-1,i-1)
    if (js == 1) { xp = xv; xmn = (m == 0) ? hal[2 * rr] : up; } else { xmn = xv; xp = (m == nyc - 1) ? hal[2 * rr + 1] : dn; }
    const double* o = stg + (size_t)(rr * 4) * TP + lane;
    const double qv = valid ? o[0] : 0.0, a5 = valid ? o[TP] : 0.0, a8 = valid ? o[2 * TP] : 0.0, gv = valid ? o[3 * TP] : 0.0;
    const double sv = -(a5 * xp + a8 * xmn);
    xv = qv + sv * gv;
    if (valid && i >= c0 && m >= mA && m <= mB) s_out[(i - c0) * G::TJ + (m - mA)] = sv;
  }
}

// barrier ids: 1,2 = slot full (producers -> Thomas) ; 3,4 = solved (Thomas -> recurrence) ; 5,6 = s ready (recurrence -> producers)
template <int NZ, int NPW, int MINB>
__global__ void __launch_bounds__((NPW + 2) * 32, MINB)
k_relax_F(LevelView L, int rb, int mirror, const double* __restrict__ x1old, double* __restrict__ x1new, int ntj, int ntiles) {
  using G = FGeom<NZ>;
  extern __shared__ __align__(16) double sm[];
  constexpr int NPROD = NPW * 32, NFULL = NPROD + 32, NSRDY = NPROD + 32;
  constexpr int sj = NZ;
  const int nx = L.nx, ny = L.ny, ny2 = ny + 2, nyc = ny >> 1;
  const int si = NZ * ny2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const s_sv = sm + 7 * G::SLOT;
  const unsigned ep = *reinterpret_cast<volatile unsigned*>(L.tctl) + 1u;   // number of this launch; advanced by the last CTA to leave
  const int Gd = gridDim.x;
  if (warp < NPW) {
    const double* __restrict__ p = L.p; const double* __restrict__ b = L.b;
    const double* __restrict__ c2 = L.cA[1]; const double* __restrict__ c3 = L.cA[2]; const double* __restrict__ c4 = L.cA[3];
    const double* __restrict__ c5 = L.cA[4]; const double* __restrict__ c6 = L.cA[5]; const double* __restrict__ c7 = L.cA[6];
    const double* __restrict__ c8 = L.cA[7]; const double* __restrict__ bet = L.bet; const double* __restrict__ g = L.g;
    double* __restrict__ pw = L.p;
    // final value of (k, column) of a solved tile: p = q + s g, Neumann mirrors, upkeep of the new k=1 plane
    auto commit = [&](int idxw, int k, int jw, int iw, double v) {
      if (!mirror) pw[idxw] = v; else store_mirrored(pw, (size_t)idxw, v, jw, iw, ny, nx, (size_t)sj, (size_t)si, true);
      if (k == 0) {
        const size_t c2d = (size_t)jw + (size_t)ny2 * iw;
        store_mirrored(x1new, c2d, v, jw, iw, ny, nx, 1, (size_t)ny2, mirror != 0);
        const int jp = ((jw & 1) ? jw + 1 : jw - 1);            // the other-colour cell of the pair keeps its value
        const size_t cp = (size_t)jp + (size_t)ny2 * iw;
        store_mirrored(x1new, cp, x1old[cp], jp, iw, ny, nx, 1, (size_t)ny2, mirror != 0);
      }
    };
    int n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += Gd, ++n) {
      double* s_rhs = sm + (n % 3) * G::SLOT;
      double* s_bet = sm + 3 * G::SLOT + (n & 1) * 2 * G::SLOT; double* s_ud = s_bet + G::SLOT;
      const bool wb = n >= 2;
      const double* s_q = sm + ((n + 1) % 3) * G::SLOT;        // slot of tile n-2
      const double* s_s = s_sv + ((n + 1) % 3) * G::TC;
      if (wb) named_bar_sync(5 + (n & 1), NSRDY);               // s of tile n-2 is known; every producer has left iteration n-1
      const int chunk = tile / ntj, c0 = 1 + chunk * G::CH, mA = (tile - chunk * ntj) * G::TJ;
      const int wtile = tile - 2 * Gd, wchunk = wtile / ntj, wc0 = 1 + wchunk * G::CH, wmA = (wtile - wchunk * ntj) * G::TJ;
      for (int e = threadIdx.x; e < G::TC * NZ; e += NPROD) {
        const int k = e & (NZ - 1), c = e / NZ, r = c / G::TJ, cj = c - r * G::TJ;
        // ---- write-back operand of tile n-2 (issued with the loads below)
        int idxw = -1, jw = 0, iw = 0; double gw = 0.0;
        if (wb) {
          iw = wc0 + r; const int mw = wmA + cj;
          if (iw <= nx && mw < nyc) { jw = 1 + ((iw + rb) & 1) + 2 * mw; idxw = k + NZ * (jw + ny2 * iw); gw = g[idxw]; }
        }
        // ---- rhs of tile n
        const int i = c0 + r, m = mA + cj;
        if (i <= nx && m < nyc) {
          const int j = 1 + ((i + rb) & 1) + 2 * m;
          const int idx = k + NZ * (j + ny2 * i);
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
          if (k == 0) {                                         // the two (i+1) diagonal terms from the old k=1 plane
            const int c2d = j + ny2 * i;
            acc = acc - L.c5b[c2d - 1 + ny2] * x1old[c2d - 1 + ny2];
            acc = acc - L.c8b[c2d + 1 + ny2] * x1old[c2d + 1 + ny2];
          }
          const int so = c * G::PITCH + k;
          s_rhs[so] = acc; s_bet[so] = bt; s_ud[so] = hi ? udv : 0.0;
        }
        if (idxw >= 0) commit(idxw, k, jw, iw, s_q[c * G::PITCH + k] + s_s[c] * gw);
      }
      __threadfence_block();
      named_bar_arrive(1 + (n & 1), NFULL);                     // slot full
    }
    // drain: the last (up to two) tiles of this CTA
    for (int m2 = (n >= 2 ? n - 2 : 0); m2 < n; ++m2) {
      named_bar_sync(5 + (m2 & 1), NSRDY);
      const double* s_q = sm + (m2 % 3) * G::SLOT; const double* s_s = s_sv + (m2 % 3) * G::TC;
      const int wtile = (int)blockIdx.x + m2 * Gd, wchunk = wtile / ntj, wc0 = 1 + wchunk * G::CH, wmA = (wtile - wchunk * ntj) * G::TJ;
      for (int e = threadIdx.x; e < G::TC * NZ; e += NPROD) {
        const int k = e & (NZ - 1), c = e / NZ, r = c / G::TJ, cj = c - r * G::TJ;
        const int iw = wc0 + r, mw = wmA + cj;
        if (iw > nx || mw >= nyc) continue;
        const int jw = 1 + ((iw + rb) & 1) + 2 * mw, idxw = k + NZ * (jw + ny2 * iw);
        commit(idxw, k, jw, iw, s_q[c * G::PITCH + k] + s_s[c] * g[idxw]);
      }
    }
  } else if (warp == NPW) {
    // ---- Thomas warp
    int n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += Gd, ++n) {
      double* s_rhs = sm + (n % 3) * G::SLOT;
      const double* s_bet = sm + 3 * G::SLOT + (n & 1) * 2 * G::SLOT; const double* s_ud = s_bet + G::SLOT;
      named_bar_sync(1 + (n & 1), NFULL);
      const int chunk = tile / ntj, c0 = 1 + chunk * G::CH, mA = (tile - chunk * ntj) * G::TJ;
      for (int c = lane; c < G::TC; c += 32) {
        const int r = c / G::TJ, cj = c - r * G::TJ;
        const int i = c0 + r, m = mA + cj;
        if (i > nx || m >= nyc) continue;
        double* x = s_rhs + c * G::PITCH; const double* bt = s_bet + c * G::PITCH; const double* ud = s_ud + c * G::PITCH;
        double xm = x[0] * bt[0];
        x[0] = xm;
        for (int k = 1; k < NZ; ++k) { xm = (x[k] - ud[k - 1] * xm) * bt[k]; x[k] = xm; }
        for (int k = NZ - 2; k >= 0; --k) { xm = x[k] - (ud[k] * bt[k]) * xm; x[k] = xm; }
        L.q1[(size_t)(1 + ((i + rb) & 1) + 2 * m) + (size_t)ny2 * i] = xm;
      }
      __threadfence();                                          // q1 is out before the flag
      __syncwarp();
      if (lane == 0) st_release_gpu_u32(L.tflags + tile, ep);
      __threadfence_block();
      named_bar_arrive(3 + (n & 1), 64);                        // solved: the recurrence warp may start on this tile
    }
  } else {
    // ---- recurrence warp
    int n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += Gd, ++n) {
      named_bar_sync(3 + (n & 1), 64);
      double* stg = sm + 3 * G::SLOT + (n & 1) * 2 * G::SLOT;   // pivots of tile n are dead after T(n); P1(n+2) waits for this warp
      fused_S<NZ>(L, rb, x1old, tile, ntj, ep, stg, s_sv + (n % 3) * G::TC);
      __syncwarp();
      __threadfence_block();
      named_bar_arrive(5 + (n & 1), NSRDY);
    }
  }
  // the last CTA to leave advances the epoch for the next launch
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); if (atomicAdd(L.tctl + 1, 1u) == (unsigned)Gd - 1u) { L.tctl[1] = 0u; __threadfence(); L.tctl[0] = ep; } }
}

struct FVariant { int npw, minb; };
FVariant fused_variant() {
  static const int v = []() { const char* e = std::getenv("MGB_F_VARIANT"); return e ? std::atoi(e) : 0; }();
  switch (v) {
    case 1: return {6, 3};      // 256 threads, 3 CTAs / SM
    case 2: return {10, 2};     // 384 threads, 2 CTAs / SM
    case 3: return {8, 2};      // 320 threads, 2 CTAs / SM
    default: return {7, 3};     // 288 threads, 3 CTAs / SM
  }
}

template <int NZ, int NPW, int MINB>
int fused_grid_cap() {
  static const int cap = []() {
    int dev = 0, sms = 148, per = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = (size_t)FGeom<NZ>::SMEM_DOUBLES * sizeof(double);
    cudaFuncSetAttribute(k_relax_F<NZ, NPW, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_relax_F<NZ, NPW, MINB>, (NPW + 2) * 32, smem);
    return per * sms;                                           // every CTA must be co-resident (the recurrence warps wait on each other)
  }();
  return cap;
}

template <int NZ, int NPW, int MINB>
void launch_F(cudaStream_t st, const LevelView& L, int rb, int mirror, const double* x1old, double* x1new) {
  using G = FGeom<NZ>;
  const int nyc = L.ny / 2, ntj = (nyc + G::TJ - 1) / G::TJ, nchunks = (L.nx + G::CH - 1) / G::CH, ntiles = ntj * nchunks;
  // The recurrence warp of tile T waits for tiles up to T + d (d <= 16) of OTHER CTAs, whose own progress needs tiles up to
  // 2 * grid earlier: the wait graph is acyclic only when 2 * grid > d, hence the floor of 32 CTAs (relax_fused_supported).
  static const int user_cap = []() { const char* e = std::getenv("MGB_F_GRID_CAP"); const int v = e ? std::atoi(e) : 0; return v >= 32 ? v : 0; }();
  int cap = fused_grid_cap<NZ, NPW, MINB>();
  if (user_cap && user_cap < cap) cap = user_cap;
  if (cap < 32 || ntiles < 32) { note_launch_error(cudaErrorLaunchOutOfResources); return; }
  const int grid = ntiles < cap ? ntiles : cap;
  launch_k(k_relax_F<NZ, NPW, MINB>, grid, (NPW + 2) * 32, (size_t)G::SMEM_DOUBLES * sizeof(double), st, L, rb, mirror, x1old, x1new, ntj, ntiles);
}

template <int NZ>
void launch_F_nz(cudaStream_t st, const LevelView& L, int rb, int mirror, const double* x1old, double* x1new) {
  const FVariant v = fused_variant();
  if (v.npw == 6) launch_F<NZ, 6, 3>(st, L, rb, mirror, x1old, x1new);
  else if (v.npw == 10) launch_F<NZ, 10, 2>(st, L, rb, mirror, x1old, x1new);
  else if (v.npw == 8) launch_F<NZ, 8, 2>(st, L, rb, mirror, x1old, x1new);
  else launch_F<NZ, 7, 3>(st, L, rb, mirror, x1old, x1new);
}

}  // namespace

void relax_fused_init() {
  const FVariant v = fused_variant();
#define MGB_F_CAP(NZv) \
  (v.npw == 6 ? fused_grid_cap<NZv, 6, 3>() : v.npw == 10 ? fused_grid_cap<NZv, 10, 2>() : v.npw == 8 ? fused_grid_cap<NZv, 8, 2>() : fused_grid_cap<NZv, 7, 3>())
  (void)MGB_F_CAP(16); (void)MGB_F_CAP(32); (void)MGB_F_CAP(64); (void)MGB_F_CAP(128);
#undef MGB_F_CAP
}
bool relax_fused_supported(const LevelView& L) {
  static const bool off = []() { const char* e = std::getenv("MGB_NO_FUSED_RELAX"); return e && e[0] == '1'; }();
  if (off || !L.tflags || !L.tctl || !L.err || !L.abort_flag) return false;
  if (!(L.nz == 16 || L.nz == 32 || L.nz == 64 || L.nz == 128)) return false;
  if (L.rb_chunk != RB_CHUNK || (L.ny & 1)) return false;
  const int TJ = 128 / L.nz, TC = RB_CHUNK * TJ;
  const int R = L.rb_W + RB_CHUNK, Mw = (R + 1) / 2, TP = TJ + 2 * Mw;
  if (TP > 32) return false;                                    // the cone of a tile must fit one warp
  if ((size_t)4 * R * TP + 2 * R > (size_t)2 * TC * (L.nz + 1)) return false;   // ... and its operands the tile's dead pivot slot
  const int ntiles = ((L.ny / 2 + TJ - 1) / TJ) * ((L.nx + RB_CHUNK - 1) / RB_CHUNK);
  // big levels only (several tiles per CTA; small ones have their own kernels); never fewer than 32 tiles (see launch_F)
  static const int min_tiles = []() { const char* e = std::getenv("MGB_FUSED_MIN_TILES"); const int v = e ? std::atoi(e) : 1024; return v < 32 ? 32 : v; }();
  return ntiles >= min_tiles;
}
size_t relax_fused_num_tiles(int nx, int ny, int nz) {
  if (!(nz == 16 || nz == 32 || nz == 64 || nz == 128)) return 0;
  const int TJ = 128 / nz;
  return (size_t)((ny / 2 + TJ - 1) / TJ) * ((nx + RB_CHUNK - 1) / RB_CHUNK);
}
void launch_relax_fused(cudaStream_t st, const LevelView& L, int rb, int mirror, const double* x1old, double* x1new) {
  switch (L.nz) {
    case 128: launch_F_nz<128>(st, L, rb, mirror, x1old, x1new); break;
    case 64: launch_F_nz<64>(st, L, rb, mirror, x1old, x1new); break;
    case 32: launch_F_nz<32>(st, L, rb, mirror, x1old, x1new); break;
    default: launch_F_nz<16>(st, L, rb, mirror, x1old, x1new); break;
  }
}

}  // namespace mgb

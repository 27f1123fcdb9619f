// kernels_relax.cu -- one colour of the red-black line relaxation in ONE persistent launch (kernels A + B + C fused).
//
// Reference: relax_3D_8_RB + relax_3D_8_heart + tridiag (src/mg_relax.f90:151-334) in the exact rank-1 form of
// DESIGN.md section 3:  q = T^-1 rhs0 (A),  s from the chunked k=1 recurrence (B),  p = q + s g (C).
// The three-launch path (kernels_solve.cu) writes q to HBM, runs B while the GPU idles, and re-reads q in C.  Here a tile
// stays in shared memory until its s is known, so p is written once: 68 + 4 B/pt (the read of g) instead of 80.
//
// A CTA is NPW producer warps + a Thomas warp + a recurrence warp and walks tiles of 8 rows (one recurrence chunk) x TJ
// columns of the colour:
//   producers   P1(n): 27 operands per point -> rhs / pivots / off-diagonal into shared memory (slot n%3 / n%2), and in the
//               same pass the write-back WB(n-2): p = q + s g of the tile solved two steps earlier (g is the only load)
//   Thomas warp T(n):  one lane per column, in place (~55 cycles per level: FP64 latency bound); q1 = q(k=1) is published to a
//               2-D plane of TAGGED 16-byte words {lo, launch number, hi, launch number}: the value is its own arrival flag
//               (as in the halo mailboxes)
//   recurrence  S(n):  the cone of the k=1 plane that the tile's s depends on (its 8 rows + the warm-up rows, widening by one
//               colour column every two rows) is fetched in two or three batches of independent loads -- q1 of this tile and
//               of the tiles around it (other CTAs; a stale tag is simply re-read), c5b, c8b, g1 -- into the tile's dead
//               pivot slot, then the chunk recurrence runs on the cone with warp shuffles -> s of the tile's columns
// Memory round trips cost 1-2 us under load (a first version with a dozen dependent trips per tile ran at half speed), so
// the recurrence warp has two or three dependent trips per tile and overlaps the Thomas solve of the next tile.
// Named barriers hand the slots over; the k=1 plane is double buffered (x1old read, x1new written completely) because
// tiles of other CTAs still read the old values.
// Same arithmetic, operation by operation, as k_relax_A4 / k_relax_B / k_relax_C.
#include "kernels.cuh"
#include "device_utils.cuh"

#include <cstdio>
#include <cstdlib>

namespace mgb {

namespace {

template <int NZ> struct FGeom {
  static constexpr int CH = RB_CHUNK;          // rows per tile = one chunk of the bottom recurrence
  static constexpr int TJ = 128 / NZ;          // colour columns per row of a tile (NZ 128 -> 1, 64 -> 2, 32 -> 4)
  static constexpr int TC = CH * TJ;           // columns per tile (TC * NZ = 1024 points)
  static constexpr int PITCH = NZ + 1;
  static constexpr int SLOT = TC * PITCH;      // doubles per array slot
  static constexpr int SV = 7 * SLOT;          // 3 rhs slots, 2 x {pivot, off-diagonal} slots, then 3 x s per column, then the cone staging
};

// half-width of the cone d rows before the tile's last row: the dependency spreads by one colour column every two rows
__device__ __host__ __forceinline__ int cone_hw(int d) { return (d + 1) >> 1; }

// A tag that never arrives (a lost CTA, a bug) must not hang the GPU: after L.timeout_ms of wall clock the wait gives up and
// raises the mapped host word (the API call fails at its end) and the device abort flag (later waits return at once).
__device__ __forceinline__ uint4 ld_tagged_raw(const uint4* src) {
  unsigned long long a, b;                                     // two single-copy-atomic halves {lo | tag << 32, hi | tag << 32}
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
  return make_uint4((unsigned)a, (unsigned)(a >> 32), (unsigned)b, (unsigned)(b >> 32));
}
// w was read from src; re-read until both tags carry the launch number
__device__ __forceinline__ double tagged_value(uint4 w, const uint4* src, unsigned ep, const LevelView& L) {
  if (w.y != ep || w.w != ep) {
    const unsigned long long start = globaltimer_ns();
    for (int spins = 1;; ++spins) {
      __nanosleep(100);
      w = ld_tagged_raw(src);
      if (w.y == ep && w.w == ep) break;
      if ((spins & 127) == 0) {
        if (*(volatile int*)L.abort_flag) break;
        if (globaltimer_ns() - start > (unsigned long long)L.timeout_ms * 1000000ull) { *L.err = 2; *(volatile int*)L.abort_flag = 1; break; }
      }
    }
  }
  return __longlong_as_double((long long)(((unsigned long long)w.z << 32) | w.x));
}
// gpu-scope relaxed stores of the two 64-bit halves {lo | tag << 32, hi | tag << 32}: each half is single-copy atomic, which is all
// the reader needs (st.volatile would be system scope: the fences behind it then wait for system-level visibility)
__device__ __forceinline__ void st_tagged(uint4* dst, double v, unsigned ep) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  const unsigned long long a = (u & 0xffffffffull) | ((unsigned long long)ep << 32), b = (u >> 32) | ((unsigned long long)ep << 32);
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1,%2};" ::"l"(dst), "l"(a), "l"(b) : "memory");
}

// geometry of the cone of one tile
struct Cone { int c0, c1, i0, R, mA, mB, Mw, TP, m0; };
template <int NZ>
__device__ __forceinline__ Cone cone_of(const LevelView& L, int tile, int ntj) {
  using G = FGeom<NZ>;
  Cone c;
  const int nyc = L.ny >> 1;
  const int chunk = tile / ntj, t = tile - chunk * ntj;
  c.c0 = 1 + chunk * G::CH; c.c1 = min(L.nx, c.c0 + G::CH - 1); c.i0 = max(1, c.c0 - L.rb_W);
  c.R = c.c1 - c.i0 + 1;
  c.mA = t * G::TJ; c.mB = min(nyc - 1, c.mA + G::TJ - 1);
  c.Mw = cone_hw(c.R); c.TP = G::TJ + 2 * c.Mw; c.m0 = c.mA - c.Mw;     // window of colour-pair indices m0 .. m0 + TP - 1 (TP <= 32)
  return c;
}

// ---- recurrence warp.  Cone entries are numbered by distance d from the tile's last row (off[d] entries lie in closer rows),
// then by colour column: entry e of row d sits in window lane Mw - hw(d) + (e - off[d]).  Staging (dedicated shared memory):
//   cof[2][3][F_CMAX]  c5b, c8b, g1 of the cone, double buffered: fetched with cp.async one tile AHEAD (they do not depend on
//                      this launch), so that they cost no registers and no round trip on the critical path
//   q1s[F_CMAX]        q1 of the cone: ONE batch of tagged loads (up to 8 per lane) after the tile's own Thomas solve
//   hal[64]            x1old(0 | ny+1, i-1) of the cone rows (windows that touch the lateral halo only)
// A dependent memory round trip costs ~4 us in this kernel (the producers keep ~150 KB in flight per SM), hence one per tile.
constexpr int F_CMAX = 256;                    // cone entries per tile at most (8 tagged loads per lane)
constexpr int F_STG_DOUBLES = 6 * F_CMAX + F_CMAX + 64;

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}

// entry e -> (window lane, 2-D index of the cell) ; false when the cell lies outside the domain (no such colour column)
__device__ __forceinline__ bool cone_entry(const LevelView& L, int rb, const Cone& c, const int* __restrict__ off, int e, int& d, int& c2d) {
  while (off[d + 1] <= e) ++d;                                  // e only grows between calls of one lane
  const int ll = c.Mw - cone_hw(d) + (e - off[d]);
  const int i = c.c1 - d, m = c.m0 + ll;
  c2d = 1 + ((i + rb) & 1) + 2 * m + (L.ny + 2) * i;
  return m >= 0 && m < (L.ny >> 1);
}
__device__ __forceinline__ void cone_prefetch_coef(const LevelView& L, int rb, const Cone& c, const int* __restrict__ off, double* __restrict__ cof, int lane) {
  const int total = off[c.R];
  int d = 0, c2d;
  for (int e = lane; e < total; e += 32)
    if (cone_entry(L, rb, c, off, e, d, c2d)) { cp_async8(cof + e, L.c5b + c2d); cp_async8(cof + F_CMAX + e, L.c8b + c2d); cp_async8(cof + 2 * F_CMAX + e, L.g1 + c2d); }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
// q1 of the cone (tagged words, retried until they carry this launch's number) -> q1s ; halo values -> hal ; returns this
// lane's entry of the row before the cone (x1old at i0 - 1).  All loads are issued before the first one is used.
__device__ __forceinline__ double cone_fetch_q1(const LevelView& L, int rb, const double* __restrict__ x1old, const Cone& c, unsigned ep,
                                                const int* __restrict__ off, double* __restrict__ q1s, double* __restrict__ hal, int lane) {
  const int ny = L.ny, ny2 = ny + 2, nyc = ny >> 1;
  const int total = off[c.R];
  const uint4* __restrict__ q1t = reinterpret_cast<const uint4*>(L.q1t);
  const int m_l = c.m0 + lane;
  const int js0 = 1 + ((c.i0 - 1 + rb) & 1);
  const double xv0 = (lane < c.TP && m_l >= 0 && m_l < nyc) ? x1old[(size_t)(js0 + 2 * m_l) + (size_t)ny2 * (c.i0 - 1)] : 0.0;
  const bool need_hal = c.m0 <= 0 || c.m0 + c.TP >= nyc;
  double h0 = 0.0, h1 = 0.0;
  if (need_hal) {
    if (lane < 2 * c.R) h0 = x1old[(size_t)((lane & 1) ? ny + 1 : 0) + (size_t)ny2 * (c.i0 - 1 + (lane >> 1))];
    if (lane + 32 < 2 * c.R) h1 = x1old[(size_t)((lane & 1) ? ny + 1 : 0) + (size_t)ny2 * (c.i0 - 1 + ((lane + 32) >> 1))];
  }
  int cc[8]; uint4 qw[8];
  int d = 0;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int e = lane + 32 * u;
    cc[u] = -1;
    if (e < total) { int c2d; if (cone_entry(L, rb, c, off, e, d, c2d)) { cc[u] = c2d; qw[u] = ld_tagged_raw(q1t + c2d); } }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) if (cc[u] >= 0) q1s[lane + 32 * u] = tagged_value(qw[u], q1t + cc[u], ep, L);
  if (need_hal) { if (lane < 2 * c.R) hal[lane] = h0; if (lane + 32 < 2 * c.R) hal[lane + 32] = h1; }     // 2 R <= 64
  return xv0;
}

// the chunk recurrence on the cone, one lane per colour-pair index, the previous row by shuffles, the operands of the next
// row loaded from shared memory before the dependent chain of the current one
template <int NZ>
__device__ __forceinline__ void cone_recur(const LevelView& L, int rb, double xv, const Cone& c, const int* __restrict__ off,
                                           const double* __restrict__ cof, const double* __restrict__ q1s, const double* __restrict__ hal,
                                           double* __restrict__ s_out) {
  using G = FGeom<NZ>;
  const int lane = threadIdx.x & 31;
  const int nyc = L.ny >> 1;
  const int m = c.m0 + lane;
  const bool inrange = lane < c.TP && m >= 0 && m < nyc;
  int js = 1 + ((c.i0 - 1 + rb) & 1);
  const bool need_hal = c.m0 <= 0 || c.m0 + c.TP >= nyc;
  auto load_row = [&](int i, bool& valid, double& qv, double& a5, double& a8, double& gv, double& hv0, double& hv1) {
    const int d = c.c1 - i, hw = cone_hw(d);
    valid = inrange && lane >= c.Mw - hw && lane < c.Mw + G::TJ + hw;       // inside the cone: staged
    qv = a5 = a8 = gv = 0.0; hv0 = hv1 = 0.0;
    if (valid) { const int e = off[d] + lane - (c.Mw - hw); qv = q1s[e]; a5 = cof[e]; a8 = cof[F_CMAX + e]; gv = cof[2 * F_CMAX + e]; }
    if (need_hal) { hv0 = hal[2 * (i - c.i0)]; hv1 = hal[2 * (i - c.i0) + 1]; }
  };
  bool valid; double qv, a5, a8, gv, hv0, hv1;
  load_row(c.i0, valid, qv, a5, a8, gv, hv0, hv1);
  for (int i = c.i0; i <= c.c1; ++i) {
    js = 3 - js;
    bool validn = false; double qn = 0.0, a5n = 0.0, a8n = 0.0, gn = 0.0, h0n = 0.0, h1n = 0.0;
    if (i < c.c1) load_row(i + 1, validn, qn, a5n, a8n, gn, h0n, h1n);
    const double up = __shfl_up_sync(0xffffffffu, xv, 1), dn = __shfl_down_sync(0xffffffffu, xv, 1);
    double xp, xmn;                                            // x1(j+1,i-1), x1(j-1,i-1)
    if (js == 1) { xp = xv; xmn = (need_hal && m == 0) ? hv0 : up; } else { xmn = xv; xp = (need_hal && m == nyc - 1) ? hv1 : dn; }
    const double sv = -(a5 * xp + a8 * xmn);
    xv = qv + sv * gv;
    if (valid && i >= c.c0 && m >= c.mA && m <= c.mB) s_out[(i - c.c0) * G::TJ + (m - c.mA)] = sv;
    valid = validn; qv = qn; a5 = a5n; a8 = a8n; gv = gn; hv0 = h0n; hv1 = h1n;
  }
}

#ifdef MGB_F_TIMING
#define FT_DECL long long ft_t0 = clock64(), ft_acc[6] = {0, 0, 0, 0, 0, 0}; int ft_spin = 0;
#define FT(i) { const long long t1_ = clock64(); ft_acc[i] += t1_ - ft_t0; ft_t0 = t1_; }
#else
#define FT_DECL
#define FT(i)
#endif
// barrier ids: 1,2 = slot full (producers -> Thomas) ; 3,4 = solved (Thomas -> recurrence) ; 5,6 = s ready (recurrence -> producers)
template <int NZ, int NPW, int MINB>
__global__ void __launch_bounds__((NPW + 2) * 32, MINB)
k_relax_F(LevelView L, int rb, int mirror, const double* __restrict__ x1old, double* __restrict__ x1new, int ntj, int ntiles, int dbg, int thv) {
  using G = FGeom<NZ>;
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_off[40];
  constexpr int NPROD = NPW * 32, NALL = NPROD + 32;
  constexpr int sj = NZ;
  const int nx = L.nx, ny = L.ny, ny2 = ny + 2, nyc = ny >> 1;
  const int si = NZ * ny2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const s_sv = sm + G::SV;
  double* const s_cof = s_sv + 3 * G::TC;                      // [2][3][F_CMAX]
  double* const s_q1 = s_cof + 6 * F_CMAX; double* const s_hal = s_q1 + F_CMAX;
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
    FT_DECL
    for (int tile = blockIdx.x; tile < ntiles; tile += Gd, ++n) {
      double* s_rhs = sm + (n % 3) * G::SLOT;
      double2* s_bu = reinterpret_cast<double2*>(sm + 3 * G::SLOT + (n & 1) * 2 * G::SLOT);   // {pivot, off-diagonal} pairs
      const bool wb = n >= 2;
      FT(0)
      const double* s_q = sm + ((n + 1) % 3) * G::SLOT;        // slot of tile n-2
      const double* s_s = s_sv + ((n + 1) % 3) * G::TC;
      if (wb) named_bar_sync(5 + (n & 1), NALL);                // s of tile n-2 is known; every producer has left iteration n-1
      FT(1)
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
          s_rhs[so] = acc; s_bu[so] = make_double2(bt, hi ? udv : 0.0);
        }
        if (idxw >= 0) commit(idxw, k, jw, iw, s_q[c * G::PITCH + k] + s_s[c] * gw);
      }
      __threadfence_block();
      named_bar_arrive(1 + (n & 1), NALL);                      // slot full
      FT(2)
    }
#ifdef MGB_F_TIMING
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == Gd / 2 || blockIdx.x == Gd - 1))
      printf("F prod cta %d tiles %d: per tile cycles wait_sready=%lld pass=%lld other=%lld\n", (int)blockIdx.x, n, ft_acc[1] / max(n, 1), ft_acc[2] / max(n, 1), ft_acc[0] / max(n, 1));
#endif
    // drain: the last (up to two) tiles of this CTA
    for (int m2 = (n >= 2 ? n - 2 : 0); m2 < n; ++m2) {
      named_bar_sync(5 + (m2 & 1), NALL);
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
    uint4* const q1t = reinterpret_cast<uint4*>(L.q1t);
    int n = 0;
    FT_DECL
    for (int tile = blockIdx.x; tile < ntiles; tile += Gd, ++n) {
      double* s_rhs = sm + (n % 3) * G::SLOT;
      const double2* s_bu = reinterpret_cast<const double2*>(sm + 3 * G::SLOT + (n & 1) * 2 * G::SLOT);
      FT(0)
      named_bar_sync(1 + (n & 1), NALL);
      FT(1)
      const int chunk = tile / ntj, c0 = 1 + chunk * G::CH, mA = (tile - chunk * ntj) * G::TJ;
      double xq[(G::TC + 31) / 32];
#pragma unroll
      for (int u = 0; u < (G::TC + 31) / 32; ++u) {
        const int c = lane + 32 * u;
        const int r = c / G::TJ, cj = c - r * G::TJ;
        const int i = c0 + r, m = mA + cj;
        xq[u] = 0.0;
        if (c < G::TC && i <= nx && m < nyc)
          xq[u] = thv ? thomas_smem<NZ>(s_rhs + c * G::PITCH, s_bu + c * G::PITCH) : thomas_smem_plain(s_rhs + c * G::PITCH, s_bu + c * G::PITCH, NZ);
      }
      FT(3)
      __threadfence_block();
      named_bar_arrive(3 + (n & 1), 64);                        // solved (shared memory only): the write-back may use the tile
      FT(4)
#pragma unroll
      for (int u = 0; u < (G::TC + 31) / 32; ++u) {             // ... and q1 goes out behind the barrier, no fence waits for it
        const int c = lane + 32 * u;
        const int r = c / G::TJ, cj = c - r * G::TJ;
        const int i = c0 + r, m = mA + cj;
        if (c < G::TC && i <= nx && m < nyc) st_tagged(q1t + (size_t)(1 + ((i + rb) & 1) + 2 * m) + (size_t)ny2 * i, xq[u], ep);
      }
      FT(2)
    }
#ifdef MGB_F_TIMING
    if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == Gd / 2 || blockIdx.x == Gd - 1))
      printf("F thomas cta %d tiles %d: per tile cycles wait_full=%lld thomas=%lld fence+arrive=%lld stores=%lld\n", (int)blockIdx.x, n, ft_acc[1] / max(n, 1), ft_acc[3] / max(n, 1), ft_acc[4] / max(n, 1), ft_acc[2] / max(n, 1));
#endif
  } else {
    // ---- recurrence warp
    if (lane == 0) { int acc = 0; for (int d = 0; d < 40; ++d) { s_off[d] = acc; acc += G::TJ + 2 * cone_hw(d); } }
    __syncwarp();
    int n = 0;
    FT_DECL
    if (!dbg && (int)blockIdx.x < ntiles) cone_prefetch_coef(L, rb, cone_of<NZ>(L, blockIdx.x, ntj), s_off, s_cof, lane);
    for (int tile = blockIdx.x; tile < ntiles; tile += Gd, ++n) {
      FT(0)
      named_bar_sync(3 + (n & 1), 64);                          // the tile's own q1 is out (no point in polling before)
      FT(1)
      if (!dbg) {
        const Cone cn = cone_of<NZ>(L, tile, ntj);
        const double xv0 = cone_fetch_q1(L, rb, x1old, cn, ep, s_off, s_q1, s_hal, lane);
        if (tile + Gd < ntiles) cone_prefetch_coef(L, rb, cone_of<NZ>(L, tile + Gd, ntj), s_off, s_cof + ((n + 1) & 1) * 3 * F_CMAX, lane);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");      // this tile's coefficients have landed (the next tile's may still fly)
        __syncwarp();
        FT(2)
        cone_recur<NZ>(L, rb, xv0, cn, s_off, s_cof + (n & 1) * 3 * F_CMAX, s_q1, s_hal, s_sv + (n % 3) * G::TC);
        FT(3)
      } else {
        for (int c = lane; c < G::TC; c += 32) s_sv[(n % 3) * G::TC + c] = 0.0;
      }
      __syncwarp();
      __threadfence_block();
      named_bar_arrive(5 + (n & 1), NALL);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#ifdef MGB_F_TIMING
    if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == Gd / 2 || blockIdx.x == Gd - 1))
      printf("F recur cta %d tiles %d: per tile cycles wait_solved=%lld fetch=%lld recur=%lld\n", (int)blockIdx.x, n, ft_acc[1] / max(n, 1), ft_acc[2] / max(n, 1), ft_acc[3] / max(n, 1));
#endif
  }
  // the last CTA to leave advances the launch number
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); if (atomicAdd(L.tctl + 1, 1u) == (unsigned)Gd - 1u) { L.tctl[1] = 0u; __threadfence(); L.tctl[0] = ep; } }
}

template <int NZ> constexpr size_t fused_smem_bytes() { return (size_t)(FGeom<NZ>::SV + 3 * FGeom<NZ>::TC + F_STG_DOUBLES) * sizeof(double); }
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
    const size_t smem = fused_smem_bytes<NZ>();
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
  // The service group of tile T waits for the q1 of tiles up to T + d (d <= 16) of OTHER CTAs, whose own progress needs tiles
  // up to 2 * grid earlier: the wait graph is acyclic only when 2 * grid > d, hence the floor of 32 CTAs (relax_fused_supported).
  static const int user_cap = []() { const char* e = std::getenv("MGB_F_GRID_CAP"); const int v = e ? std::atoi(e) : 0; return v >= 32 ? v : 0; }();
  int cap = fused_grid_cap<NZ, NPW, MINB>();
  if (user_cap && user_cap < cap) cap = user_cap;
  if (cap < 32 || ntiles < 32) { note_launch_error(cudaErrorLaunchOutOfResources); return; }
  const int grid = ntiles < cap ? ntiles : cap;
  static const int dbg = []() { const char* e = std::getenv("MGB_F_SKIP_S"); return (e && e[0] == '1') ? 1 : 0; }();   // timing experiments only: s = 0
  launch_k(k_relax_F<NZ, NPW, MINB>, grid, (NPW + 2) * 32, fused_smem_bytes<NZ>(), st, L, rb, mirror, x1old, x1new, ntj, ntiles, dbg, thomas_variant());
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
  (void)MGB_F_CAP(32); (void)MGB_F_CAP(64); (void)MGB_F_CAP(128);
#undef MGB_F_CAP
}
bool relax_fused_supported(const LevelView& L) {
  if (!L.q1t || !L.tctl || !L.err || !L.abort_flag) return false;
  if (!(L.nz == 32 || L.nz == 64 || L.nz == 128)) return false;   // levels with fewer layers are small: cooperative / single-CTA kernels
  if (L.rb_chunk != RB_CHUNK || (L.ny & 1)) return false;
  const int TJ = 128 / L.nz, TC = RB_CHUNK * TJ;
  const int R = L.rb_W + RB_CHUNK, Mw = cone_hw(R), TP = TJ + 2 * Mw;
  if (TP > 32 || R > 32) return false;                          // the cone of a tile must fit one warp (and its 2 R halo values two loads per lane)
  { int total = 0; for (int d = 0; d < R; ++d) total += TJ + 2 * cone_hw(d); if (total > F_CMAX) return false; }   // ... and its q1 one batch of loads
  (void)TC;
  const int ntiles = ((L.ny / 2 + TJ - 1) / TJ) * ((L.nx + RB_CHUNK - 1) / RB_CHUNK);
  // big levels only (several tiles per CTA; small ones have their own kernels); never fewer than 32 tiles (see launch_F)
  static const int min_tiles = []() { const char* e = std::getenv("MGB_FUSED_MIN_TILES"); const int v = e ? std::atoi(e) : 1024; return v < 32 ? 32 : v; }();
  return ntiles >= min_tiles;
}
size_t relax_fused_num_tiles(int nx, int ny, int nz) {
  if (!(nz == 32 || nz == 64 || nz == 128)) return 0;
  const int TJ = 128 / nz;
  return (size_t)((ny / 2 + TJ - 1) / TJ) * ((nx + RB_CHUNK - 1) / RB_CHUNK);
}
void launch_relax_fused(cudaStream_t st, const LevelView& L, int rb, int mirror, const double* x1old, double* x1new) {
  switch (L.nz) {
    case 128: launch_F_nz<128>(st, L, rb, mirror, x1old, x1new); break;
    case 64: launch_F_nz<64>(st, L, rb, mirror, x1old, x1new); break;
    default: launch_F_nz<32>(st, L, rb, mirror, x1old, x1new); break;
  }
}

}  // namespace mgb

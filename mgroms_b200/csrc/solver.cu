// solver.cu -- state, level hierarchy on the device, multigrid cycles, inter-GPU exchange and the C ABI
// (include/mgb.h) of libmgb.  One solver instance per process, one GPU per process.
#include "../../include/mgb.h"
#include "kernels.cuh"
#include "mgb_internal.h"

#include <cuda_runtime.h>
#include <nccl.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace mgb {

// ------------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------------
static std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define RC(call) do { int r_ = (call); if (r_ != 0) return r_; } while (0)

// ------------------------------------------------------------------------------------------------------
// NCCL, loaded lazily so that single-GPU use has no dependency on it
// ------------------------------------------------------------------------------------------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
  if (!g_nccl.h) return fail(std::string("cannot load libnccl: ") + dlerror());
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name); if (!g_nccl.field) return fail(std::string("libnccl lacks ") + name)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllReduce, "ncclAllReduce"); SYM(AllGather, "ncclAllGather");
  SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return 0;
}
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(std::string(#call) + ": " + g_nccl.GetErrorString(r_)); } while (0)

// ------------------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------------------
struct Level {
  LevelDims d;
  LevelView v{};
  double *dx = nullptr, *dy = nullptr, *zeta = nullptr, *h = nullptr, *zr = nullptr, *zw = nullptr, *cw = nullptr;
  double *dzw = nullptr, *zxdy = nullptr, *zydx = nullptr;   // level 1 only
  double *dummy3 = nullptr, *gatherbuf = nullptr;             // gathered levels only
  double *sub2d[4] = {nullptr, nullptr, nullptr, nullptr};    // coarse geometry before gathering
  double *tab = nullptr;
  double** planes_dev = nullptr;
  double *packS = nullptr, *packN = nullptr, *unpS = nullptr, *unpN = nullptr;   // S/N face staging (multi-GPU)
  double *pack8[8] = {nullptr}, *unp8[8] = {nullptr};         // nh=2 staging (setup)
  double *partials = nullptr; int npartials = 0;
  Halo1Flags hf{};
  double* rmaskl = nullptr;                                   // coarse-level rmask (ones; physical halo zeroed when bmask)
  unsigned* gridbar = nullptr;                                // barrier + exit counters of the cooperative relax kernel
  P2PBox box{};                                               // peer-memory mailbox of this level (multi-GPU, when S.p2p)
  bool bottom_diag = true;
  double rho = 0.0;
  size_t n3() const { return (size_t)d.nz * (d.ny + 2) * (d.nx + 2); }
  size_t n2() const { return (size_t)(d.ny + 2) * (d.nx + 2); }
};

struct Solver {
  bool inited = false, have_matrices = false;
  Params P;
  int rank = 0, npx = 1, npy = 1, nranks = 1, device = 0;
  std::vector<Level> L;              // L[lev-1]
  cudaStream_t st = nullptr, st_copy = nullptr;      // solver stream; PCIe copies of nhydro_solve
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_c[4] = {nullptr, nullptr, nullptr, nullptr};
  ncclComm_t comm = nullptr;
  // halo exchange over NVLink peer memory (CUDA IPC): one mailbox allocation per rank, mapped by its neighbours
  bool p2p = false;
  void* mbox = nullptr;                       // my mailbox (all levels)
  std::vector<void*> peer_mbox;               // per rank: its mailbox mapped into this process (nullptr: self / not a neighbour)
  unsigned long long* p2p_seq = nullptr; unsigned int* p2p_done = nullptr;   // per-level counters (local)
  int* err_host = nullptr; int* err_dev = nullptr;                             // mapped host word raised by a timed-out device-side wait (1: halo exchange, 2: fused relax)
  int* abort_dev = nullptr;                                                    // device word raised with it: later waits give up at once
  unsigned spin_timeout_ms = 5000;                                             // MGB_SPIN_TIMEOUT_MS
  double hlim = 0, theta_b = 0, theta_s = 0;
  // level-1 model arrays (k fastest) and staging for the (i,j,k) host layout
  double *ut = nullptr, *vt = nullptr, *wt = nullptr, *stage = nullptr, *stage_v = nullptr, *stage_w = nullptr, *flux = nullptr, *rmask = nullptr;
  double *sums = nullptr;            // device scalars: [0] residual sum, [1] bnorm sum, [2] all-reduced
  int* flag_dev = nullptr;
  std::vector<void*> allocs;
  std::vector<double> hist_res, hist_conv;
  double last_ms = 0; int last_iters = 0;
  bool fused_prolong = true;
  bool use_small = true;             // fused single-CTA relaxation on small neighbour-free levels
  bool use_mid = true;               // fused cooperative relaxation (grid barriers) on mid-size neighbour-free levels
  bool use_graph = true;             // replay one captured F-cycle (+ residual) per outer iteration
  bool fuse_halo = true;             // neighbour-free levels: relax kernels write the Neumann halo themselves
  bool fuse_exchange = true;         // small / mid levels with neighbour ranks: the relax kernel runs the peer-memory exchange itself
  bool use_fused = false;            // big levels: one persistent launch per colour (A + B + C fused)
  bool fuse_bc = true;               // neighbour-free levels: kernels B and C fused
  size_t fuse_bc_limit = 3000000;    // ... up to this many points (larger levels stream better with the separate C)
  cudaGraphExec_t cycle_graph = nullptr;
  long long cycle_graph_nodes = 0;
  int nlevs() const { return (int)L.size(); }
};
static Solver S;

// ------------------------------------------------------------------------------------------------------
// mg_tictoc (src/mg_tictoc.f90): the reference's eight tic/toc sites, timed with CUDA events on the solver stream.
// Off by default (the production solve replays a CUDA graph, which has no per-operator boundaries); mgb_tictoc_enable(1)
// or MGB_TICTOC=1 switches to plain launches and records one event pair per call.
// ------------------------------------------------------------------------------------------------------
struct TicToc {
  bool on = false;
  struct Rec { int id, lev; cudaEvent_t a, b; };
  std::vector<Rec> open;                                      // recorded, not yet accumulated
  std::vector<std::string> names;                             // in order of first tic, like the reference's subname(:)
  std::vector<std::vector<double>> time;                      // [id][lev-1] seconds
  std::vector<std::vector<long long>> calls;
};
static TicToc T;
static int tt_tic(const char* name, int lev) {
  if (!T.on) return -1;
  int id = -1;
  for (size_t q = 0; q < T.names.size(); ++q) if (T.names[q] == name) id = (int)q;
  if (id < 0) { id = (int)T.names.size(); T.names.push_back(name); T.time.emplace_back(S.nlevs(), 0.0); T.calls.emplace_back(S.nlevs(), 0); }
  TicToc::Rec r{id, lev, nullptr, nullptr};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
  cudaEventRecord(r.a, S.st);
  T.open.push_back(r);
  return (int)T.open.size() - 1;
}
static void tt_toc(int h) { if (h >= 0 && h < (int)T.open.size()) cudaEventRecord(T.open[h].b, S.st); }
static void tt_flush() {
  if (T.open.empty()) return;
  cudaStreamSynchronize(S.st);
  for (auto& r : T.open) {
    float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
    if (r.lev >= 1 && r.lev <= (int)T.time[r.id].size()) { T.time[r.id][r.lev - 1] += 1e-3 * ms; T.calls[r.id][r.lev - 1] += 1; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  T.open.clear();
}
static void tt_reset() { tt_flush(); T.names.clear(); T.time.clear(); T.calls.clear(); }
struct TicGuard { int h; TicGuard(const char* n, int lev) : h(tt_tic(n, lev)) {} ~TicGuard() { tt_toc(h); } };
static Params g_params;              // defaults + mgb_set_param overrides (used by mgb_plan_levels)
static std::vector<std::pair<std::string, std::string>> g_overrides;   // applied after the namelist file in mgb_init
static std::vector<LevelDims> g_plan;

static int dalloc(double** p, size_t n) {
  void* q = nullptr;
  if (n == 0) n = 1;
  CU(cudaMalloc(&q, n * sizeof(double)));
  CU(cudaMemsetAsync(q, 0, n * sizeof(double), S.st));
  S.allocs.push_back(q);
  *p = (double*)q;
  return 0;
}

static Halo1Flags flags_of(const LevelDims& d) {
  Halo1Flags f;
  f.south = d.neighb[0] >= 0; f.east = d.neighb[1] >= 0; f.north = d.neighb[2] >= 0; f.west = d.neighb[3] >= 0;
  f.sw = d.neighb[4] >= 0; f.se = d.neighb[5] >= 0; f.ne = d.neighb[6] >= 0; f.nw = d.neighb[7] >= 0;
  return f;
}

// end of an API call: wait for the stream and surface both execution errors and launch-configuration errors
static int finish() {
  cudaError_t e = cudaStreamSynchronize(S.st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { take_launch_error(); return fail(std::string("CUDA failure: ") + cudaGetErrorString(e)); }
  if (e == cudaSuccess) e = take_launch_error();
  if (e != cudaSuccess) return fail(std::string("CUDA launch failure: ") + cudaGetErrorString(e));
  if (S.err_host && *S.err_host) {
    const int code = *S.err_host; *S.err_host = 0;
    if (S.abort_dev) cudaMemsetAsync(S.abort_dev, 0, sizeof(int), S.st);
    return fail(code == 1 ? "halo exchange timed out waiting for a neighbour GPU (peer-memory mailbox flag never raised; MGB_SPIN_TIMEOUT_MS)"
                          : "fused relaxation kernel timed out waiting for a tile flag (MGB_SPIN_TIMEOUT_MS)");
  }
  return 0;
}

// Ranks may enter an API call seconds apart (per-rank I/O of the host model).  The peer-memory exchange spins with a
// timeout, MPI_Irecv/Waitany of the reference simply waits: absorb the skew with a blocking 1-element all-reduce on the
// stream before the first mailbox exchange of the call.
static int rank_sync() {
  if (!S.p2p || !S.comm) return 0;
  NC(g_nccl.AllReduce(S.sums + 4, S.sums + 5, 1, ncclDouble, ncclSum, S.comm, S.st));
  return 0;
}

static int check_level(int lev) {
  if (!S.inited) return fail("mgb_init has not been called");
  if (lev < 1 || lev > S.nlevs()) return fail("level out of range");
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// halo fill  (fill_halo, src/mg_mpi_exchange.f90).  nh = 1 or 2; nzl = levels of the array.
// ------------------------------------------------------------------------------------------------------
static int exchange(Level& g, double* a, int nzl, int nh) {
  // inter-GPU part: what Isend/Irecv + unpack do.  Order of directions = S,E,N,W,SW,SE,NE,NW
  const int nx = g.d.nx, ny = g.d.ny;
  const int nyh = ny + 2 * nh;
  auto at = [&](int j, int i) { return a + (size_t)nzl * ((size_t)(j - 1 + nh) + (size_t)nyh * (size_t)(i - 1 + nh)); };
  // send patches (Fortran j0,i0,rows,cols) and recv patches per direction
  struct Patch { int j0, nr, i0, nc; };
  const Patch sendp[8] = {{1, nh, 1, nx}, {1, ny, nx - nh + 1, nh}, {ny - nh + 1, nh, 1, nx}, {1, ny, 1, nh},
                          {1, nh, 1, nh}, {1, nh, nx - nh + 1, nh}, {ny - nh + 1, nh, nx - nh + 1, nh}, {ny - nh + 1, nh, 1, nh}};
  const Patch recvp[8] = {{1 - nh, nh, 1, nx}, {1, ny, nx + 1, nh}, {ny + 1, nh, 1, nx}, {1, ny, 1 - nh, nh},
                          {1 - nh, nh, 1 - nh, nh}, {1 - nh, nh, nx + 1, nh}, {ny + 1, nh, nx + 1, nh}, {ny + 1, nh, 1 - nh, nh}};
  // a patch is contiguous in memory when it spans whole columns' j-range of width 1 ... simple rule: nc == 1 && rows contiguous
  auto contiguous = [&](const Patch& p) { return p.nc == 1 || (p.j0 == 1 - nh && p.nr == nyh); };
  for (int q = 0; q < 8; ++q) {
    if (g.d.neighb[q] < 0) continue;
    if (!contiguous(sendp[q])) launch_pack_rows(S.st, a, g.pack8[q], nzl, ny, nx, nh, sendp[q].j0, sendp[q].nr, sendp[q].i0, sendp[q].nc);
  }
  NC(g_nccl.GroupStart());
  for (int q = 0; q < 8; ++q) {
    const int peer = g.d.neighb[q];
    if (peer < 0) continue;
    const size_t cnt = (size_t)nzl * sendp[q].nr * sendp[q].nc;
    const double* sb = contiguous(sendp[q]) ? at(sendp[q].j0, sendp[q].i0) : g.pack8[q];
    double* rb = contiguous(recvp[q]) ? at(recvp[q].j0, recvp[q].i0) : g.unp8[q];
    NC(g_nccl.Send(sb, cnt, ncclDouble, peer, S.comm, S.st));
    NC(g_nccl.Recv(rb, cnt, ncclDouble, peer, S.comm, S.st));
  }
  NC(g_nccl.GroupEnd());
  for (int q = 0; q < 8; ++q) {
    if (g.d.neighb[q] < 0) continue;
    if (!contiguous(recvp[q])) launch_unpack_rows(S.st, a, g.unp8[q], nzl, ny, nx, nh, recvp[q].j0, recvp[q].nr, recvp[q].i0, recvp[q].nc);
  }
  return 0;
}

// nh = 1 with neighbours: pre kernel, one NCCL group, post kernel
static int fill_halo1_multi(Level& g, double* a, int nzl, double* x1) {
  const int nx = g.d.nx, ny = g.d.ny;
  if (S.p2p) { launch_halo1_p2p(S.st, a, nzl, ny, nx, g.hf, g.box, x1); return 0; }
  auto at = [&](int j, int i) { return a + (size_t)nzl * ((size_t)j + (size_t)(ny + 2) * (size_t)i); };
  launch_halo1_pre(S.st, a, nzl, ny, nx, g.hf, g.pack8[0], g.pack8[2]);
  NC(g_nccl.GroupStart());
  const int* nb = g.d.neighb;
  const size_t row = (size_t)nzl * nx, col = (size_t)nzl * ny, cor = (size_t)nzl;
  if (nb[0] >= 0) { NC(g_nccl.Send(g.pack8[0], row, ncclDouble, nb[0], S.comm, S.st)); NC(g_nccl.Recv(g.unp8[0], row, ncclDouble, nb[0], S.comm, S.st)); }
  if (nb[1] >= 0) { NC(g_nccl.Send(at(1, nx), col, ncclDouble, nb[1], S.comm, S.st)); NC(g_nccl.Recv(at(1, nx + 1), col, ncclDouble, nb[1], S.comm, S.st)); }
  if (nb[2] >= 0) { NC(g_nccl.Send(g.pack8[2], row, ncclDouble, nb[2], S.comm, S.st)); NC(g_nccl.Recv(g.unp8[2], row, ncclDouble, nb[2], S.comm, S.st)); }
  if (nb[3] >= 0) { NC(g_nccl.Send(at(1, 1), col, ncclDouble, nb[3], S.comm, S.st)); NC(g_nccl.Recv(at(1, 0), col, ncclDouble, nb[3], S.comm, S.st)); }
  if (nb[4] >= 0) { NC(g_nccl.Send(at(1, 1), cor, ncclDouble, nb[4], S.comm, S.st)); NC(g_nccl.Recv(at(0, 0), cor, ncclDouble, nb[4], S.comm, S.st)); }
  if (nb[5] >= 0) { NC(g_nccl.Send(at(1, nx), cor, ncclDouble, nb[5], S.comm, S.st)); NC(g_nccl.Recv(at(0, nx + 1), cor, ncclDouble, nb[5], S.comm, S.st)); }
  if (nb[6] >= 0) { NC(g_nccl.Send(at(ny, nx), cor, ncclDouble, nb[6], S.comm, S.st)); NC(g_nccl.Recv(at(ny + 1, nx + 1), cor, ncclDouble, nb[6], S.comm, S.st)); }
  if (nb[7] >= 0) { NC(g_nccl.Send(at(ny, 1), cor, ncclDouble, nb[7], S.comm, S.st)); NC(g_nccl.Recv(at(ny + 1, 0), cor, ncclDouble, nb[7], S.comm, S.st)); }
  NC(g_nccl.GroupEnd());
  launch_halo1_post(S.st, a, nzl, ny, nx, g.hf, g.unp8[0], g.unp8[2], x1);
  return 0;
}

static int fill_halo(int lev, double* a, int nzl, int nh, double* x1 = nullptr) {
  Level& g = S.L[lev - 1];
  if (nh == 1 && S.comm) {
    bool any1 = false; for (int q = 0; q < 8; ++q) any1 |= g.d.neighb[q] >= 0;
    if (any1) return fill_halo1_multi(g, a, nzl, x1);
  }
  if (nh == 1) launch_halo1_phys(S.st, a, nzl, g.d.ny, g.d.nx, g.hf); else launch_halo2_phys(S.st, a, nzl, g.d.ny, g.d.nx, g.hf);
  if (nh == 1 && x1) launch_halo1_phys(S.st, x1, 1, g.d.ny, g.d.nx, g.hf);   // the k=1 plane mirrors like the field itself (its interior is kept by kernel C)
  bool any = false; for (int q = 0; q < 8; ++q) any |= g.d.neighb[q] >= 0;
  if (any) {
    if (!S.comm) return fail("multi-rank level but mgb_comm_init has not been called");
    RC(exchange(g, a, nzl, nh));
    if (nh == 1) launch_halo1_flagcorners(S.st, a, nzl, g.d.ny, g.d.nx, g.hf); else launch_halo2_flagcorners(S.st, a, nzl, g.d.ny, g.d.nx, g.hf);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// gather / split  (src/mg_gather.f90)
// ------------------------------------------------------------------------------------------------------
// x: my sub-block (nzl, nys+2, nxs+2) ; y: gathered array (nzl, ny+2, nx+2) of level lev
static int gather(int lev, const double* x, double* y, int nzl) {
  Level& g = S.L[lev - 1];
  const int nxs = g.d.nx / g.d.ngx, nys = g.d.ny / g.d.ngy;
  const size_t Ng = (size_t)nzl * (nys + 2) * (nxs + 2);
  if (!S.comm) return fail("gather needs mgb_comm_init");
  NC(g_nccl.GroupStart());
  for (int q = 0; q < g.d.nmembers; ++q) {
    const int peer = g.d.members[q];
    if (peer == S.rank) continue;
    NC(g_nccl.Send(x, Ng, ncclDouble, peer, S.comm, S.st));
    NC(g_nccl.Recv(g.gatherbuf + (size_t)q * (size_t)g.d.nz * (nys + 2) * (nxs + 2), Ng, ncclDouble, peer, S.comm, S.st));
  }
  NC(g_nccl.GroupEnd());
  for (int q = 0; q < g.d.nmembers; ++q) {
    const int l = q % g.d.ngx, m = q / g.d.ngx;
    const double* src = (g.d.members[q] == S.rank) ? x : g.gatherbuf + (size_t)q * (size_t)g.d.nz * (nys + 2) * (nxs + 2);
    launch_gather_place(S.st, src, nzl, nys, nxs, y, g.d.ny, g.d.nx, l, m);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// operators
// ------------------------------------------------------------------------------------------------------
// which kernels relax(lev, .) runs for the red-black smoother: "small" (one CTA), "mid" (cooperative), "fused" (one persistent
// launch per colour), "a+bc" / "a+b+c" (separate launches), "a" (no k=1 diagonal coupling), with "+halo" when the halo is a separate launch
static std::string relax_path(int lev) {
  Level& g = S.L[lev - 1];
  const std::string& m = S.P.relax_method;
  if (!(m == "Red-Black" || m == "RB")) return m;
  bool lonely = true; for (int q = 0; q < 8; ++q) lonely &= g.d.neighb[q] < 0;
  const bool rank1 = (S.P.cmatrix == "real") && g.bottom_diag;
  const bool inkernel_ok = lonely || (S.p2p && S.fuse_exchange);
  if (S.use_small && inkernel_ok && relax_small_supported(g.v)) return lonely ? "small" : "small+p2p";
  if (S.use_mid && inkernel_ok && relax_mid_supported(g.v)) return lonely ? "mid" : "mid+p2p";
  const int fuse = (lonely && S.fuse_halo) ? 1 : 0;
  std::string r;
  if (!rank1) r = "a";
  else if (S.use_fused && relax_fused_supported(g.v)) r = "fused";
  else if (fuse && S.fuse_bc && g.v.rb_chunk < g.d.nx && g.n3() <= S.fuse_bc_limit) r = "a+bc";
  else r = "a+b+c";
  r = fuse ? r : r + (lonely ? "+halo" : "+exchange");
  if (rank1) r += " W=" + std::to_string(g.v.rb_W);            // warm-up rows of the chunked bottom recurrence
  return r;
}

static int relax(int lev, int nsweeps) {
  Level& g = S.L[lev - 1];
  const std::string& m = S.P.relax_method;
  TicGuard tg(m == "Red-Black" || m == "RB" ? "relax_3D_8_RB" : (m == "Four-Color" || m == "FC" ? "relax_3D_8_FC" : "relax_3D_8_GS"), lev);   // src/mg_relax.f90:128,167,209
  const bool real_matrix = (S.P.cmatrix == "real");
  if (m == "Red-Black" || m == "RB") {
    bool lonely = true; for (int q = 0; q < 8; ++q) lonely &= g.d.neighb[q] < 0;
    // small and mid-size levels: the whole relax call in one launch; with neighbour ranks the kernel runs the peer-memory halo
    // exchange itself after every colour (the NCCL transport keeps the separate launches)
    const bool inkernel_ok = lonely || (S.p2p && S.fuse_exchange);
    if (S.use_small && inkernel_ok && relax_small_supported(g.v)) {
      launch_relax_small(S.st, g.v, nsweeps, (real_matrix && g.bottom_diag) ? 1 : 0, lonely ? 0 : 1, g.hf, g.box);
      return 0;
    }
    if (S.use_mid && inkernel_ok && relax_mid_supported(g.v)) {
      launch_relax_mid(S.st, g.v, nsweeps, (real_matrix && g.bottom_diag) ? 1 : 0, g.gridbar, lonely ? 0 : 1, g.hf, g.box);
      return 0;
    }
    const bool rank1 = real_matrix && g.bottom_diag;
    const int fuse = (lonely && S.fuse_halo) ? 1 : 0;
    // B+C fused: chunked recurrence only, and only where the level is latency-bound (the separate C streams better on big levels)
    const bool fuse_bc = fuse && rank1 && S.fuse_bc && g.v.rb_chunk < g.d.nx && g.n3() <= S.fuse_bc_limit;
    const int xmode = fuse ? 1 : 2;                           // 2: neighbours -> C keeps x1's interior, the exchange its halo ring
    if (rank1) launch_snapshot_x1(S.st, g.v);                 // x1 = p(1,:,:) once per call; kept current afterwards
    LevelView v = g.v;
    if (rank1 && S.use_fused && relax_fused_supported(g.v)) {  // A + B + C of a colour in one persistent launch (kernels_relax.cu)
      for (int it = 0; it < nsweeps; ++it)
        for (int rb = 1; rb <= 2; ++rb) {
          launch_relax_fused(S.st, v, rb, fuse, v.x1, v.x1b);
          { double* t = v.x1; v.x1 = v.x1b; v.x1b = t; }
          if (!fuse) RC(fill_halo(lev, v.p, g.d.nz, 1, v.x1));
        }
      return 0;
    }
    for (int it = 0; it < nsweeps; ++it)
      for (int rb = 1; rb <= 2; ++rb) {
        ColourSet cs{1, 1, 0, rb, 0, rank1 ? 1 : 0, rank1 ? xmode : fuse};
        launch_relax_A(S.st, v, cs);
        if (rank1) {
          if (fuse_bc) { launch_relax_BC(S.st, v, cs, v.x1, v.x1b); double* t = v.x1; v.x1 = v.x1b; v.x1b = t; }
          else { launch_relax_B(S.st, v, cs); launch_relax_C(S.st, v, cs); }
        }
        if (!fuse) RC(fill_halo(lev, v.p, g.d.nz, 1, rank1 ? v.x1 : nullptr));
      }
  } else if (m == "Four-Color" || m == "FC") {
    for (int it = 0; it < nsweeps; ++it)
      for (int fc1 = 1; fc1 <= 2; ++fc1) for (int fc2 = 1; fc2 <= 2; ++fc2) {
        bool alone = true; for (int q = 0; q < 8; ++q) alone &= g.d.neighb[q] < 0;
        const int fuse = (alone && S.fuse_halo) ? 1 : 0;
        ColourSet cs{1 + (fc1 - 1) % 2, 2, 1, 0, 1 + (fc2 - 1) % 2, real_matrix ? 2 : 0, fuse};
        if (cs.istart <= g.d.nx && cs.jfix <= g.d.ny) launch_relax_A(S.st, g.v, cs);
        if (!fuse) RC(fill_halo(lev, g.v.p, g.d.nz, 1));
      }
  } else if (m == "Gauss-Seidel" || m == "GS") {
    // relax_3D_8_GS (src/mg_relax.f90:116-148): lexicographic sweep, i outer, j inner, halo once per sweep.  Column (j,i)
    // needs the NEW values of (j-1,i) and of row i-1 and the OLD values of (j+1,i) and of row i+1: all columns of one
    // wavefront w = j + 2 i are independent, so a sweep is ny + 2 nx - 2 launches of the column kernel (exact, not fast).
    bool alone = true; for (int q = 0; q < 8; ++q) alone &= g.d.neighb[q] < 0;
    for (int it = 0; it < nsweeps; ++it) {
      for (int w = 3; w <= g.d.ny + 2 * g.d.nx; ++w) {
        ColourSet cs{1, 1, 2, 0, w, real_matrix ? 2 : 0, 0};
        const int ilo = std::max(1, (w - g.d.ny + 1) / 2), ihi = std::min(g.d.nx, (w - 1) / 2);
        if (ilo > ihi) continue;
        cs.istart = ilo;
        LevelView v = g.v; v.nx = ihi;                       // rows ilo..ihi only (nx is the upper bound of the row loop)
        launch_relax_A(S.st, v, cs);
      }
      RC(fill_halo(lev, g.v.p, g.d.nz, 1));
      (void)alone;
    }
  } else {
    return fail("relax_method '" + m + "' is unknown (RB, FC, GS)");
  }
  return 0;
}

// r = b - A p, halo(r); the sum of squares of this rank is left in S.sums[0]
static int residual(int lev) {
  Level& g = S.L[lev - 1];
  { TicGuard tg("residual_3D_8", lev);                        // src/mg_relax.f90:367-370
    launch_residual(S.st, g.v, S.P.cmatrix == "real", g.partials, S.sums); }
  RC(fill_halo(lev, g.v.r, g.d.nz, 1));
  return 0;
}
// global_sum (src/mg_mpi_exchange.f90:1555-1571): all-reduce + rescale for gathered levels; returns host value
static int global_sum(int lev, const double* dev_local, double* out) {
  double v = 0.0;
  if (S.nranks > 1) {
    if (!S.comm) return fail("global_sum needs mgb_comm_init");
    NC(g_nccl.AllReduce(dev_local, S.sums + 2, 1, ncclDouble, ncclSum, S.comm, S.st));
    CU(cudaMemcpyAsync(&v, S.sums + 2, sizeof(double), cudaMemcpyDeviceToHost, S.st));
  } else {
    CU(cudaMemcpyAsync(&v, dev_local, sizeof(double), cudaMemcpyDeviceToHost, S.st));
  }
  RC(finish());
  const LevelDims& d = S.L[lev - 1].d; const LevelDims& d1 = S.L[0].d;
  *out = v * (d.npx * d.npy) / (d1.npx * d1.npy);
  return 0;
}

static int fine2coarse(int lev, bool copy_b_to_r) {
  Level& f = S.L[lev - 1]; Level& c = S.L[lev];
  const int nxs = c.d.nx / c.d.ngx, nys = c.d.ny / c.d.ngy;
  bool lonely = true; for (int q = 0; q < 8; ++q) lonely &= c.d.neighb[q] < 0;
  if (lonely && !c.d.gather && S.fuse_halo) {                  // one launch: restriction + halo(b_c) + p_c = 0 (+ r_c = b_c)
    launch_restrict8_fused(S.st, f.v.r, f.d.nz, f.d.ny, f.d.nx, c.v.b, c.v.p, copy_b_to_r ? c.v.r : nullptr, c.d.nz, c.d.ny, c.d.nx);
    return 0;
  }
  double* target = c.d.gather ? c.dummy3 : c.v.b;
  launch_restrict8(S.st, f.v.r, f.d.nz, f.d.ny, f.d.nx, target, c.d.nz, nys, nxs);
  if (c.d.gather) RC(gather(lev + 1, c.dummy3, c.v.b, c.d.nz));
  RC(fill_halo(lev + 1, c.v.b, c.d.nz, 1));
  launch_fill(S.st, c.v.p, 0.0, c.n3());
  if (copy_b_to_r) launch_copy(S.st, c.v.r, c.v.b, c.n3());      // Fcycle: grid(lev+1)%r = grid(lev+1)%b
  return 0;
}

static int coarse2fine(int lev, bool fused) {
  Level& f = S.L[lev - 1]; Level& c = S.L[lev];
  const double* pc = c.v.p; int nxc = c.d.nx, nyc = c.d.ny;
  if (c.d.gather) {
    nxc = c.d.nx / c.d.ngx; nyc = c.d.ny / c.d.ngy;
    launch_split(S.st, c.v.p, c.d.nz, c.d.ny, c.d.nx, c.dummy3, nyc, nxc, c.d.key % 2, c.d.key / 2);
    pc = c.dummy3;
  }
  const int linear = (S.P.interp_type == "linear");
  if (fused) {   // p_f += P p_c on the interior, then refresh halo(p_f): identical to the reference whenever halo(p_f) was consistent
    bool lonely = true; for (int q = 0; q < 8; ++q) lonely &= f.d.neighb[q] < 0;
    if (lonely && linear && S.fuse_halo && f.n3() <= (size_t)3000000) { launch_prolong(S.st, pc, c.d.nz, nyc, nxc, f.v.r, f.v.p, linear, 2); return 0; }   // latency-bound levels: halo written by the kernel (on big levels the mirror code costs more than the launch)
    if (linear && prolong_bulk_supported(c.d.nz, nyc, nxc)) launch_prolong_bulk(S.st, pc, c.d.nz, nyc, nxc, f.v.p);
    else launch_prolong(S.st, pc, c.d.nz, nyc, nxc, f.v.r, f.v.p, linear, 1);
    RC(fill_halo(lev, f.v.p, f.d.nz, 1));
  } else {       // the reference's sequence: r_f = P p_c ; halo(r_f) ; p_f += r_f everywhere
    launch_prolong(S.st, pc, c.d.nz, nyc, nxc, f.v.r, f.v.p, linear, 0);
    RC(fill_halo(lev, f.v.r, f.d.nz, 1));
    launch_axpy_all(S.st, f.v.p, f.v.r, f.n3());
  }
  return 0;
}

static int vcycle(int lev1) {
  const int Ln = S.nlevs();
  for (int lev = lev1; lev <= Ln - 1; ++lev) { RC(relax(lev, S.P.ns_pre)); RC(residual(lev)); RC(fine2coarse(lev, false)); }
  RC(relax(Ln, S.P.ns_coarsest));
  for (int lev = Ln - 1; lev >= lev1; --lev) { RC(coarse2fine(lev, S.fused_prolong)); RC(relax(lev, S.P.ns_post)); }
  return 0;
}
static int fcycle() {
  TicGuard tg("Fcycle", 1);                                   // src/mg_solvers.f90:108,124
  const int Ln = S.nlevs();
  for (int lev = 1; lev <= Ln - 1; ++lev) RC(fine2coarse(lev, true));
  RC(relax(Ln, S.P.ns_coarsest));
  for (int lev = Ln - 1; lev >= 1; --lev) { RC(coarse2fine(lev, S.fused_prolong)); RC(vcycle(lev)); }
  return 0;
}

// One outer iteration (Fcycle + compute_residual(1)) captured into a CUDA graph: the launch sequence is fixed once
// the matrices are defined, so each iteration becomes one graph launch instead of thousands of API calls.
static int build_cycle_graph() {
  cudaGraph_t graph = nullptr;
  const long long before = g_launches;
  CU(cudaStreamBeginCapture(S.st, cudaStreamCaptureModeThreadLocal));
  int rc = fcycle();
  if (rc == 0) rc = residual(1);
  cudaError_t e = cudaStreamEndCapture(S.st, &graph);
  S.cycle_graph_nodes = g_launches - before;
  g_launches = before;
  if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (e != cudaSuccess) return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
  e = cudaGraphInstantiate(&S.cycle_graph, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
  return 0;
}
static void drop_cycle_graph() { if (S.cycle_graph) { cudaGraphExecDestroy(S.cycle_graph); S.cycle_graph = nullptr; } }

static int solve_p() {
  if (!S.have_matrices) return fail("mgb_matrices has not been called");
  Level& g = S.L[0];
  const double tol = S.P.solver_prec; const int maxite = S.P.solver_maxiter;
  RC(rank_sync());
  launch_fill(S.st, g.v.p, 0.0, g.n3());
  CU(cudaEventRecord(S.ev0, S.st));
  TicGuard tg_solve("solve", 1);                              // src/mg_solvers.f90:45,82
  int nite = 0;
  launch_sumsq(S.st, g.v, g.v.b, g.partials, S.sums + 1);
  double bsum = 0, rsum = 0;
  RC(global_sum(1, S.sums + 1, &bsum));
  const double bnorm = std::sqrt(bsum);
  RC(residual(1)); RC(global_sum(1, S.sums, &rsum));
  double rnorm = std::sqrt(rsum);
  double res0 = rnorm / bnorm;
  S.hist_res.assign(1, res0); S.hist_conv.assign(1, 0.0);
  while (nite < maxite && res0 > tol) {
    if (S.use_graph && !T.on) {
      if (!S.cycle_graph) RC(build_cycle_graph());
      CU(cudaGraphLaunch(S.cycle_graph, S.st));
      g_launches += S.cycle_graph_nodes;
    } else {
      RC(fcycle());
      RC(residual(1));
    }
    RC(global_sum(1, S.sums, &rsum));
    rnorm = std::sqrt(rsum) / bnorm;
    const double conv = res0 / rnorm;
    res0 = rnorm;
    ++nite;
    S.hist_res.push_back(rnorm); S.hist_conv.push_back(conv);
  }
  CU(cudaEventRecord(S.ev1, S.st));
  CU(cudaEventSynchronize(S.ev1));
  float ms = 0; CU(cudaEventElapsedTime(&ms, S.ev0, S.ev1));
  S.last_ms = ms; S.last_iters = nite;
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// matrices
// ------------------------------------------------------------------------------------------------------
static void vertical_tables(int nz, double hlim, double theta_b, double theta_s, std::vector<double>& tab) {
  // k-only part of src/mg_zr_zw.f90:112-166 (same expressions, host libm): cs_r, cff_r (k=1..nz), cs_w, cff_w (k=1..nz+1)
  const int nzp = nz + 1; tab.assign((size_t)4 * nzp, 0.0);
  const double cff = 1.0 / (double)nz;
  for (int k = 1; k <= nzp; ++k) {
    const double sc_w = cff * (double)(k - 1 - nz);
    const double cswf = (theta_s > 0.0) ? (1.0 - std::cosh(theta_s * sc_w)) / (std::cosh(theta_s) - 1.0) : -(sc_w * sc_w);
    const double cs_w = (theta_b > 0.0) ? (std::exp(theta_b * cswf) - 1.0) / (1.0 - std::exp(-theta_b)) : cswf;
    tab[2 * nzp + k - 1] = cs_w; tab[3 * nzp + k - 1] = hlim * sc_w;
    if (k <= nz) {
      const double sc_r = cff * ((double)(k - nz) - 0.5);
      const double csrf = (theta_s > 0.0) ? (1.0 - std::cosh(theta_s * sc_r)) / (std::cosh(theta_s) - 1.0) : -(sc_r * sc_r);
      const double cs_r = (theta_b > 0.0) ? (std::exp(theta_b * csrf) - 1.0) / (1.0 - std::exp(-theta_b)) : csrf;
      tab[k - 1] = cs_r; tab[nzp + k - 1] = hlim * sc_r;
    }
  }
}

static MatrixView matrix_view(Level& g, int lev) {
  MatrixView M{};
  M.nx = g.d.nx; M.ny = g.d.ny; M.nz = g.d.nz; M.lev1 = (lev == 1);
  M.dx = g.dx; M.dy = g.dy; M.zr = g.zr; M.zw = g.zw; M.cw = g.cw; M.dzw = g.dzw; M.zxdy = g.zxdy; M.zydx = g.zydx;
  for (int q = 0; q < 8; ++q) M.cA[q] = g.v.cA[q];
  M.bet = g.v.bet; M.g = g.v.g; M.c5b = g.v.c5b; M.c8b = g.v.c8b; M.g1 = g.v.g1;
  M.rmask = (lev == 1) ? S.rmask : g.rmaskl; M.bmask = S.P.bmask ? 1 : 0;    // level 1: the caller's mask (src/mg_define_matrix.f90:73-79)
  return M;
}

static void drop_cycle_graph();
static int define_matrices() {
  drop_cycle_graph();
  RC(rank_sync());
  for (int lev = 1; lev <= S.nlevs(); ++lev) {
    Level& g = S.L[lev - 1];
    if (lev > 1) {
      Level& f = S.L[lev - 2];
      const int nxs = g.d.nx / g.d.ngx, nys = g.d.ny / g.d.ngy;
      double* dst[4] = {g.dx, g.dy, g.zeta, g.h};
      const double* src[4] = {f.dx, f.dy, f.zeta, f.h};
      const double w[4] = {0.5, 0.5, 0.25, 0.25};
      for (int q = 0; q < 4; ++q) {
        double* t = g.d.gather ? g.sub2d[q] : dst[q];
        launch_coarsen2d(S.st, src[q], f.d.ny, f.d.nx, t, nys, nxs, w[q]);
        if (g.d.gather) RC(gather(lev, t, dst[q], 1));
      }
    }
    RC(fill_halo(lev, g.dx, 1, 1)); RC(fill_halo(lev, g.dy, 1, 1)); RC(fill_halo(lev, g.zeta, 1, 1)); RC(fill_halo(lev, g.h, 1, 1));
    std::vector<double> tab; vertical_tables(g.d.nz, S.hlim, S.theta_b, S.theta_s, tab);
    CU(cudaMemcpyAsync(g.tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, S.st));
    CU(cudaStreamSynchronize(S.st));   // tab is a stack temporary
    launch_zrzw(S.st, g.h, g.zeta, g.tab, S.hlim, g.zr, g.zw, g.d.nz, g.d.ny, g.d.nx);
    RC(fill_halo(lev, g.zr, g.d.nz, 2)); RC(fill_halo(lev, g.zw, g.d.nz + 1, 2));
    for (int q = 0; q < 8; ++q) launch_fill(S.st, g.v.cA[q], 0.0, g.n3());
    launch_fill(S.st, g.v.c5b, 0.0, g.n2()); launch_fill(S.st, g.v.c8b, 0.0, g.n2());
    if (lev > 1) launch_level_rmask(S.st, g.rmaskl, g.d.ny, g.d.nx, g.hf, S.P.bmask ? 1 : 0);
    launch_define_matrix_coef(S.st, matrix_view(g, lev));
    bool anynb = false; for (int q = 0; q < 8; ++q) anynb |= g.d.neighb[q] >= 0;
    if (S.P.bmask && anynb) {                                  // fill_halo(lev,cA), src/mg_define_matrix.f90:611-613: plain exchange, no physical fill
      if (!S.comm) return fail("multi-rank level but mgb_comm_init has not been called");
      for (int q = 0; q < 8; ++q) RC(exchange(g, g.v.cA[q], g.d.nz, 1));
    }
    launch_define_matrix_diag(S.st, matrix_view(g, lev));
    if (S.P.bmask && anynb) launch_refresh_derived(S.st, matrix_view(g, lev));   // c5b / c8b planes from the exchanged cA
    launch_any_nonzero2(S.st, g.v.c5b, g.v.c8b, g.n2(), S.flag_dev);
    int flag = 1;
    CU(cudaMemcpyAsync(&flag, S.flag_dev, sizeof(int), cudaMemcpyDeviceToHost, S.st));
    CU(cudaStreamSynchronize(S.st));
    g.bottom_diag = flag != 0;
    launch_rho(S.st, matrix_view(g, lev), S.sums + 3);
    double rho = 0.0;
    CU(cudaMemcpyAsync(&rho, S.sums + 3, sizeof(double), cudaMemcpyDeviceToHost, S.st));
    CU(cudaStreamSynchronize(S.st));
    g.rho = rho;
    g.v.rb_W = rb_warmup(rho);
    g.v.rb_chunk = (g.v.rb_W > RB_WMAX) ? g.d.nx : RB_CHUNK;
  }
  RC(finish());
  S.have_matrices = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// rhs / correction with host (model-layout) arrays
// ------------------------------------------------------------------------------------------------------
static size_t n_u() { const LevelDims& d = S.L[0].d; return (size_t)(d.nx + 1) * (d.ny + 2) * d.nz; }
static size_t n_v() { const LevelDims& d = S.L[0].d; return (size_t)(d.nx + 2) * (d.ny + 1) * d.nz; }
static size_t n_w() { const LevelDims& d = S.L[0].d; return (size_t)(d.nx + 2) * (d.ny + 2) * (d.nz + 1); }

// Host (model-layout, i fastest) <-> device (k fastest).  The PCIe copies run on their own stream through one staging buffer
// per array, so that the transposes of one array overlap the copy of the next (and, on the way back, the copy of u starts
// while v and w are still being transposed).
static int upload_uvw(const double* u, const double* v, const double* w) {
  const LevelDims& d = S.L[0].d;
  CU(cudaEventRecord(S.ev_c[3], S.st)); CU(cudaStreamWaitEvent(S.st_copy, S.ev_c[3], 0));     // staging buffers are free once earlier work is done
  CU(cudaMemcpyAsync(S.stage_w, w, n_w() * sizeof(double), cudaMemcpyHostToDevice, S.st_copy)); CU(cudaEventRecord(S.ev_c[2], S.st_copy));
  CU(cudaMemcpyAsync(S.stage, u, n_u() * sizeof(double), cudaMemcpyHostToDevice, S.st_copy));   CU(cudaEventRecord(S.ev_c[0], S.st_copy));
  CU(cudaMemcpyAsync(S.stage_v, v, n_v() * sizeof(double), cudaMemcpyHostToDevice, S.st_copy)); CU(cudaEventRecord(S.ev_c[1], S.st_copy));
  CU(cudaStreamWaitEvent(S.st, S.ev_c[2], 0)); launch_transpose_ijk_to_kji(S.st, S.stage_w, S.wt, d.nx + 2, d.ny + 2, d.nz + 1);
  CU(cudaStreamWaitEvent(S.st, S.ev_c[0], 0)); launch_transpose_ijk_to_kji(S.st, S.stage, S.ut, d.nx + 1, d.ny + 2, d.nz);
  CU(cudaStreamWaitEvent(S.st, S.ev_c[1], 0)); launch_transpose_ijk_to_kji(S.st, S.stage_v, S.vt, d.nx + 2, d.ny + 1, d.nz);
  return 0;
}
static int download_uvw(double* u, double* v, double* w) {
  const LevelDims& d = S.L[0].d;
  launch_transpose_kji_to_ijk(S.st, S.ut, S.stage, d.nx + 1, d.ny + 2, d.nz);       CU(cudaEventRecord(S.ev_c[0], S.st));
  launch_transpose_kji_to_ijk(S.st, S.vt, S.stage_v, d.nx + 2, d.ny + 1, d.nz);     CU(cudaEventRecord(S.ev_c[1], S.st));
  launch_transpose_kji_to_ijk(S.st, S.wt, S.stage_w, d.nx + 2, d.ny + 2, d.nz + 1); CU(cudaEventRecord(S.ev_c[2], S.st));
  CU(cudaStreamWaitEvent(S.st_copy, S.ev_c[0], 0)); CU(cudaMemcpyAsync(u, S.stage, n_u() * sizeof(double), cudaMemcpyDeviceToHost, S.st_copy));
  CU(cudaStreamWaitEvent(S.st_copy, S.ev_c[1], 0)); CU(cudaMemcpyAsync(v, S.stage_v, n_v() * sizeof(double), cudaMemcpyDeviceToHost, S.st_copy));
  CU(cudaStreamWaitEvent(S.st_copy, S.ev_c[2], 0)); CU(cudaMemcpyAsync(w, S.stage_w, n_w() * sizeof(double), cudaMemcpyDeviceToHost, S.st_copy));
  CU(cudaStreamSynchronize(S.st_copy));
  return finish();
}

// exchange of one flux face with the neighbour (what fill_halo(...,lbc_null) does where a neighbour exists):
// dir 'u': my flux(:,1:ny,nx+1) <- east's flux(:,1:ny,1) ; dir 'v': my flux(:,ny+1,1:nx) <- north's flux(:,1,1:nx)
static int exchange_flux_face(char dir) {
  Level& g = S.L[0]; const int nx = g.d.nx, ny = g.d.ny, nz = g.d.nz;
  const int east = g.d.neighb[1], west = g.d.neighb[3], north = g.d.neighb[2], south = g.d.neighb[0];
  auto at = [&](int j, int i) { return S.flux + (size_t)nz * ((size_t)j + (size_t)(ny + 2) * i); };
  if (dir == 'u') {
    if (east < 0 && west < 0) return 0;
    NC(g_nccl.GroupStart());
    if (west >= 0) NC(g_nccl.Send(at(1, 1), (size_t)nz * ny, ncclDouble, west, S.comm, S.st));
    if (east >= 0) NC(g_nccl.Recv(at(1, nx + 1), (size_t)nz * ny, ncclDouble, east, S.comm, S.st));
    NC(g_nccl.GroupEnd());
  } else {
    if (north < 0 && south < 0) return 0;
    if (south >= 0) launch_pack_rows(S.st, S.flux, g.pack8[0], nz, ny, nx, 1, 1, 1, 1, nx);
    NC(g_nccl.GroupStart());
    if (south >= 0) NC(g_nccl.Send(g.pack8[0], (size_t)nz * nx, ncclDouble, south, S.comm, S.st));
    if (north >= 0) NC(g_nccl.Recv(g.unp8[2], (size_t)nz * nx, ncclDouble, north, S.comm, S.st));
    NC(g_nccl.GroupEnd());
    if (north >= 0) launch_unpack_rows(S.st, S.flux, g.unp8[2], nz, ny, nx, 1, ny + 1, 1, 1, nx);
  }
  return 0;
}

static int compute_rhs_device() {
  Level& g = S.L[0];
  RhsView R{};
  R.nx = g.d.nx; R.ny = g.d.ny; R.nz = g.d.nz;
  R.dx = g.dx; R.dy = g.dy; R.zw = g.zw; R.dzw = g.dzw; R.zxdy = g.zxdy; R.zydx = g.zydx; R.cw = g.cw; R.rmask = S.rmask;
  R.ut = S.ut; R.vt = S.vt; R.wt = S.wt; R.flux = S.flux; R.b = g.v.b;
  // bmask: no fill_halo(uf/vf, lbc_null) at all (src/mg_compute_rhs.f90:170-172,271-273); umask / vmask close the walls
  const bool bm = S.P.bmask;
  R.bmask = bm ? 1 : 0;
  R.west_phys = !bm && !g.hf.west; R.east_phys = !bm && !g.hf.east; R.south_phys = !bm && !g.hf.south; R.north_phys = !bm && !g.hf.north;
  launch_fill(S.st, g.v.b, 0.0, g.n3());                      // rhs(:,:,:) = zero  (src/mg_compute_rhs.f90:91)
  launch_rhs_uf(S.st, R); if (S.nranks > 1 && !bm) RC(exchange_flux_face('u')); launch_rhs_div(S.st, R, 0);
  launch_rhs_vf(S.st, R); if (S.nranks > 1 && !bm) RC(exchange_flux_face('v')); launch_rhs_div(S.st, R, 1);
  launch_rhs_wf(S.st, R); launch_rhs_div(S.st, R, 2);
  return 0;
}
static int set_rmask(const double* rmask) {
  Level& g = S.L[0];
  if (rmask) CU(cudaMemcpyAsync(S.rmask, rmask, g.n2() * sizeof(double), cudaMemcpyHostToDevice, S.st));
  else launch_fill(S.st, S.rmask, 1.0, g.n2());
  return 0;
}
static int correct_uvw_device() {
  Level& g = S.L[0];
  CorrView C{g.d.nx, g.d.ny, g.d.nz, g.dx, g.dy, g.zr, g.zw, g.v.p, S.ut, S.vt, S.wt, S.rmask, S.P.bmask ? 1 : 0};
  launch_correct_uvw(S.st, C);
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// init / clean
// ------------------------------------------------------------------------------------------------------
static int clean() {
  if (S.st) cudaStreamSynchronize(S.st);
  tt_reset();
  drop_cycle_graph();
  for (void* p : S.peer_mbox) if (p) cudaIpcCloseMemHandle(p);
  if (S.err_host) cudaFreeHost(S.err_host);
  for (void* p : S.allocs) cudaFree(p);
  S.allocs.clear();
  if (S.comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(S.comm); }
  if (S.ev0) cudaEventDestroy(S.ev0);
  if (S.ev1) cudaEventDestroy(S.ev1);
  for (cudaEvent_t e : S.ev_c) if (e) cudaEventDestroy(e);
  if (S.st_copy) cudaStreamDestroy(S.st_copy);
  if (S.st) cudaStreamDestroy(S.st);
  S = Solver();
  return 0;
}

static int init(int nx, int ny, int nz, int npx, int npy, int rank, const char* namelist_path, int device) {
  if (S.inited) clean();
  S.P = Params();
  std::string msg;
  const char* path = (namelist_path && namelist_path[0]) ? namelist_path : "nh_namelist";
  int rc = namelist_read(path, S.P, msg);
  if (rc < 0) return fail(msg);
  for (auto& kv : g_overrides) if (namelist_set(S.P, kv.first.c_str(), kv.second.c_str(), msg) != 0) return fail(msg);
  std::vector<LevelDims> plan;
  if (plan_levels(S.P, nx, ny, nz, npx, npy, rank, plan, msg) != 0) return fail(msg);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) return fail(std::string("no usable CUDA device (libmgb has no CPU fallback): ") + cudaGetErrorString(ce));
  if (device >= 0) CU(cudaSetDevice(device));
  CU(cudaGetDevice(&S.device));
  S.rank = rank; S.npx = npx; S.npy = npy; S.nranks = npx * npy;
  CU(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&S.st_copy, cudaStreamNonBlocking));
  CU(cudaEventCreate(&S.ev0)); CU(cudaEventCreate(&S.ev1));
  for (cudaEvent_t& e : S.ev_c) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  { const char* tm = std::getenv("MGB_SPIN_TIMEOUT_MS"); if (tm && std::atoll(tm) > 0) S.spin_timeout_ms = (unsigned)std::atoll(tm); }
  CU(cudaHostAlloc((void**)&S.err_host, sizeof(int), cudaHostAllocMapped)); *S.err_host = 0;
  CU(cudaHostGetDevicePointer((void**)&S.err_dev, S.err_host, 0));
  { void* q = nullptr; CU(cudaMalloc(&q, 16)); S.allocs.push_back(q); CU(cudaMemsetAsync(q, 0, 16, S.st)); S.abort_dev = (int*)q; }
  S.L.resize(plan.size());
  for (size_t q = 0; q < plan.size(); ++q) {
    Level& g = S.L[q]; g.d = plan[q]; g.hf = flags_of(g.d);
    const LevelDims& d = g.d; const int lev = (int)q + 1;
    const size_t n3 = g.n3(), n2 = g.n2();
    g.v.nx = d.nx; g.v.ny = d.ny; g.v.nz = d.nz; g.v.rb_chunk = d.nx; g.v.rb_W = 0;
    RC(dalloc(&g.v.p, n3)); RC(dalloc(&g.v.b, n3)); RC(dalloc(&g.v.r, n3));
    for (int c = 0; c < 8; ++c) RC(dalloc(&g.v.cA[c], n3));
    RC(dalloc(&g.v.bet, n3)); RC(dalloc(&g.v.g, n3));
    RC(dalloc(&g.v.c5b, n2)); RC(dalloc(&g.v.c8b, n2)); RC(dalloc(&g.v.g1, n2)); RC(dalloc(&g.v.q1, n2)); RC(dalloc(&g.v.x1, n2)); RC(dalloc(&g.v.x1b, n2)); RC(dalloc(&g.v.s, n2));
    RC(dalloc(&g.dx, n2)); RC(dalloc(&g.dy, n2)); RC(dalloc(&g.zeta, n2)); RC(dalloc(&g.h, n2)); RC(dalloc(&g.rmaskl, n2));
    RC(dalloc(&g.zr, (size_t)d.nz * (d.ny + 4) * (d.nx + 4))); RC(dalloc(&g.zw, (size_t)(d.nz + 1) * (d.ny + 4) * (d.nx + 4)));
    RC(dalloc(&g.cw, (size_t)(d.nz + 1) * n2));
    RC(dalloc(&g.tab, (size_t)4 * (d.nz + 1)));
    if (lev == 1) { RC(dalloc(&g.dzw, (size_t)(d.nz + 1) * n2)); RC(dalloc(&g.zxdy, n3)); RC(dalloc(&g.zydx, n3)); }
    { double* q = nullptr; RC(dalloc(&q, 2)); g.gridbar = (unsigned*)q; }
    { const size_t nt = relax_fused_num_tiles(d.nx, d.ny, d.nz);           // fused relaxation: tagged q1 plane + {launch number, exit counter}
      if (nt) { double* q = nullptr; RC(dalloc(&q, 2 * n2)); g.v.q1t = q; RC(dalloc(&q, 1)); g.v.tctl = (unsigned*)q; }
      g.v.err = S.err_dev; g.v.abort_flag = S.abort_dev; g.v.timeout_ms = S.spin_timeout_ms; }
    g.npartials = residual_num_partials(g.v);
    RC(dalloc(&g.partials, (size_t)g.npartials));
    { void* pp = nullptr; CU(cudaMalloc(&pp, 8 * sizeof(double*))); S.allocs.push_back(pp); g.planes_dev = (double**)pp;
      CU(cudaMemcpy(g.planes_dev, g.v.cA, 8 * sizeof(double*), cudaMemcpyHostToDevice)); }
    if (d.gather) {
      const int nxs = d.nx / d.ngx, nys = d.ny / d.ngy;
      RC(dalloc(&g.dummy3, (size_t)d.nz * (nys + 2) * (nxs + 2)));
      RC(dalloc(&g.gatherbuf, (size_t)4 * d.nz * (nys + 2) * (nxs + 2)));
      for (int c = 0; c < 4; ++c) RC(dalloc(&g.sub2d[c], (size_t)(nys + 2) * (nxs + 2)));
    }
    bool any = false; for (int c = 0; c < 8; ++c) any |= d.neighb[c] >= 0;
    if (any) {   // staging for strided faces: sized for nh = 2 and nz+1 levels
      const size_t face = (size_t)2 * (d.nz + 1) * (size_t)(std::max(d.nx, d.ny) + 4);
      for (int c = 0; c < 8; ++c) { RC(dalloc(&g.pack8[c], face)); RC(dalloc(&g.unp8[c], face)); }
    }
  }
  { const LevelDims& d = S.L[0].d;
    RC(dalloc(&S.ut, (size_t)(d.nx + 1) * (d.ny + 2) * d.nz)); RC(dalloc(&S.vt, (size_t)(d.nx + 2) * (d.ny + 1) * d.nz));
    RC(dalloc(&S.wt, (size_t)(d.nx + 2) * (d.ny + 2) * (d.nz + 1))); RC(dalloc(&S.stage, (size_t)(d.nx + 2) * (d.ny + 2) * (d.nz + 1)));
    RC(dalloc(&S.stage_v, (size_t)(d.nx + 2) * (d.ny + 2) * (d.nz + 1))); RC(dalloc(&S.stage_w, (size_t)(d.nx + 2) * (d.ny + 2) * (d.nz + 1)));
    RC(dalloc(&S.flux, (size_t)(d.nx + 2) * (d.ny + 2) * (d.nz + 1))); RC(dalloc(&S.rmask, (size_t)(d.nx + 2) * (d.ny + 2)));
    RC(dalloc(&S.sums, 8));
    void* fp = nullptr; CU(cudaMalloc(&fp, sizeof(int))); S.allocs.push_back(fp); S.flag_dev = (int*)fp; }
  const char* ex = std::getenv("MGB_EXACT_STATE");
  S.fused_prolong = !(ex && ex[0] == '1');
  const char* ns = std::getenv("MGB_NO_SMALL_KERNEL");
  S.use_small = !(ns && ns[0] == '1');
  const char* nm = std::getenv("MGB_NO_MID_KERNEL");
  S.use_mid = !(nm && nm[0] == '1');
  const char* nf = std::getenv("MGB_NO_FUSED_HALO");
  S.fuse_halo = !(nf && nf[0] == '1');
  { const char* nfe = std::getenv("MGB_NO_FUSED_EXCHANGE"); S.fuse_exchange = !(nfe && nfe[0] == '1'); }
  // the one-launch-per-colour kernel is correct but not yet faster than A + B + C (its Thomas / recurrence warps are the
  // critical path, profiles/summary_r02.md): opt-in
  { const char* fr = std::getenv("MGB_FUSED_RELAX"); S.use_fused = fr && fr[0] == '1'; }
  const char* nb = std::getenv("MGB_NO_FUSED_BC");
  S.fuse_bc = !(nb && nb[0] == '1');
  { const char* fl = std::getenv("MGB_FUSE_BC_LIMIT"); if (fl) S.fuse_bc_limit = (size_t)std::atoll(fl); }
  { const char* tt = std::getenv("MGB_TICTOC"); T = TicToc(); T.on = tt && tt[0] == '1'; }
  const char* ng = std::getenv("MGB_NO_GRAPH");
  S.use_graph = !(ng && ng[0] == '1');
  kernels_init();
  CU(cudaStreamSynchronize(S.st));
  S.inited = true;
  g_plan = plan;
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// field access
// ------------------------------------------------------------------------------------------------------
struct FieldRef { double* ptr; size_t n; bool is_cA; };
static int field_ref(int lev, const std::string& name, FieldRef& out) {
  Level& g = S.L[lev - 1]; const LevelDims& d = g.d;
  const size_t n3 = g.n3(), n2 = g.n2();
  out.is_cA = false;
  if (name == "p") out = {g.v.p, n3, false}; else if (name == "b") out = {g.v.b, n3, false}; else if (name == "r") out = {g.v.r, n3, false};
  else if (name == "cA") out = {nullptr, 8 * n3, true};
  else if (name == "dx") out = {g.dx, n2, false}; else if (name == "dy") out = {g.dy, n2, false};
  else if (name == "zeta") out = {g.zeta, n2, false}; else if (name == "h") out = {g.h, n2, false};
  else if (name == "zr") out = {g.zr, (size_t)d.nz * (d.ny + 4) * (d.nx + 4), false};
  else if (name == "zw") out = {g.zw, (size_t)(d.nz + 1) * (d.ny + 4) * (d.nx + 4), false};
  else if (name == "cw") out = {g.cw, (size_t)(d.nz + 1) * n2, false};
  else if (name == "dzw" && lev == 1) out = {g.dzw, (size_t)(d.nz + 1) * n2, false};
  else if (name == "zxdy" && lev == 1) out = {g.zxdy, n3, false};
  else if (name == "zydx" && lev == 1) out = {g.zydx, n3, false};
  else if (name == "bet") out = {g.v.bet, n3, false}; else if (name == "g") out = {g.v.g, n3, false};
  else if (name == "s") out = {g.v.s, n2, false}; else if (name == "x1") out = {g.v.x1, n2, false}; else if (name == "q1") out = {g.v.q1, n2, false};
  else if (name == "dummy3" && d.gather) out = {g.dummy3, (size_t)d.nz * (d.ny / d.ngy + 2) * (d.nx / d.ngx + 2), false};
  else return fail("unknown field '" + name + "'");
  return 0;
}

}  // namespace mgb

// =========================================================================================================
// C ABI
// =========================================================================================================
using namespace mgb;


// ------------------------------------------------------------------------------------------------------
// peer-memory mailboxes for the nh = 1 halo exchange (replaces the Isend/Irecv pairs of src/mg_mpi_exchange.f90:396-745
// with direct NVLink stores).  Collective over all ranks; any rank that cannot map a neighbour turns it off everywhere
// (the NCCL send/recv exchange stays as the transport then).  MGB_NO_P2P=1 forces that.
// ------------------------------------------------------------------------------------------------------
static int p2p_setup() {
  S.p2p = false;
  if (S.nranks <= 1) return 0;
  const char* off = std::getenv("MGB_NO_P2P");
  int ok = !(off && off[0] == '1');
  const int nl = S.nlevs();
  // layout (identical on every rank: the decomposition is uniform): per level 2 parities of slots of 16-byte tagged words
  std::vector<size_t> data_off(nl, 0);
  size_t bytes = 0;
  auto slot_doubles = [](const LevelDims& d) { const size_t nzm = (size_t)d.nz + 1; return 2 * nzm * d.nx + 2 * nzm * d.ny + 4 * nzm; };
  for (int q = 0; q < nl; ++q) {
    bool any = false; for (int c = 0; c < 8; ++c) any |= S.L[q].d.neighb[c] >= 0;
    if (!any) continue;
    data_off[q] = bytes; bytes += ((2 * slot_doubles(S.L[q].d) * sizeof(uint4) + 255) / 256) * 256;
  }
  if (bytes == 0) return 0;
  std::vector<char> handles((size_t)64 * S.nranks, 0);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  char* hbuf = nullptr;
  { void* q = nullptr; CU(cudaMalloc(&q, (size_t)64 * S.nranks + 8)); S.allocs.push_back(q); hbuf = (char*)q; }
  { void* q = nullptr; CU(cudaMalloc(&q, bytes)); S.allocs.push_back(q); S.mbox = q; CU(cudaMemsetAsync(q, 0, bytes, S.st)); }
  { void* q = nullptr; CU(cudaMalloc(&q, (size_t)nl * 16)); S.allocs.push_back(q); CU(cudaMemsetAsync(q, 0, (size_t)nl * 16, S.st));
    S.p2p_seq = (unsigned long long*)q; S.p2p_done = (unsigned int*)((char*)q + (size_t)nl * 8); }
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, S.mbox) != cudaSuccess) { cudaGetLastError(); ok = 0; std::memset(&mine, 0, sizeof mine); }
  CU(cudaMemcpyAsync(hbuf + (size_t)64 * S.rank, &mine, 64, cudaMemcpyHostToDevice, S.st));
  NC(g_nccl.AllGather(hbuf + (size_t)64 * S.rank, hbuf, 64, ncclChar, S.comm, S.st));
  CU(cudaMemcpyAsync(handles.data(), hbuf, handles.size(), cudaMemcpyDeviceToHost, S.st));
  CU(cudaStreamSynchronize(S.st));
  S.peer_mbox.assign(S.nranks, nullptr);
  if (ok) {
    for (int q = 0; q < nl && ok; ++q)
      for (int c = 0; c < 8 && ok; ++c) {
        const int nb = S.L[q].d.neighb[c];
        if (nb < 0 || nb == S.rank || S.peer_mbox[nb]) continue;
        cudaIpcMemHandle_t h; std::memcpy(&h, handles.data() + (size_t)64 * nb, 64);
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        S.peer_mbox[nb] = ptr;
      }
  }
  // agree: everyone or no one
  double* flag = (double*)(hbuf + (size_t)64 * S.nranks);   // 8 spare bytes behind the handles
  const double mineok = ok ? 1.0 : 0.0; double allok = 0.0;
  CU(cudaMemcpyAsync(flag, &mineok, sizeof(double), cudaMemcpyHostToDevice, S.st));
  NC(g_nccl.AllReduce(flag, flag, 1, ncclDouble, ncclMin, S.comm, S.st));
  CU(cudaMemcpyAsync(&allok, flag, sizeof(double), cudaMemcpyDeviceToHost, S.st));
  CU(cudaStreamSynchronize(S.st));
  if (allok < 0.5) {
    if (!(off && off[0] == '1')) std::fprintf(stderr, "[mgb] rank %d: peer-memory halo exchange unavailable, using NCCL send/recv\n", S.rank);
    return 0;
  }
  for (int q = 0; q < nl; ++q) {
    Level& g = S.L[q]; P2PBox& B = g.box;
    bool any = false; for (int c = 0; c < 8; ++c) any |= g.d.neighb[c] >= 0;
    if (!any) continue;
    for (int c = 0; c < 8; ++c) {
      const int nb = g.d.neighb[c];
      char* base = nb < 0 ? nullptr : (nb == S.rank ? (char*)S.mbox : (char*)S.peer_mbox[nb]);
      B.peer[c] = base ? (uint4*)(base + data_off[q]) : nullptr;
    }
    B.mine = (const uint4*)((char*)S.mbox + data_off[q]);
    B.seq = S.p2p_seq + q; B.done = S.p2p_done + q; B.err = S.err_dev; B.abort_flag = S.abort_dev; B.timeout_ms = S.spin_timeout_ms; B.nzmax = g.d.nz + 1;
  }
  S.p2p = true;
  return 0;
}

extern "C" {

const char* mgb_last_error(void) { return g_err.c_str(); }
const char* mgb_version(void) { return "mgroms_b200 0.1 (sm_100a, FP64, -fmad=false)"; }

int mgb_set_param(const char* key, const char* value) {
  std::string m; if (namelist_set(g_params, key, value, m) != 0) return fail(m);
  g_overrides.push_back({key, value}); return 0;
}
int mgb_reset_params(void) { g_params = Params(); g_overrides.clear(); return 0; }

int mgb_namelist_probe(const char* path, char* out, int cap) {
  Params P; std::string m;
  int rc = namelist_read((path && path[0]) ? path : "nh_namelist", P, m);
  if (rc < 0) return fail(m);
  for (auto& kv : g_overrides) if (namelist_set(P, kv.first.c_str(), kv.second.c_str(), m) != 0) return fail(m);
  char buf[1024];
  std::snprintf(buf, sizeof buf, "solver_prec=%.17g;solver_maxiter=%d;nsmall=%d;ns_coarsest=%d;ns_pre=%d;ns_post=%d;cmatrix=%s;relax_method=%s;"
                "interp_type=%s;restrict_type=%s;netcdf_output=%d;aggressive=%d;bmask=%d;file=%d",
                P.solver_prec, P.solver_maxiter, P.nsmall, P.ns_coarsest, P.ns_pre, P.ns_post, P.cmatrix.c_str(), P.relax_method.c_str(),
                P.interp_type.c_str(), P.restrict_type.c_str(), (int)P.netcdf_output, (int)P.aggressive, (int)P.bmask, rc == 0 ? 1 : 0);
  std::strncpy(out, buf, cap); if (cap > 0) out[cap - 1] = 0;
  return 0;
}

int mgb_plan_levels(int nx, int ny, int nz, int npx, int npy, int rank) {
  std::string m; Params P = g_params;
  if (plan_levels(P, nx, ny, nz, npx, npy, rank, g_plan, m) != 0) return fail(m);
  return 0;
}
int mgb_nlevs(void) { return (int)g_plan.size(); }
int mgb_level_info(int lev, int* info) {
  if (lev < 1 || lev > (int)g_plan.size()) return fail("level out of range");
  const LevelDims& d = g_plan[lev - 1];
  const int v[12] = {d.nx, d.ny, d.nz, d.npx, d.npy, d.incx, d.incy, d.gather, d.ngx, d.ngy, d.key, d.color};
  for (int q = 0; q < 12; ++q) info[q] = v[q];
  for (int q = 0; q < 8; ++q) info[12 + q] = d.neighb[q];
  return 0;
}
int mgb_gather_members(int lev, int* out4) {
  if (lev < 1 || lev > (int)g_plan.size()) return -1;
  const LevelDims& d = g_plan[lev - 1];
  for (int q = 0; q < d.nmembers; ++q) out4[q] = d.members[q];
  return d.nmembers;
}

int mgb_init(int nx, int ny, int nz, int npx, int npy, int rank, const char* namelist_path, int device) {
  return init(nx, ny, nz, npx, npy, rank, namelist_path, device);
}
int mgb_clean(void) { return clean(); }

int mgb_comm_unique_id(void* id128) {
  RC(nccl_load());
  ncclUniqueId id; NC(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(id128, &id, 128);
  return 0;
}
int mgb_comm_init(const void* id128) {
  if (!S.inited) return fail("mgb_init first");
  RC(nccl_load());
  ncclUniqueId id; std::memcpy(&id, id128, 128);
  NC(g_nccl.CommInitRank(&S.comm, S.nranks, id, S.rank));
  return p2p_setup();
}

int mgb_matrices(const double* dx, const double* dy, const double* zeta, const double* h, const double* rmask,
                 double hc, double theta_b, double theta_s) {
  if (!S.inited) return fail("mgb_init has not been called");
  Level& g = S.L[0]; const size_t n2 = g.n2() * sizeof(double);
  S.hlim = hc; S.theta_b = theta_b; S.theta_s = theta_s;
  CU(cudaMemcpyAsync(g.dx, dx, n2, cudaMemcpyHostToDevice, S.st)); CU(cudaMemcpyAsync(g.dy, dy, n2, cudaMemcpyHostToDevice, S.st));
  CU(cudaMemcpyAsync(g.zeta, zeta, n2, cudaMemcpyHostToDevice, S.st)); CU(cudaMemcpyAsync(g.h, h, n2, cudaMemcpyHostToDevice, S.st));
  RC(set_rmask(rmask));
  CU(cudaStreamSynchronize(S.st));
  return define_matrices();
}

int mgb_compute_rhs(const double* rmask, const double* u, const double* v, const double* w) {
  if (!S.have_matrices) return fail("mgb_matrices has not been called");
  RC(set_rmask(rmask)); RC(upload_uvw(u, v, w)); RC(compute_rhs_device());
  return finish();
}
int mgb_check_nondivergence(const double* rmask, const double* u, const double* v, const double* w) { return mgb_compute_rhs(rmask, u, v, w); }
int mgb_solve_p(void) { if (!S.inited) return fail("mgb_init has not been called"); return solve_p(); }
int mgb_correct_uvw(double* u, double* v, double* w) {
  if (!S.have_matrices) return fail("mgb_matrices has not been called");
  RC(correct_uvw_device()); return download_uvw(u, v, w);
}
int mgb_solve(const double* rmask, double* u, double* v, double* w) {
  if (!S.have_matrices) return fail("mgb_matrices has not been called");
  TicGuard tg("nhydro_solve", 1);                             // src/nhydro.f90:67,100
  RC(set_rmask(rmask)); RC(upload_uvw(u, v, w));
  { TicGuard tr("compute_rhs", 1); RC(compute_rhs_device()); }   // src/nhydro.f90:81-83
  RC(solve_p());
  RC(correct_uvw_device()); return download_uvw(u, v, w);
}

int mgb_tictoc_enable(int on) {
  if (!S.inited) return fail("mgb_init has not been called");
  tt_reset();
  T.on = on != 0;
  return 0;
}
// print_tictoc (src/mg_tictoc.f90:114-153): per timer one line of seconds (Total + one column per level, E9.3) and one of calls (I9)
int mgb_write_tictoc(const char* path) {
  if (!S.inited) return fail("mgb_init has not been called");
  tt_flush();
  char def[32]; std::snprintf(def, sizeof def, "fort.%d", 10 + S.rank);
  FILE* f = std::fopen((path && path[0]) ? path : def, "w");
  if (!f) return fail("cannot open the tictoc file");
  auto e93 = [](double v, char* out) {       // Fortran E9.3: 0.dddE+xx
    if (v == 0.0) { std::snprintf(out, 32, "0.000E+00"); return; }
    int ex = (int)std::floor(std::log10(std::fabs(v))) + 1;
    double m = v / std::pow(10.0, ex);
    if (std::fabs(m) >= 0.9995) { m /= 10.0; ++ex; }
    std::snprintf(out, 32, "%.3fE%c%02d", m, ex < 0 ? '-' : '+', std::abs(ex) % 100);
  };
  const int nl = S.nlevs(); char buf[32];
  std::fprintf(f, "%21s%10s", "", "Total");
  for (int l = 1; l <= nl; ++l) std::fprintf(f, " %9d", l);
  std::fprintf(f, " \n");
  for (size_t q = 0; q < T.names.size(); ++q) {
    double tot = 0; long long ctot = 0;
    for (int l = 0; l < nl; ++l) { tot += T.time[q][l]; ctot += T.calls[q][l]; }
    e93(tot, buf); std::fprintf(f, " %20s %s", T.names[q].c_str(), buf);
    for (int l = 0; l < nl; ++l) { e93(T.time[q][l], buf); std::fprintf(f, " %s", buf); }
    std::fprintf(f, " \n%21s %9lld", "", ctot);
    for (int l = 0; l < nl; ++l) std::fprintf(f, " %9lld", T.calls[q][l]);
    std::fprintf(f, " \n");
  }
  std::fclose(f);
  return 0;
}

int mgb_get_history(double* res, double* conv, int cap) {
  const int n = (int)S.hist_res.size();
  for (int q = 0; q < n && q < cap; ++q) { res[q] = S.hist_res[q]; if (conv) conv[q] = S.hist_conv[q]; }
  return n;
}
int mgb_write_fort100(const char* path) {
  FILE* f = std::fopen(path, "w"); if (!f) return fail("cannot open fort.100 target");
  for (size_t q = 0; q < S.hist_res.size(); ++q) {
    if (q == 0) std::fprintf(f, "  %24.16E %11d\n", S.hist_res[0], 0);
    else std::fprintf(f, "  %24.16E  %24.16E\n", S.hist_res[q], S.hist_conv[q]);
  }
  std::fclose(f); return 0;
}
double mgb_last_solve_ms(void) { return S.last_ms; }
int mgb_last_solve_iters(void) { return S.last_iters; }

int mgb_relax(int lev, int nsweeps) { RC(check_level(lev)); RC(rank_sync()); RC(relax(lev, nsweeps)); return finish(); }
int mgb_residual(int lev, double* res) {
  RC(check_level(lev)); RC(rank_sync()); RC(residual(lev));
  double s = 0; RC(global_sum(lev, S.sums, &s)); *res = std::sqrt(s); return 0;
}
int mgb_fine2coarse(int lev) { RC(check_level(lev)); if (lev >= S.nlevs()) return fail("no coarser level"); RC(rank_sync()); RC(fine2coarse(lev, false)); return finish(); }
int mgb_coarse2fine(int lev) { RC(check_level(lev)); if (lev >= S.nlevs()) return fail("no coarser level"); RC(rank_sync()); RC(coarse2fine(lev, false)); return finish(); }
int mgb_fcycle(void) { RC(check_level(1)); RC(rank_sync()); RC(fcycle()); return finish(); }
int mgb_vcycle(int lev1) { RC(check_level(lev1)); RC(rank_sync()); RC(vcycle(lev1)); return finish(); }
int mgb_fill_halo(int lev, const char* name) {
  RC(check_level(lev)); FieldRef f; RC(field_ref(lev, name, f));
  RC(rank_sync());
  if (f.is_cA) {                                               // fill_halo_4D, src/mg_mpi_exchange.f90:1245-1530
    Level& g = S.L[lev - 1];
    bool anynb = false; for (int q = 0; q < 8; ++q) anynb |= g.d.neighb[q] >= 0;
    if (anynb) { if (!S.comm) return fail("mgb_comm_init has not been called"); for (int q = 0; q < 8; ++q) RC(exchange(g, g.v.cA[q], g.d.nz, 1)); }
    if (!S.P.bmask) {                                          // Neumann branch of the reference: zero halo where there is no neighbour
      return fail("fill_halo(cA) without bmask zeroes the physical halo of cA in the reference; it is never called there and is not offered");
    }
    return finish();
  }
  const LevelDims& d = S.L[lev - 1].d; const std::string n = name;
  if (n == "zr") RC(fill_halo(lev, f.ptr, d.nz, 2)); else if (n == "zw") RC(fill_halo(lev, f.ptr, d.nz + 1, 2));
  else if (n == "dx" || n == "dy" || n == "zeta" || n == "h") RC(fill_halo(lev, f.ptr, 1, 1));
  else if (n == "cw" || n == "dzw") RC(fill_halo(lev, f.ptr, d.nz + 1, 1));
  else RC(fill_halo(lev, f.ptr, d.nz, 1));
  CU(cudaStreamSynchronize(S.st)); return 0;
}

long long mgb_field_size(int lev, const char* name) {
  if (check_level(lev) != 0) return -1;
  FieldRef f; if (field_ref(lev, name, f) != 0) return -1;
  return (long long)f.n;
}
int mgb_get_field(int lev, const char* name, double* out) {
  RC(check_level(lev)); FieldRef f; RC(field_ref(lev, name, f));
  if (f.is_cA) {
    double* tmp = nullptr; CU(cudaMalloc((void**)&tmp, f.n * sizeof(double)));
    launch_soa_to_aos8(S.st, S.L[lev - 1].planes_dev, tmp, f.n / 8);
    cudaError_t e = cudaMemcpyAsync(out, tmp, f.n * sizeof(double), cudaMemcpyDeviceToHost, S.st);
    cudaStreamSynchronize(S.st); cudaFree(tmp);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e));
    return 0;
  }
  CU(cudaMemcpyAsync(out, f.ptr, f.n * sizeof(double), cudaMemcpyDeviceToHost, S.st));
  CU(cudaStreamSynchronize(S.st)); return 0;
}
int mgb_set_field(int lev, const char* name, const double* in) {
  RC(check_level(lev)); FieldRef f; RC(field_ref(lev, name, f));
  if (f.is_cA) {
    Level& g = S.L[lev - 1];
    double* tmp = nullptr; CU(cudaMalloc((void**)&tmp, f.n * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(tmp, in, f.n * sizeof(double), cudaMemcpyHostToDevice, S.st);
    launch_aos8_to_soa(S.st, tmp, g.planes_dev, f.n / 8);
    cudaStreamSynchronize(S.st); cudaFree(tmp);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e));
    drop_cycle_graph();
    launch_refresh_derived(S.st, matrix_view(g, lev));     // bet, g, c5b, c8b follow cA
    launch_rho(S.st, matrix_view(g, lev), S.sums + 3);
    double rho = 0.0;
    CU(cudaMemcpyAsync(&rho, S.sums + 3, sizeof(double), cudaMemcpyDeviceToHost, S.st));
    CU(cudaStreamSynchronize(S.st));
    g.rho = rho; g.v.rb_W = rb_warmup(rho); g.v.rb_chunk = (g.v.rb_W > RB_WMAX) ? g.d.nx : RB_CHUNK;
    launch_any_nonzero2(S.st, g.v.c5b, g.v.c8b, g.n2(), S.flag_dev);   // the installed matrix may (or may no longer) couple the k=1 diagonals
    int flag = 1;
    CU(cudaMemcpyAsync(&flag, S.flag_dev, sizeof(int), cudaMemcpyDeviceToHost, S.st));
    CU(cudaStreamSynchronize(S.st));
    g.bottom_diag = flag != 0;
    return 0;
  }
  CU(cudaMemcpyAsync(f.ptr, in, f.n * sizeof(double), cudaMemcpyHostToDevice, S.st));
  CU(cudaStreamSynchronize(S.st)); return 0;
}

int mgb_relax_path(int lev, char* out, int cap) {
  RC(check_level(lev));
  if (!S.have_matrices) return fail("mgb_matrices has not been called");
  const std::string r = relax_path(lev);
  std::strncpy(out, r.c_str(), cap); if (cap > 0) out[cap - 1] = 0;
  return 0;
}
int mgb_halo_transport(void) { return !S.comm ? 0 : (S.p2p ? 1 : 2); }
long long mgb_launch_count(void) { return g_launches; }
void mgb_reset_launch_count(void) { g_launches = 0; }

int mgb_profile_halo(int lev, int reps, double* us_per_exchange) {
  RC(check_level(lev));
  RC(rank_sync());
  Level& g = S.L[lev - 1];
  cudaEvent_t a, b; CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
  float ms = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int n = pass == 0 ? 3 : reps;
    CU(cudaEventRecord(a, S.st));
    for (int r = 0; r < n; ++r) RC(fill_halo(lev, g.v.p, g.d.nz, 1));
    CU(cudaEventRecord(b, S.st)); CU(cudaEventSynchronize(b)); CU(cudaEventElapsedTime(&ms, a, b));
    *us_per_exchange = 1e3 * ms / n;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return finish();
}

int mgb_profile_level1(int reps, double* out4) {
  if (!S.have_matrices) return fail("mgb_matrices has not been called");
  if (S.nlevs() < 2) return fail("need at least two levels");
  Level& g = S.L[0]; Level& c = S.L[1];
  RC(rank_sync());
  cudaEvent_t a, b; CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
  float ms = 0;
  // relax: `reps` full red-black sweeps at level 1 through the production path; reported per colour half-sweep
  for (int pass = 0; pass < 2; ++pass) {
    const int n = pass == 0 ? 1 : reps;
    CU(cudaEventRecord(a, S.st));
    RC(relax(1, n));
    CU(cudaEventRecord(b, S.st)); CU(cudaEventSynchronize(b)); CU(cudaEventElapsedTime(&ms, a, b));
    out4[0] = ms / (2 * n);
  }
  for (int pass = 0; pass < 2; ++pass) {
    const int n = pass == 0 ? 2 : reps;
    CU(cudaEventRecord(a, S.st));
    for (int r = 0; r < n; ++r) RC(residual(1));
    CU(cudaEventRecord(b, S.st)); CU(cudaEventSynchronize(b)); CU(cudaEventElapsedTime(&ms, a, b));
    out4[1] = ms / n;
  }
  if (!c.d.gather) {
    for (int pass = 0; pass < 2; ++pass) {
      const int n = pass == 0 ? 2 : reps;
      CU(cudaEventRecord(a, S.st));
      for (int r = 0; r < n; ++r) RC(fine2coarse(1, false));
      CU(cudaEventRecord(b, S.st)); CU(cudaEventSynchronize(b)); CU(cudaEventElapsedTime(&ms, a, b));
      out4[2] = ms / n;
    }
    for (int pass = 0; pass < 2; ++pass) {
      const int n = pass == 0 ? 2 : reps;
      CU(cudaEventRecord(a, S.st));
      for (int r = 0; r < n; ++r) RC(coarse2fine(1, true));
      CU(cudaEventRecord(b, S.st)); CU(cudaEventSynchronize(b)); CU(cudaEventElapsedTime(&ms, a, b));
      out4[3] = ms / n;
    }
  } else { out4[2] = out4[3] = 0.0; }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return 0;
}

}  // extern "C"

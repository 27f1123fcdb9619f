// mgoracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT)
//
// A plain host restatement of the multigrid pressure-solve path of CESR-lab/mgroms 0.3.9
// (Fortran 90 + MPI).  It exists only so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs can check and time "what the reference computes".
// Nothing under mgroms_b200/ may include, link or call this file.
//
// PARITY: PINNED AGAINST THE REFERENCE'S OWN SOURCE TEXT, ONE RANK AND DECOMPOSED.  The reference ships no golden vectors
// and the image has no Fortran compiler / MPI, so no native reference build exists; instead the reference's Fortran
// routines -- the solver AND its MPI layer (fill_halo_*, gather / split, global_sum, define_grids / neighbours / gather
// families) -- are translated statement by statement to C (oracle/f90toc.py, oracle/build_ref.py -> oracle/_ref/libmgref.so)
// and run by oracle/refdriver.py with the MPI library calls emulated between threads (oracle/ref/ref_mpi.c).  With
// relax_order = 0 this restatement reproduces that output BIT FOR BIT on 1, 2x1, 2x2 and 4x2 ranks (level plan, coefficients
// of every level, rhs, the whole residual history, p, u/v/w; tests/test_cpu_ref_pin.py, vectors in tests/golden/ref_golden.*).
// Independent properties are tested as well (A.1 = 0, projection identity, SciPy sparse assembly of the sweep, halo = global
// Neumann extension, decomposition invariance, the surveyor's NumPy history; tests/test_cpu_oracle.py and friends).
//
// MPI ranks are emulated: a World holds npx*npy Rank objects in ONE process; every operator is a
// bulk-synchronous loop over ranks (OpenMP thread per rank), and halo exchange / gather copy
// between the ranks' arrays with exactly the reference's edge, corner and ordering rules.
//
// Each routine cites the reference file:line it follows (paths relative to /root/reference/src).
// Expressions keep the reference's operation order; compile with -ffp-contract=off.

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// array views (Fortran index conventions, column-major, k fastest)
// ---------------------------------------------------------------------------------------------
struct A3 {            // a(nzl, 1-nh:ny+nh, 1-nh:nx+nh)
  double* d = nullptr; int nzl = 0, ny = 0, nx = 0, nh = 1;
  inline double& operator()(int k, int j, int i) const {
    return d[(size_t)(k - 1) + (size_t)nzl * ((size_t)(j - 1 + nh) + (size_t)(ny + 2 * nh) * (size_t)(i - 1 + nh))];
  }
  size_t size() const { return (size_t)nzl * (ny + 2 * nh) * (nx + 2 * nh); }
};
struct A2 {            // a(0:ny+1, 0:nx+1)
  double* d = nullptr; int ny = 0, nx = 0;
  inline double& operator()(int j, int i) const { return d[(size_t)j + (size_t)(ny + 2) * (size_t)i]; }
  size_t size() const { return (size_t)(ny + 2) * (nx + 2); }
};
struct A4 {            // cA(8, nz, 0:ny+1, 0:nx+1)   mg_grids.f90:221
  double* d = nullptr; int nz = 0, ny = 0, nx = 0;
  inline double& operator()(int c, int k, int j, int i) const {
    return d[(size_t)(c - 1) + 8 * ((size_t)(k - 1) + (size_t)nz * ((size_t)j + (size_t)(ny + 2) * (size_t)i))];
  }
  size_t size() const { return (size_t)8 * nz * (ny + 2) * (nx + 2); }
};

struct Store {         // owns memory for views
  std::vector<std::vector<double>> bufs;
  double* get(size_t n) { bufs.emplace_back(n, 0.0); return bufs.back().data(); }
};

// ---------------------------------------------------------------------------------------------
// namelist  (mg_namelist.f90:11-50)
// ---------------------------------------------------------------------------------------------
struct Params {
  int nsmall = 8, ns_coarsest = 40, ns_pre = 3, ns_post = 2;
  double solver_prec = 1e-6; int solver_maxiter = 50;
  std::string cmatrix = "real", relax_method = "RB", interp_type = "linear", restrict_type = "avg";
  bool aggressive = false, netcdf_output = false, bmask = false;
};

static std::string lower(std::string s) { for (auto& c : s) c = (char)tolower(c); return s; }
static std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
static bool parse_logical(const std::string& v) { std::string l = lower(v); return l.find('t') != std::string::npos && l.find('t') < 3; }
static double parse_real(std::string v) { for (auto& c : v) if (c == 'd' || c == 'D') c = 'e'; return atof(v.c_str()); }

// Minimal Fortran namelist reader for group &nhparam (key = value, "!" comments, "/" terminator).
static int read_namelist(const char* path, Params& P) {
  FILE* f = fopen(path, "r");
  if (!f) return 0;                       // absent file => defaults (mg_namelist.f90:72-83)
  std::string txt; char buf[4096]; size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) txt.append(buf, n);
  fclose(f);
  std::string clean;
  { size_t pos = 0;
    while (pos < txt.size()) {
      size_t e = txt.find('\n', pos); if (e == std::string::npos) e = txt.size();
      std::string line = txt.substr(pos, e - pos);
      bool inq = false; size_t cut = line.size();
      for (size_t q = 0; q < line.size(); ++q) { if (line[q] == '\'') inq = !inq; if (line[q] == '!' && !inq) { cut = q; break; } }
      clean += line.substr(0, cut) + "\n"; pos = e + 1;
    } }
  size_t g = lower(clean).find("&nhparam");
  if (g == std::string::npos) return -1;
  size_t end = clean.find('/', g); if (end == std::string::npos) end = clean.size();
  std::string body = clean.substr(g + 8, end - g - 8);
  for (auto& c : body) if (c == '\n') c = ',';
  size_t pos = 0;
  while (pos < body.size()) {
    size_t e = body.find(',', pos); if (e == std::string::npos) e = body.size();
    std::string item = trim(body.substr(pos, e - pos)); pos = e + 1;
    if (item.empty()) continue;
    size_t eq = item.find('='); if (eq == std::string::npos) return -2;
    std::string key = lower(trim(item.substr(0, eq))), val = trim(item.substr(eq + 1));
    if (val.size() >= 2 && val.front() == '\'' && val.back() == '\'') val = val.substr(1, val.size() - 2);
    if      (key == "solver_prec")    P.solver_prec = parse_real(val);
    else if (key == "solver_maxiter") P.solver_maxiter = atoi(val.c_str());
    else if (key == "nsmall")         P.nsmall = atoi(val.c_str());
    else if (key == "ns_coarsest")    P.ns_coarsest = atoi(val.c_str());
    else if (key == "ns_pre")         P.ns_pre = atoi(val.c_str());
    else if (key == "ns_post")        P.ns_post = atoi(val.c_str());
    else if (key == "cmatrix")        P.cmatrix = trim(val);
    else if (key == "relax_method")   P.relax_method = trim(val);
    else if (key == "interp_type")    P.interp_type = trim(val);
    else if (key == "restrict_type")  P.restrict_type = trim(val);
    else if (key == "netcdf_output")  P.netcdf_output = parse_logical(val);
    else if (key == "aggressive")     P.aggressive = parse_logical(val);
    else if (key == "bmask")          P.bmask = parse_logical(val);
    else return -3;                       // unknown key aborts the read in Fortran too
  }
  return 1;
}

// ---------------------------------------------------------------------------------------------
// grid hierarchy   (mg_grids.f90:24-65)
// ---------------------------------------------------------------------------------------------
struct Grid {
  int nx = 0, ny = 0, nz = 0, npx = 1, npy = 1, incx = 1, incy = 1;
  int gather = 0, ngx = 1, ngy = 1, key = 0, color = 0;
  int neighb[8] = {-1, -1, -1, -1, -1, -1, -1, -1};   // S,E,N,W,SW,SE,NE,NW ; -1 = MPI_PROC_NULL
  std::vector<int> members;                            // ranks of my gather communicator sorted by key
  A4 cA; A3 p, b, r, zr, zw, cw, dzw, zxdy, zydx, dummy3, dummy3Dnz, dummy3Dnzp;
  A2 dx, dy, zeta, h;
};

struct Rank {
  int rank = 0, pi = 0, pj = 0;
  std::vector<Grid> grid;      // grid[lev], lev = 1..nlevs (index 0 unused)
  Store mem;
  A2 rmask; bool bmask = false;
  std::vector<double> u, v, w; // model layout (i,j,k), i fastest (nhydro.f90:55-59)
};

struct World {
  Params P;
  int npx = 1, npy = 1, nranks = 1, nlevs = 1;
  double hlim = 0, theta_b = 0, theta_s = 0;
  std::vector<Rank> R;
  std::vector<double> hist_res, hist_conv;
  double bnorm = 0, solve_seconds = 0; int nite = 0;
  int relax_order = 0;         // 0 = reference sequential sweep order, 1 = rank-1 corrected parallel form
  std::string err;
};

static A3 mk3(Store& s, int nzl, int ny, int nx, int nh) { A3 a; a.nzl = nzl; a.ny = ny; a.nx = nx; a.nh = nh; a.d = s.get(a.size()); return a; }
static A2 mk2(Store& s, int ny, int nx) { A2 a; a.ny = ny; a.nx = nx; a.d = s.get(a.size()); return a; }

// mg_grids.f90:468-498  find_grid_levels
static int find_grid_levels(int npxg, int npyg, int nx, int ny, int nz) {
  int nxg = npxg * nx, nyg = npyg * ny, nzg = nz;
  int ncoarsest = 4, nzmin = 2;
  int nhoriz = std::min(nxg, nyg);
  int nl1 = 1 + (int)std::floor(std::log(nhoriz * 1.0 / ncoarsest * 1.0) / std::log(2.0));
  int nl2 = 1 + (int)std::floor(std::log(nzg * 1.0 / nzmin * 1.0) / std::log(2.0));
  return std::min(nl1, nl2);
}

// mg_grids.f90:503-577 define_grid_dims ; :580-661 define_neighbours ; :664-738 gather info
static int define_grids(World& W, int nxl, int nyl, int nzl) {
  const Params& P = W.P;
  W.nlevs = find_grid_levels(W.npx, W.npy, nxl, nyl, nzl);
  if (W.nlevs < 1) { W.err = "grid too small: nlevs < 1"; return 1; }
  if (P.aggressive) { W.err = "aggressive coarsening is unfinished in the reference (mg_intergrids.f90:243)"; return 1; }
  for (int r = 0; r < W.nranks; ++r) {
    Rank& K = W.R[r];
    K.rank = r; K.pj = r / W.npx; K.pi = r % W.npx;          // mg_grids.f90:593-594
    K.grid.assign(W.nlevs + 1, Grid());
    int nx = nxl, ny = nyl, nz = nzl, npx = W.npx, npy = W.npy, incx = 1, incy = 1;
    for (int lev = 1; lev <= W.nlevs; ++lev) {
      Grid& g = K.grid[lev];
      if (lev > 1) {
        if (nz == 1) { nx /= 2; ny /= 2; } else { nx /= 2; ny /= 2; nz /= 2; }
        g.gather = 0; g.ngx = 1; g.ngy = 1;
        if (std::min(nx, ny) < P.nsmall && npx * npy > 1) {
          g.gather = 1;
          if (npx > 1) { npx /= 2; nx *= 2; g.ngx = 2; }
          if (npy > 1) { npy /= 2; ny *= 2; g.ngy = 2; }
          incx *= 2; incy *= 2;
        }
      }
      g.nx = nx; g.ny = ny; g.nz = nz; g.npx = npx; g.npy = npy; g.incx = incx; g.incy = incy;
      if (nx < 1 || ny < 1 || nz < 2) { W.err = "level with nz<2 or empty: unsupported (reference 2-D path is dead code)"; return 1; }
      // neighbours
      int pi = K.pi, pj = K.pj, NPX = W.npx, NPY = W.npy;
      g.neighb[0] = (pj >= incy) ? (pj - incy) * NPX + pi : -1;                                  // south
      g.neighb[1] = (pi < NPX - incx) ? pj * NPX + pi + incx : -1;                               // east
      g.neighb[2] = (pj < NPY - incy) ? (pj + incy) * NPX + pi : -1;                             // north
      g.neighb[3] = (pi >= incx) ? pj * NPX + pi - incx : -1;                                    // west
      g.neighb[4] = (pj >= incy && pi >= incx) ? (pj - incy) * NPX + pi - incx : -1;             // south west
      g.neighb[5] = (pj >= incy && pi < NPX - incx) ? (pj - incy) * NPX + pi + incx : -1;        // south east
      g.neighb[6] = (pj < NPY - incy && pi < NPX - incx) ? (pj + incy) * NPX + pi + incx : -1;   // north east
      g.neighb[7] = (pj < NPY - incy && pi >= incx) ? (pj + incy) * NPX + pi - incx : -1;        // north west
      // gather colour / key  (mg_grids.f90:681-718)
      if (g.gather == 1) {
        int ix = g.incx / 2, iy = g.incy / 2;
        int family = (pi / ix) * ix * iy + NPX * iy * (pj / iy);
        int nextfamily = (pi / (2 * ix)) * ix * iy * 4 + NPX * 2 * iy * (pj / (iy * 2));
        g.color = nextfamily + (pi % ix) + (pj % iy) * ix;
        int N = ix * NPX;
        g.key = ((family % N) / (ix * iy)) % 2 + 2 * ((family / N) % 2);
      }
      // allocations (mg_grids.f90:179-221)
      Store& S = K.mem;
      g.dx = mk2(S, ny, nx); g.dy = mk2(S, ny, nx); g.zeta = mk2(S, ny, nx); g.h = mk2(S, ny, nx);
      g.zr = mk3(S, nz, ny, nx, 2); g.zw = mk3(S, nz + 1, ny, nx, 2); g.cw = mk3(S, nz + 1, ny, nx, 1);
      g.p = mk3(S, nz, ny, nx, 1); g.b = mk3(S, nz, ny, nx, 1); g.r = mk3(S, nz, ny, nx, 1);
      g.cA.nz = nz; g.cA.ny = ny; g.cA.nx = nx; g.cA.d = S.get(g.cA.size());
      if (lev == 1) {
        g.dzw = mk3(S, nz + 1, ny, nx, 1); g.zxdy = mk3(S, nz, ny, nx, 1); g.zydx = mk3(S, nz, ny, nx, 1);
        g.dummy3Dnz = mk3(S, nz, ny, nx, 1); g.dummy3Dnzp = mk3(S, nz + 1, ny, nx, 1);
      }
      if (g.gather == 1) g.dummy3 = mk3(S, nz, ny / g.ngy, nx / g.ngx, 1);
    }
    K.rmask = mk2(K.mem, nyl, nxl);
    K.u.assign((size_t)(nxl + 1) * (nyl + 2) * nzl, 0.0);
    K.v.assign((size_t)(nxl + 2) * (nyl + 1) * nzl, 0.0);
    K.w.assign((size_t)(nxl + 2) * (nyl + 2) * (nzl + 1), 0.0);
  }
  // communicators of MPI_COMM_SPLIT(color,key): members with equal colour, ordered by key
  for (int lev = 2; lev <= W.nlevs; ++lev) {
    if (!W.R[0].grid[lev].gather) continue;
    for (int r = 0; r < W.nranks; ++r) {
      Grid& g = W.R[r].grid[lev];
      std::vector<std::pair<int, int>> m;
      for (int q = 0; q < W.nranks; ++q) if (W.R[q].grid[lev].color == g.color) m.push_back({W.R[q].grid[lev].key, q});
      std::sort(m.begin(), m.end());
      g.members.clear(); for (auto& kv : m) g.members.push_back(kv.second);
      if ((int)g.members.size() != g.ngx * g.ngy) { W.err = "gather communicator size mismatch"; return 1; }
    }
  }
  return 0;
}

#define FOR_RANKS(W) _Pragma("omp parallel for schedule(static)") for (int rk = 0; rk < (W).nranks; ++rk)

// ---------------------------------------------------------------------------------------------
// halo exchange  (mg_mpi_exchange.f90:23-352 2D, :396-745 3D_relax, :750-1242 3D incl. nh=2, lbc_null)
// Three bulk phases: (1) physical edges / corners / lbc zeroing, local; (2) copy neighbour
// interiors into my halo (what Isend/Irecv + unpack do); (3) "flag" corners from just-received halos.
// ---------------------------------------------------------------------------------------------
typedef A3 (*FieldSel)(Grid&);

static void halo_phase1(const Grid& g, const A3& p, char lbc) {
  const int nx = g.nx, ny = g.ny, nh = p.nh, nzl = p.nzl;
  const int south = g.neighb[0], east = g.neighb[1], north = g.neighb[2], west = g.neighb[3];
  const int sw = g.neighb[4], se = g.neighb[5], ne = g.neighb[6], nw = g.neighb[7];
  if (south < 0) {
    if (lbc == 'v') { for (int i = 0; i <= nx + 1; ++i) for (int k = 1; k <= nzl; ++k) p(k, 1, i) = 0.0; }
    else for (int i = 1; i <= nx; ++i) for (int k = 1; k <= nzl; ++k) {
      p(k, 0, i) = p(k, 1, i);
      if (nh == 2) p(k, -1, i) = 2.0 * p(k, 1, i) - p(k, 2, i);
    }
  }
  if (east < 0) {
    if (lbc == 'u') { for (int j = 0; j <= ny + 1; ++j) for (int k = 1; k <= nzl; ++k) p(k, j, nx + 1) = 0.0; }
    else for (int j = 1; j <= ny; ++j) for (int k = 1; k <= nzl; ++k) {
      p(k, j, nx + 1) = p(k, j, nx);
      if (nh == 2) p(k, j, nx + 2) = 2.0 * p(k, j, nx) - p(k, j, nx - 1);
    }
  }
  if (north < 0) {
    if (lbc == 'v') { for (int i = 0; i <= nx + 1; ++i) for (int k = 1; k <= nzl; ++k) p(k, ny + 1, i) = 0.0; }
    else for (int i = 1; i <= nx; ++i) for (int k = 1; k <= nzl; ++k) {
      p(k, ny + 1, i) = p(k, ny, i);
      if (nh == 2) p(k, ny + 2, i) = 2.0 * p(k, ny, i) - p(k, ny - 1, i);
    }
  }
  if (west < 0) {
    if (lbc == 'u') { for (int j = 0; j <= ny + 1; ++j) for (int k = 1; k <= nzl; ++k) p(k, j, 1) = 0.0; }
    else for (int j = 1; j <= ny; ++j) for (int k = 1; k <= nzl; ++k) {
      p(k, j, 0) = p(k, j, 1);
      if (nh == 2) p(k, j, -1) = 2.0 * p(k, j, 1) - p(k, j, 2);
    }
  }
  // corners (only the branches that act before the exchange)
  auto zero_c = [&](int j0, int i0) { for (int a = 0; a < nh; ++a) for (int c = 0; c < nh; ++c) for (int k = 1; k <= nzl; ++k) p(k, j0 + a, i0 + c) = 0.0; };
  // SW
  if (sw >= 0) {} else if (lbc == 'u' && west < 0) zero_c(1 - nh, 1 - nh);
  else if (south >= 0) {} else if (west >= 0) {}
  else for (int a = 1; a <= nh; ++a) for (int c = 1; c <= nh; ++c) for (int k = 1; k <= nzl; ++k) p(k, 1 - a, 1 - c) = p(k, a, c);
  // SE
  if (se >= 0) {} else if (lbc == 'u' && east < 0) zero_c(1 - nh, nx + 1);
  else if (south >= 0) {} else if (east >= 0) {}
  else for (int a = 1; a <= nh; ++a) for (int c = 1; c <= nh; ++c) for (int k = 1; k <= nzl; ++k) p(k, 1 - a, nx + c) = p(k, a, nx + 1 - c);
  // NE
  if (ne >= 0) {} else if ((lbc == 'u' && east < 0) || lbc == 'v') zero_c(ny + 1, nx + 1);
  else if (north >= 0) {} else if (east >= 0) {}
  else for (int a = 1; a <= nh; ++a) for (int c = 1; c <= nh; ++c) for (int k = 1; k <= nzl; ++k) p(k, ny + a, nx + c) = p(k, ny + 1 - a, nx + 1 - c);
  // NW
  if (nw >= 0) {} else if ((lbc == 'u' && west < 0) || lbc == 'v') zero_c(ny + 1, 1 - nh);
  else if (north >= 0) {} else if (west >= 0) {}
  else for (int a = 1; a <= nh; ++a) for (int c = 1; c <= nh; ++c) for (int k = 1; k <= nzl; ++k) p(k, ny + a, 1 - c) = p(k, ny + 1 - a, c);
}

static void halo_phase2(const Grid& g, const A3& p, World& W, int lev, FieldSel sel) {
  const int nx = g.nx, ny = g.ny, nh = p.nh, nzl = p.nzl;
  auto nb = [&](int q) -> A3 { return sel(W.R[q].grid[lev]); };
  if (g.neighb[0] >= 0) { A3 q = nb(g.neighb[0]);   // south: their north rows ny-nh+1..ny -> my 1-nh..0
    for (int i = 1; i <= nx; ++i) for (int a = 0; a < nh; ++a) for (int k = 1; k <= nzl; ++k) p(k, 1 - nh + a, i) = q(k, ny - nh + 1 + a, i); }
  if (g.neighb[1] >= 0) { A3 q = nb(g.neighb[1]);   // east: their west cols 1..nh -> my nx+1..nx+nh
    for (int c = 0; c < nh; ++c) for (int j = 1; j <= ny; ++j) for (int k = 1; k <= nzl; ++k) p(k, j, nx + 1 + c) = q(k, j, 1 + c); }
  if (g.neighb[2] >= 0) { A3 q = nb(g.neighb[2]);   // north: their south rows 1..nh -> my ny+1..ny+nh
    for (int i = 1; i <= nx; ++i) for (int a = 0; a < nh; ++a) for (int k = 1; k <= nzl; ++k) p(k, ny + 1 + a, i) = q(k, 1 + a, i); }
  if (g.neighb[3] >= 0) { A3 q = nb(g.neighb[3]);   // west: their east cols nx-nh+1..nx -> my 1-nh..0
    for (int c = 0; c < nh; ++c) for (int j = 1; j <= ny; ++j) for (int k = 1; k <= nzl; ++k) p(k, j, 1 - nh + c) = q(k, j, nx - nh + 1 + c); }
  if (g.neighb[4] >= 0) { A3 q = nb(g.neighb[4]);   // SW: their NE block
    for (int c = 0; c < nh; ++c) for (int a = 0; a < nh; ++a) for (int k = 1; k <= nzl; ++k) p(k, 1 - nh + a, 1 - nh + c) = q(k, ny - nh + 1 + a, nx - nh + 1 + c); }
  if (g.neighb[5] >= 0) { A3 q = nb(g.neighb[5]);   // SE: their NW block
    for (int c = 0; c < nh; ++c) for (int a = 0; a < nh; ++a) for (int k = 1; k <= nzl; ++k) p(k, 1 - nh + a, nx + 1 + c) = q(k, ny - nh + 1 + a, 1 + c); }
  if (g.neighb[6] >= 0) { A3 q = nb(g.neighb[6]);   // NE: their SW block
    for (int c = 0; c < nh; ++c) for (int a = 0; a < nh; ++a) for (int k = 1; k <= nzl; ++k) p(k, ny + 1 + a, nx + 1 + c) = q(k, 1 + a, 1 + c); }
  if (g.neighb[7] >= 0) { A3 q = nb(g.neighb[7]);   // NW: their SE block
    for (int c = 0; c < nh; ++c) for (int a = 0; a < nh; ++a) for (int k = 1; k <= nzl; ++k) p(k, ny + 1 + a, 1 - nh + c) = q(k, 1 + a, nx - nh + 1 + c); }
}

static void halo_phase3(const Grid& g, const A3& p, char lbc) {
  const int nx = g.nx, ny = g.ny, nh = p.nh, nzl = p.nzl;
  const int south = g.neighb[0], east = g.neighb[1], north = g.neighb[2], west = g.neighb[3];
  const int sw = g.neighb[4], se = g.neighb[5], ne = g.neighb[6], nw = g.neighb[7];
  // flags exactly as mg_mpi_exchange.f90:1026-1101 (lbc zero branches take precedence over flags)
  bool f_sw_s = false, f_sw_w = false, f_se_s = false, f_se_e = false, f_ne_n = false, f_ne_e = false, f_nw_n = false, f_nw_w = false;
  if (sw >= 0) {} else if (lbc == 'u' && west < 0) {} else if (south >= 0) f_sw_s = true; else if (west >= 0) f_sw_w = true;
  if (se >= 0) {} else if (lbc == 'u' && east < 0) {} else if (south >= 0) f_se_s = true; else if (east >= 0) f_se_e = true;
  if (ne >= 0) {} else if ((lbc == 'u' && east < 0) || lbc == 'v') {} else if (north >= 0) f_ne_n = true; else if (east >= 0) f_ne_e = true;
  if (nw >= 0) {} else if ((lbc == 'u' && west < 0) || lbc == 'v') {} else if (north >= 0) f_nw_n = true; else if (west >= 0) f_nw_w = true;
  for (int k = 1; k <= nzl; ++k) for (int a = 1; a <= nh; ++a) for (int c = 1; c <= nh; ++c) {
    // a indexes halo row distance / c halo column distance (1..nh)
    if (f_sw_s) p(k, 1 - a, 1 - c) = p(k, 1 - a, c); else if (f_sw_w) p(k, 1 - a, 1 - c) = p(k, a, 1 - c);
    if (f_se_s) p(k, 1 - a, nx + c) = p(k, 1 - a, nx + 1 - c); else if (f_se_e) p(k, 1 - a, nx + c) = p(k, a, nx + c);
    if (f_ne_n) p(k, ny + a, nx + c) = p(k, ny + a, nx + 1 - c); else if (f_ne_e) p(k, ny + a, nx + c) = p(k, ny + 1 - a, nx + c);
    if (f_nw_n) p(k, ny + a, 1 - c) = p(k, ny + a, c); else if (f_nw_w) p(k, ny + a, 1 - c) = p(k, ny + 1 - a, 1 - c);
  }
}

static void fill_halo(World& W, int lev, FieldSel sel, char lbc = 0) {
  FOR_RANKS(W) { Grid& g = W.R[rk].grid[lev]; halo_phase1(g, sel(g), lbc); }
  if (W.nranks > 1) {
    FOR_RANKS(W) { Grid& g = W.R[rk].grid[lev]; halo_phase2(g, sel(g), W, lev, sel); }
    FOR_RANKS(W) { Grid& g = W.R[rk].grid[lev]; halo_phase3(g, sel(g), lbc); }
  }
}
// 2-D fields ride through the same code as 1-level 3-D views (mg_mpi_exchange.f90:23-352)
static A3 as3(const A2& a) { A3 v; v.d = a.d; v.nzl = 1; v.ny = a.ny; v.nx = a.nx; v.nh = 1; return v; }
static A3 sel_p(Grid& g) { return g.p; }
static A3 sel_b(Grid& g) { return g.b; }
static A3 sel_r(Grid& g) { return g.r; }
static A3 sel_zr(Grid& g) { return g.zr; }
static A3 sel_zw(Grid& g) { return g.zw; }
static A3 sel_dx(Grid& g) { return as3(g.dx); }
static A3 sel_dy(Grid& g) { return as3(g.dy); }
static A3 sel_zeta(Grid& g) { return as3(g.zeta); }
static A3 sel_h(Grid& g) { return as3(g.h); }
static A3 sel_d3nz(Grid& g) { return g.dummy3Dnz; }
static A3 sel_cA(Grid& g) { A3 v; v.d = g.cA.d; v.nzl = 8 * g.nz; v.ny = g.ny; v.nx = g.nx; v.nh = 1; return v; }   // cA(8,nz,j,i) as an (8 nz)-level field

// mg_mpi_exchange.f90:1555-1571 global_sum
static double global_sum(World& W, int lev, const std::vector<double>& loc) {
  double s = 0.0; for (int r = 0; r < W.nranks; ++r) s += loc[r];   // MPI_SUM in rank order
  const Grid& g = W.R[0].grid[lev]; const Grid& g1 = W.R[0].grid[1];
  return s * (g.npx * g.npy) / (g1.npx * g1.npy);
}

// ---------------------------------------------------------------------------------------------
// gather / split  (mg_gather.f90:18-92, 95-174, 177-220)
// ---------------------------------------------------------------------------------------------
typedef A3 (*SubSel)(Grid&);
static void gather3(World& W, int lev, SubSel xsel, FieldSel ysel) {
  FOR_RANKS(W) {
    Grid& g = W.R[rk].grid[lev]; A3 y = ysel(g);
    int nx = g.nx / g.ngx, ny = g.ny / g.ngy, nzl = y.nzl;
    for (int m = 0; m < g.ngy; ++m) for (int l = 0; l < g.ngx; ++l) {
      int src = g.members[l + g.ngx * m];
      A3 x = xsel(W.R[src].grid[lev]);
      for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny; ++j) for (int k = 1; k <= nzl; ++k)
        y(k, j + m * ny, i + l * nx) = x(k, j, i);
    }
  }
}
static void split3(const Grid& g, const A3& x, const A3& y) {
  int nx = g.nx / g.ngx, ny = g.ny / g.ngy, nz = g.nz;
  int l = g.key % 2, m = g.key / 2;
  for (int i = 0; i <= nx + 1; ++i) for (int j = 0; j <= ny + 1; ++j) for (int k = 1; k <= nz; ++k)
    y(k, j, i) = x(k, j + m * ny, i + l * nx);
}
static A3 sel_dummy3(Grid& g) { return g.dummy3; }

// ---------------------------------------------------------------------------------------------
// s-coordinate  (mg_zr_zw.f90:98-170, coord_type 'new_s_coord')
// ---------------------------------------------------------------------------------------------
static void setup_zr_zw(const World& W, Grid& g) {
  const int nx = g.nx, ny = g.ny, nz = g.nz;
  const double hlim = W.hlim, theta_b = W.theta_b, theta_s = W.theta_s;
  for (int i = 0; i <= nx + 1; ++i) for (int j = 0; j <= ny + 1; ++j) {
    double cff = 1.0 / (double)nz;
    double h = g.h(j, i), zeta = g.zeta(j, i);
    double hinv = 1.0 / (h + hlim);
    for (int k = 1; k <= nz + 1; ++k) {
      double sc_w = cff * (double)(k - 1 - nz);
      double cswf = (theta_s > 0.0) ? (1.0 - std::cosh(theta_s * sc_w)) / (std::cosh(theta_s) - 1.0) : -(sc_w * sc_w);
      double cs_w = (theta_b > 0.0) ? (std::exp(theta_b * cswf) - 1.0) / (1.0 - std::exp(-theta_b)) : cswf;
      double cff_w = hlim * sc_w;
      double z_w0 = cff_w + cs_w * h;
      g.zw(k, j, i) = z_w0 * h * hinv + zeta * (1.0 + z_w0 * hinv);
      if (k <= nz) {
        double sc_r = cff * ((double)(k - nz) - 0.5);
        double csrf = (theta_s > 0.0) ? (1.0 - std::cosh(theta_s * sc_r)) / (std::cosh(theta_s) - 1.0) : -(sc_r * sc_r);
        double cs_r = (theta_b > 0.0) ? (std::exp(theta_b * csrf) - 1.0) / (1.0 - std::exp(-theta_b)) : csrf;
        double cff_r = hlim * sc_r;
        double z_r0 = cff_r + cs_r * h;
        g.zr(k, j, i) = z_r0 * h * hinv + zeta * (1.0 + z_r0 * hinv);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// define_matrix  (mg_define_matrix.f90:211-680); umask/vmask = 1 unless bmask (:255-275)
// ---------------------------------------------------------------------------------------------
// rm: the level's rmask (0:ny+1,0:nx+1) -- the caller's at level 1, ones with a zeroed physical halo below (:155-161)
static void define_matrix_coef(Grid& g, int lev, const A2& rm, bool bmask) {
  const int nx = g.nx, ny = g.ny, nz = g.nz;
  auto UM = [&](int j, int i) -> double { if (!bmask) return 1.0; return (i >= 1 && i <= nx + 1 && j >= 0 && j <= ny + 1) ? rm(j, i - 1) * rm(j, i) : 0.0; };
  auto VM = [&](int j, int i) -> double { if (!bmask) return 1.0; return (i >= 0 && i <= nx + 1 && j >= 1 && j <= ny + 1) ? rm(j - 1, i) * rm(j, i) : 0.0; };
  const A2 &dx = g.dx, &dy = g.dy; const A3 &zr = g.zr, &zw = g.zw, &cw = g.cw; const A4& cA = g.cA;
  const double hlf = 0.5, qrt = 0.25, one = 1.0;
  auto ZYDX = [&](int k, int j, int i) { return ((hlf * (zr(k, j + 1, i) - zr(k, j - 1, i))) / dy(j, i)) * dx(j, i); };
  auto ZXDY = [&](int k, int j, int i) { return ((hlf * (zr(k, j, i + 1) - zr(k, j, i - 1))) / dx(j, i)) * dy(j, i); };
  auto ARY = [&](int k, int j, int i) { return (qrt * (zw(k + 1, j, i) - zw(k, j, i) + zw(k + 1, j - 1, i) - zw(k, j - 1, i))) * (dx(j, i) + dx(j - 1, i)); };
  auto ARX = [&](int k, int j, int i) { return (qrt * (zw(k + 1, j, i) - zw(k, j, i) + zw(k + 1, j, i - 1) - zw(k, j, i - 1))) * (dy(j, i) + dy(j, i - 1)); };
  auto SQ = [](double x) { return x * x; };

  if (lev == 1) {                                            // :283-306
    for (int i = 0; i <= nx + 1; ++i) for (int j = 0; j <= ny + 1; ++j) {
      g.dzw(1, j, i) = zr(1, j, i) - zw(1, j, i);
      for (int k = 2; k <= nz; ++k) g.dzw(k, j, i) = zr(k, j, i) - zr(k - 1, j, i);
      g.dzw(nz + 1, j, i) = zw(nz + 1, j, i) - zr(nz, j, i);
      for (int k = 1; k <= nz; ++k) {
        g.zydx(k, j, i) = hlf * ((zr(k, j + 1, i) - zr(k, j - 1, i)) / dy(j, i)) * dx(j, i);
        g.zxdy(k, j, i) = hlf * ((zr(k, j, i + 1) - zr(k, j, i - 1)) / dx(j, i)) * dy(j, i);
      }
    }
  }
  for (int i = 0; i <= nx + 1; ++i) for (int j = 0; j <= ny + 1; ++j) {   // cw :309-336
    double Arz = dx(j, i) * dy(j, i);
    for (int k = 1; k <= nz + 1; ++k) {
      double dz = (k == 1) ? (zr(1, j, i) - zw(1, j, i)) : (k <= nz ? (zr(k, j, i) - zr(k - 1, j, i)) : (zw(k, j, i) - zr(k - 1, j, i)));
      cw(k, j, i) = (Arz / dz) * (one + SQ(hlf * (zw(k, j, i + 1) - zw(k, j, i - 1)) / dx(j, i)) + SQ(hlf * (zw(k, j + 1, i) - zw(k, j - 1, i)) / dy(j, i)));
    }
  }
  // ---- k = 1 (:352-485)
  { const int k = 1;
    for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny + 1; ++j) {
      cA(3, k, j, i) = qrt * (ZYDX(k + 1, j, i) + ZYDX(k, j - 1, i)) * VM(j, i);
      cA(4, k, j, i) = ARY(k, j, i) / (hlf * (dy(j, i) + dy(j - 1, i)))
                       - (SQ(ZYDX(k, j, i)) / (cw(k, j, i) + cw(k + 1, j, i)) + SQ(ZYDX(k, j - 1, i)) / (cw(k, j - 1, i) + cw(k + 1, j - 1, i)))
                       - qrt * (ZYDX(k, j - 1, i) - ZYDX(k, j, i));
      if (bmask)   // :375-389, i-j cross terms at a lateral wall; the second slope is written with /dy*dx in the reference (kept)
        cA(4, k, j, i) = (cA(4, k, j, i)
                          - (hlf * ZXDY(k, j - 1, i) * ZYDX(k, j - 1, i) / (cw(k, j - 1, i) + cw(k + 1, j - 1, i)) * (UM(j - 1, i + 1) - UM(j - 1, i))
                             - hlf * (((hlf * (zr(k, j, i + 1) - zr(k, j, i - 1))) / dy(j, i)) * dx(j, i)) * ZYDX(k, j, i) / (cw(k, j, i) + cw(k + 1, j, i)) * (UM(j, i + 1) - UM(j, i)))) * VM(j, i);
    }
    for (int i = 1; i <= nx + 1; ++i) for (int j = 1; j <= ny; ++j) {
      cA(6, k, j, i) = qrt * (ZXDY(k + 1, j, i) + ZXDY(k, j, i - 1)) * UM(j, i);
      cA(7, k, j, i) = ARX(k, j, i) / (hlf * (dx(j, i) + dx(j, i - 1)))
                       - (SQ(ZXDY(k, j, i)) / (cw(k, j, i) + cw(k + 1, j, i)) + SQ(ZXDY(k, j, i - 1)) / (cw(k, j, i - 1) + cw(k + 1, j, i - 1)))
                       - qrt * (ZXDY(k, j, i - 1) - ZXDY(k, j, i));
      if (bmask)   // :417-433
        cA(7, k, j, i) = (cA(7, k, j, i)
                          - (hlf * ZXDY(k, j, i - 1) * ZYDX(k, j, i - 1) / (cw(k, j, i - 1) + cw(k + 1, j, i - 1)) * (VM(j + 1, i - 1) - VM(j, i - 1))
                             - hlf * (((hlf * (zr(k, j, i + 1) - zr(k, j, i - 1))) / dy(j, i)) * dx(j, i)) * ZYDX(k, j, i) / (cw(k, j, i) + cw(k + 1, j, i)) * (VM(j + 1, i) - VM(j, i)))) * UM(j, i);
    }
    for (int i = 1; i <= nx + 1; ++i) for (int j = 0; j <= ny; ++j)       // couples (1,j,i) <-> (1,j+1,i-1)
      cA(5, k, j, i) = hlf * ZXDY(k, j + 1, i) * ZYDX(k, j + 1, i) / (cw(k, j + 1, i) + cw(k + 1, j + 1, i)) * UM(j + 1, i) * VM(j + 1, i)
                     + hlf * ZXDY(k, j, i - 1) * ZYDX(k, j, i - 1) / (cw(k, j, i - 1) + cw(k + 1, j, i - 1)) * UM(j, i) * VM(j + 1, i - 1);
    for (int i = 1; i <= nx + 1; ++i) for (int j = 1; j <= ny + 1; ++j)   // couples (1,j,i) <-> (1,j-1,i-1)
      cA(8, k, j, i) = -(hlf * ZXDY(k, j - 1, i) * ZYDX(k, j - 1, i) / (cw(k, j - 1, i) + cw(k + 1, j - 1, i)) * UM(j - 1, i) * VM(j, i))
                       - hlf * ZXDY(k, j, i - 1) * ZYDX(k, j, i - 1) / (cw(k, j, i - 1) + cw(k + 1, j, i - 1)) * UM(j, i) * VM(j, i - 1);
  }
  // ---- k = 2 .. nz-1 (:491-560)
  for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny; ++j) for (int k = 2; k <= nz - 1; ++k) {
    cA(2, k, j, i) = cw(k, j, i);
    if (bmask)     // :497-509, i-k and j-k cross terms at a lateral wall
      cA(2, k, j, i) = cA(2, k, j, i) - qrt * (ZXDY(k - 1, j, i) - ZXDY(k, j, i)) * (UM(j, i + 1) - UM(j, i))
                                      - qrt * (ZYDX(k - 1, j, i) - ZYDX(k, j, i)) * (VM(j + 1, i) - VM(j, i));
  }
  for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny + 1; ++j) for (int k = 2; k <= nz - 1; ++k) {
    cA(3, k, j, i) = qrt * (ZYDX(k + 1, j, i) + ZYDX(k, j - 1, i)) * VM(j, i);
    cA(4, k, j, i) = ARY(k, j, i) / (hlf * (dy(j, i) + dy(j - 1, i))) * VM(j, i);
    cA(5, k, j, i) = -qrt * (ZYDX(k - 1, j, i) + ZYDX(k, j - 1, i)) * VM(j, i);
  }
  for (int i = 1; i <= nx + 1; ++i) for (int j = 1; j <= ny; ++j) for (int k = 2; k <= nz - 1; ++k) {
    cA(6, k, j, i) = qrt * (ZXDY(k + 1, j, i) + ZXDY(k, j, i - 1)) * UM(j, i);
    cA(7, k, j, i) = ARX(k, j, i) / (hlf * (dx(j, i) + dx(j, i - 1))) * UM(j, i);
    cA(8, k, j, i) = -qrt * (ZXDY(k - 1, j, i) + ZXDY(k, j, i - 1)) * UM(j, i);
  }
  // ---- k = nz (:566-609)
  { const int k = nz;
    for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny; ++j) cA(2, k, j, i) = cw(k, j, i);
    for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny + 1; ++j) {
      cA(4, k, j, i) = (ARY(k, j, i) / (hlf * (dy(j, i) + dy(j - 1, i))) + qrt * (-ZYDX(k, j - 1, i) + ZYDX(k, j, i))) * VM(j, i);
      cA(5, k, j, i) = -qrt * (ZYDX(k - 1, j, i) + ZYDX(k, j - 1, i)) * VM(j, i);
    }
    for (int i = 1; i <= nx + 1; ++i) for (int j = 1; j <= ny; ++j) {
      cA(7, k, j, i) = (ARX(k, j, i) / (hlf * (dx(j, i) + dx(j, i - 1))) + qrt * (-ZXDY(k, j, i - 1) + ZXDY(k, j, i))) * UM(j, i);
      cA(8, k, j, i) = -qrt * (ZXDY(k - 1, j, i) + ZXDY(k, j, i - 1)) * UM(j, i);
    }
  }
}
// the diagonal (:616-657), after the cA halo exchange that bmask adds (:611-613)
static void define_matrix_diag(Grid& g) {
  const int nx = g.nx, ny = g.ny, nz = g.nz;
  const A2 &dx = g.dx, &dy = g.dy; const A3 &zr = g.zr, &cw = g.cw; const A4& cA = g.cA;
  const double hlf = 0.5;
  // ---- diagonal (:616-657)
  for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny; ++j) {
    int k = 1;
    cA(1, k, j, i) = -cA(2, k + 1, j, i) - cA(4, k, j, i) - cA(4, k, j + 1, i) - cA(7, k, j, i) - cA(7, k, j, i + 1)
                     - cA(6, k, j, i) - cA(8, k + 1, j, i + 1) - cA(3, k, j, i) - cA(5, k + 1, j + 1, i)
                     - cA(5, k, j, i) - cA(5, k, j - 1, i + 1) - cA(8, k, j, i) - cA(8, k, j + 1, i + 1);
    for (k = 2; k <= nz - 1; ++k)
      cA(1, k, j, i) = -cA(2, k, j, i) - cA(2, k + 1, j, i) - cA(4, k, j, i) - cA(4, k, j + 1, i) - cA(7, k, j, i) - cA(7, k, j, i + 1)
                       - cA(6, k, j, i) - cA(6, k - 1, j, i + 1) - cA(8, k, j, i) - cA(8, k + 1, j, i + 1)
                       - cA(3, k, j, i) - cA(3, k - 1, j + 1, i) - cA(5, k, j, i) - cA(5, k + 1, j + 1, i);
    k = nz;
    cA(1, k, j, i) = -cA(2, k, j, i) - cw(k + 1, j, i)
                     + hlf * (hlf * (zr(k, j, i + 2) - zr(k, j, i)) / dx(j, i + 1)) * dy(j, i + 1)
                     - hlf * (hlf * (zr(k, j, i) - zr(k, j, i - 2)) / dx(j, i - 1)) * dy(j, i - 1)
                     + hlf * (hlf * (zr(k, j + 2, i) - zr(k, j, i)) / dy(j + 1, i)) * dx(j + 1, i)
                     - hlf * (hlf * (zr(k, j, i) - zr(k, j - 2, i)) / dy(j - 1, i)) * dx(j - 1, i)
                     - cA(4, k, j, i) - cA(4, k, j + 1, i) - cA(7, k, j, i) - cA(7, k, j, i + 1)
                     - cA(6, k - 1, j, i + 1) - cA(8, k, j, i) - cA(3, k - 1, j + 1, i) - cA(5, k, j, i);
  }
}

// mg_define_matrix.f90:28-208 define_matrices_topo
static void define_matrices(World& W) {
  for (int lev = 1; lev <= W.nlevs; ++lev) {
    if (lev > 1) {
      const bool gath = W.R[0].grid[lev].gather == 1;
      // coarsen (into the sub-block when gathering; we reuse dummy storage per rank)
      std::vector<std::vector<double>> tmp(gath ? W.nranks * 4 : 0);
      FOR_RANKS(W) {
        Grid& gc = W.R[rk].grid[lev]; Grid& gf = W.R[rk].grid[lev - 1];
        int nxc = gc.nx / gc.ngx, nyc = gc.ny / gc.ngy;
        A2 dxc = gc.dx, dyc = gc.dy, zc = gc.zeta, hc = gc.h;
        if (gath) {
          for (int q = 0; q < 4; ++q) tmp[rk * 4 + q].assign((size_t)(nyc + 2) * (nxc + 2), 0.0);
          dxc.d = tmp[rk * 4 + 0].data(); dyc.d = tmp[rk * 4 + 1].data(); zc.d = tmp[rk * 4 + 2].data(); hc.d = tmp[rk * 4 + 3].data();
          dxc.ny = dyc.ny = zc.ny = hc.ny = nyc; dxc.nx = dyc.nx = zc.nx = hc.nx = nxc;
        }
        for (int i = 1; i <= nxc; ++i) for (int j = 1; j <= nyc; ++j) {
          int jf = 2 * j - 1, If = 2 * i - 1;
          dxc(j, i) = 0.5 * (gf.dx(jf, If) + gf.dx(jf + 1, If) + gf.dx(jf, If + 1) + gf.dx(jf + 1, If + 1));
          dyc(j, i) = 0.5 * (gf.dy(jf, If) + gf.dy(jf + 1, If) + gf.dy(jf, If + 1) + gf.dy(jf + 1, If + 1));
          zc(j, i) = 0.25 * (gf.zeta(jf, If) + gf.zeta(jf + 1, If) + gf.zeta(jf, If + 1) + gf.zeta(jf + 1, If + 1));
          hc(j, i) = 0.25 * (gf.h(jf, If) + gf.h(jf + 1, If) + gf.h(jf, If + 1) + gf.h(jf + 1, If + 1));
        }
      }
      if (gath) {                                            // gather_2D x4 (mg_gather.f90:18-92)
        FOR_RANKS(W) {
          Grid& g = W.R[rk].grid[lev];
          int nxs = g.nx / g.ngx, nys = g.ny / g.ngy;
          A2 dst[4] = {g.dx, g.dy, g.zeta, g.h};
          for (int m = 0; m < g.ngy; ++m) for (int l = 0; l < g.ngx; ++l) {
            int src = g.members[l + g.ngx * m];
            for (int q = 0; q < 4; ++q) {
              A2 x; x.d = tmp[src * 4 + q].data(); x.ny = nys; x.nx = nxs;
              for (int i = 1; i <= nxs; ++i) for (int j = 1; j <= nys; ++j) dst[q](j + m * nys, i + l * nxs) = x(j, i);
            }
          }
        }
      }
    }
    fill_halo(W, lev, sel_dx); fill_halo(W, lev, sel_dy); fill_halo(W, lev, sel_zeta); fill_halo(W, lev, sel_h);
    FOR_RANKS(W) setup_zr_zw(W, W.R[rk].grid[lev]);
    fill_halo(W, lev, sel_zr); fill_halo(W, lev, sel_zw);
    const bool bmask = W.P.bmask;
    FOR_RANKS(W) {
      Rank& K = W.R[rk]; Grid& g = K.grid[lev];
      std::vector<double> ones;
      A2 rm = K.rmask;                                       // level 1: the caller's mask (:73-79)
      if (lev > 1) {                                         // :155-161 ones, physical halo zeroed by fill_halo_2D_bmask (mg_mpi_exchange.f90:357-391)
        ones.assign((size_t)(g.ny + 2) * (g.nx + 2), 1.0);
        rm.d = ones.data(); rm.ny = g.ny; rm.nx = g.nx;
        if (bmask) {
          if (g.neighb[0] < 0) for (int i = 0; i <= g.nx + 1; ++i) rm(0, i) = 0.0;
          if (g.neighb[1] < 0) for (int j = 0; j <= g.ny + 1; ++j) rm(j, g.nx + 1) = 0.0;
          if (g.neighb[2] < 0) for (int i = 0; i <= g.nx + 1; ++i) rm(g.ny + 1, i) = 0.0;
          if (g.neighb[3] < 0) for (int j = 0; j <= g.ny + 1; ++j) rm(j, 0) = 0.0;
        }
      }
      define_matrix_coef(g, lev, rm, bmask);
    }
    if (bmask && W.nranks > 1)                               // fill_halo_4D (mg_mpi_exchange.f90:1245-1530): plain exchange, no physical fill
      FOR_RANKS(W) { Grid& g = W.R[rk].grid[lev]; halo_phase2(g, sel_cA(g), W, lev, sel_cA); }
    FOR_RANKS(W) define_matrix_diag(W.R[rk].grid[lev]);
  }
}

// ---------------------------------------------------------------------------------------------
// compute_rhs (mg_compute_rhs.f90:14-379) and correct_uvw (mg_correct_uvw.f90:15-115); umask = vmask = 1 unless bmask
// ---------------------------------------------------------------------------------------------
struct UVW {
  double *u, *v, *w; int nx, ny, nz;
  inline double& U(int i, int j, int k) const { return u[(size_t)(i - 1) + (size_t)(nx + 1) * ((size_t)j + (size_t)(ny + 2) * (size_t)(k - 1))]; }
  inline double& V(int i, int j, int k) const { return v[(size_t)i + (size_t)(nx + 2) * ((size_t)(j - 1) + (size_t)(ny + 1) * (size_t)(k - 1))]; }
  inline double& Wv(int i, int j, int k) const { return w[(size_t)i + (size_t)(nx + 2) * ((size_t)j + (size_t)(ny + 2) * (size_t)k)]; }
};

static void rhs_uf(Rank& K) {
  Grid& g = K.grid[1]; const int nx = g.nx, ny = g.ny, nz = g.nz;
  UVW m{K.u.data(), K.v.data(), K.w.data(), nx, ny, nz};
  const A2& rm_ = K.rmask; const bool bm_ = K.bmask;      // umask / vmask (mg_compute_rhs.f90:53-72)
  auto UM = [&](int j, int i) -> double { if (!bm_) return 1.0; return (i >= 1 && i <= nx + 1 && j >= 0 && j <= ny + 1) ? rm_(j, i - 1) * rm_(j, i) : 0.0; };
  auto VM = [&](int j, int i) -> double { if (!bm_) return 1.0; return (i >= 0 && i <= nx + 1 && j >= 1 && j <= ny + 1) ? rm_(j - 1, i) * rm_(j, i) : 0.0; };
  const A2 &dx = g.dx, &dy = g.dy, &rmask = K.rmask; const A3 &zw = g.zw, &dzw = g.dzw, &zxdy = g.zxdy, &zydx = g.zydx, &cw = g.cw;
  A3 uf = g.dummy3Dnz; const double qrt = 0.25, hlf = 0.5, two = 2.0;
  std::fill(g.b.d, g.b.d + g.b.size(), 0.0);
  for (int i = 1; i <= nx + 1; ++i) for (int j = 1; j <= ny; ++j) {
    int k = 1;
    uf(k, j, i) = (qrt * (zw(k + 1, j, i) - zw(k, j, i) + zw(k + 1, j, i - 1) - zw(k, j, i - 1)) * (dy(j, i) + dy(j, i - 1)) * m.U(i, j, k)
      - qrt * (+zxdy(k, j, i) * dzw(k + 1, j, i) * m.Wv(i, j, k + 1 - 1) * rmask(j, i) + zxdy(k, j, i - 1) * dzw(k + 1, j, i - 1) * m.Wv(i - 1, j, k + 1 - 1) * rmask(j, i - 1))
      - (+zxdy(k, j, i) * zxdy(k, j, i) / (cw(k, j, i) + cw(k + 1, j, i)) + zxdy(k, j, i - 1) * zxdy(k, j, i - 1) / (cw(k, j, i - 1) + cw(k + 1, j, i - 1)))
          * (hlf * (dx(j, i) + dx(j, i - 1))) * m.U(i, j, k)
      - (+zxdy(k, j, i) * zydx(k, j, i) / (cw(k, j, i) + cw(k + 1, j, i))
            * hlf * (hlf * (dy(j, i) + dy(j - 1, i)) * m.V(i, j, k) * VM(j, i) + hlf * (dy(j + 1, i) + dy(j, i)) * m.V(i, j + 1, k) * VM(j + 1, i))
         + zxdy(k, j, i - 1) * zydx(k, j, i - 1) / (cw(k, j, i - 1) + cw(k + 1, j, i - 1))
            * hlf * (hlf * (dy(j, i - 1) + dy(j - 1, i - 1)) * m.V(i - 1, j, k) * VM(j, i - 1) + hlf * (dy(j + 1, i - 1) + dy(j, i - 1)) * m.V(i - 1, j + 1, k) * VM(j + 1, i - 1)))) * UM(j, i);
    for (k = 2; k <= nz - 1; ++k)
      uf(k, j, i) = (qrt * (zw(k + 1, j, i) - zw(k, j, i) + zw(k + 1, j, i - 1) - zw(k, j, i - 1)) * (dy(j, i) + dy(j, i - 1)) * m.U(i, j, k)
        - qrt * (+zxdy(k, j, i) * dzw(k, j, i) * m.Wv(i, j, k - 1) * rmask(j, i) + zxdy(k, j, i) * dzw(k + 1, j, i) * m.Wv(i, j, k + 1 - 1) * rmask(j, i)
                 + zxdy(k, j, i - 1) * dzw(k, j, i - 1) * m.Wv(i - 1, j, k - 1) * rmask(j, i - 1) + zxdy(k, j, i - 1) * dzw(k + 1, j, i - 1) * m.Wv(i - 1, j, k + 1 - 1) * rmask(j, i - 1))) * UM(j, i);
    k = nz;
    uf(k, j, i) = (qrt * (zw(k + 1, j, i) - zw(k, j, i) + zw(k + 1, j, i - 1) - zw(k, j, i - 1)) * (dy(j, i) + dy(j, i - 1)) * m.U(i, j, k)
      - qrt * (+zxdy(k, j, i) * dzw(k, j, i) * m.Wv(i, j, k - 1) * rmask(j, i) + zxdy(k, j, i) * two * dzw(k + 1, j, i) * m.Wv(i, j, k + 1 - 1) * rmask(j, i)
               + zxdy(k, j, i - 1) * dzw(k, j, i - 1) * m.Wv(i - 1, j, k - 1) * rmask(j, i - 1) + zxdy(k, j, i - 1) * two * dzw(k + 1, j, i - 1) * m.Wv(i - 1, j, k + 1 - 1) * rmask(j, i - 1))) * UM(j, i);
  }
}
static void rhs_uf_div(Rank& K) {
  Grid& g = K.grid[1]; A3 uf = g.dummy3Dnz;
  for (int i = 1; i <= g.nx; ++i) for (int j = 1; j <= g.ny; ++j) for (int k = 1; k <= g.nz; ++k) g.b(k, j, i) = uf(k, j, i + 1) - uf(k, j, i);
}
static void rhs_vf(Rank& K) {
  Grid& g = K.grid[1]; const int nx = g.nx, ny = g.ny, nz = g.nz;
  UVW m{K.u.data(), K.v.data(), K.w.data(), nx, ny, nz};
  const A2& rm_ = K.rmask; const bool bm_ = K.bmask;      // umask / vmask (mg_compute_rhs.f90:53-72)
  auto UM = [&](int j, int i) -> double { if (!bm_) return 1.0; return (i >= 1 && i <= nx + 1 && j >= 0 && j <= ny + 1) ? rm_(j, i - 1) * rm_(j, i) : 0.0; };
  auto VM = [&](int j, int i) -> double { if (!bm_) return 1.0; return (i >= 0 && i <= nx + 1 && j >= 1 && j <= ny + 1) ? rm_(j - 1, i) * rm_(j, i) : 0.0; };
  const A2 &dx = g.dx, &dy = g.dy, &rmask = K.rmask; const A3 &zw = g.zw, &dzw = g.dzw, &zxdy = g.zxdy, &zydx = g.zydx, &cw = g.cw;
  A3 vf = g.dummy3Dnz; const double qrt = 0.25, hlf = 0.5, two = 2.0;
  for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny + 1; ++j) {
    int k = 1;
    vf(k, j, i) = (qrt * (zw(k + 1, j, i) - zw(k, j, i) + zw(k + 1, j - 1, i) - zw(k, j - 1, i)) * (dx(j, i) + dx(j - 1, i)) * m.V(i, j, k)
      - qrt * (+zydx(k, j, i) * dzw(k + 1, j, i) * m.Wv(i, j, k + 1 - 1) * rmask(j, i) + zydx(k, j - 1, i) * dzw(k + 1, j - 1, i) * m.Wv(i, j - 1, k + 1 - 1) * rmask(j - 1, i))
      - (+zydx(k, j, i) * zydx(k, j, i) / (cw(k, j, i) + cw(k + 1, j, i)) + zydx(k, j - 1, i) * zydx(k, j - 1, i) / (cw(k, j - 1, i) + cw(k + 1, j - 1, i)))
          * hlf * (dy(j, i) + dy(j - 1, i)) * m.V(i, j, k)
      - (+zxdy(k, j, i) * zydx(k, j, i) / (cw(k, j, i) + cw(k + 1, j, i))
            * hlf * (hlf * (dx(j, i) + dx(j, i - 1)) * m.U(i, j, k) * UM(j, i) + hlf * (dx(j, i + 1) + dx(j, i)) * m.U(i + 1, j, k) * UM(j, i + 1))
         + zxdy(k, j - 1, i) * zydx(k, j - 1, i) / (cw(k, j - 1, i) + cw(k + 1, j - 1, i))
            * hlf * (hlf * (dx(j - 1, i) + dx(j - 1, i - 1)) * m.U(i, j - 1, k) * UM(j - 1, i) + hlf * (dx(j - 1, i + 1) + dx(j - 1, i)) * m.U(i + 1, j - 1, k) * UM(j - 1, i + 1)))) * VM(j, i);
    for (k = 2; k <= nz - 1; ++k)
      vf(k, j, i) = (qrt * (zw(k + 1, j, i) - zw(k, j, i) + zw(k + 1, j - 1, i) - zw(k, j - 1, i)) * (dx(j, i) + dx(j - 1, i)) * m.V(i, j, k)
        - qrt * (+zydx(k, j, i) * dzw(k, j, i) * m.Wv(i, j, k - 1) * rmask(j, i) + zydx(k, j, i) * dzw(k + 1, j, i) * m.Wv(i, j, k + 1 - 1) * rmask(j, i)
                 + zydx(k, j - 1, i) * dzw(k, j - 1, i) * m.Wv(i, j - 1, k - 1) * rmask(j - 1, i) + zydx(k, j - 1, i) * dzw(k + 1, j - 1, i) * m.Wv(i, j - 1, k + 1 - 1) * rmask(j - 1, i))) * VM(j, i);
    k = nz;
    vf(k, j, i) = (qrt * (zw(k + 1, j, i) - zw(k, j, i) + zw(k + 1, j - 1, i) - zw(k, j - 1, i)) * (dx(j, i) + dx(j - 1, i)) * m.V(i, j, k)
      - qrt * (+zydx(k, j, i) * dzw(k, j, i) * m.Wv(i, j, k - 1) * rmask(j, i) + zydx(k, j, i) * two * dzw(k + 1, j, i) * m.Wv(i, j, k + 1 - 1) * rmask(j, i)
               + zydx(k, j - 1, i) * dzw(k, j - 1, i) * m.Wv(i, j - 1, k - 1) * rmask(j - 1, i) + zydx(k, j - 1, i) * two * dzw(k + 1, j - 1, i) * m.Wv(i, j - 1, k + 1 - 1) * rmask(j - 1, i))) * VM(j, i);
  }
}
static void rhs_vf_div(Rank& K) {
  Grid& g = K.grid[1]; A3 vf = g.dummy3Dnz;
  for (int i = 1; i <= g.nx; ++i) for (int j = 1; j <= g.ny; ++j) for (int k = 1; k <= g.nz; ++k) g.b(k, j, i) = g.b(k, j, i) + vf(k, j + 1, i) - vf(k, j, i);
}
static void rhs_wf(Rank& K) {
  Grid& g = K.grid[1]; const int nx = g.nx, ny = g.ny, nz = g.nz;
  UVW m{K.u.data(), K.v.data(), K.w.data(), nx, ny, nz};
  const A2& rm_ = K.rmask; const bool bm_ = K.bmask;      // umask / vmask (mg_compute_rhs.f90:53-72)
  auto UM = [&](int j, int i) -> double { if (!bm_) return 1.0; return (i >= 1 && i <= nx + 1 && j >= 0 && j <= ny + 1) ? rm_(j, i - 1) * rm_(j, i) : 0.0; };
  auto VM = [&](int j, int i) -> double { if (!bm_) return 1.0; return (i >= 0 && i <= nx + 1 && j >= 1 && j <= ny + 1) ? rm_(j - 1, i) * rm_(j, i) : 0.0; };
  const A2 &dx = g.dx, &dy = g.dy; const A3 &dzw = g.dzw, &zxdy = g.zxdy, &zydx = g.zydx, &cw = g.cw;
  A3 wf = g.dummy3Dnzp; const double qrt = 0.25, hlf = 0.5;
  for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny; ++j) {
    wf(1, j, i) = 0.0;
    for (int k = 2; k <= nz; ++k) {
      wf(k, j, i) = cw(k, j, i) * dzw(k, j, i) * m.Wv(i, j, k - 1)
        - qrt * hlf * (+zxdy(k, j, i) * (dx(j, i) + dx(j, i - 1)) * m.U(i, j, k) * UM(j, i) + zxdy(k, j, i) * (dx(j, i + 1) + dx(j, i)) * m.U(i + 1, j, k) * UM(j, i + 1)
                       + zxdy(k - 1, j, i) * (dx(j, i) + dx(j, i - 1)) * m.U(i, j, k - 1) * UM(j, i) + zxdy(k - 1, j, i) * (dx(j, i + 1) + dx(j, i)) * m.U(i + 1, j, k - 1) * UM(j, i + 1));
      wf(k, j, i) = wf(k, j, i)
        - qrt * hlf * (+zydx(k, j, i) * (dy(j, i) + dy(j - 1, i)) * m.V(i, j, k) * VM(j, i) + zydx(k, j, i) * (dy(j + 1, i) + dy(j, i)) * m.V(i, j + 1, k) * VM(j + 1, i)
                       + zydx(k - 1, j, i) * (dy(j, i) + dy(j - 1, i)) * m.V(i, j, k - 1) * VM(j, i) + zydx(k - 1, j, i) * (dy(j + 1, i) + dy(j, i)) * m.V(i, j + 1, k - 1) * VM(j + 1, i));
    }
    int k = nz + 1;
    wf(k, j, i) = cw(k, j, i) * dzw(k, j, i) * m.Wv(i, j, k - 1)
      - hlf * hlf * (+zxdy(k - 1, j, i) * (dx(j, i) + dx(j, i - 1)) * m.U(i, j, k - 1) * UM(j, i) + zxdy(k - 1, j, i) * (dx(j, i + 1) + dx(j, i)) * m.U(i + 1, j, k - 1) * UM(j, i + 1))
      - hlf * hlf * (+zydx(k - 1, j, i) * (dy(j, i) + dy(j - 1, i)) * m.V(i, j, k - 1) * VM(j, i) + zydx(k - 1, j, i) * (dy(j + 1, i) + dy(j, i)) * m.V(i, j + 1, k - 1) * VM(j + 1, i));
    for (k = 1; k <= nz; ++k) g.b(k, j, i) = g.b(k, j, i) + wf(k + 1, j, i) - wf(k, j, i);
  }
}
static void compute_rhs(World& W) {
  FOR_RANKS(W) { W.R[rk].bmask = W.P.bmask; rhs_uf(W.R[rk]); }
  if (!W.P.bmask) fill_halo(W, 1, sel_d3nz, 'u');          // mg_compute_rhs.f90:170-172: with bmask the masks close the walls
  FOR_RANKS(W) rhs_uf_div(W.R[rk]);
  FOR_RANKS(W) rhs_vf(W.R[rk]);
  if (!W.P.bmask) fill_halo(W, 1, sel_d3nz, 'v');
  FOR_RANKS(W) rhs_vf_div(W.R[rk]);
  FOR_RANKS(W) rhs_wf(W.R[rk]);
}
static void correct_uvw(World& W) {
  FOR_RANKS(W) {
    Rank& K = W.R[rk]; Grid& g = K.grid[1]; const int nx = g.nx, ny = g.ny, nz = g.nz;
    UVW m{K.u.data(), K.v.data(), K.w.data(), nx, ny, nz};
    const A3& p = g.p; const double hlf = 0.5, one = 1.0;
    const A2& rm_ = K.rmask; const bool bm_ = W.P.bmask;    // mg_correct_uvw.f90:48-68
    auto UM = [&](int j, int i) -> double { if (!bm_) return 1.0; return (i >= 1 && i <= nx + 1 && j >= 0 && j <= ny + 1) ? rm_(j, i - 1) * rm_(j, i) : 0.0; };
    auto VM = [&](int j, int i) -> double { if (!bm_) return 1.0; return (i >= 0 && i <= nx + 1 && j >= 1 && j <= ny + 1) ? rm_(j - 1, i) * rm_(j, i) : 0.0; };
    for (int i = 1; i <= nx + 1; ++i) for (int j = 0; j <= ny + 1; ++j) for (int k = 1; k <= nz; ++k) {
      double dxu = hlf * (g.dx(j, i) + g.dx(j, i - 1));
      m.U(i, j, k) = m.U(i, j, k) - one / dxu * (p(k, j, i) - p(k, j, i - 1)) * UM(j, i);
    }
    for (int i = 0; i <= nx + 1; ++i) for (int j = 1; j <= ny + 1; ++j) for (int k = 1; k <= nz; ++k) {
      double dyv = hlf * (g.dy(j, i) + g.dy(j - 1, i));
      m.V(i, j, k) = m.V(i, j, k) - one / dyv * (p(k, j, i) - p(k, j - 1, i)) * VM(j, i);
    }
    for (int i = 0; i <= nx + 1; ++i) for (int j = 0; j <= ny + 1; ++j) {
      for (int k = 2; k <= nz; ++k) {
        double dzw = g.zr(k, j, i) - g.zr(k - 1, j, i);
        m.Wv(i, j, k - 1) = m.Wv(i, j, k - 1) - one / dzw * (p(k, j, i) - p(k - 1, j, i));
      }
      int k = nz + 1;
      double dzw = g.zw(nz + 1, j, i) - g.zr(nz, j, i);
      m.Wv(i, j, k - 1) = m.Wv(i, j, k - 1) - one / dzw * (-p(k - 1, j, i));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// relaxation  (mg_relax.f90:16-47 relax, :151-190 RB, :116-148 GS, :193-234 FC, :237-305 heart, :308-334 tridiag)
// ---------------------------------------------------------------------------------------------
static inline void tridiag(int l, const double* d, const double* dd, const double* b, double* xc, double* gam) {
  double bet = 1.0 / d[0];
  xc[0] = b[0] * bet;
  for (int k = 1; k < l; ++k) {
    gam[k] = dd[k - 1] * bet;
    bet = 1.0 / (d[k] - dd[k - 1] * gam[k]);
    xc[k] = (b[k] - dd[k - 1] * xc[k - 1]) * bet;
  }
  for (int k = l - 2; k >= 0; --k) xc[k] = xc[k] - gam[k + 1] * xc[k + 1];
}

// rhs of the column system; with_bottom_diag selects which of the four same-colour k=1 terms enter:
// bit0: (i+1) terms  cA(5,1,j-1,i+1) p(1,j-1,i+1), cA(8,1,j+1,i+1) p(1,j+1,i+1)
// bit1: (i-1) terms  cA(5,1,j,i) p(1,j+1,i-1),     cA(8,1,j,i) p(1,j-1,i-1)
static inline void column_rhs(const Grid& g, int i, int j, bool real_matrix, double* rhs, double* d, double* ud) {
  const A3 &p = g.p, &b = g.b; const A4& cA = g.cA; const int nz = g.nz;
  int k = 1;
  rhs[0] = b(k, j, i) - cA(3, k, j, i) * p(k + 1, j - 1, i)
           - cA(4, k, j, i) * p(k, j - 1, i) - cA(4, k, j + 1, i) * p(k, j + 1, i)
           - cA(5, k + 1, j + 1, i) * p(k + 1, j + 1, i)
           - cA(6, k, j, i) * p(k + 1, j, i - 1)
           - cA(7, k, j, i) * p(k, j, i - 1) - cA(7, k, j, i + 1) * p(k, j, i + 1)
           - cA(8, k + 1, j, i + 1) * p(k + 1, j, i + 1);
  if (real_matrix)
    rhs[0] = rhs[0] - cA(5, k, j, i) * p(k, j + 1, i - 1) - cA(5, k, j - 1, i + 1) * p(k, j - 1, i + 1)
                    - cA(8, k, j, i) * p(k, j - 1, i - 1) - cA(8, k, j + 1, i + 1) * p(k, j + 1, i + 1);
  d[0] = cA(1, k, j, i); ud[0] = cA(2, k + 1, j, i);
  for (k = 2; k <= nz - 1; ++k) {
    rhs[k - 1] = b(k, j, i)
      - cA(3, k, j, i) * p(k + 1, j - 1, i) - cA(3, k - 1, j + 1, i) * p(k - 1, j + 1, i)
      - cA(4, k, j, i) * p(k, j - 1, i) - cA(4, k, j + 1, i) * p(k, j + 1, i)
      - cA(5, k, j, i) * p(k - 1, j - 1, i) - cA(5, k + 1, j + 1, i) * p(k + 1, j + 1, i)
      - cA(6, k, j, i) * p(k + 1, j, i - 1) - cA(6, k - 1, j, i + 1) * p(k - 1, j, i + 1)
      - cA(7, k, j, i) * p(k, j, i - 1) - cA(7, k, j, i + 1) * p(k, j, i + 1)
      - cA(8, k, j, i) * p(k - 1, j, i - 1) - cA(8, k + 1, j, i + 1) * p(k + 1, j, i + 1);
    d[k - 1] = cA(1, k, j, i); ud[k - 1] = cA(2, k + 1, j, i);
  }
  k = nz;
  rhs[k - 1] = b(k, j, i) - cA(3, k - 1, j + 1, i) * p(k - 1, j + 1, i)
    - cA(4, k, j, i) * p(k, j - 1, i) - cA(4, k, j + 1, i) * p(k, j + 1, i)
    - cA(5, k, j, i) * p(k - 1, j - 1, i)
    - cA(6, k - 1, j, i + 1) * p(k - 1, j, i + 1)
    - cA(7, k, j, i) * p(k, j, i - 1) - cA(7, k, j, i + 1) * p(k, j, i + 1)
    - cA(8, k, j, i) * p(k - 1, j, i - 1);
  d[k - 1] = cA(1, k, j, i);
}

static inline void relax_heart(const Grid& g, int i, int j, bool real_matrix, double* ws) {
  const int nz = g.nz; double *rhs = ws, *d = ws + nz, *ud = ws + 2 * nz, *gam = ws + 3 * nz;
  column_rhs(g, i, j, real_matrix, rhs, d, ud);
  tridiag(nz, d, ud, rhs, &g.p(1, j, i), gam);
}

// One colour of the red-black sweep in the REFERENCE order (i outer ascending, j inner) -- mg_relax.f90:173-179
static void rb_colour_reference(const Grid& g, int rb, bool real_matrix) {
  std::vector<double> ws(4 * g.nz);
  for (int i = 1; i <= g.nx; ++i) for (int j = 1 + (i + rb) % 2; j <= g.ny; j += 2) relax_heart(g, i, j, real_matrix, ws.data());
}

// Same colour in the rank-1 corrected parallel form; equal to the reference order up to
// round-off.  This is the formulation the CUDA path uses, restated here so tests can compare bit-for-bit.
//   A: q = T^-1 rhs0, rhs0 has OLD values in the two (i+1) bottom-diagonal terms and omits the (i-1) ones
//   B: for i ascending: s = -(cA5(1,j,i) x1(j+1,i-1) + cA8(1,j,i) x1(j-1,i-1)); x1(j,i) = q1 + s g1
//   C: p = q + s g, g = T^-1 e1
// B is a contraction (rho = max |g1 cA5| + |g1 cA8| is ~1e-2), so it is evaluated in independent chunks of
// RB_CHUNK columns, each warmed up over W columns starting from the OLD x1: the start-up error decays like
// rho^W <= 1e-20, i.e. far below one ulp.  W > RB_WMAX falls back to one sequential chunk.
static const int RB_CHUNK = 8, RB_WMAX = 24;
static int rb_warmup(double rho) { int w = 0; double acc = 1.0; while (acc > 1e-20 && w < 64) { acc *= rho; ++w; } return w; }
static void rb_colour_rank1(const Grid& g, int rb, bool real_matrix) {
  const int nx = g.nx, ny = g.ny, nz = g.nz; const A3& p = g.p; const A4& cA = g.cA;
  const size_t n2 = (size_t)(ny + 2) * (nx + 2);
  std::vector<double> x1(n2), q1(n2, 0.0), s(n2, 0.0), g1(n2, 0.0), gfull((size_t)nz * n2, 0.0), ws(5 * nz);
  auto I2 = [&](int j, int i) { return (size_t)j + (size_t)(ny + 2) * i; };
  for (int i = 0; i <= nx + 1; ++i) for (int j = 0; j <= ny + 1; ++j) x1[I2(j, i)] = p(1, j, i);     // snapshot
  double *rhs = ws.data(), *d = rhs + nz, *ud = d + nz, *gam = ud + nz;
  for (int i = 1; i <= nx; ++i) for (int j = 1 + (i + rb) % 2; j <= ny; j += 2) {                     // A
    column_rhs(g, i, j, false, rhs, d, ud);
    if (real_matrix) rhs[0] = rhs[0] - cA(5, 1, j - 1, i + 1) * x1[I2(j - 1, i + 1)] - cA(8, 1, j + 1, i + 1) * x1[I2(j + 1, i + 1)];
    tridiag(nz, d, ud, rhs, &p(1, j, i), gam);
    q1[I2(j, i)] = p(1, j, i);
  }
  if (!real_matrix) return;
  // g = T^-1 e1 for every interior column, rho over ALL interior columns (both colours), as at matrix-build time
  std::vector<double> e1(nz);
  double rho = 0.0; bool any = false;
  for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny; ++j) {
    for (int k = 1; k <= nz; ++k) { d[k - 1] = cA(1, k, j, i); if (k < nz) ud[k - 1] = cA(2, k + 1, j, i); e1[k - 1] = 0.0; }
    e1[0] = 1.0;
    tridiag(nz, d, ud, e1.data(), &gfull[(size_t)nz * I2(j, i)], gam);
    g1[I2(j, i)] = gfull[(size_t)nz * I2(j, i)];
    rho = std::max(rho, std::fabs(g1[I2(j, i)] * cA(5, 1, j, i)) + std::fabs(g1[I2(j, i)] * cA(8, 1, j, i)));
  }
  for (int i = 1; i <= nx + 1 && !any; ++i) for (int j = 0; j <= ny + 1; ++j) if (cA(5, 1, j, i) != 0.0 || cA(8, 1, j, i) != 0.0) { any = true; break; }
  if (!any) return;                                            // flat bottom: no same-colour coupling
  const int W = rb_warmup(rho);
  const int chunk = (W > RB_WMAX) ? nx : RB_CHUNK;
  std::vector<double> prev(ny + 2), cur(ny + 2);
  for (int c0 = 1; c0 <= nx; c0 += chunk) {                                                           // B
    const int c1 = std::min(nx, c0 + chunk - 1), i0 = std::max(1, c0 - W);
    for (int j = 0; j <= ny + 1; ++j) prev[j] = x1[I2(j, i0 - 1)];
    for (int i = i0; i <= c1; ++i) {
      for (int j = 1 + (i + rb) % 2; j <= ny; j += 2) {
        const double sv = -(cA(5, 1, j, i) * prev[j + 1] + cA(8, 1, j, i) * prev[j - 1]);
        cur[j] = q1[I2(j, i)] + sv * g1[I2(j, i)];
        if (i >= c0) s[I2(j, i)] = sv;
      }
      cur[0] = x1[I2(0, i)]; cur[ny + 1] = x1[I2(ny + 1, i)];
      std::swap(prev, cur);
    }
  }
  for (int i = 1; i <= nx; ++i) for (int j = 1 + (i + rb) % 2; j <= ny; j += 2)                       // C
    for (int k = 1; k <= nz; ++k) p(k, j, i) = p(k, j, i) + s[I2(j, i)] * gfull[(size_t)nz * I2(j, i) + k - 1];
}

static void relax(World& W, int lev, int nsweeps) {
  const std::string& m = W.P.relax_method; const bool real_matrix = (W.P.cmatrix == "real");
  for (int it = 1; it <= nsweeps; ++it) {
    if (m == "Red-Black" || m == "RB") {
      for (int rb = 1; rb <= 2; ++rb) {
        FOR_RANKS(W) { if (W.relax_order == 0) rb_colour_reference(W.R[rk].grid[lev], rb, real_matrix); else rb_colour_rank1(W.R[rk].grid[lev], rb, real_matrix); }
        fill_halo(W, lev, sel_p);
      }
    } else if (m == "Gauss-Seidel" || m == "GS") {
      FOR_RANKS(W) { const Grid& g = W.R[rk].grid[lev]; std::vector<double> ws(4 * g.nz);
        for (int i = 1; i <= g.nx; ++i) for (int j = 1; j <= g.ny; ++j) relax_heart(g, i, j, real_matrix, ws.data()); }
      fill_halo(W, lev, sel_p);
    } else if (m == "Four-Color" || m == "FC") {
      for (int fc1 = 1; fc1 <= 2; ++fc1) for (int fc2 = 1; fc2 <= 2; ++fc2) {
        FOR_RANKS(W) { const Grid& g = W.R[rk].grid[lev]; std::vector<double> ws(4 * g.nz);
          for (int i = 1 + (fc1 - 1) % 2; i <= g.nx; i += 2) for (int j = 1 + (fc2 - 1) % 2; j <= g.ny; j += 2) relax_heart(g, i, j, real_matrix, ws.data()); }
        fill_halo(W, lev, sel_p);
      }
    }
  }
}

// mg_relax.f90:337-383 compute_residual, :421-515 compute_residual_3D_8
static double residual_local(const Grid& g, bool real_matrix) {
  const A3 &p = g.p, &b = g.b, &r = g.r; const A4& cA = g.cA; const int nx = g.nx, ny = g.ny, nz = g.nz;
  double res = 0.0;
  for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny; ++j) {
    int k = 1;
    r(k, j, i) = b(k, j, i) - cA(1, k, j, i) * p(k, j, i) - cA(2, k + 1, j, i) * p(k + 1, j, i)
      - cA(3, k, j, i) * p(k + 1, j - 1, i)
      - cA(4, k, j, i) * p(k, j - 1, i) - cA(4, k, j + 1, i) * p(k, j + 1, i)
      - cA(5, k + 1, j + 1, i) * p(k + 1, j + 1, i)
      - cA(6, k, j, i) * p(k + 1, j, i - 1)
      - cA(7, k, j, i) * p(k, j, i - 1) - cA(7, k, j, i + 1) * p(k, j, i + 1)
      - cA(8, k + 1, j, i + 1) * p(k + 1, j, i + 1);
    if (real_matrix)
      r(k, j, i) = r(k, j, i) - cA(5, k, j, i) * p(k, j + 1, i - 1) - cA(5, k, j - 1, i + 1) * p(k, j - 1, i + 1)
                              - cA(8, k, j, i) * p(k, j - 1, i - 1) - cA(8, k, j + 1, i + 1) * p(k, j + 1, i + 1);
    res = res + r(k, j, i) * r(k, j, i);
    for (k = 2; k <= nz - 1; ++k) {
      r(k, j, i) = b(k, j, i) - cA(1, k, j, i) * p(k, j, i)
        - cA(2, k, j, i) * p(k - 1, j, i) - cA(2, k + 1, j, i) * p(k + 1, j, i)
        - cA(3, k, j, i) * p(k + 1, j - 1, i) - cA(3, k - 1, j + 1, i) * p(k - 1, j + 1, i)
        - cA(4, k, j, i) * p(k, j - 1, i) - cA(4, k, j + 1, i) * p(k, j + 1, i)
        - cA(5, k, j, i) * p(k - 1, j - 1, i) - cA(5, k + 1, j + 1, i) * p(k + 1, j + 1, i)
        - cA(6, k, j, i) * p(k + 1, j, i - 1) - cA(6, k - 1, j, i + 1) * p(k - 1, j, i + 1)
        - cA(7, k, j, i) * p(k, j, i - 1) - cA(7, k, j, i + 1) * p(k, j, i + 1)
        - cA(8, k, j, i) * p(k - 1, j, i - 1) - cA(8, k + 1, j, i + 1) * p(k + 1, j, i + 1);
      res = res + r(k, j, i) * r(k, j, i);
    }
    k = nz;
    r(k, j, i) = b(k, j, i) - cA(1, k, j, i) * p(k, j, i) - cA(2, k, j, i) * p(k - 1, j, i)
      - cA(3, k - 1, j + 1, i) * p(k - 1, j + 1, i)
      - cA(4, k, j, i) * p(k, j - 1, i) - cA(4, k, j + 1, i) * p(k, j + 1, i)
      - cA(5, k, j, i) * p(k - 1, j - 1, i)
      - cA(6, k - 1, j, i + 1) * p(k - 1, j, i + 1)
      - cA(7, k, j, i) * p(k, j, i - 1) - cA(7, k, j, i + 1) * p(k, j, i + 1)
      - cA(8, k, j, i) * p(k - 1, j, i - 1);
    res = res + r(k, j, i) * r(k, j, i);
  }
  return res;
}
static double compute_residual(World& W, int lev) {
  std::vector<double> loc(W.nranks); const bool real_matrix = (W.P.cmatrix == "real");
  FOR_RANKS(W) loc[rk] = residual_local(W.R[rk].grid[lev], real_matrix);
  fill_halo(W, lev, sel_r);
  return std::sqrt(global_sum(W, lev, loc));
}

// ---------------------------------------------------------------------------------------------
// intergrid transfers  (mg_intergrids.f90:16-72,139-162 fine2coarse ; :167-228,336-450 coarse2fine)
// ---------------------------------------------------------------------------------------------
static void fine2coarse(World& W, int lev) {
  const bool gath = W.R[0].grid[lev + 1].gather == 1;
  FOR_RANKS(W) {
    Grid& gf = W.R[rk].grid[lev]; Grid& gc = W.R[rk].grid[lev + 1];
    A3 x = gf.r; A3 y = gath ? gc.dummy3 : gc.b;
    int nx = gc.nx / gc.ngx, ny = gc.ny / gc.ngy, nz = gc.nz;
    for (int i2 = 1; i2 <= nx; ++i2) { int i = 2 * i2 - 1;
      for (int j2 = 1; j2 <= ny; ++j2) { int j = 2 * j2 - 1;
        for (int k2 = 1; k2 <= nz; ++k2) { int k = 2 * k2 - 1;
          double z = x(k, j, i) + x(k, j, i + 1) + x(k, j + 1, i) + x(k, j + 1, i + 1)
                   + x(k + 1, j, i) + x(k + 1, j, i + 1) + x(k + 1, j + 1, i) + x(k + 1, j + 1, i + 1);
          y(k2, j2, i2) = z; } } }
  }
  if (gath) gather3(W, lev + 1, sel_dummy3, sel_b);
  fill_halo(W, lev + 1, sel_b);
  FOR_RANKS(W) { Grid& gc = W.R[rk].grid[lev + 1]; std::fill(gc.p.d, gc.p.d + gc.p.size(), 0.0); }
}

static void c2f_linear(const A3& xf, const A3& xc, int nx, int ny, int nz) {
  const double a = 9.0 / 16.0, b = 3.0 / 16.0, c = 1.0 / 16.0, d = 27.0 / 64.0, e = 9.0 / 64.0, f = 3.0 / 64.0, g = 1.0 / 64.0;
  for (int i2 = 1; i2 <= nx; ++i2) { int i = 2 * i2 - 1;
    for (int j2 = 1; j2 <= ny; ++j2) { int j = 2 * j2 - 1;
      int k = 1, k2 = 1;
      xf(k, j, i)         = +a * xc(k2, j2, i2) + c * xc(k2, j2 - 1, i2 - 1) + b * xc(k2, j2 - 1, i2) + b * xc(k2, j2, i2 - 1);
      xf(k, j + 1, i)     = +a * xc(k2, j2, i2) + c * xc(k2, j2 + 1, i2 - 1) + b * xc(k2, j2 + 1, i2) + b * xc(k2, j2, i2 - 1);
      xf(k, j, i + 1)     = +a * xc(k2, j2, i2) + c * xc(k2, j2 - 1, i2 + 1) + b * xc(k2, j2 - 1, i2) + b * xc(k2, j2, i2 + 1);
      xf(k, j + 1, i + 1) = +a * xc(k2, j2, i2) + c * xc(k2, j2 + 1, i2 + 1) + b * xc(k2, j2 + 1, i2) + b * xc(k2, j2, i2 + 1);
      for (k = 2; k <= nz * 2 - 1; ++k) {
        k2 = (k + 1) / 2; int kp = k2 - ((k % 2) * 2 - 1);
        xf(k, j, i)         = +d * xc(k2, j2, i2) + f * xc(k2, j2 - 1, i2 - 1) + e * xc(k2, j2 - 1, i2) + e * xc(k2, j2, i2 - 1)
                              + e * xc(kp, j2, i2) + g * xc(kp, j2 - 1, i2 - 1) + f * xc(kp, j2 - 1, i2) + f * xc(kp, j2, i2 - 1);
        xf(k, j + 1, i)     = +d * xc(k2, j2, i2) + f * xc(k2, j2 + 1, i2 - 1) + e * xc(k2, j2 + 1, i2) + e * xc(k2, j2, i2 - 1)
                              + e * xc(kp, j2, i2) + g * xc(kp, j2 + 1, i2 - 1) + f * xc(kp, j2 + 1, i2) + f * xc(kp, j2, i2 - 1);
        xf(k, j, i + 1)     = +d * xc(k2, j2, i2) + f * xc(k2, j2 - 1, i2 + 1) + e * xc(k2, j2 - 1, i2) + e * xc(k2, j2, i2 + 1)
                              + e * xc(kp, j2, i2) + g * xc(kp, j2 - 1, i2 + 1) + f * xc(kp, j2 - 1, i2) + f * xc(kp, j2, i2 + 1);
        xf(k, j + 1, i + 1) = +d * xc(k2, j2, i2) + f * xc(k2, j2 + 1, i2 + 1) + e * xc(k2, j2 + 1, i2) + e * xc(k2, j2, i2 + 1)
                              + e * xc(kp, j2, i2) + g * xc(kp, j2 + 1, i2 + 1) + f * xc(kp, j2 + 1, i2) + f * xc(kp, j2, i2 + 1);
      }
      k = nz * 2; k2 = nz;     // k2 keeps its last loop value (= nz) in the reference (:434)
      xf(k, j, i)         = 0.5 * (a * xc(k2, j2, i2) + c * xc(k2, j2 - 1, i2 - 1) + b * xc(k2, j2 - 1, i2) + b * xc(k2, j2, i2 - 1));
      xf(k, j + 1, i)     = 0.5 * (a * xc(k2, j2, i2) + c * xc(k2, j2 + 1, i2 - 1) + b * xc(k2, j2 + 1, i2) + b * xc(k2, j2, i2 - 1));
      xf(k, j, i + 1)     = 0.5 * (a * xc(k2, j2, i2) + c * xc(k2, j2 - 1, i2 + 1) + b * xc(k2, j2 - 1, i2) + b * xc(k2, j2, i2 + 1));
      xf(k, j + 1, i + 1) = 0.5 * (a * xc(k2, j2, i2) + c * xc(k2, j2 + 1, i2 + 1) + b * xc(k2, j2 + 1, i2) + b * xc(k2, j2, i2 + 1));
    } }
}
static void c2f_nearest(const A3& xf, const A3& xc, int nx, int ny, int nz) {   // :336-363
  for (int i2 = 1; i2 <= nx; ++i2) { int i = 2 * i2 - 1;
    for (int j2 = 1; j2 <= ny; ++j2) { int j = 2 * j2 - 1;
      for (int k2 = 1; k2 <= nz; ++k2) { int k = 2 * k2 - 1; double v = xc(k2, j2, i2);
        xf(k, j, i) = v; xf(k + 1, j, i) = v; xf(k, j + 1, i) = v; xf(k + 1, j + 1, i) = v;
        xf(k, j, i + 1) = v; xf(k + 1, j, i + 1) = v; xf(k, j + 1, i + 1) = v; xf(k + 1, j + 1, i + 1) = v; } } }
}
static void coarse2fine(World& W, int lev) {
  const bool gath = W.R[0].grid[lev + 1].gather == 1; const bool linear = (W.P.interp_type == "linear");
  FOR_RANKS(W) {
    Grid& gf = W.R[rk].grid[lev]; Grid& gc = W.R[rk].grid[lev + 1];
    A3 pc = gc.p; int nxc = gc.nx, nyc = gc.ny, nzc = gc.nz;
    if (gath) { split3(gc, gc.p, gc.dummy3); pc = gc.dummy3; nxc = gc.nx / gc.ngx; nyc = gc.ny / gc.ngy; }
    if (linear) c2f_linear(gf.r, pc, nxc, nyc, nzc); else c2f_nearest(gf.r, pc, nxc, nyc, nzc);
  }
  fill_halo(W, lev, sel_r);
  FOR_RANKS(W) { Grid& gf = W.R[rk].grid[lev]; size_t n = gf.p.size(); for (size_t q = 0; q < n; ++q) gf.p.d[q] = gf.p.d[q] + gf.r.d[q]; }
}

// ---------------------------------------------------------------------------------------------
// cycles and solver  (mg_solvers.f90:17-101 solve_p, :104-126 Fcycle, :129-151 Vcycle)
// ---------------------------------------------------------------------------------------------
static void Vcycle(World& W, int lev1) {
  const int L = W.nlevs;
  for (int lev = lev1; lev <= L - 1; ++lev) { relax(W, lev, W.P.ns_pre); (void)compute_residual(W, lev); fine2coarse(W, lev); }
  relax(W, L, W.P.ns_coarsest);
  for (int lev = L - 1; lev >= lev1; --lev) { coarse2fine(W, lev); relax(W, lev, W.P.ns_post); }
}
static void Fcycle(World& W) {
  const int L = W.nlevs;
  for (int lev = 1; lev <= L - 1; ++lev) {
    fine2coarse(W, lev);
    FOR_RANKS(W) { Grid& g = W.R[rk].grid[lev + 1]; std::memcpy(g.r.d, g.b.d, g.b.size() * sizeof(double)); }
  }
  relax(W, L, W.P.ns_coarsest);
  for (int lev = L - 1; lev >= 1; --lev) { coarse2fine(W, lev); Vcycle(W, lev); }
}
static void solve_p(World& W) {
  const double tol = W.P.solver_prec; const int maxite = W.P.solver_maxiter;
  FOR_RANKS(W) { Grid& g = W.R[rk].grid[1]; std::fill(g.p.d, g.p.d + g.p.size(), 0.0); }
  auto t0 = std::chrono::steady_clock::now();
  int nite = 0;
  std::vector<double> loc(W.nranks);
  FOR_RANKS(W) { const Grid& g = W.R[rk].grid[1]; double s = 0.0;
    for (int i = 1; i <= g.nx; ++i) for (int j = 1; j <= g.ny; ++j) for (int k = 1; k <= g.nz; ++k) s += g.b(k, j, i) * g.b(k, j, i);
    loc[rk] = s; }
  double bnorm = std::sqrt(global_sum(W, 1, loc));
  double rnorm = compute_residual(W, 1);
  double res0 = rnorm / bnorm;
  W.hist_res.clear(); W.hist_conv.clear();
  W.hist_res.push_back(res0); W.hist_conv.push_back(0.0);
  while (nite < maxite && res0 > tol) {
    Fcycle(W);
    rnorm = compute_residual(W, 1);
    rnorm = rnorm / bnorm;
    double conv = res0 / rnorm;
    res0 = rnorm;
    nite = nite + 1;
    W.hist_res.push_back(rnorm); W.hist_conv.push_back(conv);
  }
  W.solve_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  W.bnorm = bnorm; W.nite = nite;
}

static A3* field3(Grid& g, const std::string& n) {
  if (n == "p") return &g.p; if (n == "b") return &g.b; if (n == "r") return &g.r;
  if (n == "zr") return &g.zr; if (n == "zw") return &g.zw; if (n == "cw") return &g.cw;
  if (n == "dzw") return &g.dzw; if (n == "zxdy") return &g.zxdy; if (n == "zydx") return &g.zydx;
  if (n == "dummy3") return &g.dummy3;
  return nullptr;
}
static A2* field2(Grid& g, const std::string& n) {
  if (n == "dx") return &g.dx; if (n == "dy") return &g.dy; if (n == "zeta") return &g.zeta; if (n == "h") return &g.h;
  return nullptr;
}

}  // namespace

// =============================================================================================
// C API (ctypes).  Levels are 1-based like the reference.  Return 0 on success.
// =============================================================================================
extern "C" {

void* mgo_create(int nx, int ny, int nz, int npx, int npy, const char* namelist_path) {
  World* W = new World();
  W->npx = npx; W->npy = npy; W->nranks = npx * npy; W->R.resize(W->nranks);
  if (namelist_path && namelist_path[0]) {
    int rc = read_namelist(namelist_path, W->P);
    if (rc < 0) { W->err = "namelist parse error"; return W; }
  }
  return W;
}
int mgo_set_param(void* h, const char* key, const char* val) {
  World* W = (World*)h; std::string k = lower(key), v = val; Params& P = W->P;
  if (k == "solver_prec") P.solver_prec = parse_real(v); else if (k == "solver_maxiter") P.solver_maxiter = atoi(val);
  else if (k == "nsmall") P.nsmall = atoi(val); else if (k == "ns_coarsest") P.ns_coarsest = atoi(val);
  else if (k == "ns_pre") P.ns_pre = atoi(val); else if (k == "ns_post") P.ns_post = atoi(val);
  else if (k == "cmatrix") P.cmatrix = v; else if (k == "relax_method") P.relax_method = v;
  else if (k == "interp_type") P.interp_type = v; else if (k == "restrict_type") P.restrict_type = v;
  else if (k == "netcdf_output") P.netcdf_output = parse_logical(v); else if (k == "aggressive") P.aggressive = parse_logical(v);
  else if (k == "bmask") P.bmask = parse_logical(v);
  else if (k == "relax_order") W->relax_order = atoi(val);          // oracle-only switch, not a namelist key
  else return 1;
  return 0;
}
int mgo_init(void* h, int nx, int ny, int nz) { World* W = (World*)h; if (!W->err.empty()) return 1; return define_grids(*W, nx, ny, nz); }
const char* mgo_last_error(void* h) { return ((World*)h)->err.c_str(); }
void mgo_destroy(void* h) { delete (World*)h; }
int mgo_nlevs(void* h) { return ((World*)h)->nlevs; }
// info[0..11] = nx,ny,nz,npx,npy,incx,incy,gather,ngx,ngy,key,color ; info[12..19] = neighbours
int mgo_level_info(void* h, int rank, int lev, int* info) {
  World* W = (World*)h; if (lev < 1 || lev > W->nlevs || rank < 0 || rank >= W->nranks) return 1;
  const Grid& g = W->R[rank].grid[lev];
  int v[12] = {g.nx, g.ny, g.nz, g.npx, g.npy, g.incx, g.incy, g.gather, g.ngx, g.ngy, g.key, g.color};
  for (int q = 0; q < 12; ++q) info[q] = v[q];
  for (int q = 0; q < 8; ++q) info[12 + q] = g.neighb[q];
  return 0;
}
int mgo_gather_members(void* h, int rank, int lev, int* out) {
  World* W = (World*)h; const Grid& g = W->R[rank].grid[lev];
  for (size_t q = 0; q < g.members.size(); ++q) out[q] = g.members[q];
  return (int)g.members.size();
}
// geometry of one rank: arrays (0:ny+1,0:nx+1), j fastest (nhydro.f90:36-40)
int mgo_set_geometry(void* h, int rank, const double* dx, const double* dy, const double* zeta, const double* hh, const double* rmask) {
  World* W = (World*)h; Rank& K = W->R[rank]; Grid& g = K.grid[1]; size_t n = g.dx.size();
  std::memcpy(g.dx.d, dx, n * 8); std::memcpy(g.dy.d, dy, n * 8); std::memcpy(g.zeta.d, zeta, n * 8); std::memcpy(g.h.d, hh, n * 8);
  if (rmask) std::memcpy(K.rmask.d, rmask, n * 8); else std::fill(K.rmask.d, K.rmask.d + n, 1.0);
  return 0;
}
int mgo_matrices(void* h, double hc, double theta_b, double theta_s) {
  World* W = (World*)h; W->hlim = hc; W->theta_b = theta_b; W->theta_s = theta_s; define_matrices(*W); return 0;
}
int mgo_set_uvw(void* h, int rank, const double* u, const double* v, const double* w) {
  Rank& K = ((World*)h)->R[rank];
  std::memcpy(K.u.data(), u, K.u.size() * 8); std::memcpy(K.v.data(), v, K.v.size() * 8); std::memcpy(K.w.data(), w, K.w.size() * 8); return 0;
}
int mgo_get_uvw(void* h, int rank, double* u, double* v, double* w) {
  Rank& K = ((World*)h)->R[rank];
  std::memcpy(u, K.u.data(), K.u.size() * 8); std::memcpy(v, K.v.data(), K.v.size() * 8); std::memcpy(w, K.w.data(), K.w.size() * 8); return 0;
}
int mgo_compute_rhs(void* h) { compute_rhs(*(World*)h); return 0; }
int mgo_correct_uvw(void* h) { correct_uvw(*(World*)h); return 0; }
int mgo_solve_p(void* h) { solve_p(*(World*)h); return 0; }
int mgo_nhydro_solve(void* h) { World* W = (World*)h; compute_rhs(*W); solve_p(*W); correct_uvw(*W); return 0; }   // nhydro.f90:80-98
int mgo_relax(void* h, int lev, int nsweeps) { relax(*(World*)h, lev, nsweeps); return 0; }
int mgo_residual(void* h, int lev, double* res) { *res = compute_residual(*(World*)h, lev); return 0; }
int mgo_fine2coarse(void* h, int lev) { fine2coarse(*(World*)h, lev); return 0; }
int mgo_coarse2fine(void* h, int lev) { coarse2fine(*(World*)h, lev); return 0; }
int mgo_fcycle(void* h) { Fcycle(*(World*)h); return 0; }
int mgo_vcycle(void* h, int lev1) { Vcycle(*(World*)h, lev1); return 0; }
int mgo_fill_halo(void* h, int lev, const char* name) {
  World* W = (World*)h; std::string n = name;
  if (n == "p") fill_halo(*W, lev, sel_p); else if (n == "b") fill_halo(*W, lev, sel_b); else if (n == "r") fill_halo(*W, lev, sel_r);
  else if (n == "zr") fill_halo(*W, lev, sel_zr); else if (n == "zw") fill_halo(*W, lev, sel_zw);
  else if (n == "h") fill_halo(*W, lev, sel_h); else if (n == "r:u") fill_halo(*W, lev, sel_r, 'u'); else if (n == "r:v") fill_halo(*W, lev, sel_r, 'v');
  else return 1;
  return 0;
}
int mgo_gather(void* h, int lev) { gather3(*(World*)h, lev, sel_dummy3, sel_b); return 0; }
int mgo_split(void* h, int lev) { World* W = (World*)h; FOR_RANKS(*W) { Grid& g = W->R[rk].grid[lev]; split3(g, g.p, g.dummy3); } return 0; }
long long mgo_field_size(void* h, int rank, int lev, const char* name) {
  Grid& g = ((World*)h)->R[rank].grid[lev]; std::string n = name;
  if (n == "cA") return (long long)g.cA.size();
  if (A3* a = field3(g, n)) return a->d ? (long long)a->size() : -1;
  if (A2* a = field2(g, n)) return (long long)a->size();
  return -1;
}
int mgo_get_field(void* h, int rank, int lev, const char* name, double* out) {
  Grid& g = ((World*)h)->R[rank].grid[lev]; std::string n = name;
  if (n == "cA") { std::memcpy(out, g.cA.d, g.cA.size() * 8); return 0; }
  if (A3* a = field3(g, n)) { if (!a->d) return 1; std::memcpy(out, a->d, a->size() * 8); return 0; }
  if (A2* a = field2(g, n)) { std::memcpy(out, a->d, a->size() * 8); return 0; }
  return 1;
}
int mgo_set_field(void* h, int rank, int lev, const char* name, const double* in) {
  Grid& g = ((World*)h)->R[rank].grid[lev]; std::string n = name;
  if (n == "cA") { std::memcpy(g.cA.d, in, g.cA.size() * 8); return 0; }
  if (A3* a = field3(g, n)) { if (!a->d) return 1; std::memcpy(a->d, in, a->size() * 8); return 0; }
  if (A2* a = field2(g, n)) { std::memcpy(a->d, in, a->size() * 8); return 0; }
  return 1;
}
int mgo_get_history(void* h, double* res, double* conv, int cap) {
  World* W = (World*)h; int n = (int)W->hist_res.size();
  for (int q = 0; q < n && q < cap; ++q) { res[q] = W->hist_res[q]; if (conv) conv[q] = W->hist_conv[q]; }
  return n;
}
double mgo_solve_seconds(void* h) { return ((World*)h)->solve_seconds; }
double mgo_bnorm(void* h) { return ((World*)h)->bnorm; }
int mgo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void mgo_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

}  // extern "C"

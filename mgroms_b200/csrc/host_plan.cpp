// host_plan.cpp -- host-only logic of libmgb: the nh_namelist reader and the multigrid level planner.
// Mirrors the behaviour of src/mg_namelist.f90:55-127 and src/mg_grids.f90:468-738 of the reference.
#include "mgb_internal.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>

namespace mgb {

namespace {

std::string strip(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}
std::string lowercase(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  return s;
}
bool to_logical(const std::string& v, bool& out) {
  std::string l = lowercase(strip(v));
  if (l == ".true." || l == "t" || l == "true" || l == ".t.") { out = true; return true; }
  if (l == ".false." || l == "f" || l == "false" || l == ".f.") { out = false; return true; }
  return false;
}
bool to_real(std::string v, double& out) {
  v = strip(v);
  for (char& c : v) if (c == 'd' || c == 'D') c = 'e';      // Fortran 1.d-12
  char* end = nullptr;
  out = std::strtod(v.c_str(), &end);
  return end && *end == '\0' && !v.empty();
}
bool to_int(const std::string& v, int& out) {
  std::string s = strip(v);
  char* end = nullptr;
  long x = std::strtol(s.c_str(), &end, 10);
  out = (int)x;
  return end && *end == '\0' && !s.empty();
}
std::string unquote(const std::string& v) {
  std::string s = strip(v);
  if (s.size() >= 2 && (s.front() == '\'' || s.front() == '"') && s.back() == s.front()) return strip(s.substr(1, s.size() - 2));
  return s;
}

}  // namespace

int namelist_set(Params& P, const char* key_c, const char* value_c, std::string& msg) {
  const std::string key = lowercase(strip(key_c ? key_c : "")), val = value_c ? value_c : "";
  bool ok = true;
  if      (key == "solver_prec")    ok = to_real(val, P.solver_prec);
  else if (key == "solver_maxiter") ok = to_int(val, P.solver_maxiter);
  else if (key == "nsmall")         ok = to_int(val, P.nsmall);
  else if (key == "ns_coarsest")    ok = to_int(val, P.ns_coarsest);
  else if (key == "ns_pre")         ok = to_int(val, P.ns_pre);
  else if (key == "ns_post")        ok = to_int(val, P.ns_post);
  else if (key == "cmatrix")        P.cmatrix = unquote(val);
  else if (key == "relax_method")   P.relax_method = unquote(val);
  else if (key == "interp_type")    P.interp_type = unquote(val);
  else if (key == "restrict_type")  P.restrict_type = unquote(val);
  else if (key == "netcdf_output")  ok = to_logical(val, P.netcdf_output);
  else if (key == "aggressive")     ok = to_logical(val, P.aggressive);
  else if (key == "bmask")          ok = to_logical(val, P.bmask);
  else { msg = "namelist: unknown key '" + key + "' in group nhparam"; return -1; }
  if (!ok) { msg = "namelist: bad value '" + val + "' for key '" + key + "'"; return -2; }
  return 0;
}

// Reads group &nhparam ... / from a Fortran namelist file.  Comments start at '!' outside quotes; items are
// separated by commas and/or newlines.  Unknown keys are an error, like a Fortran namelist read.
int namelist_read(const char* path, Params& P, std::string& msg) {
  std::ifstream f(path);
  if (!f.good()) return 1;   // "Read namelist file if it is present, else use default values"
  std::ostringstream body;
  std::string line;
  bool in_group = false, done = false;
  while (!done && std::getline(f, line)) {
    std::string kept;
    char quote = 0;
    for (char c : line) {
      if (quote) { kept += c; if (c == quote) quote = 0; continue; }
      if (c == '\'' || c == '"') { quote = c; kept += c; continue; }
      if (c == '!') break;
      kept += c;
    }
    std::string t = strip(kept);
    if (!in_group) {
      std::string l = lowercase(t);
      size_t pos = l.find("&nhparam");
      if (pos == std::string::npos) continue;
      in_group = true;
      t = strip(t.substr(pos + 8));
    }
    // a '/' outside quotes terminates the group
    quote = 0;
    std::string upto;
    for (char c : t) {
      if (quote) { upto += c; if (c == quote) quote = 0; continue; }
      if (c == '\'' || c == '"') { quote = c; upto += c; continue; }
      if (c == '/') { done = true; break; }
      upto += c;
    }
    body << upto << ",";
  }
  if (!in_group) { msg = std::string("namelist: group &nhparam not found in ") + path; return -3; }
  std::string item;
  std::istringstream items(body.str());
  while (std::getline(items, item, ',')) {
    item = strip(item);
    if (item.empty()) continue;
    size_t eq = item.find('=');
    if (eq == std::string::npos) { msg = "namelist: expected key = value, got '" + item + "'"; return -4; }
    int rc = namelist_set(P, item.substr(0, eq).c_str(), item.substr(eq + 1).c_str(), msg);
    if (rc != 0) return rc;
  }
  // src/mg_namelist.f90:94-98
  if (P.interp_type == "linear" && P.restrict_type == "linear") { msg = "linear interp + linear restrict is not permitted"; return -5; }
  return 0;
}

static int ilog2_floor_ratio(int num, int den) {
  // 1 + floor(log(num/den)/log(2)) as src/mg_grids.f90:493-495 evaluates it (floating point)
  return 1 + (int)std::floor(std::log((double)num * 1.0 / (double)den * 1.0) / std::log(2.0));
}

int plan_levels(const Params& P, int nx, int ny, int nz, int npx, int npy, int rank,
                std::vector<LevelDims>& out, std::string& msg) {
  out.clear();
  if (nx < 1 || ny < 1 || nz < 1 || npx < 1 || npy < 1) { msg = "mgb_init: sizes must be positive"; return 1; }
  if (rank < 0 || rank >= npx * npy) { msg = "mgb_init: rank outside the npx*npy process grid"; return 1; }
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  if (!pow2(npx) || !pow2(npy) || !pow2(nx) || !pow2(ny)) { msg = "horizontal dimensions and process counts must be powers of 2 (reference 'assumptions')"; return 1; }
  if (P.aggressive) { msg = "aggressive coarsening is not available (the reference stops: src/mg_intergrids.f90:243)"; return 1; }

  const int nxg = npx * nx, nyg = npy * ny;
  const int nl1 = ilog2_floor_ratio(std::min(nxg, nyg), 4);   // ncoarsest = 4
  const int nl2 = ilog2_floor_ratio(nz, 2);                    // nzmin = 2
  const int nlevs = std::min(nl1, nl2);
  if (nlevs < 1) { msg = "grid too small for even one level (need min(nxg,nyg) >= 4 and nz >= 2)"; return 1; }

  const int pj = rank / npx, pi = rank % npx;
  int cnx = nx, cny = ny, cnz = nz, cpx = npx, cpy = npy, incx = 1, incy = 1;
  for (int lev = 1; lev <= nlevs; ++lev) {
    LevelDims d;
    if (lev > 1) {
      cnx /= 2; cny /= 2; if (cnz > 1) cnz /= 2;
      if (std::min(cnx, cny) < P.nsmall && cpx * cpy > 1) {
        d.gather = 1;
        if (cpx > 1) { cpx /= 2; cnx *= 2; d.ngx = 2; }
        if (cpy > 1) { cpy /= 2; cny *= 2; d.ngy = 2; }
        incx *= 2; incy *= 2;
      }
    }
    if (cnx < 1 || cny < 1) { msg = "a level has an empty local grid; increase nsmall or use fewer ranks"; return 1; }
    if (cnz < 2) { msg = "a level has nz < 2: the reference's 2-D path is dead code and is not built"; return 1; }
    if ((cnx % 2 || cny % 2) && lev < nlevs) { msg = "odd local size on a non-coarsest level"; return 1; }
    d.nx = cnx; d.ny = cny; d.nz = cnz; d.npx = cpx; d.npy = cpy; d.incx = incx; d.incy = incy;
    const bool S = pj >= incy, N = pj < npy - incy, W = pi >= incx, E = pi < npx - incx;
    d.neighb[0] = S ? (pj - incy) * npx + pi : -1;
    d.neighb[1] = E ? pj * npx + pi + incx : -1;
    d.neighb[2] = N ? (pj + incy) * npx + pi : -1;
    d.neighb[3] = W ? pj * npx + pi - incx : -1;
    d.neighb[4] = (S && W) ? (pj - incy) * npx + pi - incx : -1;
    d.neighb[5] = (S && E) ? (pj - incy) * npx + pi + incx : -1;
    d.neighb[6] = (N && E) ? (pj + incy) * npx + pi + incx : -1;
    d.neighb[7] = (N && W) ? (pj + incy) * npx + pi - incx : -1;
    out.push_back(d);
  }
  // gather families: what MPI_COMM_SPLIT(world, color, key) would build (src/mg_grids.f90:681-718)
  for (int lev = 2; lev <= nlevs; ++lev) {
    LevelDims& d = out[lev - 1];
    if (!d.gather) continue;
    auto color_key = [&](int r, int& color, int& key) {
      const int qj = r / npx, qi = r % npx;
      const int ix = d.incx / 2, iy = d.incy / 2;
      const int family = (qi / ix) * ix * iy + npx * iy * (qj / iy);
      const int nextfamily = (qi / (2 * ix)) * ix * iy * 4 + npx * 2 * iy * (qj / (iy * 2));
      color = nextfamily + (qi % ix) + (qj % iy) * ix;
      const int N = ix * npx;
      key = ((family % N) / (ix * iy)) % 2 + 2 * ((family / N) % 2);
    };
    color_key(rank, d.color, d.key);
    std::vector<std::pair<int, int>> fam;
    for (int r = 0; r < npx * npy; ++r) { int c, k; color_key(r, c, k); if (c == d.color) fam.push_back({k, r}); }
    std::sort(fam.begin(), fam.end());
    if ((int)fam.size() != d.ngx * d.ngy || fam.size() > 4) { msg = "gather family size does not match ngx*ngy"; return 1; }
    d.nmembers = (int)fam.size();
    for (int q = 0; q < d.nmembers; ++q) d.members[q] = fam[q].second;
  }
  return 0;
}

}  // namespace mgb

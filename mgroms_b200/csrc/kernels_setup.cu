// kernels_setup.cu -- FP64 kernels outside the F-cycle: geometry coarsening, s-coordinate depths,
// stencil-matrix construction (define_matrix), rhs (flux-form divergence) and the velocity correction.
// Expressions keep the reference's operation order (and -fmad=false) so results round like the CPU oracle's.
#include "kernels.cuh"

namespace mgb {

static inline unsigned blocks_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

// ---------------------------------------------------------------------------------------------------------
// 2x2 coarsening of a 2-D field (src/mg_define_matrix.f90:116-138): c = w * (f00 + f10 + f01 + f11)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_coarsen2d(const double* __restrict__ f, int nyf, double* __restrict__ c, int nyc, int nxc, double w) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nyc * nxc) return;
  const int j = 1 + e % nyc, i = 1 + e / nyc;
  const int jf = 2 * j - 1, i_f = 2 * i - 1;
  const size_t o = (size_t)jf + (size_t)(nyf + 2) * i_f, si = nyf + 2;
  c[(size_t)j + (size_t)(nyc + 2) * i] = w * (f[o] + f[o + 1] + f[o + si] + f[o + 1 + si]);
}
void launch_coarsen2d(cudaStream_t st, const double* f, int nyf, int nxf, double* c, int nyc, int nxc, double w) {
  (void)nxf;
  k_coarsen2d<<<blocks_for((size_t)nyc * nxc, 256), 256, 0, st>>>(f, nyf, c, nyc, nxc, w); ++g_launches;
}

// ---------------------------------------------------------------------------------------------------------
// zr, zw from h, zeta (src/mg_zr_zw.f90:98-170, 'new_s_coord').  The k-only stretching tables (which involve
// cosh/exp) are evaluated on the host with libm so that device and CPU use bit-identical curve values.
// tab = [cs_r(1..nz) | cff_r(1..nz) | cs_w(1..nz+1) | cff_w(1..nz+1)], each padded to nz+1 entries
// ---------------------------------------------------------------------------------------------------------
__global__ void k_zrzw(const double* __restrict__ h, const double* __restrict__ zeta, const double* __restrict__ tab, double hlim,
                       double* __restrict__ zr, double* __restrict__ zw, int nz, int ny, int nx) {
  const int nzp = nz + 1;
  const size_t n = (size_t)nzp * (ny + 2) * (nx + 2);
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int k = (int)(e % nzp); size_t t = e / nzp; const int j = (int)(t % (ny + 2)), i = (int)(t / (ny + 2));
  const double hh = h[(size_t)j + (size_t)(ny + 2) * i], ze = zeta[(size_t)j + (size_t)(ny + 2) * i];
  const double hinv = 1.0 / (hh + hlim);
  const double cs_w = tab[2 * nzp + k], cff_w = tab[3 * nzp + k];
  const double z_w0 = cff_w + cs_w * hh;
  zw[(size_t)k + (size_t)nzp * ((size_t)(j + 1) + (size_t)(ny + 4) * (size_t)(i + 1))] = z_w0 * hh * hinv + ze * (1.0 + z_w0 * hinv);
  if (k < nz) {
    const double cs_r = tab[k], cff_r = tab[nzp + k];
    const double z_r0 = cff_r + cs_r * hh;
    zr[(size_t)k + (size_t)nz * ((size_t)(j + 1) + (size_t)(ny + 4) * (size_t)(i + 1))] = z_r0 * hh * hinv + ze * (1.0 + z_r0 * hinv);
  }
}
void launch_zrzw(cudaStream_t st, const double* h, const double* zeta, const double* tab, double hlim, double* zr, double* zw, int nz, int ny, int nx) {
  size_t n = (size_t)(nz + 1) * (ny + 2) * (nx + 2);
  k_zrzw<<<blocks_for(n, 256), 256, 0, st>>>(h, zeta, tab, hlim, zr, zw, nz, ny, nx); ++g_launches;
}

// ---------------------------------------------------------------------------------------------------------
// define_matrix (src/mg_define_matrix.f90:211-680); umask = vmask = 1 unless bmask (:255-275)
// Accessors take FORTRAN indices (K = 1..nz, j, i incl. halos)
// ---------------------------------------------------------------------------------------------------------
struct MAcc {
  int nz, ny, nx; const double *dx, *dy, *zr, *zw, *cw; const double* rm; int bmask;
  __device__ __forceinline__ double RMK(int j, int i) const { return rm[(size_t)j + (size_t)(ny + 2) * i]; }
  __device__ __forceinline__ double UM(int j, int i) const { if (!bmask) return 1.0; return (i >= 1 && i <= nx + 1 && j >= 0 && j <= ny + 1) ? RMK(j, i - 1) * RMK(j, i) : 0.0; }
  __device__ __forceinline__ double VM(int j, int i) const { if (!bmask) return 1.0; return (i >= 0 && i <= nx + 1 && j >= 1 && j <= ny + 1) ? RMK(j - 1, i) * RMK(j, i) : 0.0; }
  __device__ __forceinline__ double DX(int j, int i) const { return dx[(size_t)j + (size_t)(ny + 2) * i]; }
  __device__ __forceinline__ double DY(int j, int i) const { return dy[(size_t)j + (size_t)(ny + 2) * i]; }
  __device__ __forceinline__ double ZR(int K, int j, int i) const { return zr[(size_t)(K - 1) + (size_t)nz * ((size_t)(j + 1) + (size_t)(ny + 4) * (size_t)(i + 1))]; }
  __device__ __forceinline__ double ZW(int K, int j, int i) const { return zw[(size_t)(K - 1) + (size_t)(nz + 1) * ((size_t)(j + 1) + (size_t)(ny + 4) * (size_t)(i + 1))]; }
  __device__ __forceinline__ double CW(int K, int j, int i) const { return cw[(size_t)(K - 1) + (size_t)(nz + 1) * ((size_t)j + (size_t)(ny + 2) * i)]; }
  __device__ __forceinline__ double ZYDX(int K, int j, int i) const { return ((0.5 * (ZR(K, j + 1, i) - ZR(K, j - 1, i))) / DY(j, i)) * DX(j, i); }
  __device__ __forceinline__ double ZXDY(int K, int j, int i) const { return ((0.5 * (ZR(K, j, i + 1) - ZR(K, j, i - 1))) / DX(j, i)) * DY(j, i); }
  __device__ __forceinline__ double ARY(int K, int j, int i) const { return (0.25 * (ZW(K + 1, j, i) - ZW(K, j, i) + ZW(K + 1, j - 1, i) - ZW(K, j - 1, i))) * (DX(j, i) + DX(j - 1, i)); }
  __device__ __forceinline__ double ARX(int K, int j, int i) const { return (0.25 * (ZW(K + 1, j, i) - ZW(K, j, i) + ZW(K + 1, j, i - 1) - ZW(K, j, i - 1))) * (DY(j, i) + DY(j, i - 1)); }
  __device__ __forceinline__ double CWS(int j, int i) const { return CW(1, j, i) + CW(2, j, i); }
};
__device__ __forceinline__ double sq(double x) { return x * x; }

// cw (all levels) and dzw, zxdy, zydx (level 1)  -- :283-336
__global__ void k_cw_metrics(MatrixView M) {
  const int nz = M.nz, ny = M.ny, nx = M.nx, nzp = nz + 1;
  const size_t n = (size_t)nzp * (ny + 2) * (nx + 2);
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int K = 1 + (int)(e % nzp); size_t t = e / nzp; const int j = (int)(t % (ny + 2)), i = (int)(t / (ny + 2));
  MAcc A{nz, ny, nx, M.dx, M.dy, M.zr, M.zw, M.cw, M.rmask, M.bmask};
  const double Arz = A.DX(j, i) * A.DY(j, i);
  const double dz = (K == 1) ? (A.ZR(1, j, i) - A.ZW(1, j, i)) : (K <= nz ? (A.ZR(K, j, i) - A.ZR(K - 1, j, i)) : (A.ZW(K, j, i) - A.ZR(K - 1, j, i)));
  M.cw[e] = (Arz / dz) * (1.0 + sq(0.5 * (A.ZW(K, j, i + 1) - A.ZW(K, j, i - 1)) / A.DX(j, i)) + sq(0.5 * (A.ZW(K, j + 1, i) - A.ZW(K, j - 1, i)) / A.DY(j, i)));
  if (M.lev1) {
    M.dzw[e] = dz;
    if (K <= nz) {
      const size_t o = (size_t)(K - 1) + (size_t)nz * ((size_t)j + (size_t)(ny + 2) * i);
      M.zydx[o] = 0.5 * ((A.ZR(K, j + 1, i) - A.ZR(K, j - 1, i)) / A.DY(j, i)) * A.DX(j, i);
      M.zxdy[o] = 0.5 * ((A.ZR(K, j, i + 1) - A.ZR(K, j, i - 1)) / A.DX(j, i)) * A.DY(j, i);
    }
  }
}

// off-diagonal couplings cA(2..8)  -- :352-609
__global__ void k_coef(MatrixView M) {
  const int nz = M.nz, ny = M.ny, nx = M.nx;
  const size_t n = (size_t)nz * (ny + 2) * (nx + 2);
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int K = 1 + (int)(e % nz); size_t t = e / nz; const int j = (int)(t % (ny + 2)), i = (int)(t / (ny + 2));
  MAcc A{nz, ny, nx, M.dx, M.dy, M.zr, M.zw, M.cw, M.rmask, M.bmask};
  const double qrt = 0.25, hlf = 0.5;
  const bool jt = (i >= 1 && i <= nx && j >= 1 && j <= ny + 1);
  const bool it = (i >= 1 && i <= nx + 1 && j >= 1 && j <= ny);
  if (K >= 2 && i >= 1 && i <= nx && j >= 1 && j <= ny) {
    double v = A.CW(K, j, i);
    if (M.bmask && K <= nz - 1)     // :497-509, i-k and j-k cross terms at a lateral wall (not at k = nz in the reference)
      v = v - qrt * (A.ZXDY(K - 1, j, i) - A.ZXDY(K, j, i)) * (A.UM(j, i + 1) - A.UM(j, i))
            - qrt * (A.ZYDX(K - 1, j, i) - A.ZYDX(K, j, i)) * (A.VM(j + 1, i) - A.VM(j, i));
    M.cA[1][e] = v;
  }
  if (jt) {
    const double vm = A.VM(j, i);
    if (K <= nz - 1) M.cA[2][e] = qrt * (A.ZYDX(K + 1, j, i) + A.ZYDX(K, j - 1, i)) * vm;
    const double base = A.ARY(K, j, i) / (hlf * (A.DY(j, i) + A.DY(j - 1, i)));
    if (K == 1) {
      double v = base - (sq(A.ZYDX(K, j, i)) / A.CWS(j, i) + sq(A.ZYDX(K, j - 1, i)) / A.CWS(j - 1, i)) - qrt * (A.ZYDX(K, j - 1, i) - A.ZYDX(K, j, i));
      if (M.bmask)    // :375-389 ; the second slope is written /dy*dx in the reference (kept)
        v = (v - (hlf * A.ZXDY(K, j - 1, i) * A.ZYDX(K, j - 1, i) / A.CWS(j - 1, i) * (A.UM(j - 1, i + 1) - A.UM(j - 1, i))
                  - hlf * (((hlf * (A.ZR(K, j, i + 1) - A.ZR(K, j, i - 1))) / A.DY(j, i)) * A.DX(j, i)) * A.ZYDX(K, j, i) / A.CWS(j, i) * (A.UM(j, i + 1) - A.UM(j, i)))) * vm;
      M.cA[3][e] = v;
    } else if (K == nz)
      M.cA[3][e] = (base + qrt * (-A.ZYDX(K, j - 1, i) + A.ZYDX(K, j, i))) * vm;
    else
      M.cA[3][e] = base * vm;
    if (K >= 2) M.cA[4][e] = -qrt * (A.ZYDX(K - 1, j, i) + A.ZYDX(K, j - 1, i)) * vm;
  }
  if (it) {
    const double um = A.UM(j, i);
    if (K <= nz - 1) M.cA[5][e] = qrt * (A.ZXDY(K + 1, j, i) + A.ZXDY(K, j, i - 1)) * um;
    const double base = A.ARX(K, j, i) / (hlf * (A.DX(j, i) + A.DX(j, i - 1)));
    if (K == 1) {
      double v = base - (sq(A.ZXDY(K, j, i)) / A.CWS(j, i) + sq(A.ZXDY(K, j, i - 1)) / A.CWS(j, i - 1)) - qrt * (A.ZXDY(K, j, i - 1) - A.ZXDY(K, j, i));
      if (M.bmask)    // :417-433
        v = (v - (hlf * A.ZXDY(K, j, i - 1) * A.ZYDX(K, j, i - 1) / A.CWS(j, i - 1) * (A.VM(j + 1, i - 1) - A.VM(j, i - 1))
                  - hlf * (((hlf * (A.ZR(K, j, i + 1) - A.ZR(K, j, i - 1))) / A.DY(j, i)) * A.DX(j, i)) * A.ZYDX(K, j, i) / A.CWS(j, i) * (A.VM(j + 1, i) - A.VM(j, i)))) * um;
      M.cA[6][e] = v;
    } else if (K == nz)
      M.cA[6][e] = (base + qrt * (-A.ZXDY(K, j, i - 1) + A.ZXDY(K, j, i))) * um;
    else
      M.cA[6][e] = base * um;
    if (K >= 2) M.cA[7][e] = -qrt * (A.ZXDY(K - 1, j, i) + A.ZXDY(K, j, i - 1)) * um;
  }
  if (K == 1) {
    const size_t c2d = (size_t)j + (size_t)(ny + 2) * i;
    if (i >= 1 && i <= nx + 1 && j >= 0 && j <= ny) {         // couples (1,j,i) <-> (1,j+1,i-1)
      const double v = hlf * A.ZXDY(1, j + 1, i) * A.ZYDX(1, j + 1, i) / A.CWS(j + 1, i) * A.UM(j + 1, i) * A.VM(j + 1, i)
                     + hlf * A.ZXDY(1, j, i - 1) * A.ZYDX(1, j, i - 1) / A.CWS(j, i - 1) * A.UM(j, i) * A.VM(j + 1, i - 1);
      M.cA[4][e] = v; M.c5b[c2d] = v;
    }
    if (i >= 1 && i <= nx + 1 && j >= 1 && j <= ny + 1) {     // couples (1,j,i) <-> (1,j-1,i-1)
      const double v = -(hlf * A.ZXDY(1, j - 1, i) * A.ZYDX(1, j - 1, i) / A.CWS(j - 1, i) * A.UM(j - 1, i) * A.VM(j, i))
                       - hlf * A.ZXDY(1, j, i - 1) * A.ZYDX(1, j, i - 1) / A.CWS(j, i - 1) * A.UM(j, i) * A.VM(j, i - 1);
      M.cA[7][e] = v; M.c8b[c2d] = v;
    }
  }
}

// diagonal cA(1)  -- :616-657 ; thread per interior point
__global__ void k_diag(MatrixView M) {
  const int nz = M.nz, ny = M.ny, nx = M.nx, ny2 = ny + 2;
  const size_t n = (size_t)nz * ny * nx;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int k = (int)(e % nz); size_t t = e / nz; const int j = 1 + (int)(t % ny), i = 1 + (int)(t / ny);
  const size_t idx = (size_t)k + (size_t)nz * ((size_t)j + (size_t)ny2 * i);
  const int sj = nz; const size_t si = (size_t)nz * ny2;
  const double *c2 = M.cA[1], *c3 = M.cA[2], *c4 = M.cA[3], *c5 = M.cA[4], *c6 = M.cA[5], *c7 = M.cA[6], *c8 = M.cA[7];
  double v;
  if (k == 0) {
    v = -c2[idx + 1] - c4[idx] - c4[idx + sj] - c7[idx] - c7[idx + si]
        - c6[idx] - c8[idx + si + 1] - c3[idx] - c5[idx + sj + 1]
        - c5[idx] - c5[idx - sj + si] - c8[idx] - c8[idx + sj + si];
  } else if (k < nz - 1) {
    v = -c2[idx] - c2[idx + 1] - c4[idx] - c4[idx + sj] - c7[idx] - c7[idx + si]
        - c6[idx] - c6[idx + si - 1] - c8[idx] - c8[idx + si + 1]
        - c3[idx] - c3[idx + sj - 1] - c5[idx] - c5[idx + sj + 1];
  } else {
    MAcc A{nz, ny, nx, M.dx, M.dy, M.zr, M.zw, M.cw, M.rmask, M.bmask};
    const int K = nz; const double hlf = 0.5;
    v = -c2[idx] - A.CW(K + 1, j, i)
        + hlf * (hlf * (A.ZR(K, j, i + 2) - A.ZR(K, j, i)) / A.DX(j, i + 1)) * A.DY(j, i + 1)
        - hlf * (hlf * (A.ZR(K, j, i) - A.ZR(K, j, i - 2)) / A.DX(j, i - 1)) * A.DY(j, i - 1)
        + hlf * (hlf * (A.ZR(K, j + 2, i) - A.ZR(K, j, i)) / A.DY(j + 1, i)) * A.DX(j + 1, i)
        - hlf * (hlf * (A.ZR(K, j, i) - A.ZR(K, j - 2, i)) / A.DY(j - 1, i)) * A.DX(j - 1, i)
        - c4[idx] - c4[idx + sj] - c7[idx] - c7[idx + si]
        - c6[idx + si - 1] - c8[idx] - c3[idx + sj - 1] - c5[idx];
  }
  M.cA[0][idx] = v;
}

// per interior column: reciprocal Thomas pivots `bet` and g = T^-1 e1 (same recurrences as tridiag, src/mg_relax.f90:308-334)
__global__ void k_thomas_setup(MatrixView M) {
  const int nz = M.nz, ny = M.ny, nx = M.nx, ny2 = ny + 2;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ny * nx) return;
  const int j = 1 + e % ny, i = 1 + e / ny;
  const size_t o = (size_t)nz * ((size_t)j + (size_t)ny2 * i);
  const double* d = M.cA[0] + o; const double* c2 = M.cA[1] + o;    // ud(k) = cA(2,k+1) = c2[k+1] (0-based k)
  double* bet = M.bet + o; double* g = M.g + o;
  double b = 1.0 / d[0];
  bet[0] = b;
  double x = 1.0 * b;
  g[0] = x;
  for (int k = 1; k < nz; ++k) {
    const double dd = c2[k];
    const double gam = dd * b;
    b = 1.0 / (d[k] - dd * gam);
    bet[k] = b;
    x = (0.0 - dd * x) * b;
    g[k] = x;
  }
  for (int k = nz - 2; k >= 0; --k) { x = g[k] - (c2[k + 1] * bet[k]) * x; g[k] = x; }
  M.g1[(size_t)j + (size_t)ny2 * i] = x;
}

void launch_define_matrix_coef(cudaStream_t st, const MatrixView& M) {
  const size_t np1 = (size_t)(M.nz + 1) * (M.ny + 2) * (M.nx + 2), n = (size_t)M.nz * (M.ny + 2) * (M.nx + 2);
  k_cw_metrics<<<blocks_for(np1, 256), 256, 0, st>>>(M); ++g_launches;
  k_coef<<<blocks_for(n, 256), 256, 0, st>>>(M); ++g_launches;
}
void launch_define_matrix_diag(cudaStream_t st, const MatrixView& M) {
  k_diag<<<blocks_for((size_t)M.nz * M.ny * M.nx, 256), 256, 0, st>>>(M); ++g_launches;
  k_thomas_setup<<<blocks_for((size_t)M.ny * M.nx, 128), 128, 0, st>>>(M); ++g_launches;
}
void launch_define_matrix(cudaStream_t st, const MatrixView& M) { launch_define_matrix_coef(st, M); launch_define_matrix_diag(st, M); }
__global__ void k_level_rmask(double* __restrict__ rm, int ny, int nx, Halo1Flags f, int bmask) {
  const int ny2 = ny + 2;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ny2 * (nx + 2)) return;
  const int j = e % ny2, i = e / ny2;
  double v = 1.0;
  if (bmask && ((j == 0 && !f.south) || (i == nx + 1 && !f.east) || (j == ny + 1 && !f.north) || (i == 0 && !f.west))) v = 0.0;
  rm[e] = v;
}
void launch_level_rmask(cudaStream_t st, double* rm, int ny, int nx, Halo1Flags f, int bmask) {
  k_level_rmask<<<blocks_for((size_t)(ny + 2) * (nx + 2), 256), 256, 0, st>>>(rm, ny, nx, f, bmask); ++g_launches;
}

__global__ void k_copy_bottom(MatrixView M) {
  const size_t n = (size_t)(M.ny + 2) * (M.nx + 2);
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  M.c5b[e] = M.cA[4][e * M.nz]; M.c8b[e] = M.cA[7][e * M.nz];
}
void launch_refresh_derived(cudaStream_t st, const MatrixView& M) {
  k_copy_bottom<<<blocks_for((size_t)(M.ny + 2) * (M.nx + 2), 256), 256, 0, st>>>(M); ++g_launches;
  k_thomas_setup<<<blocks_for((size_t)M.ny * M.nx, 128), 128, 0, st>>>(M); ++g_launches;
}

// contraction factor of the bottom recurrence: rho = max over interior columns of |g1 c5b| + |g1 c8b|
__global__ void __launch_bounds__(1024) k_rho(MatrixView M, double* out) {
  __shared__ double s_red[32];
  const int ny = M.ny, nx = M.nx, ny2 = ny + 2;
  double v = 0.0;
  for (int e = threadIdx.x; e < ny * nx; e += blockDim.x) {
    const int j = 1 + e % ny, i = 1 + e / ny; const size_t c = (size_t)j + (size_t)ny2 * i;
    v = fmax(v, fabs(M.g1[c] * M.c5b[c]) + fabs(M.g1[c] * M.c8b[c]));
  }
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = s_red[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) *out = v;
  }
}
void launch_rho(cudaStream_t st, const MatrixView& M, double* out_dev) { k_rho<<<1, 1024, 0, st>>>(M, out_dev); ++g_launches; }

__global__ void k_any_nonzero2(const double* __restrict__ a, const double* __restrict__ b, size_t n, int* flag) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  int any = 0;
  for (; e < n; e += stride) any |= (a[e] != 0.0) | (b[e] != 0.0);
  if (any) atomicOr(flag, 1);
}
void launch_any_nonzero2(cudaStream_t st, const double* a, const double* b, size_t n, int* flag_dev) {
  cudaMemsetAsync(flag_dev, 0, sizeof(int), st);
  unsigned nb = blocks_for(n, 256); if (nb > 1024) nb = 1024;
  k_any_nonzero2<<<nb, 256, 0, st>>>(a, b, n, flag_dev); ++g_launches;
}

// ---------------------------------------------------------------------------------------------------------
// layout transposes between the model's (i,j,k) i-fastest arrays and the solver's k-fastest arrays
// ---------------------------------------------------------------------------------------------------------
__global__ void k_transpose_ijk_to_kji(const double* __restrict__ src, double* __restrict__ dst, int ni, int nj, int nk) {
  __shared__ double tile[32][33];
  const int j = blockIdx.z;
  const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, i = i0 + threadIdx.x;
    if (k < nk && i < ni) tile[r][threadIdx.x] = src[(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k)];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, k = k0 + threadIdx.x;
    if (k < nk && i < ni) dst[(size_t)k + (size_t)nk * ((size_t)j + (size_t)nj * i)] = tile[threadIdx.x][r];
  }
}
__global__ void k_transpose_kji_to_ijk(const double* __restrict__ src, double* __restrict__ dst, int ni, int nj, int nk) {
  __shared__ double tile[32][33];
  const int j = blockIdx.z;
  const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, k = k0 + threadIdx.x;
    if (k < nk && i < ni) tile[r][threadIdx.x] = src[(size_t)k + (size_t)nk * ((size_t)j + (size_t)nj * i)];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, i = i0 + threadIdx.x;
    if (k < nk && i < ni) dst[(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k)] = tile[threadIdx.x][r];
  }
}
void launch_transpose_ijk_to_kji(cudaStream_t st, const double* src, double* dst, int ni, int nj, int nk) {
  dim3 grid((ni + 31) / 32, (nk + 31) / 32, nj), block(32, 8);
  k_transpose_ijk_to_kji<<<grid, block, 0, st>>>(src, dst, ni, nj, nk); ++g_launches;
}
void launch_transpose_kji_to_ijk(cudaStream_t st, const double* src, double* dst, int ni, int nj, int nk) {
  dim3 grid((ni + 31) / 32, (nk + 31) / 32, nj), block(32, 8);
  k_transpose_kji_to_ijk<<<grid, block, 0, st>>>(src, dst, ni, nj, nk); ++g_launches;
}

// ---------------------------------------------------------------------------------------------------------
// compute_rhs (src/mg_compute_rhs.f90:14-379); umask = vmask = 1 unless bmask (:53-72)
// ---------------------------------------------------------------------------------------------------------
struct RAcc {
  int nz, ny, nx; const double *dx, *dy, *zw, *dzw, *zxdy, *zydx, *cw, *rmask, *ut, *vt, *wt; int bmask;
  __device__ __forceinline__ double UM(int j, int i) const { if (!bmask) return 1.0; return (i >= 1 && i <= nx + 1 && j >= 0 && j <= ny + 1) ? RM(j, i - 1) * RM(j, i) : 0.0; }
  __device__ __forceinline__ double VM(int j, int i) const { if (!bmask) return 1.0; return (i >= 0 && i <= nx + 1 && j >= 1 && j <= ny + 1) ? RM(j - 1, i) * RM(j, i) : 0.0; }
  __device__ __forceinline__ double DX(int j, int i) const { return dx[(size_t)j + (size_t)(ny + 2) * i]; }
  __device__ __forceinline__ double DY(int j, int i) const { return dy[(size_t)j + (size_t)(ny + 2) * i]; }
  __device__ __forceinline__ double RM(int j, int i) const { return rmask[(size_t)j + (size_t)(ny + 2) * i]; }
  __device__ __forceinline__ double ZW(int K, int j, int i) const { return zw[(size_t)(K - 1) + (size_t)(nz + 1) * ((size_t)(j + 1) + (size_t)(ny + 4) * (size_t)(i + 1))]; }
  __device__ __forceinline__ double DZW(int K, int j, int i) const { return dzw[(size_t)(K - 1) + (size_t)(nz + 1) * ((size_t)j + (size_t)(ny + 2) * i)]; }
  __device__ __forceinline__ double CW(int K, int j, int i) const { return cw[(size_t)(K - 1) + (size_t)(nz + 1) * ((size_t)j + (size_t)(ny + 2) * i)]; }
  __device__ __forceinline__ double ZX(int K, int j, int i) const { return zxdy[(size_t)(K - 1) + (size_t)nz * ((size_t)j + (size_t)(ny + 2) * i)]; }
  __device__ __forceinline__ double ZY(int K, int j, int i) const { return zydx[(size_t)(K - 1) + (size_t)nz * ((size_t)j + (size_t)(ny + 2) * i)]; }
  __device__ __forceinline__ double U(int i, int j, int K) const { return ut[(size_t)(K - 1) + (size_t)nz * ((size_t)j + (size_t)(ny + 2) * (size_t)(i - 1))]; }
  __device__ __forceinline__ double V(int i, int j, int K) const { return vt[(size_t)(K - 1) + (size_t)nz * ((size_t)(j - 1) + (size_t)(ny + 1) * (size_t)i)]; }
  __device__ __forceinline__ double W(int i, int j, int Kw) const { return wt[(size_t)Kw + (size_t)(nz + 1) * ((size_t)j + (size_t)(ny + 2) * (size_t)i)]; }
};
__device__ __forceinline__ RAcc racc(const RhsView& R) { return RAcc{R.nz, R.ny, R.nx, R.dx, R.dy, R.zw, R.dzw, R.zxdy, R.zydx, R.cw, R.rmask, R.ut, R.vt, R.wt, R.bmask}; }

// uf(k,j,i), i = 1..nx+1, j = 1..ny  -> flux(k,j,i) with stride nz ; wall faces zeroed (lbc_null='u')
__global__ void k_rhs_uf(RhsView R) {
  const int nz = R.nz, ny = R.ny, nx = R.nx;
  const size_t n = (size_t)nz * ny * (nx + 1);
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int k = 1 + (int)(e % nz); size_t t = e / nz; const int j = 1 + (int)(t % ny), i = 1 + (int)(t / ny);
  const RAcc A = racc(R);
  const double qrt = 0.25, hlf = 0.5, two = 2.0;
  double v;
  const double arx_u = qrt * (A.ZW(k + 1, j, i) - A.ZW(k, j, i) + A.ZW(k + 1, j, i - 1) - A.ZW(k, j, i - 1)) * (A.DY(j, i) + A.DY(j, i - 1)) * A.U(i, j, k);
  if (k == 1) {
    v = (arx_u
         - qrt * (+A.ZX(k, j, i) * A.DZW(k + 1, j, i) * A.W(i, j, k + 1 - 1) * A.RM(j, i) + A.ZX(k, j, i - 1) * A.DZW(k + 1, j, i - 1) * A.W(i - 1, j, k + 1 - 1) * A.RM(j, i - 1))
         - (+A.ZX(k, j, i) * A.ZX(k, j, i) / (A.CW(k, j, i) + A.CW(k + 1, j, i)) + A.ZX(k, j, i - 1) * A.ZX(k, j, i - 1) / (A.CW(k, j, i - 1) + A.CW(k + 1, j, i - 1)))
               * (hlf * (A.DX(j, i) + A.DX(j, i - 1))) * A.U(i, j, k)
         - (+A.ZX(k, j, i) * A.ZY(k, j, i) / (A.CW(k, j, i) + A.CW(k + 1, j, i))
                  * hlf * (hlf * (A.DY(j, i) + A.DY(j - 1, i)) * A.V(i, j, k) * A.VM(j, i) + hlf * (A.DY(j + 1, i) + A.DY(j, i)) * A.V(i, j + 1, k) * A.VM(j + 1, i))
            + A.ZX(k, j, i - 1) * A.ZY(k, j, i - 1) / (A.CW(k, j, i - 1) + A.CW(k + 1, j, i - 1))
                  * hlf * (hlf * (A.DY(j, i - 1) + A.DY(j - 1, i - 1)) * A.V(i - 1, j, k) * A.VM(j, i - 1) + hlf * (A.DY(j + 1, i - 1) + A.DY(j, i - 1)) * A.V(i - 1, j + 1, k) * A.VM(j + 1, i - 1)))) * A.UM(j, i);
  } else if (k < nz) {
    v = (arx_u
         - qrt * (+A.ZX(k, j, i) * A.DZW(k, j, i) * A.W(i, j, k - 1) * A.RM(j, i) + A.ZX(k, j, i) * A.DZW(k + 1, j, i) * A.W(i, j, k + 1 - 1) * A.RM(j, i)
                  + A.ZX(k, j, i - 1) * A.DZW(k, j, i - 1) * A.W(i - 1, j, k - 1) * A.RM(j, i - 1) + A.ZX(k, j, i - 1) * A.DZW(k + 1, j, i - 1) * A.W(i - 1, j, k + 1 - 1) * A.RM(j, i - 1))) * A.UM(j, i);
  } else {
    v = (arx_u
         - qrt * (+A.ZX(k, j, i) * A.DZW(k, j, i) * A.W(i, j, k - 1) * A.RM(j, i) + A.ZX(k, j, i) * two * A.DZW(k + 1, j, i) * A.W(i, j, k + 1 - 1) * A.RM(j, i)
                  + A.ZX(k, j, i - 1) * A.DZW(k, j, i - 1) * A.W(i - 1, j, k - 1) * A.RM(j, i - 1) + A.ZX(k, j, i - 1) * two * A.DZW(k + 1, j, i - 1) * A.W(i - 1, j, k + 1 - 1) * A.RM(j, i - 1))) * A.UM(j, i);
  }
  if ((i == 1 && R.west_phys) || (i == nx + 1 && R.east_phys)) v = 0.0;
  R.flux[(size_t)(k - 1) + (size_t)nz * ((size_t)j + (size_t)(ny + 2) * i)] = v;
}
// vf(k,j,i), i = 1..nx, j = 1..ny+1
__global__ void k_rhs_vf(RhsView R) {
  const int nz = R.nz, ny = R.ny, nx = R.nx;
  const size_t n = (size_t)nz * (ny + 1) * nx;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int k = 1 + (int)(e % nz); size_t t = e / nz; const int j = 1 + (int)(t % (ny + 1)), i = 1 + (int)(t / (ny + 1));
  const RAcc A = racc(R);
  const double qrt = 0.25, hlf = 0.5, two = 2.0;
  double v;
  const double ary_v = qrt * (A.ZW(k + 1, j, i) - A.ZW(k, j, i) + A.ZW(k + 1, j - 1, i) - A.ZW(k, j - 1, i)) * (A.DX(j, i) + A.DX(j - 1, i)) * A.V(i, j, k);
  if (k == 1) {
    v = (ary_v
         - qrt * (+A.ZY(k, j, i) * A.DZW(k + 1, j, i) * A.W(i, j, k + 1 - 1) * A.RM(j, i) + A.ZY(k, j - 1, i) * A.DZW(k + 1, j - 1, i) * A.W(i, j - 1, k + 1 - 1) * A.RM(j - 1, i))
         - (+A.ZY(k, j, i) * A.ZY(k, j, i) / (A.CW(k, j, i) + A.CW(k + 1, j, i)) + A.ZY(k, j - 1, i) * A.ZY(k, j - 1, i) / (A.CW(k, j - 1, i) + A.CW(k + 1, j - 1, i)))
               * hlf * (A.DY(j, i) + A.DY(j - 1, i)) * A.V(i, j, k)
         - (+A.ZX(k, j, i) * A.ZY(k, j, i) / (A.CW(k, j, i) + A.CW(k + 1, j, i))
                  * hlf * (hlf * (A.DX(j, i) + A.DX(j, i - 1)) * A.U(i, j, k) * A.UM(j, i) + hlf * (A.DX(j, i + 1) + A.DX(j, i)) * A.U(i + 1, j, k) * A.UM(j, i + 1))
            + A.ZX(k, j - 1, i) * A.ZY(k, j - 1, i) / (A.CW(k, j - 1, i) + A.CW(k + 1, j - 1, i))
                  * hlf * (hlf * (A.DX(j - 1, i) + A.DX(j - 1, i - 1)) * A.U(i, j - 1, k) * A.UM(j - 1, i) + hlf * (A.DX(j - 1, i + 1) + A.DX(j - 1, i)) * A.U(i + 1, j - 1, k) * A.UM(j - 1, i + 1)))) * A.VM(j, i);
  } else if (k < nz) {
    v = (ary_v
         - qrt * (+A.ZY(k, j, i) * A.DZW(k, j, i) * A.W(i, j, k - 1) * A.RM(j, i) + A.ZY(k, j, i) * A.DZW(k + 1, j, i) * A.W(i, j, k + 1 - 1) * A.RM(j, i)
                  + A.ZY(k, j - 1, i) * A.DZW(k, j - 1, i) * A.W(i, j - 1, k - 1) * A.RM(j - 1, i) + A.ZY(k, j - 1, i) * A.DZW(k + 1, j - 1, i) * A.W(i, j - 1, k + 1 - 1) * A.RM(j - 1, i))) * A.VM(j, i);
  } else {
    v = (ary_v
         - qrt * (+A.ZY(k, j, i) * A.DZW(k, j, i) * A.W(i, j, k - 1) * A.RM(j, i) + A.ZY(k, j, i) * two * A.DZW(k + 1, j, i) * A.W(i, j, k + 1 - 1) * A.RM(j, i)
                  + A.ZY(k, j - 1, i) * A.DZW(k, j - 1, i) * A.W(i, j - 1, k - 1) * A.RM(j - 1, i) + A.ZY(k, j - 1, i) * two * A.DZW(k + 1, j - 1, i) * A.W(i, j - 1, k + 1 - 1) * A.RM(j - 1, i))) * A.VM(j, i);
  }
  if ((j == 1 && R.south_phys) || (j == ny + 1 && R.north_phys)) v = 0.0;
  R.flux[(size_t)(k - 1) + (size_t)nz * ((size_t)j + (size_t)(ny + 2) * i)] = v;
}
// wf(k,j,i), k = 1..nz+1, interior columns -> flux with stride nz+1
__global__ void k_rhs_wf(RhsView R) {
  const int nz = R.nz, ny = R.ny, nx = R.nx, nzp = nz + 1;
  const size_t n = (size_t)nzp * ny * nx;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int k = 1 + (int)(e % nzp); size_t t = e / nzp; const int j = 1 + (int)(t % ny), i = 1 + (int)(t / ny);
  const RAcc A = racc(R);
  const double qrt = 0.25, hlf = 0.5;
  double v;
  if (k == 1) {
    v = 0.0;
  } else if (k <= nz) {
    v = A.CW(k, j, i) * A.DZW(k, j, i) * A.W(i, j, k - 1)
        - qrt * hlf * (+A.ZX(k, j, i) * (A.DX(j, i) + A.DX(j, i - 1)) * A.U(i, j, k) * A.UM(j, i) + A.ZX(k, j, i) * (A.DX(j, i + 1) + A.DX(j, i)) * A.U(i + 1, j, k) * A.UM(j, i + 1)
                       + A.ZX(k - 1, j, i) * (A.DX(j, i) + A.DX(j, i - 1)) * A.U(i, j, k - 1) * A.UM(j, i) + A.ZX(k - 1, j, i) * (A.DX(j, i + 1) + A.DX(j, i)) * A.U(i + 1, j, k - 1) * A.UM(j, i + 1));
    v = v
        - qrt * hlf * (+A.ZY(k, j, i) * (A.DY(j, i) + A.DY(j - 1, i)) * A.V(i, j, k) * A.VM(j, i) + A.ZY(k, j, i) * (A.DY(j + 1, i) + A.DY(j, i)) * A.V(i, j + 1, k) * A.VM(j + 1, i)
                       + A.ZY(k - 1, j, i) * (A.DY(j, i) + A.DY(j - 1, i)) * A.V(i, j, k - 1) * A.VM(j, i) + A.ZY(k - 1, j, i) * (A.DY(j + 1, i) + A.DY(j, i)) * A.V(i, j + 1, k - 1) * A.VM(j + 1, i));
  } else {
    v = A.CW(k, j, i) * A.DZW(k, j, i) * A.W(i, j, k - 1)
        - hlf * hlf * (+A.ZX(k - 1, j, i) * (A.DX(j, i) + A.DX(j, i - 1)) * A.U(i, j, k - 1) * A.UM(j, i) + A.ZX(k - 1, j, i) * (A.DX(j, i + 1) + A.DX(j, i)) * A.U(i + 1, j, k - 1) * A.UM(j, i + 1))
        - hlf * hlf * (+A.ZY(k - 1, j, i) * (A.DY(j, i) + A.DY(j - 1, i)) * A.V(i, j, k - 1) * A.VM(j, i) + A.ZY(k - 1, j, i) * (A.DY(j + 1, i) + A.DY(j, i)) * A.V(i, j + 1, k - 1) * A.VM(j + 1, i));
  }
  R.flux[(size_t)(k - 1) + (size_t)nzp * ((size_t)j + (size_t)(ny + 2) * i)] = v;
}
__global__ void k_rhs_div(RhsView R, int dir) {
  const int nz = R.nz, ny = R.ny, nx = R.nx, ny2 = ny + 2;
  const size_t n = (size_t)nz * ny * nx;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int k = (int)(e % nz); size_t t = e / nz; const int j = 1 + (int)(t % ny), i = 1 + (int)(t / ny);
  const size_t idx = (size_t)k + (size_t)nz * ((size_t)j + (size_t)ny2 * i);
  if (dir == 0) R.b[idx] = R.flux[idx + (size_t)nz * ny2] - R.flux[idx];
  else if (dir == 1) R.b[idx] = R.b[idx] + R.flux[idx + nz] - R.flux[idx];
  else { const size_t w = (size_t)k + (size_t)(nz + 1) * ((size_t)j + (size_t)ny2 * i); R.b[idx] = R.b[idx] + R.flux[w + 1] - R.flux[w]; }
}
void launch_rhs_uf(cudaStream_t st, const RhsView& R) { size_t n = (size_t)R.nz * R.ny * (R.nx + 1); k_rhs_uf<<<blocks_for(n, 128), 128, 0, st>>>(R); ++g_launches; }
void launch_rhs_vf(cudaStream_t st, const RhsView& R) { size_t n = (size_t)R.nz * (R.ny + 1) * R.nx; k_rhs_vf<<<blocks_for(n, 128), 128, 0, st>>>(R); ++g_launches; }
void launch_rhs_wf(cudaStream_t st, const RhsView& R) { size_t n = (size_t)(R.nz + 1) * R.ny * R.nx; k_rhs_wf<<<blocks_for(n, 128), 128, 0, st>>>(R); ++g_launches; }
void launch_rhs_div(cudaStream_t st, const RhsView& R, int dir) { size_t n = (size_t)R.nz * R.ny * R.nx; k_rhs_div<<<blocks_for(n, 256), 256, 0, st>>>(R, dir); ++g_launches; }

// ---------------------------------------------------------------------------------------------------------
// correct_uvw (src/mg_correct_uvw.f90:68-110): u,v,w -= grad p  (k-fastest copies of the model arrays)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_correct_uvw(CorrView C) {
  const int nz = C.nz, ny = C.ny, nx = C.nx, ny2 = ny + 2, nzp = nz + 1;
  const size_t ncol = (size_t)ny2 * (nx + 2);
  const size_t n = (size_t)nzp * ncol;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int kk = (int)(e % nzp); size_t t = e / nzp; const int j = (int)(t % ny2), i = (int)(t / ny2);
  const double hlf = 0.5, one = 1.0;
  auto P = [&](int K, int jj, int ii) { return C.p[(size_t)(K - 1) + (size_t)nz * ((size_t)jj + (size_t)ny2 * ii)]; };
  auto DX = [&](int jj, int ii) { return C.dx[(size_t)jj + (size_t)ny2 * ii]; };
  auto DY = [&](int jj, int ii) { return C.dy[(size_t)jj + (size_t)ny2 * ii]; };
  auto ZR = [&](int K, int jj, int ii) { return C.zr[(size_t)(K - 1) + (size_t)nz * ((size_t)(jj + 1) + (size_t)(ny + 4) * (size_t)(ii + 1))]; };
  auto RM = [&](int jj, int ii) { return C.rmask[(size_t)jj + (size_t)ny2 * ii]; };
  auto UM = [&](int jj, int ii) -> double { if (!C.bmask) return 1.0; return (ii >= 1 && ii <= nx + 1 && jj >= 0 && jj <= ny + 1) ? RM(jj, ii - 1) * RM(jj, ii) : 0.0; };   // src/mg_correct_uvw.f90:48-68
  auto VM = [&](int jj, int ii) -> double { if (!C.bmask) return 1.0; return (ii >= 0 && ii <= nx + 1 && jj >= 1 && jj <= ny + 1) ? RM(jj - 1, ii) * RM(jj, ii) : 0.0; };
  if (kk >= 1) {
    const int K = kk;    // K = 1..nz
    if (i >= 1) {        // u(i,j,K), i = 1..nx+1, j = 0..ny+1
      const double dxu = hlf * (DX(j, i) + DX(j, i - 1));
      double* u = C.ut + ((size_t)(K - 1) + (size_t)nz * ((size_t)j + (size_t)ny2 * (size_t)(i - 1)));
      *u = *u - one / dxu * (P(K, j, i) - P(K, j, i - 1)) * UM(j, i);
    }
    if (j >= 1) {        // v(i,j,K), i = 0..nx+1, j = 1..ny+1
      const double dyv = hlf * (DY(j, i) + DY(j - 1, i));
      double* v = C.vt + ((size_t)(K - 1) + (size_t)nz * ((size_t)(j - 1) + (size_t)(ny + 1) * (size_t)i));
      *v = *v - one / dyv * (P(K, j, i) - P(K, j - 1, i)) * VM(j, i);
    }
  }
  // w(i,j,kw), kw = 1..nz ; (kw = 0, the bottom, is not corrected)
  if (kk >= 1) {
    const int kw = kk;
    double* w = C.wt + ((size_t)kw + (size_t)nzp * ((size_t)j + (size_t)ny2 * (size_t)i));
    if (kw < nz) {
      const int K = kw + 1;
      const double dzw = ZR(K, j, i) - ZR(K - 1, j, i);
      *w = *w - one / dzw * (P(K, j, i) - P(K - 1, j, i));
    } else {
      const double zwtop = C.zw[(size_t)nz + (size_t)nzp * ((size_t)(j + 1) + (size_t)(ny + 4) * (size_t)(i + 1))];
      const double dzw = zwtop - ZR(nz, j, i);
      *w = *w - one / dzw * (-P(nz, j, i));
    }
  }
}
void launch_correct_uvw(cudaStream_t st, const CorrView& C) {
  size_t n = (size_t)(C.nz + 1) * (C.ny + 2) * (C.nx + 2);
  k_correct_uvw<<<blocks_for(n, 256), 256, 0, st>>>(C); ++g_launches;
}

}  // namespace mgb

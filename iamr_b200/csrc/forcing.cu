// forcing.cu -- the HIT tutorial's turbulent forcing (Tutorials/HIT/NS_getForce.cpp:205-640 with the mode tables of
// TurbulentForcing_def.H:20-366): a sum over low-wavenumber Fourier modes with random phases, amplitudes and
// frequencies, optionally divergence free, added to the velocity forcing as rho * f(x, t).
//
// The reference evaluates 9 sin / cos pairs per mode in every cell (or, with USE_FAST_FORCE, on a coarse grid it then
// interpolates).  Every term is a product of one factor per direction, so this implementation tabulates the factors
// along the three axes of the box once per call (a few thousand sincos instead of billions) and the cell kernel
// multiplies table entries: x factors are coalesced loads, y / z factors are warp-uniform.  The full mode set of the
// reference's exact path is kept (no coarse-grid interpolation).
#include <cmath>
#include <vector>
#include "kernels.h"

namespace ix {
namespace k {

// host: select the active modes exactly as the reference's two loops do and fold the time factor into the amplitudes
// forcedata: TurbulentForcing::forcedata, 17 arrays of array_size^3, i fastest (FTX TAT FPX FPY FPZ FAX FAY FAZ FPXX FPXY FPXZ
// FPYX FPYY FPYZ FPZX FPZY FPZZ, TurbulentForcing_def.H:66-82)
int turb_modes(const TurbParams& tp, const double* fd, const double L[3], double time, std::vector<TurbMode>& out) {
  const int as = tp.array_size;
  const size_t ne = (size_t)as * as * as;
  auto at = [&](int arr, int kx, int ky, int kz) { return fd[(size_t)arr * ne + kx + (size_t)as * (ky + (size_t)as * kz)]; };
  const double Lmin = std::min(L[0], std::min(L[1], L[2]));
  const int xstep = (int)(L[0] / Lmin + 0.5), ystep = (int)(L[1] / Lmin + 0.5), zstep = (int)(L[2] / Lmin + 0.5);
  const double kappaMax = tp.nmodes / Lmin + 1.0e-8;
  const double twopi = 2.0 * M_PI;
  auto add = [&](int kx, int ky, int kz) -> int {
    if (kx >= as || ky >= as || kz >= as) return IAMRX_ERR_ARG;
    const double kappa = std::sqrt((kx * kx) / (L[0] * L[0]) + (ky * ky) / (L[1] * L[1]) + (kz * kz) / (L[2] * L[2]));
    if (!(kappa <= kappaMax)) return IAMRX_OK;
    TurbMode m{};
    const double xT = std::cos(at(0, kx, ky, kz) * time + at(1, kx, ky, kz));
    m.w[0] = twopi * kx / L[0]; m.w[1] = twopi * ky / L[1]; m.w[2] = twopi * kz / L[2];
    const double fax = at(5, kx, ky, kz), fay = at(6, kx, ky, kz), faz = at(7, kx, ky, kz);
    if (tp.div_free) {
      // phase[c][d]: component c = X, Y, Z of the vector potential, direction d (FPcd)
      for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) m.ph[c][d] = at(8 + 3 * c + d, kx, ky, kz);
      m.a[0] = xT * fax; m.a[1] = xT * fay; m.a[2] = xT * faz;
    } else {
      for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) m.ph[c][d] = at(2 + d, kx, ky, kz);   // FPX, FPY, FPZ for every component
      m.a[0] = xT * fax; m.a[1] = xT * fay; m.a[2] = xT * faz;
    }
    out.push_back(m);
    return IAMRX_OK;
  };
  out.clear();
  for (int kz = tp.mode_start * zstep; kz <= tp.nmodes * zstep; kz += zstep)
    for (int ky = tp.mode_start * ystep; ky <= tp.nmodes * ystep; ky += ystep)
      for (int kx = tp.mode_start * xstep; kx <= tp.nmodes * xstep; kx += xstep) { const int rc = add(kx, ky, kz); if (rc) return rc; }
  // high aspect ratio domains: the extra modes that break the symmetry at a low level (NS_getForce.cpp:436-440)
  for (int kz = 1; kz <= zstep - 1; ++kz)
    for (int ky = tp.mode_start; ky <= tp.nmodes * ystep; ++ky)
      for (int kx = tp.mode_start; kx <= tp.nmodes * xstep; ++kx) { const int rc = add(kx, ky, kz); if (rc) return rc; }
  return IAMRX_OK;
}

namespace {
constexpr int TX = 128, TY = 2;

// tab[((d * nm + m) * 3 + c) * 2 + s][q] over q = 0 .. len-1 (row pitch `pitch`): s = 0 sin, 1 cos of w_d * x_q + ph[c][d]
__global__ void turb_table_kernel(double* tab, const TurbMode* modes, int nm, int pitch, int l0, int l1, int l2, int n0, int n1, int n2,
                                  double x0, double y0, double z0, double hx, double hy, double hz) {
  const int row = blockIdx.y;                 // (d, m, c)
  const int c = row % 3, m = (row / 3) % nm, d = row / (3 * nm);
  const int len = d == 0 ? n0 : (d == 1 ? n1 : n2);
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= len) return;
  const double x = d == 0 ? x0 + hx * (l0 + q + 0.5) : (d == 1 ? y0 + hy * (l1 + q + 0.5) : z0 + hz * (l2 + q + 0.5));
  double sv, cv;
  sincos(modes[m].w[d] * x + modes[m].ph[c][d], &sv, &cv);
  tab[(size_t)(row * 2) * pitch + q] = sv;
  tab[(size_t)(row * 2 + 1) * pitch + q] = cv;
}

// frc(i,j,k,n) (+)= rho * f_n.  One thread per (i, j) and KP consecutive planes, all modes: the x and y factors of a mode are
// loaded once and combined into the six plane-independent products, each plane then costs five warp-uniform z loads and six
// FMAs.  T(d,m,c,s) = table row.
constexpr int KP = 4;
template <bool DIVFREE>
__global__ void __launch_bounds__(TX* TY)
turb_force_kernel(Bx bx, V4 frc, C4 rho, const double* __restrict__ tab, const TurbMode* __restrict__ modes, int nm, int pitch, int accumulate) {
  const int k0 = bx.lo[2] + KP * (int)blockIdx.z;
  const int j = bx.lo[1] + blockIdx.y * TY + threadIdx.y;
  const int i = bx.lo[0] + blockIdx.x * TX + threadIdx.x;
  if (j > bx.hi[1] || i > bx.hi[0]) return;
  const int qi = i - bx.lo[0], qj = j - bx.lo[1], qk0 = k0 - bx.lo[2];
  const int nk = (bx.hi[2] - k0 + 1 < KP) ? bx.hi[2] - k0 + 1 : KP;
  double f1[KP], f2[KP], f3[KP];
#pragma unroll
  for (int p = 0; p < KP; ++p) { f1[p] = 0.0; f2[p] = 0.0; f3[p] = 0.0; }
  for (int m = 0; m < nm; ++m) {
#define TT(d, c, s, q) tab[(size_t)((((d) * nm + m) * 3 + (c)) * 2 + (s)) * pitch + (q)]
    const double ax = modes[m].a[0], ay = modes[m].a[1], az = modes[m].a[2];
    if (DIVFREE) {
      const double wx = modes[m].w[0], wy = modes[m].w[1], wz = modes[m].w[2];
      // curl of the vector potential (A_x, A_y, A_z) sin sin sin with per-component phases (NS_getForce.cpp:376-401)
      const double sXx = TT(0, 0, 0, qi), sXy = TT(1, 0, 0, qj), cXy = TT(1, 0, 1, qj);
      const double sYx = TT(0, 1, 0, qi), cYx = TT(0, 1, 1, qi), sYy = TT(1, 1, 0, qj);
      const double sZx = TT(0, 2, 0, qi), cZx = TT(0, 2, 1, qi), sZy = TT(1, 2, 0, qj), cZy = TT(1, 2, 1, qj);
      const double p1a = az * wy * sZx * cZy, p1b = ay * wz * sYx * sYy;   // f1 = p1a sZz - p1b cYz
      const double p2a = ax * wz * sXx * sXy, p2b = az * wx * cZx * sZy;   // f2 = p2a cXz - p2b sZz
      const double p3a = ay * wx * cYx * sYy, p3b = ax * wy * sXx * cXy;   // f3 = p3a sYz - p3b sXz
#pragma unroll
      for (int p = 0; p < KP; ++p) {
        if (p < nk) {
          const int qk = qk0 + p;
          const double sXz = TT(2, 0, 0, qk), cXz = TT(2, 0, 1, qk), sYz = TT(2, 1, 0, qk), cYz = TT(2, 1, 1, qk), sZz = TT(2, 2, 0, qk);
          f1[p] += p1a * sZz - p1b * cYz;
          f2[p] += p2a * cXz - p2b * sZz;
          f3[p] += p3a * sYz - p3b * sXz;
        }
      }
    } else {
      const double sx = TT(0, 0, 0, qi), cx = TT(0, 0, 1, qi), sy = TT(1, 0, 0, qj), cy = TT(1, 0, 1, qj);
      const double q1 = ax * cx * sy, q2 = ay * sx * cy, q3 = az * sx * sy;
#pragma unroll
      for (int p = 0; p < KP; ++p) {
        if (p < nk) {
          const int qk = qk0 + p;
          const double sz = TT(2, 0, 0, qk), cz = TT(2, 0, 1, qk);
          f1[p] += q1 * sz;
          f2[p] += q2 * sz;
          f3[p] += q3 * cz;
        }
      }
    }
#undef TT
  }
#pragma unroll
  for (int p = 0; p < KP; ++p) {
    if (p < nk) {
      const int k = k0 + p;
      const double r = rho.ok() ? rho(i, j, k) : 1.0;
      if (accumulate) { frc(i, j, k, 0) += r * f1[p]; frc(i, j, k, 1) += r * f2[p]; frc(i, j, k, 2) += r * f3[p]; }
      else { frc(i, j, k, 0) = r * f1[p]; frc(i, j, k, 1) = r * f2[p]; frc(i, j, k, 2) = r * f3[p]; }
    }
  }
}
}  // namespace

// scratch: device buffer of at least turb_scratch_doubles(bx, nm) doubles (tables) -- plus the mode list, uploaded by the caller
size_t turb_scratch_doubles(const Bx& bx, int nm) {
  const int pitch = std::max(bx.nx(), std::max(bx.ny(), bx.nz()));
  return (size_t)3 * nm * 3 * 2 * pitch;
}

int turb_force(const Bx& bx, V4 frc, C4 rho, const iamrx_geom& g, const TurbMode* d_modes, int nm, int div_free, double* d_tab,
               int accumulate, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  if (nm <= 0) {
    if (!accumulate) return setval(bx, frc, 3, 0.0, s);
    return IAMRX_OK;
  }
  const int pitch = std::max(bx.nx(), std::max(bx.ny(), bx.nz()));
  IX_LAUNCH(turb_table_kernel, dim3(cdiv(pitch, 128), 3 * nm * 3, 1), dim3(128, 1, 1), 0, s, d_tab, d_modes, nm, pitch,
            bx.lo[0] - g.domain.lo[0], bx.lo[1] - g.domain.lo[1], bx.lo[2] - g.domain.lo[2], bx.nx(), bx.ny(), bx.nz(),
            g.prob_lo[0], g.prob_lo[1], g.prob_lo[2], g.dx[0], g.dx[1], g.dx[2]);
  int rc = check_launch("turb_table");
  if (rc) return rc;
  const dim3 grd(cdiv(bx.nx(), TX), cdiv(bx.ny(), TY), cdiv(bx.nz(), KP));
  if (div_free) IX_LAUNCH((turb_force_kernel<true>), grd, dim3(TX, TY, 1), 0, s, bx, frc, rho, d_tab, d_modes, nm, pitch, accumulate);
  else IX_LAUNCH((turb_force_kernel<false>), grd, dim3(TX, TY, 1), 0, s, bx, frc, rho, d_tab, d_modes, nm, pitch, accumulate);
  return check_launch("turb_force");
}

}  // namespace k
}  // namespace ix

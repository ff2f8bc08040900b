// dune-gdt_b200/csrc/fv_system.hpp -- first-order FV advection operator for SYSTEMS of conservation laws (m > 1): the
// Euler equations with the Vijayasundaram / Lax-Friedrichs numerical fluxes (fv_system.cu), and the host side of
// estimate_dt_for_hyperbolic_system for systems.
#pragma once
#include <algorithm>
#include <cmath>

#include "kernels.hpp"

namespace gdtb {

struct FvSysParams
{
  GridDev g;
  int m;       // components per cell (Euler: d + 2); DoF layout [cell][component] (spaces/mapper/finite-volume.hh:92-97)
  int numflux; // GDTB_NUMFLUX_VIJAYASUNDARAM / GDTB_NUMFLUX_LAX_FRIEDRICHS
  double gamma;
  double half_over_lambda; // Lax-Friedrichs: 0.5 / lambda (lax-friedrichs.hh:84)
  const double* inv_ext[3]; // 1 / cell extent per axis
  int euler;                // 0: out = L(u), 1: out = u - dt L(u)
  double dt;
  // impermeable walls on non-periodic domain sides (bit 2k + s): wall flux (0, p n, 0) / mirrored ghost state
  unsigned wall_mask, mirror_mask;
};

int launch_fvsys_apply(Launch& L, const FvSysParams& p, const double* u, double* out);
// per-component minimum / maximum of the state: partial[block][2 m] (min_0 .. min_{m-1}, max_0 .. max_{m-1})
int launch_fvsys_minmax(Launch& L, const FvSysParams& p, const double* u, double* partial, int blocks);

// max over the space directions of the infinity norm (largest absolute row sum) of the Euler flux jacobian at the
// state w (EulerTools<d>::flux_jacobian, tools/euler.hh:262-316; hyperbolic.hh:71-73)
inline double euler_jacobian_inf_norm(const int d, const double gamma, const double* w)
{
  const int m = d + 2;
  const double rho = w[0], E = w[m - 1];
  double v[2] = {0., 0.};
  for (int i = 0; i < d; ++i)
    v[i] = w[1 + i] / w[0];
  const double gamma_1 = gamma - 1.;
  const double vnorm2 = v[0] * v[0] + (d > 1 ? v[1] * v[1] : 0.);
  const double ek = 0.5 * vnorm2;
  double J[2][4][4] = {};
  if (d == 1) {
    J[0][0][0] = 0., J[0][0][1] = 1., J[0][0][2] = 0.;
    J[0][1][0] = gamma_1 * ek - v[0] * v[0], J[0][1][1] = (3. - gamma) * v[0], J[0][1][2] = gamma_1;
    J[0][2][0] = v[0] * (gamma_1 * vnorm2 - (gamma * E) / rho);
    J[0][2][1] = ((gamma * E) / rho) - gamma_1 * v[0] * v[0] - gamma_1 * ek;
    J[0][2][2] = gamma * v[0];
  } else {
    J[0][0][1] = 1.;
    J[0][1][0] = gamma_1 * ek - v[0] * v[0];
    J[0][1][1] = (3. - gamma) * v[0];
    J[0][1][2] = -1. * gamma_1 * v[1];
    J[0][1][3] = gamma_1;
    J[0][2][0] = -1. * v[0] * v[1];
    J[0][2][1] = v[1];
    J[0][2][2] = v[0];
    J[0][3][0] = v[0] * (gamma_1 * vnorm2 - (gamma * E) / rho);
    J[0][3][1] = ((gamma * E) / rho) - gamma_1 * v[0] * v[0] - gamma_1 * ek;
    J[0][3][2] = -1. * gamma_1 * v[0] * v[1];
    J[0][3][3] = gamma * v[0];
    J[1][0][2] = 1.;
    J[1][1][0] = -1. * v[0] * v[1];
    J[1][1][1] = v[1];
    J[1][1][2] = v[0];
    J[1][2][0] = 0.5 * gamma_1 * vnorm2 - v[1] * v[1];
    J[1][2][1] = -1. * gamma_1 * v[0];
    J[1][2][2] = (3. - gamma) * v[1];
    J[1][2][3] = gamma_1;
    J[1][3][0] = v[1] * (gamma_1 * vnorm2 - ((gamma * E) / rho));
    J[1][3][1] = -1. * gamma_1 * v[0] * v[1];
    J[1][3][2] = ((gamma * E) / rho) - gamma_1 * v[1] * v[1] - gamma_1 * ek;
    J[1][3][3] = gamma * v[1];
  }
  double ret = 0.;
  for (int s = 0; s < d; ++s)
    for (int r = 0; r < m; ++r) {
      double sum = 0.;
      for (int c = 0; c < m; ++c)
        sum += std::fabs(J[s][r][c]);
      ret = std::max(ret, sum);
    }
  return ret;
}

} // namespace gdtb

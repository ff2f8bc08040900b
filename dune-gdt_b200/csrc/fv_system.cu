// dune-gdt_b200/csrc/fv_system.cu -- AdvectionFvOperator::apply for a finite volume space with m > 1 components
// (make_finite_volume_space<m>, spaces/l2/finite-volume.hh:208-230): the Euler equations of the reference's 2d_euler driver
// (examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:381-434) and its 1D EOC test
// (test/inviscid-compressible-flow/base.hh), d = 1, 2.
//
// Replaces LocalizableOperator::apply (operators/localizable-operator.hh:352-387) with
// LocalAdvectionFvCouplingOperator::apply (local/operators/advection-fv.hh:127-153: g = numerical_flux(u, v, n), the
// inside element gets + g |I| / |E_in|, the outside one - g |I| / |E_out|, per component) and the numerical fluxes
//   NumericalVijayasundaramFlux::apply (local/numerical-fluxes/vijayasundaram.hh:111-133):  g = P^+ u + P^- v with
//     P^+- = T diag(max / min(lambda_i, 0)) T^{-1}, eigendecomposition of the flux jacobian at (u + v) / 2 from
//     EulerTools (tools/euler.hh:325-462, Kroener's M T and (M T)^{-1});
//   NumericalLaxFriedrichsFlux::apply (lax-friedrichs.hh:66-88) with the caller's lambda:
//     g = sum_s (f_s(u) + f_s(v)) n_s / 2 + (u - v) / (2 lambda).
// Formulation: cell gather like fv.cu -- a thread owns one cell, evaluates the numerical flux G along +e_k through its
// lower and upper face of every axis (inside = the cell with the smaller coordinate, normal +e_k) and writes
// sum_k (G_up - G_low) / ext_k once: deterministic, no read-modify-write.  On the periodic wrap face the reference's
// inside element is cell 0 with normal -e_k; P(w, -n) = -P(w, n) makes that the same flux up to rounding.
// P^+- u is evaluated as T (Lambda^+- (T^{-1} u)) (two m x m mat-vecs per state instead of two mat-mat products).
// State layout [cell][component]: a cell's state is one 24 / 32-byte record, neighbour records come through L1 / L2.
#include "fv_system.hpp"

#include "common.cuh"

namespace gdtb {

namespace {

template <int D>
struct EulerState
{
  static constexpr int M = D + 2;
  double rho, v[D], p, E, a, v2;
  __device__ __forceinline__ EulerState(const double gamma, const double (&w)[M])
  {
    rho = w[0];
    v2 = 0.;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      v[i] = w[1 + i] / w[0];
      v2 += v[i] * v[i];
    }
    E = w[M - 1];
    p = (gamma - 1.) * (E - 0.5 * rho * v2); // tools/euler.hh:136-147
    a = sqrt(gamma * p / rho);               // speed_of_sound (:170-178)
  }
};

// f_k(w) (EulerTools::flux, tools/euler.hh:212-236)
template <int D>
__device__ __forceinline__ void euler_flux_k(const double gamma, const double (&w)[D + 2], const int k, double (&f)[D + 2])
{
  constexpr int M = D + 2;
  const EulerState<D> s(gamma, w);
  f[0] = s.rho * s.v[k];
#pragma unroll
  for (int i = 0; i < D; ++i)
    f[1 + i] = s.rho * s.v[i] * s.v[k] + (i == k ? 1. : 0.) * s.p;
  f[M - 1] = (s.E + s.p) * s.v[k];
}

// g = P^+(wbar, e_k) u + P^-(wbar, e_k) v  (vijayasundaram.hh:111-133 with tools/euler.hh:325-462)
template <int D>
__device__ __forceinline__ void vijayasundaram_k(const double gamma, const double (&u)[D + 2], const double (&v)[D + 2],
                                                 const int k, double (&g)[D + 2])
{
  constexpr int M = D + 2;
  double w[M];
#pragma unroll
  for (int i = 0; i < M; ++i)
    w[i] = 0.5 * (u[i] + v[i]);
  const EulerState<D> s(gamma, w);
  const double n0 = k == 0 ? 1. : 0., n1 = k == 1 ? 1. : 0.;
  const double a = s.a, rho = s.rho;
  const double H = (s.E + s.p) / rho; // enthalpy (:197-205)
  const double rho_over_2a = rho / (2 * a);
  const double ek = 0.5 * s.v2;
  const double vn = s.v[k];
  const double mach = sqrt(s.v2) / a; // mach_number (:191-195)
  const double gamma_1 = gamma - 1.;
  const double gM = (gamma_1 / 2.) * mach * mach;
  double ev[M], T[M][M], Ti[M][M];
  if (D == 1) {
    ev[0] = vn, ev[1] = vn + a, ev[2] = vn - a;
    T[0][0] = 1., T[0][1] = rho_over_2a, T[0][2] = rho_over_2a;
    T[1][0] = s.v[0], T[1][1] = rho_over_2a * (s.v[0] + a * n0), T[1][2] = rho_over_2a * (s.v[0] - a * n0);
    T[2][0] = ek, T[2][1] = rho_over_2a * (H + a * vn), T[2][2] = rho_over_2a * (H - a * vn);
    Ti[0][0] = 1. - gM, Ti[0][1] = gamma_1 * s.v[0] / (a * a), Ti[0][2] = -gamma_1 / (a * a);
    Ti[1][0] = (a / rho) * (gM - vn / a), Ti[1][1] = (1. / rho) * (n0 - gamma_1 * (s.v[0] / a)), Ti[1][2] = gamma_1 / (rho * a);
    Ti[2][0] = (a / rho) * (gM + vn / a), Ti[2][1] = (-1. / rho) * (n0 + gamma_1 * (s.v[0] / a)), Ti[2][2] = gamma_1 / (rho * a);
  } else {
    const double v0 = s.v[0], v1 = s.v[D - 1];
    ev[0] = vn, ev[1] = vn, ev[2] = vn + a, ev[M - 1] = vn - a;
    constexpr int L = M - 1; // index 3
    T[0][0] = 1., T[0][1] = 0., T[0][2] = rho_over_2a, T[0][L] = rho_over_2a;
    T[1][0] = v0, T[1][1] = rho * n1, T[1][2] = rho_over_2a * (v0 + a * n0), T[1][L] = rho_over_2a * (v0 - a * n0);
    T[2][0] = v1, T[2][1] = -rho * n0, T[2][2] = rho_over_2a * (v1 + a * n1), T[2][L] = rho_over_2a * (v1 - a * n1);
    T[L][0] = ek, T[L][1] = rho * (v0 * n1 - v1 * n0), T[L][2] = rho_over_2a * (H + a * vn), T[L][L] = rho_over_2a * (H - a * vn);
    Ti[0][0] = 1. - gM, Ti[0][1] = gamma_1 * v0 / (a * a), Ti[0][2] = gamma_1 * v1 / (a * a), Ti[0][L] = -gamma_1 / (a * a);
    Ti[1][0] = (1. / rho) * (v1 * n0 - v0 * n1), Ti[1][1] = n1 / rho, Ti[1][2] = -n0 / rho, Ti[1][L] = 0.;
    Ti[2][0] = (a / rho) * (gM - vn / a), Ti[2][1] = (1. / rho) * (n0 - gamma_1 * (v0 / a));
    Ti[2][2] = (1. / rho) * (n1 - gamma_1 * (v1 / a)), Ti[2][L] = gamma_1 / (rho * a);
    Ti[L][0] = (a / rho) * (gM + vn / a), Ti[L][1] = (-1. / rho) * (n0 + gamma_1 * (v0 / a));
    Ti[L][2] = (-1. / rho) * (n1 + gamma_1 * (v1 / a)), Ti[L][L] = gamma_1 / (rho * a);
  }
  // c = Lambda^+ (T^{-1} u) + Lambda^- (T^{-1} v), g = T c
  double c[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double y = 0., z = 0.;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      y = fma(Ti[i][j], u[j], y);
      z = fma(Ti[i][j], v[j], z);
    }
    c[i] = fmax(ev[i], 0.) * y + fmin(ev[i], 0.) * z;
  }
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double r = 0.;
#pragma unroll
    for (int j = 0; j < M; ++j)
      r = fma(T[i][j], c[j], r);
    g[i] = r;
  }
}

template <int D, int NUMFLUX>
__device__ __forceinline__ void numflux_k(const FvSysParams& p, const double (&u)[D + 2], const double (&v)[D + 2], const int k,
                                          double (&g)[D + 2])
{
  constexpr int M = D + 2;
  if (NUMFLUX == GDTB_NUMFLUX_VIJAYASUNDARAM)
    vijayasundaram_k<D>(p.gamma, u, v, k, g);
  else {
    double fu[M], fv[M];
    euler_flux_k<D>(p.gamma, u, k, fu);
    euler_flux_k<D>(p.gamma, v, k, fv);
#pragma unroll
    for (int i = 0; i < M; ++i)
      g[i] = (fu[i] + fv[i]) * 0.5 + (u[i] - v[i]) * p.half_over_lambda;
  }
}

template <int M>
__device__ __forceinline__ void load_state(const double* __restrict__ q, double (&w)[M])
{
  if (M == 4) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(q)), b = __ldg(reinterpret_cast<const double2*>(q) + 1);
    w[0] = a.x, w[1] = a.y, w[2] = b.x, w[M - 1] = b.y;
  } else {
#pragma unroll
    for (int i = 0; i < M; ++i)
      w[i] = __ldg(q + i);
  }
}

template <int D, int NUMFLUX>
__global__ void __launch_bounds__(128) k_fvsys_apply(const __grid_constant__ FvSysParams p, const double* __restrict__ u,
                                                     double* __restrict__ out)
{
  constexpr int M = D + 2;
  const GridDev& g = p.g;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne)
    return;
  const int n0 = (int)g.n[0];
  int idx[2] = {int(e % n0), D > 1 ? int(e / n0) : 0};
  const long long stride[2] = {1, n0};
  double uc[M], acc[M];
  load_state<M>(u + e * M, uc);
#pragma unroll
  for (int i = 0; i < M; ++i)
    acc[i] = 0.;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const int nk = (int)g.n[k];
    const bool per = ((g.periodic >> k) & 1) && nk > 1;
    const double rk = __ldg(p.inv_ext[k] + idx[k]);
    // impermeable walls on a domain side without a neighbour (test/inviscid-compressible-flow/base.hh:187-241).  The
    // reference adds g(u, n) |I| / |E| with the outer normal n = -+e_k; along +e_k that is -g for the lower and +g for the
    // upper face.  Wall flux: g = (0, p n, 0), i.e. (0, p e_k, 0) along +e_k on either side.  Mirror: ghost state with the
    // normal velocity reflected, g = numerical_flux(u, v, n) = (by P(w, -n) = -P(w, n)) the flux along +e_k with the ghost
    // on the side of the wall.
    if (!per && (idx[k] == 0 || idx[k] == nk - 1) && ((p.wall_mask | p.mirror_mask) != 0u)) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if ((s == 0 ? idx[k] != 0 : idx[k] != nk - 1))
          continue;
        const unsigned bit = 1u << (2 * k + s);
        const double sign = s ? 1. : -1.;
        if (p.wall_mask & bit) {
          const EulerState<D> st(p.gamma, uc);
          acc[1 + k] += sign * st.p * rk;
        }
        if (p.mirror_mask & bit) {
          double vg[M], G[M];
#pragma unroll
          for (int i = 0; i < M; ++i)
            vg[i] = uc[i];
          {
            // conservative(rho, v - 2 (v . n) n, p): only the momentum along k changes sign, the energy is unchanged up to
            // rounding; evaluated like the reference (primitives -> reflected velocity -> conservative)
            const EulerState<D> st(p.gamma, uc);
            double vel[D], v2 = 0.;
#pragma unroll
            for (int i = 0; i < D; ++i) {
              vel[i] = st.v[i] - (i == k ? 2. * st.v[k] : 0.);
              v2 += vel[i] * vel[i];
            }
#pragma unroll
            for (int i = 0; i < D; ++i)
              vg[1 + i] = st.rho * vel[i];
            vg[M - 1] = st.p / (p.gamma - 1.) + 0.5 * st.rho * v2;
          }
          if (s)
            numflux_k<D, NUMFLUX>(p, uc, vg, k, G);
          else
            numflux_k<D, NUMFLUX>(p, vg, uc, k, G);
#pragma unroll
          for (int i = 0; i < M; ++i)
            acc[i] += sign * G[i] * rk;
        }
      }
    }
    // lower face: inside = the lower neighbour, outside = this cell
    if (idx[k] > 0 || per) {
      const long long en = e + (idx[k] > 0 ? -stride[k] : (long long)(nk - 1) * stride[k]);
      double un[M], G[M];
      load_state<M>(u + en * M, un);
      numflux_k<D, NUMFLUX>(p, un, uc, k, G);
#pragma unroll
      for (int i = 0; i < M; ++i)
        acc[i] -= G[i] * rk;
    }
    // upper face: inside = this cell
    if (idx[k] < nk - 1 || per) {
      const long long en = e + (idx[k] < nk - 1 ? stride[k] : -(long long)(nk - 1) * stride[k]);
      double un[M], G[M];
      load_state<M>(u + en * M, un);
      numflux_k<D, NUMFLUX>(p, uc, un, k, G);
#pragma unroll
      for (int i = 0; i < M; ++i)
        acc[i] += G[i] * rk;
    }
  }
#pragma unroll
  for (int i = 0; i < M; ++i)
    out[e * M + i] = p.euler ? uc[i] - acc[i] * p.dt : acc[i]; // u_n - L(u_n) dt (examples/mpi...cc:154)
}

template <int M>
__global__ void __launch_bounds__(256) k_fvsys_minmax(const long long ne, const double* __restrict__ u, double* __restrict__ partial)
{
  __shared__ double s[2 * M][256];
  double lo[M], hi[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    lo[i] = 1.7976931348623157e308;
    hi[i] = -1.7976931348623157e308;
  }
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x)
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const double v = __ldg(u + e * M + i);
      lo[i] = fmin(lo[i], v);
      hi[i] = fmax(hi[i], v);
    }
#pragma unroll
  for (int i = 0; i < M; ++i) {
    s[i][threadIdx.x] = lo[i];
    s[M + i][threadIdx.x] = hi[i];
  }
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w)
#pragma unroll
      for (int i = 0; i < M; ++i) {
        s[i][threadIdx.x] = fmin(s[i][threadIdx.x], s[i][threadIdx.x + w]);
        s[M + i][threadIdx.x] = fmax(s[M + i][threadIdx.x], s[M + i][threadIdx.x + w]);
      }
    __syncthreads();
  }
  if (threadIdx.x < 2 * M)
    partial[(long long)blockIdx.x * 2 * M + threadIdx.x] = s[threadIdx.x][0];
}

} // namespace

int launch_fvsys_apply(Launch& L, const FvSysParams& p, const double* u, double* out)
{
  if (p.g.d != 1 && p.g.d != 2)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "Euler equations: d = 1, 2 (tools/euler.hh: 3d is not implemented in the reference either)");
  if (p.m != p.g.d + 2)
    return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "Euler equations need a finite volume space with m = d + 2 components");
  const unsigned grid = (unsigned)((p.g.ne + 127) / 128);
  const bool vij = p.numflux == GDTB_NUMFLUX_VIJAYASUNDARAM;
  time_begin(L, KF_FV_APPLY);
  if (p.g.d == 1) {
    if (vij)
      k_fvsys_apply<1, GDTB_NUMFLUX_VIJAYASUNDARAM><<<grid, 128, 0, L.stream>>>(p, u, out);
    else
      k_fvsys_apply<1, GDTB_NUMFLUX_LAX_FRIEDRICHS><<<grid, 128, 0, L.stream>>>(p, u, out);
  } else {
    if (vij)
      k_fvsys_apply<2, GDTB_NUMFLUX_VIJAYASUNDARAM><<<grid, 128, 0, L.stream>>>(p, u, out);
    else
      k_fvsys_apply<2, GDTB_NUMFLUX_LAX_FRIEDRICHS><<<grid, 128, 0, L.stream>>>(p, u, out);
  }
  time_end(L, KF_FV_APPLY);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

int launch_fvsys_minmax(Launch& L, const FvSysParams& p, const double* u, double* partial, int blocks)
{
  if (p.m == 3)
    k_fvsys_minmax<3><<<blocks, 256, 0, L.stream>>>(p.g.ne, u, partial);
  else if (p.m == 4)
    k_fvsys_minmax<4><<<blocks, 256, 0, L.stream>>>(p.g.ne, u, partial);
  else
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "fv systems: m = 3, 4");
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

} // namespace gdtb

// dune-gdt_b200/csrc/fv.cu -- explicit first-order finite-volume advection operator apply.
//
// Replaces AdvectionFvOperator / LocalizableOperator::apply (dune/gdt/operators/advection-fv.hh:66-83,
// operators/localizable-operator.hh:352-387), LocalIntersectionOperatorApplicator::apply_local
// (local/assembler/operator-applicators.hh:219-227), LocalAdvectionFvCouplingOperator::apply
// (local/operators/advection-fv.hh:127-153) and the numerical fluxes (local/numerical-fluxes/upwind.hh:61-73,
// lax-friedrichs.hh:60-88), plus the explicit Euler update of examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:152-157.
//
// The reference walks the faces once and read-modify-writes both adjacent cells (2 RMW per face behind a lock).
// Here each cell gathers the fluxes of its 2d faces, so every range value is written exactly once (16 B/cell of
// HBM traffic: u in, L(u) out).  Each face flux is evaluated with the reference's roles (inside = the element with
// the smaller index, its outer normal) and the per-cell sum runs in the order in which the walker would have touched
// the cell, so results differ from the face-once walk only by FMA contraction (a few ulp).  Cell extents come from
// per-axis tables computed once per operator with YaspGrid's own formula.
#include <algorithm>

#include "common.cuh"
#include "kernels.hpp"

namespace gdtb {

namespace {

// flux function along axis k and its derivative (scalar conservation law, m = 1)
__device__ __forceinline__ double flux_k(const gdtb_flux& fl, int k, double w)
{
  return fl.kind == GDTB_FLUX_LINEAR ? fl.p[k] * w : 0.5 * w * w;
}

// numerical flux g(u, v, n) * 1 for the axis-aligned unit normal n = sign * e_k
// NumericalUpwindFlux<I,d,1>::apply (upwind.hh:67-72) / NumericalLaxFriedrichsFlux::apply (lax-friedrichs.hh:70-87);
// with n = +-e_k all terms of the reference's dot products except the k-th are exact zeros.
template <int D>
__device__ __forceinline__ double numerical_flux_axis(const gdtb_flux& fl, double lf_lambda_linear, int k, double sign,
                                                      double u, double v)
{
  if (fl.numflux == GDTB_NUMFLUX_UPWIND) {
    const double ubar = (u + v) / 2.;
    const double dfk = fl.kind == GDTB_FLUX_LINEAR ? fl.p[k] : ubar;
    return (sign * dfk > 0 ? flux_k(fl, k, u) : flux_k(fl, k, v)) * sign;
  }
  // lambda = 1 / max_j(|f_j'(u)|, |f_j'(v)|); ret = (f(u) + f(v)) . n / 2 + (u - v) / (2 lambda)
  const double inv_lambda = fl.kind == GDTB_FLUX_LINEAR ? lf_lambda_linear : fmax(fabs(u), fabs(v));
  return (flux_k(fl, k, u) + flux_k(fl, k, v)) * (sign * 0.5) + (u - v) * (0.5 * inv_lambda);
}

// One thread per cell, launched on a (x-chunks, y, z) grid so that no index needs a division.  Every cell gathers
// the fluxes of its 2 D faces: per face the reference's roles are kept (inside = element with the smaller index, its
// outer normal, advection-fv.hh:135-152) and the contributions are summed in the order the face-once walk would have
// added them to this cell.
template <int D>
__global__ void __launch_bounds__(256) k_fv_apply(const __grid_constant__ FvParams p, const double* __restrict__ u,
                                                  double* __restrict__ out)
{
  const GridDev& g = p.g;
  constexpr int last = D - 1;
  const int n0 = (int)g.n[0], n1 = D > 1 ? (int)g.n[1] : 1, n2 = D > 2 ? (int)g.n[2] : 1;
  const int layer_lo = (int)g.layer_lo;
  int idx[3];
  idx[0] = blockIdx.x * blockDim.x + threadIdx.x;
  idx[1] = D > 1 ? (int)blockIdx.y : 0;
  idx[2] = D > 2 ? (int)blockIdx.z : 0;
  if (D == 1)
    idx[0] += layer_lo;
  if (idx[0] >= (D == 1 ? (int)g.layer_hi : n0))
    return;
  idx[last] += (D == 1) ? 0 : layer_lo; // global coordinate along the partitioned direction
  const int n[3] = {n0, n1, n2};
  // local linear index: x fastest; along the last direction the local layer (+1 ghost layer below)
  long long stride[3] = {1, n0, (long long)n0 * n1};
  const long long plane = stride[last];
  const long long loc = (long long)idx[0] * (D == 1 ? 1 : 1) + (D > 1 ? (long long)idx[1] * stride[1] : 0)
                        + (D > 2 ? (long long)idx[2] * stride[2] : 0) - (long long)layer_lo * plane
                        + (p.ghosted ? plane : 0);
  const long long e = (long long)idx[0] + (long long)n0 * ((D > 1 ? idx[1] : 0) + (long long)n1 * (D > 2 ? idx[2] : 0));
  const double ue = u[loc];
  double ext[3];
#pragma unroll
  for (int k = 0; k < D; ++k)
    ext[k] = __ldg(p.ext[k] + idx[k]);
  double vol = ext[0];
  if (D > 1)
    vol *= ext[1];
  if (D > 2)
    vol *= ext[2];
  const double hinv = 1. / vol;
  double hI[3]; // |intersection| normal to axis k
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double a = 1.;
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (j != k)
        a *= ext[j];
    hI[k] = a;
  }

  bool interior = true;
#pragma unroll
  for (int k = 0; k < D; ++k)
    interior = interior && idx[k] > 0 && idx[k] < n[k] - 1;

  double acc = 0.;
  if (interior) {
    // lower neighbours are the inside elements (visited earlier, ascending index: z-, y-, x-), then this cell's
    // own upper faces in intersection order (x+, y+, z+)
#pragma unroll
    for (int k = D - 1; k >= 0; --k) {
      const double un = __ldg(u + loc - stride[k]);
      const double gf = numerical_flux_axis<D>(p.flux, p.lf_lambda_linear, k, 1., un, ue);
      acc += -(gf * hI[k]) * hinv; // advection-fv.hh:151
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double un = __ldg(u + loc + stride[k]);
      const double gf = numerical_flux_axis<D>(p.flux, p.lf_lambda_linear, k, 1., ue, un);
      acc += (gf * hI[k]) * hinv; // advection-fv.hh:150
    }
  } else {
    double contrib[2 * D];
    long long key[2 * D];
    int nc = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        int t = idx[k] + (s ? 1 : -1);
        long long nloc;
        if (t < 0 || t >= n[k]) {
          if (!(g.periodic & (1 << k)))
            continue; // domain boundary without neighbour: no coupling operator (advection-fv.hh:77-82)
          t = (t + n[k]) % n[k];
        }
        if (t == idx[k])
          continue; // single periodic cell: filtered by index(inside) < index(outside)
        if (k == last && p.ghosted)
          nloc = loc + (s ? plane : -plane); // ghost layers carry the (periodic) neighbours' data
        else
          nloc = loc + (long long)(t - idx[k]) * stride[k];
        const long long en = e + (long long)(t - idx[k]) * stride[k];
        const double un = __ldg(u + nloc);
        if (e < en) { // this cell is the inside element, the face is its (k, s) intersection
          const double gf = numerical_flux_axis<D>(p.flux, p.lf_lambda_linear, k, s ? 1. : -1., ue, un);
          contrib[nc] = (gf * hI[k]) * hinv;
          key[nc] = e * 8 + (2 * k + s);
        } else { // the neighbour is the inside element, the face is its (k, 1-s) intersection
          const double gf = numerical_flux_axis<D>(p.flux, p.lf_lambda_linear, k, s ? -1. : 1., un, ue);
          contrib[nc] = -(gf * hI[k]) * hinv;
          key[nc] = en * 8 + (2 * k + (1 - s));
        }
        ++nc;
      }
    }
    // sum in walker order: ascending (inside element, intersection index)
    for (int a = 0; a < nc; ++a) {
      int best = -1;
      long long bk = 0x7fffffffffffffffLL;
#pragma unroll
      for (int b = 0; b < 2 * D; ++b)
        if (b < nc && key[b] < bk) {
          bk = key[b];
          best = b;
        }
#pragma unroll
      for (int b = 0; b < 2 * D; ++b)
        if (b == best) {
          acc += contrib[b];
          key[b] = 0x7fffffffffffffffLL;
        }
    }
  }
  out[loc] = p.euler ? ue - acc * p.dt : acc; // u_n - L(u_n) * dt (examples/mpi_2019_02...cc:154)
}

template <int D>
__global__ void __launch_bounds__(256) k_fv_interpolate(const GridDev g, const FnDev f, int m,
                                                        const double* __restrict__ qx, const double* __restrict__ qw,
                                                        double* __restrict__ u)
{
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < g.ne;
       e += (long long)gridDim.x * blockDim.x) {
    long long idx[3];
    elem_coords(g, e, idx);
    double lower[3], ext[3];
    cell_geometry(g, idx, lower, ext);
    double vol = 1.;
#pragma unroll
    for (int k = 0; k < D; ++k)
      vol *= ext[k];
    const int my = D > 1 ? m : 1, mz = D > 2 ? m : 1;
    double integral = 0.;
    for (int qz = 0; qz < mz; ++qz)
      for (int qy = 0; qy < my; ++qy)
        for (int q0 = 0; q0 < m; ++q0) {
          const int q[3] = {q0, qy, qz};
          double x[3] = {0., 0., 0.}, w = 1.;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            x[k] = lower[k] + qx[q[k]] * ext[k];
            w *= qw[q[k]];
          }
          integral += fn_scalar(f, D, e, x) * vol * w;
        }
    u[e] = integral / vol; // spaces/basis/finite-volume.hh:249-250
  }
}

} // namespace

int launch_fv_apply(Launch& L, const FvParams& p, const double* u, double* out)
{
  const GridDev& g = p.g;
  const long long layers = g.layer_hi - g.layer_lo;
  const int block = 256;
  dim3 grid(1, 1, 1);
  if (g.d == 1)
    grid.x = (unsigned)((layers + block - 1) / block);
  else {
    grid.x = (unsigned)((g.n[0] + block - 1) / block);
    grid.y = (unsigned)(g.d == 2 ? layers : g.n[1]);
    grid.z = (unsigned)(g.d == 3 ? layers : 1);
    if (grid.y > 65535 || grid.z > 65535)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "fv: more than 65535 cells in the second / third direction");
  }
  time_begin(L, KF_FV_APPLY);
  switch (g.d) {
    case 1: k_fv_apply<1><<<grid, block, 0, L.stream>>>(p, u, out); break;
    case 2: k_fv_apply<2><<<grid, block, 0, L.stream>>>(p, u, out); break;
    case 3: k_fv_apply<3><<<grid, block, 0, L.stream>>>(p, u, out); break;
    default: return fail(GDTB_ERR_INVALID_ARGUMENT, "fv: dimension must be 1, 2 or 3");
  }
  time_end(L, KF_FV_APPLY);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

int launch_fv_interpolate(Launch& L, const GridDev& g, const FnDev& f, int m, const double* qx, const double* qw,
                          double* u)
{
  const int block = 256;
  const unsigned grid =
      (unsigned)std::max<long long>(1, std::min<long long>((g.ne + block - 1) / block, (long long)L.sm_count * 16));
  switch (g.d) {
    case 1: k_fv_interpolate<1><<<grid, block, 0, L.stream>>>(g, f, m, qx, qw, u); break;
    case 2: k_fv_interpolate<2><<<grid, block, 0, L.stream>>>(g, f, m, qx, qw, u); break;
    case 3: k_fv_interpolate<3><<<grid, block, 0, L.stream>>>(g, f, m, qx, qw, u); break;
    default: return fail(GDTB_ERR_INVALID_ARGUMENT, "fv: dimension must be 1, 2 or 3");
  }
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

} // namespace gdtb

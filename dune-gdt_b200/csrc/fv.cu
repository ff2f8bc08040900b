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
// HBM traffic: u in, L(u) out).  To stay bit-identical to the face-once walk, each face flux is evaluated with
// the reference's roles (inside = the element with the smaller index, its outer normal) and the per-cell sum
// runs in the order in which the walker would have touched the cell; FMA contraction is disabled explicitly.
#include <algorithm>

#include "common.cuh"
#include "kernels.hpp"

namespace gdtb {

namespace {

__device__ __forceinline__ double mul(double a, double b)
{
  return __dmul_rn(a, b);
}
__device__ __forceinline__ double add(double a, double b)
{
  return __dadd_rn(a, b);
}

template <int D>
__device__ __forceinline__ double dot_n(const double* a, const double* n)
{
  double s = 0.;
#pragma unroll
  for (int k = 0; k < D; ++k)
    s = add(s, mul(a[k], n[k]));
  return s;
}

template <int D>
__device__ __forceinline__ void flux_eval(const gdtb_flux& fl, double u, double* f)
{
#pragma unroll
  for (int k = 0; k < D; ++k)
    f[k] = fl.kind == GDTB_FLUX_LINEAR ? mul(fl.p[k], u) : mul(mul(0.5, u), u);
}

template <int D>
__device__ __forceinline__ void flux_jac(const gdtb_flux& fl, double u, double* df)
{
#pragma unroll
  for (int k = 0; k < D; ++k)
    df[k] = fl.kind == GDTB_FLUX_LINEAR ? fl.p[k] : u;
}

template <int D>
__device__ __forceinline__ double numerical_flux(const gdtb_flux& fl, double u, double v, const double* n)
{
  double a[3], b[3];
  if (fl.numflux == GDTB_NUMFLUX_UPWIND) { // upwind.hh:67-72
    flux_jac<D>(fl, add(u, v) / 2., a);
    if (dot_n<D>(n, a) > 0) {
      flux_eval<D>(fl, u, b);
      return dot_n<D>(b, n);
    }
    flux_eval<D>(fl, v, b);
    return dot_n<D>(b, n);
  }
  // lax-friedrichs.hh:70-87 with lambda_ = 0
  double lambda = 0.;
  flux_jac<D>(fl, u, a);
  flux_jac<D>(fl, v, b);
#pragma unroll
  for (int k = 0; k < D; ++k) {
    lambda = fmax(lambda, fabs(a[k]));
    lambda = fmax(lambda, fabs(b[k]));
  }
  lambda = 1. / lambda;
  flux_eval<D>(fl, u, a);
  flux_eval<D>(fl, v, b);
  double ret = 0.;
#pragma unroll
  for (int k = 0; k < D; ++k)
    ret = add(ret, mul(add(a[k], b[k]), mul(n[k], 0.5)));
  ret = add(ret, mul(add(u, -v), 0.5 / lambda));
  return ret;
}

__device__ __forceinline__ void cell_geometry_strict(const GridDev& g, const long long* idx, double* ext)
{
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < g.d) {
      const double lower = add(g.lo[k], mul(double(idx[k]), g.h[k]));
      const double upper = add(g.lo[k], mul(double(idx[k] + 1), g.h[k]));
      ext[k] = add(upper, -lower);
    } else
      ext[k] = 1.;
  }
}

template <int D>
__device__ __forceinline__ double volume(const double* ext)
{
  double v = 1.;
#pragma unroll
  for (int k = 0; k < D; ++k)
    v = mul(v, ext[k]);
  return v;
}

template <int D>
__global__ void __launch_bounds__(256) k_fv_apply(const __grid_constant__ FvParams p, const double* __restrict__ u,
                                                  double* __restrict__ out)
{
  const GridDev& g = p.g;
  const int last = D - 1;
  const long long layers = g.layer_hi - g.layer_lo;
  long long plane = 1; // cells per layer of the last direction
#pragma unroll
  for (int k = 0; k < D - 1; ++k)
    plane *= g.n[k];
  const long long owned = plane * layers;
  const long long shift = p.ghosted ? plane : 0; // local offset of the first owned cell

  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < owned;
       t += (long long)gridDim.x * blockDim.x) {
    // global coordinates of this cell
    long long idx[3] = {0, 0, 0};
    {
      long long r = t;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const long long nk = (k == last) ? layers : g.n[k];
        idx[k] = r % nk;
        r /= nk;
      }
      idx[last] += g.layer_lo;
    }
    const long long e = elem_index(g, idx);
    const long long loc = t + shift;
    const double ue = u[loc];
    double ext_e[3];
    cell_geometry_strict(g, idx, ext_e);
    const double hinv_e = 1. / volume<D>(ext_e);

    double contrib[2 * D];
    long long key[2 * D];
    int nc = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        long long nb[3];
        bool boundary;
        if (!face_neighbor(g, idx, k, s, nb, &boundary))
          continue;
        const long long en = elem_index(g, nb);
        if (en == e)
          continue; // degenerate periodic direction with a single cell: filtered by index(in) < index(out)
        // local position of the neighbour's value
        long long stride = 1;
#pragma unroll
        for (int j = 0; j < k; ++j)
          stride *= g.n[j];
        long long nloc;
        if (k == last && p.ghosted)
          nloc = loc + (s ? plane : -plane); // ghost layers carry the (periodic) neighbours' data
        else
          nloc = loc + (nb[k] - idx[k]) * stride;
        const double un = __ldg(u + nloc);
        double ext_n[3];
        cell_geometry_strict(g, nb, ext_n);
        double normal[3] = {0., 0., 0.};
        double c;
        if (e < en) {
          // this cell is the inside element, the face is its (k, s) intersection
          normal[k] = s ? 1. : -1.;
          double hI = 1.;
#pragma unroll
          for (int j = 0; j < D; ++j)
            if (j != k)
              hI = mul(hI, ext_e[j]);
          const double gf = numerical_flux<D>(p.flux, ue, un, normal);
          c = mul(mul(gf, hI), hinv_e); // advection-fv.hh:149-150
          key[nc] = e * 8 + (2 * k + s);
        } else {
          // the neighbour is the inside element, the face is its (k, 1-s) intersection
          normal[k] = s ? -1. : 1.;
          double hI = 1.;
#pragma unroll
          for (int j = 0; j < D; ++j)
            if (j != k)
              hI = mul(hI, ext_n[j]);
          const double gf = numerical_flux<D>(p.flux, un, ue, normal);
          c = mul(-mul(gf, hI), hinv_e); // advection-fv.hh:151
          key[nc] = en * 8 + (2 * k + (1 - s));
        }
        contrib[nc] = c;
        ++nc;
      }
    }
    // sum in walker order: ascending (inside element, intersection index)
    double acc = 0.;
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
          acc = add(acc, contrib[b]);
          key[b] = 0x7fffffffffffffffLL;
        }
    }
    if (p.euler)
      out[loc] = add(ue, -mul(acc, p.dt)); // u_n - L(u_n) * dt
    else
      out[loc] = acc;
  }
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
  long long owned = g.layer_hi - g.layer_lo;
  for (int k = 0; k < g.d - 1; ++k)
    owned *= g.n[k];
  const int block = 256;
  long long want = (owned + block - 1) / block;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)L.sm_count * 16));
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

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
#include <cstdlib>

#include "common.cuh"
#include "fv_tma.hpp"
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
    idx[0] += (int)p.apply_lo;
  if (idx[0] >= (D == 1 ? (int)p.apply_hi : n0))
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
          if (!(g.periodic & (1 << k))) {
            // domain boundary without neighbour: no coupling operator (advection-fv.hh:77-82), but the appended
            // boundary treatments (local/operators/advection-fv.hh:281-296, 418-443): g * (|I| / |E|)
            const int side = 2 * k + s;
            double gb = 0.;
            bool any = false;
            if (p.bnd_ext_mask >> side & 1) {
              const double v = p.bnd_ext_a[side] * ue + p.bnd_ext_b[side];
              gb += numerical_flux_axis<D>(p.flux, p.lf_lambda_linear, k, s ? 1. : -1., ue, v);
              any = true;
            }
            if (p.bnd_nf_mask >> side & 1) {
              gb += p.bnd_nf_a[side] * (flux_k(p.flux, k, ue) * (s ? 1. : -1.)) + p.bnd_nf_b[side];
              any = true;
            }
            if (any) {
              contrib[nc] = gb * (hI[k] / vol);
              key[nc] = e * 8 + side;
              ++nc;
            }
            continue;
          }
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

// Flux through the face between the cells L (lower index along axis k) and U = L + 1 (or U = 0, L = n_k - 1 for a
// periodic wrap face), measured along +e_k.  In the reference the inside element is the one with the smaller index
// (ApplyOn::InnerIntersectionsOnce / PeriodicBoundaryIntersectionsOnce) and the flux is evaluated with its outer
// normal: +e_k for an inner face (inside = L), -e_k for the wrap face (inside = U = cell 0).  Written out for both
// orientations the upwind flux (upwind.hh:67-72) along +e_k is f_k(u_L) if f_k'(ubar) > 0 and f_k(u_U) if
// f_k'(ubar) < 0 for either orientation.  Only the tie f_k'(ubar) == 0 depends on who is inside, and for the two flux
// families here the tie does not matter: a_k == 0 makes both candidates a_k u = 0, and ubar == 0 means
// u_L = -u_U, i.e. u_L^2 / 2 == u_U^2 / 2.  The Lax-Friedrichs flux (lax-friedrichs.hh:70-87) is orientation
// independent.  Hence one formula serves inner and wrap faces.
template <int NUMFLUX, int KIND>
__device__ __forceinline__ double flux_plus(const FvParams& p, int k, double uL, double uU)
{
  if (NUMFLUX == GDTB_NUMFLUX_UPWIND) {
    if (KIND == GDTB_FLUX_LINEAR) {
      const double a = p.flux.p[k];
      return a * (a > 0. ? uL : uU);
    }
    const double w = (uL + uU) > 0. ? uL : uU; // sign of ubar = (u + v) / 2
    return 0.5 * w * w;
  }
  if (KIND == GDTB_FLUX_LINEAR)
    return (p.flux.p[k] * uL + p.flux.p[k] * uU) * 0.5 + (uL - uU) * (0.5 * p.lf_lambda_linear);
  return (0.5 * uL * uL + 0.5 * uU * uU) * 0.5 + (uL - uU) * (0.5 * fmax(fabs(uL), fabs(uU)));
}

// Flux along +e_k through the domain boundary face on side s (0: lower, 1: upper) of a cell with value uc, produced by
// the boundary treatments of that side.  The reference evaluates g with the outer normal n = -+e_k and adds
// g |I| / |E| to the cell (local/operators/advection-fv.hh:293, 440); along +e_k that is g for s = 1 and -g for s = 0.
// Extrapolation: the ghost value v = a uc + b takes the neighbour's place in flux_plus (same orientation argument).
template <int NUMFLUX, int KIND>
__device__ __forceinline__ double bnd_flux_plus(const FvParams& p, int k, int s, double uc)
{
  const int side = 2 * k + s;
  double G = 0.;
  if (p.bnd_ext_mask >> side & 1) {
    const double v = p.bnd_ext_a[side] * uc + p.bnd_ext_b[side];
    G += s ? flux_plus<NUMFLUX, KIND>(p, k, uc, v) : flux_plus<NUMFLUX, KIND>(p, k, v, uc);
  }
  if (p.bnd_nf_mask >> side & 1) {
    const double fk = KIND == GDTB_FLUX_LINEAR ? p.flux.p[k] * uc : 0.5 * uc * uc;
    const double sgn = s ? 1. : -1.;
    G += (p.bnd_nf_a[side] * (fk * sgn) + p.bnd_nf_b[side]) * sgn;
  }
  return G;
}

#ifndef GDTB_FV_BATCH
#define GDTB_FV_BATCH 4
#endif
constexpr int FV_BATCH = GDTB_FV_BATCH;

template <int C>
__device__ __forceinline__ void ldg_cols(const double* q, double (&v)[C])
{
  if (C == 2) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(q));
    v[0] = t.x;
    v[C - 1] = t.y;
  } else
    v[0] = __ldg(q);
}

// the same through L2 only: layers another GPU may have written during this kernel (peer-memory ghost exchange)
template <int C>
__device__ __forceinline__ void ldcg_cols(const double* q, double (&v)[C])
{
  if (C == 2) {
    const double2 t = __ldcg(reinterpret_cast<const double2*>(q));
    v[0] = t.x;
    v[C - 1] = t.y;
  } else
    v[0] = __ldcg(q);
}

// bounded wait for a counter another GPU raises in this GPU's memory (system-scope acquire); gives up after ~1 s
__device__ __forceinline__ void fv_wait_counter(const int* counter, int expect, int* timeout_flag)
{
  for (long long spin = 0; spin < (1LL << 23); ++spin) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v - expect >= 0)
      return;
    __nanosleep(100);
  }
  atomicExch(timeout_flag, 1);
}

template <int C>
__device__ __forceinline__ void st_cols(double* q, const double (&v)[C])
{
  if (C == 2)
    *reinterpret_cast<double2*>(q) = make_double2(v[0], v[C - 1]);
  else
    *q = v[0];
}

// D >= 2: the thread block covers a tile of the first D-1 axes (x: 2D, x-y: 3D) and marches along the last axis
// over `rows` layers; a thread owns C adjacent cells in x (C = 2: 16-byte loads and stores).  The layers u[j], u[j+1]
// of a thread's cells live in registers, so every value of u is read from L2/HBM once per tile (plus the two halo
// layers); the neighbours along the tile axes are read through L1.  The flux through the upper face of layer j is
// reused as the lower-face flux of layer j+1, the flux between the thread's two cells is shared.  Every cell sums
// (G_up - G_low) / ext_k over its axes, G = flux along +e_k through the face (flux_plus); (g |I|) (1/|E|) of
// advection-fv.hh:147-152 equals g / ext_k on axis-aligned cells up to rounding.  Faces that do not exist (domain
// boundary without periodicity) get the coefficient 0.  Loads are issued FV_BATCH layers at a time before the first
// flux of the batch is evaluated (memory-level parallelism).
template <int D, int NUMFLUX, int KIND, int C, bool BND, bool P2P = false, int NS = -1>
__global__ void __launch_bounds__(256) k_fv_march(const __grid_constant__ FvParams p, const double* __restrict__ u,
                                                  double* __restrict__ out, int rows)
{
  static_assert(D == 2 || D == 3, "marching kernel is for 2D / 3D grids");
  static_assert(NS < 0 || (!BND && !P2P), "the fused stage combination covers the plain periodic / inner kernel");
  // NS >= 0: Runge-Kutta stage fused into the apply: every value of the source is u + sum_{j < NS} stage_v[j] stage_c[j]
  // (terms added in order like the separate axpy pass), formed where it is loaded
  auto ldc = [&](const double* q, double(&v)[C]) {
    ldg_cols<C>(q, v);
    if (NS > 0) {
      const long long o = q - u;
#pragma unroll
      for (int j = 0; j < (NS > 0 ? NS : 0); ++j) {
        double t[C];
        ldg_cols<C>(p.stage_v[j] + o, t);
#pragma unroll
        for (int c = 0; c < C; ++c)
          v[c] += t[c] * p.stage_c[j];
      }
    }
  };
  auto ld1 = [&](const double* q) {
    double a = __ldg(q);
    if (NS > 0) {
      const long long o = q - u;
#pragma unroll
      for (int j = 0; j < (NS > 0 ? NS : 0); ++j)
        a += __ldg(p.stage_v[j] + o) * p.stage_c[j];
    }
    return a;
  };
  // what a cell's result is: L(u_i), the fused Euler update, or (out_mode 1) the Runge-Kutta update of the step
  auto result = [&](const double* po_, const double uc_, const double acc) {
    if (NS >= 0 && p.out_mode == 1) {
      const long long o = po_ - out;
      double t = __ldg(u + o);
      for (int j = 0; j < p.n_out; ++j)
        t += __ldg(p.out_v[j] + o) * p.out_c[j];
      t += acc * p.out_cL;
      return t;
    }
    return p.euler ? uc_ - acc * p.dt : acc; // u_n - L(u_n) dt (examples/mpi...cc:154)
  };
  const GridDev& g = p.g;
  constexpr int last = D - 1;
  constexpr int R = FV_BATCH;
  const int n0 = (int)g.n[0], n1 = (int)g.n[1], nl = (int)g.n[last];
  const int layer_lo = (int)g.layer_lo;
  const int ix = (blockIdx.x * blockDim.x + threadIdx.x) * C; // first of the thread's cells
  const int iy = D == 3 ? blockIdx.y * blockDim.y + threadIdx.y : 0;
  if (ix >= n0 || (D == 3 && iy >= n1))
    return;
  const long long plane = D == 3 ? (long long)n0 * n1 : n0; // cells per layer
  const long long col = D == 3 ? (long long)iy * n0 + ix : ix;
  const int j0 = (int)p.apply_lo + (int)(D == 3 ? blockIdx.z : blockIdx.y) * rows;
  const int j1 = min(j0 + rows, (int)p.apply_hi);
  if (j0 >= j1)
    return;
  // local layer index of global layer j: j - layer_lo (+1 ghost layer below on a slab)
  const int shift = p.ghosted ? 1 - layer_lo : -layer_lo;
  // peer-memory exchange: the blocks that touch the slab's first / last layer read a ghost layer the neighbour GPU
  // filled during its previous step and hand their own boundary layer over at the end
  const bool edge_lo = P2P && j0 == layer_lo, edge_hi = P2P && j1 == (int)g.layer_hi;
  if (P2P) {
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      if (edge_lo && p.wait_lo)
        fv_wait_counter(p.my_flags + 0, p.expect, p.timeout_flag);
      if (edge_hi && p.wait_hi)
        fv_wait_counter(p.my_flags + 1, p.expect, p.timeout_flag);
    }
    if (edge_lo || edge_hi)
      __syncthreads();
  }

  // tile axes: offsets to the neighbour cells (periodic wrap, or 0 = the cell itself where there is no face) and
  // the face coefficients 1 / ext (0 where there is no face)
  const bool per0 = (g.periodic & 1) && n0 > 1, per1 = D == 3 && (g.periodic & 2) && n1 > 1;
  const bool perl = (g.periodic & (1 << last)) && nl > 1;
  const bool x_lo_edge = ix == 0, x_hi_edge = ix + C - 1 == n0 - 1;
  const bool x_lo = !x_lo_edge || per0, x_hi = !x_hi_edge || per0;
  const int d_xm = x_lo ? (x_lo_edge ? n0 - 1 : -1) : 0;              // relative to the first cell
  const int d_xp = C - 1 + (x_hi ? (x_hi_edge ? 1 - n0 : 1) : 0);    // relative to the first cell
  double rx[C];
  ldg_cols<C>(p.inv_ext[0] + ix, rx);
  const unsigned bnd_mask = BND ? (p.bnd_ext_mask | p.bnd_nf_mask) : 0u; // sides with a boundary treatment
  const bool bx_lo = BND && !x_lo && (bnd_mask & 1u), bx_hi = BND && !x_hi && (bnd_mask & 2u);
  const double cxl = (x_lo || bx_lo) ? rx[0] : 0., cxh = (x_hi || bx_hi) ? rx[C - 1] : 0.;
  long long d_ym = 0, d_yp = 0;
  double cyl = 0., cyh = 0.;
  bool by_lo = false, by_hi = false;
  if (D == 3) {
    const bool y_lo_edge = iy == 0, y_hi_edge = iy == n1 - 1;
    const bool y_lo = !y_lo_edge || per1, y_hi = !y_hi_edge || per1;
    by_lo = BND && !y_lo && (bnd_mask & 4u);
    by_hi = BND && !y_hi && (bnd_mask & 8u);
    d_ym = y_lo ? (y_lo_edge ? (long long)(n1 - 1) * n0 : -(long long)n0) : 0;
    d_yp = y_hi ? (y_hi_edge ? -(long long)(n1 - 1) * n0 : (long long)n0) : 0;
    const double ry = __ldg(p.inv_ext[1] + iy);
    cyl = (y_lo || by_lo) ? ry : 0.;
    cyh = (y_hi || by_hi) ? ry : 0.;
  }

  const double* pc = u + (long long)(j0 + shift) * plane + col; // own cells, layer j
  double* po = out + (long long)(j0 + shift) * plane + col;
  const double* prl = p.inv_ext[last] + j0;
  double uc[C], G_low[C];
  ldc(pc, uc);
  // flux through the lower face of the first layer (on a slab the ghost layer below carries the neighbour); 0 if
  // there is no face
  {
    const bool has = j0 > 0 || perl;
    const long long off = (j0 > 0 || p.ghosted) ? -plane : (has ? (long long)(nl - 1) * plane : 0);
    double ub[C];
    if (P2P)
      ldcg_cols<C>(pc + off, ub);
    else
      ldc(pc + off, ub);
#pragma unroll
    for (int c = 0; c < C; ++c)
      G_low[c] = has ? flux_plus<NUMFLUX, KIND>(p, last, ub[c], uc[c])
                     : (BND ? bnd_flux_plus<NUMFLUX, KIND>(p, last, 0, uc[c]) : 0.);
  }
  // layers whose upper neighbour is simply the next layer in memory: all but the top layer of an unpartitioned grid
  // (on a slab the ghost layer above carries the periodic neighbour; without periodicity the top face does not exist)
  const int j_plain = (p.ghosted && perl) ? j1 : min(j1, nl - 1);
  int jb = j0;
  for (; jb + R <= j_plain; jb += R) {
    double un[R][C], xl[R], xr[R], yl[R][C], yr[R][C], rl[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double* q = pc + (long long)r * plane;
      if (P2P)
        ldcg_cols<C>(q + plane, un[r]);
      else
        ldc(q + plane, un[r]);
      xl[r] = ld1(q + d_xm);
      xr[r] = ld1(q + d_xp);
      if (D == 3) {
        ldc(q + d_ym, yl[r]);
        ldc(q + d_yp, yr[r]);
      }
      rl[r] = __ldg(prl + r);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double res[C], gx[C + 1];
      gx[0] = bx_lo ? bnd_flux_plus<NUMFLUX, KIND>(p, 0, 0, uc[0]) : flux_plus<NUMFLUX, KIND>(p, 0, xl[r], uc[0]);
      if (C == 2)
        gx[1] = flux_plus<NUMFLUX, KIND>(p, 0, uc[0], uc[C - 1]);
      gx[C] = bx_hi ? bnd_flux_plus<NUMFLUX, KIND>(p, 0, 1, uc[C - 1]) : flux_plus<NUMFLUX, KIND>(p, 0, uc[C - 1], xr[r]);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const double G_up = flux_plus<NUMFLUX, KIND>(p, last, uc[c], un[r][c]);
        // faces between the thread's own cells always exist; the outer ones carry the (possibly zero) coefficient
        double acc = gx[c + 1] * (c == C - 1 ? cxh : rx[c]) - gx[c] * (c == 0 ? cxl : rx[c]);
        if (D == 3)
          acc += (by_hi ? bnd_flux_plus<NUMFLUX, KIND>(p, 1, 1, uc[c]) : flux_plus<NUMFLUX, KIND>(p, 1, uc[c], yr[r][c])) * cyh
                 - (by_lo ? bnd_flux_plus<NUMFLUX, KIND>(p, 1, 0, uc[c]) : flux_plus<NUMFLUX, KIND>(p, 1, yl[r][c], uc[c])) * cyl;
        acc += (G_up - G_low[c]) * rl[r];
        res[c] = result(po + (long long)r * plane + c, uc[c], acc);
        G_low[c] = G_up;
        uc[c] = un[r][c];
      }
      st_cols<C>(po + (long long)r * plane, res);
      if (P2P) {
        if (jb + r == layer_lo && p.peer_lo_ghost)
          st_cols<C>(p.peer_lo_ghost + col, res);
        if (jb + r == (int)g.layer_hi - 1 && p.peer_hi_ghost)
          st_cols<C>(p.peer_hi_ghost + col, res);
      }
    }
    pc += (long long)R * plane;
    po += (long long)R * plane;
    prl += R;
  }
  // remaining layers one at a time (batch tail and the top layer of the grid: periodic wrap or no upper face)
  for (int j = jb; j < j1; ++j) {
    const bool top = j == nl - 1;
    const bool has_up = !top || perl;
    double un[C], yl[C], yr[C], res[C], gx[C + 1];
    if (P2P)
      ldcg_cols<C>(pc + plane, un);
    else
      ldc(pc + ((top && !p.ghosted) ? (has_up ? (1 - (long long)nl) * plane : 0) : plane), un);
    gx[0] = bx_lo ? bnd_flux_plus<NUMFLUX, KIND>(p, 0, 0, uc[0]) : flux_plus<NUMFLUX, KIND>(p, 0, ld1(pc + d_xm), uc[0]);
    if (C == 2)
      gx[1] = flux_plus<NUMFLUX, KIND>(p, 0, uc[0], uc[C - 1]);
    gx[C] = bx_hi ? bnd_flux_plus<NUMFLUX, KIND>(p, 0, 1, uc[C - 1])
                  : flux_plus<NUMFLUX, KIND>(p, 0, uc[C - 1], ld1(pc + d_xp));
    if (D == 3) {
      ldc(pc + d_ym, yl);
      ldc(pc + d_yp, yr);
    }
    const double rl = __ldg(prl);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const double G_up = has_up ? flux_plus<NUMFLUX, KIND>(p, last, uc[c], un[c])
                                 : (BND ? bnd_flux_plus<NUMFLUX, KIND>(p, last, 1, uc[c]) : 0.);
      double acc = gx[c + 1] * (c == C - 1 ? cxh : rx[c]) - gx[c] * (c == 0 ? cxl : rx[c]);
      if (D == 3)
        acc += (by_hi ? bnd_flux_plus<NUMFLUX, KIND>(p, 1, 1, uc[c]) : flux_plus<NUMFLUX, KIND>(p, 1, uc[c], yr[c])) * cyh
               - (by_lo ? bnd_flux_plus<NUMFLUX, KIND>(p, 1, 0, uc[c]) : flux_plus<NUMFLUX, KIND>(p, 1, yl[c], uc[c])) * cyl;
      acc += (G_up - G_low[c]) * rl;
      res[c] = result(po + c, uc[c], acc);
      G_low[c] = G_up;
      uc[c] = un[c];
    }
    st_cols<C>(po, res);
    if (P2P) {
      if (j == layer_lo && p.peer_lo_ghost)
        st_cols<C>(p.peer_lo_ghost + col, res);
      if (j == (int)g.layer_hi - 1 && p.peer_hi_ghost)
        st_cols<C>(p.peer_hi_ghost + col, res);
    }
    pc += plane;
    po += plane;
    prl += 1;
  }
  if (P2P && (edge_lo || edge_hi)) {
    // make this block's remote stores visible system-wide; the last edge block of the launch raises the neighbour's
    // counter by one (so the counter counts steps, whatever the tiling)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      const int tiles = D == 3 ? (int)(gridDim.x * gridDim.y) : (int)gridDim.x;
      if (edge_lo && p.peer_lo_flag && atomicAdd(p.edge_count + 0, 1) == tiles - 1) {
        atomicExch(p.edge_count + 0, 0);
        __threadfence_system();
        atomicAdd_system(p.peer_lo_flag, 1);
      }
      if (edge_hi && p.peer_hi_flag && atomicAdd(p.edge_count + 1, 1) == tiles - 1) {
        atomicExch(p.edge_count + 1, 0);
        __threadfence_system();
        atomicAdd_system(p.peer_hi_flag, 1);
      }
    }
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
          const double xh[3] = {qx[q0], D > 1 ? qx[qy] : 0., D > 2 ? qx[qz] : 0.};
          const EvalPt pt = {q0 + m * (qy + my * qz), idx, xh};
          integral += fn_scalar(f, g, e, x, pt) * vol * w;
        }
    u[e] = integral / vol; // spaces/basis/finite-volume.hh:249-250
  }
}

// ExplicitRungeKuttaTimeStepper::step (tools/timestepper/explicit-rungekutta.hh:248-263): the stage vector
// u_i = u_n + sum_j k_j (dt r A_ij) and the update u_n += sum_i k_i (r dt b_i), terms added in the reference's order
// (one axpy per term there, one fused pass here: each vector is read once, the result written once).
template <int NV, int W>
__global__ void __launch_bounds__(256) k_rk_axpy(const __grid_constant__ RkAxpyParams p, const double* __restrict__ base,
                                                 double* __restrict__ out)
{
  const long long nw = p.n / W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (long long)gridDim.x * blockDim.x) {
    if (W == 2) {
      double2 a = __ldg(reinterpret_cast<const double2*>(base) + i);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const double2 k = __ldg(reinterpret_cast<const double2*>(p.v[j]) + i);
        a.x += k.x * p.c[j];
        a.y += k.y * p.c[j];
      }
      reinterpret_cast<double2*>(out)[i] = a;
    } else {
      double a = __ldg(base + i);
#pragma unroll
      for (int j = 0; j < NV; ++j)
        a += __ldg(p.v[j] + i) * p.c[j];
      out[i] = a;
    }
  }
  if (W == 2 && (p.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    double a = base[p.n - 1];
#pragma unroll
    for (int j = 0; j < NV; ++j)
      a += p.v[j][p.n - 1] * p.c[j];
    out[p.n - 1] = a;
  }
}

// Stage-vector hand-over of the slab Runge-Kutta loop: first owned layer -> lower neighbour's upper ghost layer, last
// owned layer -> upper neighbour's lower ghost layer; every block fences system-wide, the last one raises the counters.
__global__ void __launch_bounds__(256) k_p2p_send_layers(const __grid_constant__ P2pSendParams p)
{
  const double* first = p.src + p.plane;
  const double* last = p.src + p.layers * p.plane;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < p.plane; c += (long long)gridDim.x * blockDim.x) {
    if (p.peer_lo_ghost)
      p.peer_lo_ghost[c] = first[c];
    if (p.peer_hi_ghost)
      p.peer_hi_ghost[c] = last[c];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(p.edge_count, 1) == (int)gridDim.x - 1) {
    atomicExch(p.edge_count, 0);
    __threadfence_system();
    if (p.peer_lo_flag)
      atomicAdd_system(p.peer_lo_flag, 1);
    if (p.peer_hi_flag)
      atomicAdd_system(p.peer_hi_flag, 1);
  }
}

// estimate_dt_for_hyperbolic_system (tools/hyperbolic.hh:47-60, 75-82): data range of the (elementwise constant) state
// and max over the elements of perimeter / volume; block partials {min, max, pov}, finished on the host.
template <int D>
__global__ void __launch_bounds__(256) k_fv_dt_reduce(const __grid_constant__ FvParams p, const double* __restrict__ u,
                                                      double* __restrict__ partial)
{
  const GridDev& g = p.g;
  double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308, pov = 0.;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < g.ne;
       e += (long long)gridDim.x * blockDim.x) {
    const double v = __ldg(u + e);
    mn = fmin(mn, v);
    mx = fmax(mx, v);
    long long idx[3];
    elem_coords(g, e, idx);
    double ext[3] = {1., 1., 1.};
#pragma unroll
    for (int k = 0; k < D; ++k)
      ext[k] = __ldg(p.ext[k] + idx[k]);
    double vol = ext[0], perimeter = 0.;
    if (D > 1)
      vol *= ext[1];
    if (D > 2)
      vol *= ext[2];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      double a = 1.;
#pragma unroll
      for (int j = 0; j < D; ++j)
        if (j != k)
          a *= ext[j];
      perimeter += a; // the two faces normal to e_k, in intersection order
      perimeter += a;
    }
    pov = fmax(pov, perimeter / vol);
  }
  __shared__ double s[3][256];
  s[0][threadIdx.x] = mn;
  s[1][threadIdx.x] = mx;
  s[2][threadIdx.x] = pov;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
      s[0][threadIdx.x] = fmin(s[0][threadIdx.x], s[0][threadIdx.x + w]);
      s[1][threadIdx.x] = fmax(s[1][threadIdx.x], s[1][threadIdx.x + w]);
      s[2][threadIdx.x] = fmax(s[2][threadIdx.x], s[2][threadIdx.x + w]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x + 0] = s[0][0];
    partial[3 * blockIdx.x + 1] = s[1][0];
    partial[3 * blockIdx.x + 2] = s[2][0];
  }
}

} // namespace

// Runge-Kutta stage combination fused into the apply: NS stage vectors are combined with the source at every load
template <int D, int C, int NS>
static void launch_fv_march_fused(const FvParams& p, const double* u, double* out, int rows, dim3 grid, dim3 block,
                                  cudaStream_t stream)
{
  const int variant = (p.flux.numflux == GDTB_NUMFLUX_LAX_FRIEDRICHS ? 2 : 0) + (p.flux.kind == GDTB_FLUX_BURGERS ? 1 : 0);
  switch (variant) {
    case 0: k_fv_march<D, GDTB_NUMFLUX_UPWIND, GDTB_FLUX_LINEAR, C, false, false, NS><<<grid, block, 0, stream>>>(p, u, out, rows); break;
    case 1: k_fv_march<D, GDTB_NUMFLUX_UPWIND, GDTB_FLUX_BURGERS, C, false, false, NS><<<grid, block, 0, stream>>>(p, u, out, rows); break;
    case 2: k_fv_march<D, GDTB_NUMFLUX_LAX_FRIEDRICHS, GDTB_FLUX_LINEAR, C, false, false, NS><<<grid, block, 0, stream>>>(p, u, out, rows); break;
    default: k_fv_march<D, GDTB_NUMFLUX_LAX_FRIEDRICHS, GDTB_FLUX_BURGERS, C, false, false, NS><<<grid, block, 0, stream>>>(p, u, out, rows); break;
  }
}

template <int D, int C, bool P2P>
static void launch_fv_march_p(const FvParams& p, const double* u, double* out, int rows, dim3 grid, dim3 block,
                              cudaStream_t stream)
{
  const int variant = (p.flux.numflux == GDTB_NUMFLUX_LAX_FRIEDRICHS ? 2 : 0) + (p.flux.kind == GDTB_FLUX_BURGERS ? 1 : 0)
                      + ((p.bnd_ext_mask | p.bnd_nf_mask) ? 4 : 0);
  if (!P2P && p.n_stage >= 0 && variant < 4) {
    switch (p.n_stage) {
      case 0: launch_fv_march_fused<D, C, 0>(p, u, out, rows, grid, block, stream); break;
      case 1: launch_fv_march_fused<D, C, 1>(p, u, out, rows, grid, block, stream); break;
      case 2: launch_fv_march_fused<D, C, 2>(p, u, out, rows, grid, block, stream); break;
      default: launch_fv_march_fused<D, C, 3>(p, u, out, rows, grid, block, stream); break;
    }
    return;
  }
  switch (variant) {
    case 0: k_fv_march<D, GDTB_NUMFLUX_UPWIND, GDTB_FLUX_LINEAR, C, false, P2P><<<grid, block, 0, stream>>>(p, u, out, rows); break;
    case 1: k_fv_march<D, GDTB_NUMFLUX_UPWIND, GDTB_FLUX_BURGERS, C, false, P2P><<<grid, block, 0, stream>>>(p, u, out, rows); break;
    case 2: k_fv_march<D, GDTB_NUMFLUX_LAX_FRIEDRICHS, GDTB_FLUX_LINEAR, C, false, P2P><<<grid, block, 0, stream>>>(p, u, out, rows); break;
    case 3: k_fv_march<D, GDTB_NUMFLUX_LAX_FRIEDRICHS, GDTB_FLUX_BURGERS, C, false, P2P><<<grid, block, 0, stream>>>(p, u, out, rows); break;
    case 4: k_fv_march<D, GDTB_NUMFLUX_UPWIND, GDTB_FLUX_LINEAR, C, true, P2P><<<grid, block, 0, stream>>>(p, u, out, rows); break;
    case 5: k_fv_march<D, GDTB_NUMFLUX_UPWIND, GDTB_FLUX_BURGERS, C, true, P2P><<<grid, block, 0, stream>>>(p, u, out, rows); break;
    case 6: k_fv_march<D, GDTB_NUMFLUX_LAX_FRIEDRICHS, GDTB_FLUX_LINEAR, C, true, P2P><<<grid, block, 0, stream>>>(p, u, out, rows); break;
    default: k_fv_march<D, GDTB_NUMFLUX_LAX_FRIEDRICHS, GDTB_FLUX_BURGERS, C, true, P2P><<<grid, block, 0, stream>>>(p, u, out, rows); break;
  }
}

template <int D, int C>
static void launch_fv_march(const FvParams& p, const double* u, double* out, int rows, dim3 grid, dim3 block,
                            cudaStream_t stream)
{
  if (p.p2p)
    launch_fv_march_p<D, C, true>(p, u, out, rows, grid, block, stream);
  else
    launch_fv_march_p<D, C, false>(p, u, out, rows, grid, block, stream);
}

int launch_fv_apply(Launch& L, const FvParams& p, const double* u, double* out)
{
  const GridDev& g = p.g;
  const long long layers = p.apply_hi - p.apply_lo;
  if (layers <= 0)
    return GDTB_OK;
  uintptr_t fused_bits = 0;
  if (p.n_stage >= 0) {
    if (g.d == 1 || p.p2p || (p.bnd_ext_mask | p.bnd_nf_mask) || p.n_stage > 3 || p.n_out > 3)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "fv: the fused Runge-Kutta stage needs the plain 2D / 3D kernel");
    for (int j = 0; j < p.n_stage; ++j)
      fused_bits |= reinterpret_cast<uintptr_t>(p.stage_v[j]);
  }
  time_begin(L, KF_FV_APPLY);
  // GDTB_FV_TMA=1: source staged in shared memory by TMA bulk loads (fv_tma.cu: warp-specialised loader / marching warps,
  // mbarrier ring).  Measured on B200 at 4096^2: 54.7 - 58.9 us against 50.9 us for the register-marching kernel below (the
  // L2 -> SM path caps both the same way, B300_MICROARCH "LTS throughput cap ... LDG == TMA"), so it is opt-in; read per
  // launch so that one process can compare both.
  const char* tma_env = std::getenv("GDTB_FV_TMA");
  if (tma_env && tma_env[0] == '1' && fv_tma_eligible(p, u, out)) {
    GDTB_TRY(launch_fv_tma(L, p, u, out));
  } else if (g.d == 1) {
    const int block = 256;
    k_fv_apply<1><<<(unsigned)((layers + block - 1) / block), block, 0, L.stream>>>(p, u, out);
  } else {
    // two cells per thread (16-byte accesses) when every row starts 16-byte aligned
    const bool two = g.n[0] % 2 == 0 && ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(out)) & 15) == 0
                     && (reinterpret_cast<uintptr_t>(p.inv_ext[0]) & 15) == 0
                     && ((reinterpret_cast<uintptr_t>(p.peer_lo_ghost) | reinterpret_cast<uintptr_t>(p.peer_hi_ghost)) & 15) == 0
                     && (fused_bits & 15) == 0;
    const long long nx = two ? g.n[0] / 2 : g.n[0]; // threads along x
    dim3 block(1, 1, 1), grid(1, 1, 1);
    long long tiles;
    if (g.d == 2) {
      static const int bx_env = std::getenv("GDTB_FV_BLOCK") ? std::atoi(std::getenv("GDTB_FV_BLOCK")) : 0; // A/B knob
      const long long bx = (bx_env == 64 || bx_env == 128 || bx_env == 256) ? bx_env : 128;
      block.x = (unsigned)std::min<long long>(bx, ((nx + 31) / 32) * 32);
      grid.x = (unsigned)((nx + block.x - 1) / block.x);
      tiles = grid.x;
    } else {
      block.x = (unsigned)std::min<long long>(32, ((nx + 31) / 32) * 32);
      block.y = 128 / block.x;
      grid.x = (unsigned)((nx + block.x - 1) / block.x);
      grid.y = (unsigned)((g.n[1] + block.y - 1) / block.y);
      tiles = (long long)grid.x * grid.y;
    }
    // layers per block along the marching axis: long enough to amortise the two halo layers, short enough for a few
    // waves of blocks over the SMs
    int rows = p.rows_per_block;
    if (rows <= 0) {
      const long long want = (long long)L.sm_count * 8 * 4;
      rows = (int)std::max<long long>(8, std::min<long long>(64, layers * tiles / want));
    }
    rows = (rows + FV_BATCH - 1) / FV_BATCH * FV_BATCH; // whole register batches
    const long long chunks = (layers + rows - 1) / rows;
    if (chunks > 65535)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "fv: too many layers along the last direction");
    (g.d == 2 ? grid.y : grid.z) = (unsigned)chunks;
    if (g.d == 2) {
      if (two)
        launch_fv_march<2, 2>(p, u, out, rows, grid, block, L.stream);
      else
        launch_fv_march<2, 1>(p, u, out, rows, grid, block, L.stream);
    } else {
      if (two)
        launch_fv_march<3, 2>(p, u, out, rows, grid, block, L.stream);
      else
        launch_fv_march<3, 1>(p, u, out, rows, grid, block, L.stream);
    }
  }
  time_end(L, KF_FV_APPLY);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

template <int W>
static void launch_rk_axpy_w(const RkAxpyParams& p, const double* base, double* out, unsigned grid, cudaStream_t st)
{
  switch (p.nv) {
    case 0: k_rk_axpy<0, W><<<grid, 256, 0, st>>>(p, base, out); break;
    case 1: k_rk_axpy<1, W><<<grid, 256, 0, st>>>(p, base, out); break;
    case 2: k_rk_axpy<2, W><<<grid, 256, 0, st>>>(p, base, out); break;
    case 3: k_rk_axpy<3, W><<<grid, 256, 0, st>>>(p, base, out); break;
    default: k_rk_axpy<4, W><<<grid, 256, 0, st>>>(p, base, out); break;
  }
}

int launch_rk_axpy(Launch& L, const RkAxpyParams& p, const double* base, double* out)
{
  if (p.n <= 0)
    return GDTB_OK;
  if (p.nv < 0 || p.nv > RK_MAX_TERMS)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "rk_axpy: too many terms");
  uintptr_t bits = reinterpret_cast<uintptr_t>(base) | reinterpret_cast<uintptr_t>(out);
  for (int j = 0; j < p.nv; ++j)
    bits |= reinterpret_cast<uintptr_t>(p.v[j]);
  const bool two = (bits & 15) == 0 && p.n >= 2;
  const long long work = two ? p.n / 2 : p.n;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((work + 255) / 256, (long long)L.sm_count * 8));
  if (two)
    launch_rk_axpy_w<2>(p, base, out, grid, L.stream);
  else
    launch_rk_axpy_w<1>(p, base, out, grid, L.stream);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

int launch_p2p_send_layers(Launch& L, const P2pSendParams& p)
{
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((p.plane + 255) / 256, (long long)L.sm_count * 2));
  k_p2p_send_layers<<<grid, 256, 0, L.stream>>>(p);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

int launch_fv_dt_reduce(Launch& L, const FvParams& p, const double* u, double* partial, int blocks)
{
  switch (p.g.d) {
    case 1: k_fv_dt_reduce<1><<<blocks, 256, 0, L.stream>>>(p, u, partial); break;
    case 2: k_fv_dt_reduce<2><<<blocks, 256, 0, L.stream>>>(p, u, partial); break;
    case 3: k_fv_dt_reduce<3><<<blocks, 256, 0, L.stream>>>(p, u, partial); break;
    default: return fail(GDTB_ERR_INVALID_ARGUMENT, "fv: dimension must be 1, 2 or 3");
  }
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

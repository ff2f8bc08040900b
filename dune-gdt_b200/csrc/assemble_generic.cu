// dune-gdt_b200/csrc/assemble_generic.cu -- quadrature-faithful assembly kernels with deterministic,
// atomic-free coloured scatter.
//
// These kernels restate the reference's local loops one-to-one and are the path every configuration can
// take (any supported space / order / integrand sum / coefficient kind):
//   k_element_matrix  : LocalElementBilinearFormAssembler::apply_local (local/assembler/bilinear-form-assemblers.hh:110-128)
//                       + LocalElementIntegralBilinearForm::apply2 (local/bilinear-forms/integrals.hh:97-134)
//                       + LocalLaplaceIntegrand / LocalElementProductIntegrand / sums
//                         (local/integrands/laplace.hh:81-102, product.hh:104-130, combined.hh:212-227)
//   k_element_vector  : LocalElementFunctionalAssembler::apply_local (local/assembler/functional-assemblers.hh:77-86)
//                       + LocalElementIntegralFunctional::apply (local/functionals/integrals.hh:72-98)
//                       + LocalBinaryToUnaryElementIntegrand (local/integrands/conversion.hh:97-117)
//   k_coupling_matrix : LocalCouplingIntersectionBilinearFormAssembler::apply_local (bilinear-form-assemblers.hh:238-278)
//                       + LocalCouplingIntersectionIntegralBilinearForm::apply2 (integrals.hh:197-267)
//                       + InnerCoupling (laplace-ipdg.hh:107-186) / InnerPenalty (ipdg.hh:113-171) / sums (combined.hh:381-431)
//   k_boundary_matrix : LocalIntersectionBilinearFormAssembler::apply_local (bilinear-form-assemblers.hh:380-396)
//                       + LocalIntersectionIntegralBilinearForm::apply2 (integrals.hh:338-369)
//                       + DirichletCoupling (laplace-ipdg.hh:340-368) / BoundaryPenalty (ipdg.hh:254-282)
// Geometry (J^{-T}, integration element) is evaluated per quadrature point in FP64 from the cell's own
// corners.  One thread owns one row of one local matrix, so the scatter into CSR needs no atomics as long
// as no two entities of a launch share a global row: CG element forms are launched once per parity colour
// (2^d colours), DG face forms once per (direction, parity[, periodic wrap]) class.  The colour order is
// fixed, hence results are run-to-run bit-identical.
#include <algorithm>

#include "common.cuh"
#include "kernels.hpp"
#include "local_forms.cuh"

namespace gdtb {

namespace {

// entity enumeration of one launch: idx_k = first[k] + stride[k] * j_k, j_k in [0, cnt[k])
struct Range
{
  long long first[3];
  long long stride[3];
  long long cnt[3];
  __host__ __device__ long long total() const
  {
    return cnt[0] * cnt[1] * cnt[2];
  }
};

__device__ inline void decode(const Range& r, long long t, long long* idx)
{
  const long long j0 = t % r.cnt[0];
  const long long j1 = (t / r.cnt[0]) % r.cnt[1];
  const long long j2 = t / (r.cnt[0] * r.cnt[1]);
  idx[0] = r.first[0] + r.stride[0] * j0;
  idx[1] = r.first[1] + r.stride[1] * j1;
  idx[2] = r.first[2] + r.stride[2] * j2;
}

// ------------------------------------------------------------------------------------------------
// element bilinear forms
// ------------------------------------------------------------------------------------------------
template <int D, int K>
__global__ void __launch_bounds__(128)
    k_element_matrix(const GridDev g, const SpaceDev sp, const __grid_constant__ FormDev f, const Range range,
                     const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                     double* __restrict__ values, int* __restrict__ error_flag)
{
  using L = Loc<D, K>;
  constexpr int N = L::N;
  __shared__ Tables tab;
  load_tables(tab, f);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= range.total() * N)
    return;
  const int i = int(t % N);
  long long idx[3];
  decode(range, t / N, idx);
  double acc[N];
#pragma unroll
  for (int j = 0; j < N; ++j)
    acc[j] = 0.;
  element_row<D, K>(g, f, tab, idx, i, acc);

  // scatter (bilinear-form-assemblers.hh:122-127)
  const long long row = global_index(g, sp, idx, i);
  const long long rb = rowptr[row], re = rowptr[row + 1];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const long long col = global_index(g, sp, idx, j);
    const long long pos = csr_find(colidx, rb, re, (int)col);
    if (pos < 0)
      *error_flag = 1;
    else
      values[pos] += f.scaling * acc[j];
  }
}

// ------------------------------------------------------------------------------------------------
// element functionals (right-hand side)
// ------------------------------------------------------------------------------------------------
template <int D, int K>
__global__ void __launch_bounds__(128)
    k_element_vector(const GridDev g, const SpaceDev sp, const __grid_constant__ FormDev f, const Range range,
                     double* __restrict__ vec, const RowMap rows)
{
  using L = Loc<D, K>;
  constexpr int N = L::N;
  __shared__ Tables tab;
  load_tables(tab, f);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= range.total() * N)
    return;
  const int i = int(t % N);
  long long idx[3];
  decode(range, t / N, idx);
  const long long e = elem_index(g, idx);
  double lower[3], ext[3], inv_ext[3];
  cell_geometry(g, idx, lower, ext);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    inv_ext[k] = 1. / ext[k];
  const double ie = ext[0] * (D > 1 ? ext[1] : 1.) * (D > 2 ? ext[2] : 1.);
  const int ai0 = L::a(i, 0), ai1 = L::a(i, 1), ai2 = L::a(i, 2);
  const IntegrandDev& term = f.terms[0];
  double l = 0.;
  const int m = f.m;
  const int my = D > 1 ? m : 1, mz = D > 2 ? m : 1;
  for (int qz = 0; qz < mz; ++qz)
    for (int qy = 0; qy < my; ++qy)
      for (int qx = 0; qx < m; ++qx) {
        const int q[3] = {qx, qy, qz};
        const double* pv[3];
        const double* pd[3];
        double x[3];
        double wq = 1.;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          pv[k] = tab.phi[k < D ? q[k] : 0];
          pd[k] = tab.dphi[k < D ? q[k] : 0];
          x[k] = k < D ? lower[k] + f.qx[q[k]] * ext[k] : 0.;
          if (k < D)
            wq *= f.qw[q[k]];
        }
        double vi, gi[3];
        basis_at<D>(pv, pd, inv_ext, ai0, ai1, ai2, vi, gi);
        const double xh[3] = {f.qx[qx], D > 1 ? f.qx[qy] : 0., D > 2 ? f.qx[qz] : 0.};
        const EvalPt pt = {qx + m * (qy + my * qz), idx, xh};
        const double w = fn_scalar(term.diffusion, g, e, x, pt);
        const double fx = fn_scalar(term.weight, g, e, x, pt);
        const double v = (w * vi) * fx; // conversion.hh:109-116 -> product.hh:126-128
        l += v * ie * wq;               // local/functionals/integrals.hh:96
      }
  long long row = global_index(g, sp, idx, i);
  if (rows.n > 0) { // slab: only the owned rows live in the local vector
    long long local = -1;
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < rows.n && row >= rows.begin[r] && row < rows.end[r])
        local = rows.local[r] + (row - rows.begin[r]);
    if (local < 0)
      return;
    row = local;
  }
  vec[row] += l; // functional-assemblers.hh:84-85
}

// ------------------------------------------------------------------------------------------------
// intersections
// ------------------------------------------------------------------------------------------------
// One thread per (face, row ii in [0, 2N)): ii < N is row ii of the inside element (blocks in_in, in_out),
// ii >= N is row ii-N of the outside element (blocks out_in, out_out).
template <int D, int K>
__global__ void __launch_bounds__(128)
    k_coupling_matrix(const GridDev g, const SpaceDev sp, const __grid_constant__ FormDev f, const Range range, int k,
                      int wrap, const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                      double* __restrict__ values, int* __restrict__ error_flag)
{
  using L = Loc<D, K>;
  constexpr int N = L::N;
  __shared__ Tables tab;
  load_tables(tab, f);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= range.total() * 2 * N)
    return;
  const int ii = int(t % (2 * N));
  const bool row_inside = ii < N;
  const int i = row_inside ? ii : ii - N;
  long long idx_in[3], idx_out[3];
  decode(range, t / (2 * N), idx_in);
  // inner face: inside = lower element, seen through its upper face (s = 1);
  // periodic wrap face: inside = element with e_k = 0 (the lower index), seen through its lower face (s = 0)
  const int s = wrap ? 0 : 1;
  idx_out[0] = idx_in[0];
  idx_out[1] = idx_in[1];
  idx_out[2] = idx_in[2];
  idx_out[k] = wrap ? g.n[k] - 1 : idx_in[k] + 1;
  double acc_a[N], acc_b[N]; // columns of the inside / outside element
#pragma unroll
  for (int j = 0; j < N; ++j)
    acc_a[j] = acc_b[j] = 0.;
  coupling_row<D, K>(g, f, tab, idx_in, idx_out, k, s, row_inside, i, acc_a, acc_b);

  const long long row = global_index(g, sp, row_inside ? idx_in : idx_out, i);
  const long long rb = rowptr[row], re = rowptr[row + 1];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const long long ca = global_index(g, sp, idx_in, j);
    const long long cb = global_index(g, sp, idx_out, j);
    const long long pa = csr_find(colidx, rb, re, (int)ca);
    const long long pb = csr_find(colidx, rb, re, (int)cb);
    if (pa < 0 || pb < 0)
      *error_flag = 1;
    else {
      values[pa] += f.scaling * acc_a[j];
      values[pb] += f.scaling * acc_b[j];
    }
  }
}

// One thread per (boundary face, row i of the inside element).
template <int D, int K>
__global__ void __launch_bounds__(128)
    k_boundary_matrix(const GridDev g, const SpaceDev sp, const __grid_constant__ FormDev f, const Range range, int k,
                      int s, const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                      double* __restrict__ values, int* __restrict__ error_flag)
{
  using L = Loc<D, K>;
  constexpr int N = L::N;
  __shared__ Tables tab;
  load_tables(tab, f);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= range.total() * N)
    return;
  const int i = int(t % N);
  long long idx[3];
  decode(range, t / N, idx);
  double acc[N];
#pragma unroll
  for (int j = 0; j < N; ++j)
    acc[j] = 0.;
  boundary_row<D, K>(g, f, tab, idx, k, s, i, acc);
  const long long row = global_index(g, sp, idx, i);
  const long long rb = rowptr[row], re = rowptr[row + 1];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const long long col = global_index(g, sp, idx, j);
    const long long pos = csr_find(colidx, rb, re, (int)col);
    if (pos < 0)
      *error_flag = 1;
    else
      values[pos] += f.scaling * acc[j];
  }
}

// ------------------------------------------------------------------------------------------------
// host-side launch helpers
// ------------------------------------------------------------------------------------------------
inline unsigned blocks_for(long long threads, int block)
{
  return (unsigned)((threads + block - 1) / block);
}

// element ranges: all elements of the process' layers, optionally restricted to one parity colour
bool element_range(const GridDev& g, bool coloured, int colour, Range& r)
{
  const int last = g.d - 1;
  for (int k = 0; k < 3; ++k) {
    long long lo = 0, hi = k < g.d ? g.n[k] : 1;
    if (k == last) {
      lo = g.layer_lo;
      hi = g.layer_hi;
    }
    if (coloured && k < g.d) {
      const int c = (colour >> k) & 1;
      long long first = lo + (((lo & 1) != c) ? 1 : 0);
      r.first[k] = first;
      r.stride[k] = 2;
      r.cnt[k] = first < hi ? (hi - first + 1) / 2 : 0;
    } else {
      r.first[k] = lo;
      r.stride[k] = 1;
      r.cnt[k] = hi - lo;
    }
  }
  return r.total() > 0;
}

template <template <int, int> class Launcher, class... Args>
int dispatch_dk(int d, int K, Args&&... args)
{
  switch (d * 10 + K) {
    case 10: return Launcher<1, 0>::run(args...);
    case 11: return Launcher<1, 1>::run(args...);
    case 12: return Launcher<1, 2>::run(args...);
    case 13: return Launcher<1, 3>::run(args...);
    case 20: return Launcher<2, 0>::run(args...);
    case 21: return Launcher<2, 1>::run(args...);
    case 22: return Launcher<2, 2>::run(args...);
    case 23: return Launcher<2, 3>::run(args...);
    case 30: return Launcher<3, 0>::run(args...);
    case 31: return Launcher<3, 1>::run(args...);
    case 32: return Launcher<3, 2>::run(args...);
    case 33: return Launcher<3, 3>::run(args...);
    default:
      return fail(GDTB_ERR_FINITE_ELEMENT, "unsupported (dimension, order) combination for the generic kernels");
  }
}

template <int D, int K>
struct ElementMatrixLauncher
{
  static int run(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, const long long* rowptr,
                 const int* colidx, double* values, int* error_flag)
  {
    time_begin(L, KF_ELEMENT_MATRIX);
    const bool coloured = sp.kind == GDTB_SPACE_CG;
    const int ncol = coloured ? (1 << g.d) : 1;
    for (int c = 0; c < ncol; ++c) {
      Range r;
      if (!element_range(g, coloured, c, r))
        continue;
      const long long threads = r.total() * Loc<D, K>::N;
      k_element_matrix<D, K><<<blocks_for(threads, 128), 128, 0, L.stream>>>(g, sp, f, r, rowptr, colidx, values,
                                                                             error_flag);
      L.count++;
    }
    time_end(L, KF_ELEMENT_MATRIX);
    GDTB_CUDA(cudaGetLastError());
    return GDTB_OK;
  }
};

template <int D, int K>
struct ElementVectorLauncher
{
  static int run(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, double* vec, const RowMap& rows)
  {
    time_begin(L, KF_ELEMENT_VECTOR);
    const bool coloured = sp.kind == GDTB_SPACE_CG;
    const int ncol = coloured ? (1 << g.d) : 1;
    for (int c = 0; c < ncol; ++c) {
      Range r;
      if (!element_range(g, coloured, c, r))
        continue;
      const long long threads = r.total() * Loc<D, K>::N;
      k_element_vector<D, K><<<blocks_for(threads, 128), 128, 0, L.stream>>>(g, sp, f, r, vec, rows);
      L.count++;
    }
    time_end(L, KF_ELEMENT_VECTOR);
    GDTB_CUDA(cudaGetLastError());
    return GDTB_OK;
  }
};

template <int D, int K>
struct CouplingLauncher
{
  static int run(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, int filter,
                 const long long* rowptr, const int* colidx, double* values, int* error_flag)
  {
    time_begin(L, KF_COUPLING_MATRIX);
    // classes per direction: inner faces with even / odd lower element, then the periodic wrap faces
    for (int k = 0; k < g.d; ++k)
      for (int cls = 0; cls < 3; ++cls) {
        const bool wrap = cls == 2;
        if (wrap && !(filter == GDTB_FILTER_INNER_AND_PERIODIC_ONCE && (g.periodic & (1 << k))))
          continue;
        Range r;
        element_range(g, false, 0, r);
        if (wrap) {
          if (g.n[k] < 2)
            continue;
          r.first[k] = 0;
          r.stride[k] = 1;
          r.cnt[k] = 1;
        } else {
          r.first[k] = cls;
          r.stride[k] = 2;
          const long long nfaces = g.n[k] - 1; // lower elements 0 .. n-2
          r.cnt[k] = nfaces > cls ? (nfaces - cls + 1) / 2 : 0;
        }
        if (r.total() <= 0)
          continue;
        const long long threads = r.total() * 2 * Loc<D, K>::N;
        k_coupling_matrix<D, K><<<blocks_for(threads, 128), 128, 0, L.stream>>>(g, sp, f, r, k, wrap ? 1 : 0, rowptr,
                                                                                colidx, values, error_flag);
        L.count++;
      }
    time_end(L, KF_COUPLING_MATRIX);
    GDTB_CUDA(cudaGetLastError());
    return GDTB_OK;
  }
};

template <int D, int K>
struct BoundaryLauncher
{
  static int run(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, const long long* rowptr,
                 const int* colidx, double* values, int* error_flag)
  {
    time_begin(L, KF_BOUNDARY_MATRIX);
    for (int k = 0; k < g.d; ++k) {
      if (g.periodic & (1 << k))
        continue; // periodic intersections have a neighbour
      for (int s = 0; s < 2; ++s) {
        Range r;
        element_range(g, false, 0, r);
        r.first[k] = s ? g.n[k] - 1 : 0;
        r.stride[k] = 1;
        r.cnt[k] = 1;
        if (r.total() <= 0)
          continue;
        const long long threads = r.total() * Loc<D, K>::N;
        k_boundary_matrix<D, K><<<blocks_for(threads, 128), 128, 0, L.stream>>>(g, sp, f, r, k, s, rowptr, colidx,
                                                                                values, error_flag);
        L.count++;
      }
    }
    time_end(L, KF_BOUNDARY_MATRIX);
    GDTB_CUDA(cudaGetLastError());
    return GDTB_OK;
  }
};

struct SampleRule
{
  double qx[MAX_Q1D];
};

// GridFunction -> one value (tensor) per quadrature point of the tensor Gauss rule with m points per direction: what
// LocalLaplaceIntegrand / LocalElementProductIntegrand::evaluate ask the bound local function for at every point
// (laplace.hh:96, product.hh:124), x_q = lower + xhat_q * ext like geometry.global(xhat_q) [EXT]
__global__ void __launch_bounds__(256)
    k_sample_function(const GridDev g, const FnDev f, const int m, const SampleRule rule, const int tensor,
                      const long long e_begin, const long long e_end, double* __restrict__ out)
{
  const double* qx = rule.qx;
  const int d = g.d;
  const int nq = m * (d > 1 ? m : 1) * (d > 2 ? m : 1);
  const long long total = (e_end - e_begin) * nq;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long el = t / nq;
    const int q = int(t - el * nq);
    const long long e = e_begin + el;
    long long idx[3];
    elem_coords(g, e, idx);
    double lower[3], ext[3];
    cell_geometry(g, idx, lower, ext);
    const int qk[3] = {q % m, d > 1 ? (q / m) % m : 0, d > 2 ? q / (m * m) : 0};
    double xh[3], x[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      xh[k] = k < d ? qx[qk[k]] : 0.;
      x[k] = k < d ? lower[k] + xh[k] * ext[k] : 0.;
    }
    const EvalPt pt = {q, idx, xh};
    if (tensor) {
      double T[9];
      fn_tensor(f, g, e, x, pt, T);
      double* dst = out + t * (d * d);
      for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c)
          dst[r * d + c] = T[r * 3 + c];
    } else
      out[t] = fn_scalar(f, g, e, x, pt);
  }
}

} // namespace

int launch_sample_function(Launch& L, const GridDev& g, const FnDev& f, int m, const double* qx, int tensor,
                           long long e_begin, long long e_end, double* out)
{
  SampleRule rule;
  for (int q = 0; q < MAX_Q1D; ++q)
    rule.qx[q] = q < m ? qx[q] : 0.;
  const int nq = m * (g.d > 1 ? m : 1) * (g.d > 2 ? m : 1);
  const long long total = (e_end - e_begin) * nq;
  if (total <= 0)
    return GDTB_OK;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)L.sm_count * 32);
  k_sample_function<<<grid, 256, 0, L.stream>>>(g, f, m, rule, tensor, e_begin, e_end, out);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

namespace {
} // namespace

int launch_element_matrix(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, const long long* rowptr,
                          const int* colidx, double* values, int* error_flag)
{
  return dispatch_dk<ElementMatrixLauncher>(g.d, sp.K, L, g, sp, f, rowptr, colidx, values, error_flag);
}

int launch_element_vector(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, double* vec,
                          const RowMap& rows)
{
  return dispatch_dk<ElementVectorLauncher>(g.d, sp.K, L, g, sp, f, vec, rows);
}

int launch_coupling_matrix(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, int filter,
                           const long long* rowptr, const int* colidx, double* values, int* error_flag)
{
  return dispatch_dk<CouplingLauncher>(g.d, sp.K, L, g, sp, f, filter, rowptr, colidx, values, error_flag);
}

int launch_boundary_matrix(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, const long long* rowptr,
                           const int* colidx, double* values, int* error_flag)
{
  return dispatch_dk<BoundaryLauncher>(g.d, sp.K, L, g, sp, f, rowptr, colidx, values, error_flag);
}

} // namespace gdtb

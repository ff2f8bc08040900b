// dune-gdt_b200/csrc/solve.cu -- the callers on either side of the assembly hot path, on device:
//   n1  DirichletConstraints: collection of the Dirichlet DoFs + apply(matrix, vector)
//       (dune/gdt/tools/dirichlet-constraints.hh:85-110, 122-184)
//   n2  ConstMatrixOperator::apply = CSR mat-vec, apply_inverse = Krylov solve (operators/matrix-based.hh:121-159);
//       [EXT] XT::LA::make_solver -- here CG and BiCGStab with optional Jacobi preconditioning, the whole iteration
//       resident on the device: scalars never visit the host, check_every iterations are replayed as one CUDA graph
//       between two convergence checks
//   n4  BilinearForm::apply2 with (error, error) for the H^1-semi / L^2 norms (operators/bilinear-form.hh:340-440,
//       examples/stationary-heat-equation.cc:116-127) and default_interpolation into Lagrange spaces
//       (interpolations/default.hh:40-83)
// All reductions are two-stage with a fixed launch geometry, so results are run-to-run bit-identical.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "handles.hpp"

struct gdtb_dirichlet
{
  gdtb_ctx* ctx;
  GridDev grid;
  SpaceDev space;
  unsigned mask;
  long long n_dofs;
  long long* d_dofs;
  unsigned char* d_flags;
};

namespace {

constexpr int RED_BLOCK = 256;  // threads per block of every reducing kernel
constexpr int MAX_PARTIALS = 2048;

// ------------------------------------------------------------------------------------------------------------------
// n1 Dirichlet constraints
// ------------------------------------------------------------------------------------------------------------------
// DirichletConstraints::apply_local (dirichlet-constraints.hh:85-110): for every boundary intersection of Dirichlet
// type, the local DoFs whose local key lies on the intersection or one of its sub-entities, i.e. (Lagrange, cube) the
// lattice points with a_k = 0 / K on face (k, s); set semantics -> idempotent flag writes.
__global__ void k_dirichlet_flags(const GridDev g, const SpaceDev sp, unsigned mask, unsigned char* __restrict__ flags)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne)
    return;
  long long idx[3];
  elem_coords(g, e, idx);
  const int K = sp.K, n1 = K + 1;
  for (int k = 0; k < g.d; ++k)
    for (int s = 0; s < 2; ++s) {
      const bool on_boundary = s ? idx[k] == g.n[k] - 1 : idx[k] == 0;
      if (!on_boundary || (g.periodic >> k & 1) || !(mask >> (2 * k + s) & 1))
        continue;
      for (int i = 0; i < sp.nloc; ++i) {
        const int a = k == 0 ? i % n1 : (k == 1 ? (i / n1) % n1 : i / (n1 * n1));
        if (a == (s ? K : 0))
          flags[global_index(g, sp, idx, i)] = 1;
      }
    }
}

// unit_row / unit_col (clear_row / clear_col) of every flagged DoF in one pass over the rows [row_begin, row_end):
// LANES consecutive threads share a row
template <int LANES>
__global__ void k_dirichlet_apply(long long row_begin, long long row_end, const long long* __restrict__ rowptr,
                                  const int* __restrict__ colidx, long long value_offset, double* __restrict__ values,
                                  const unsigned char* __restrict__ flags, int only_clear, int ensure_symmetry,
                                  int* __restrict__ error_flag)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = row_begin + t / LANES;
  const int lane = int(t % LANES);
  if (row >= row_end)
    return;
  const bool rflag = flags[row] != 0;
  if (!rflag && !ensure_symmetry)
    return;
  const long long b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  for (long long p = b + lane; p < e; p += LANES) {
    const int c = __ldg(colidx + p);
    if (rflag) {
      const bool is_diag = c == row;
      values[p - value_offset] = (is_diag && !only_clear) ? 1. : 0.;
    } else if (flags[c])
      values[p - value_offset] = 0.;
  }
  // XT::LA unit_row requires the diagonal entry to be part of the pattern
  if (rflag && !only_clear && lane == 0 && csr_find(colidx, b, e, (int)row) < 0)
    atomicExch(error_flag, 2);
}

__global__ void k_dirichlet_vector(const long long* __restrict__ dofs, long long n, long long row_begin,
                                   long long row_end, double* __restrict__ vec)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n)
    return;
  const long long dof = dofs[t];
  if (dof >= row_begin && dof < row_end)
    vec[dof - row_begin] = 0.;
}

// ------------------------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v)
{
  __shared__ double warp_sums[RED_BLOCK / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads(); // protects warp_sums against the previous call
  if ((threadIdx.x & 31) == 0)
    warp_sums[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.;
#pragma unroll
  for (int w = 0; w < RED_BLOCK / 32; ++w)
    s += warp_sums[w];
  return s; // every thread holds the block sum
}

// fixed-order sum of n <= MAX_PARTIALS partial results, identical in every thread of the block
__device__ __forceinline__ double sum_partials(const double* __restrict__ partial, int n)
{
  double v = 0.;
  for (int i = threadIdx.x; i < n; i += RED_BLOCK)
    v += partial[i];
  return block_sum(v);
}

__global__ void __launch_bounds__(RED_BLOCK) k_finish_sum(const double* __restrict__ partial, int n,
                                                          double* __restrict__ out)
{
  const double s = sum_partials(partial, n);
  if (threadIdx.x == 0)
    *out = s;
}

// ------------------------------------------------------------------------------------------------------------------
// n2 CSR mat-vec (+ fused dot with a second vector)
// ------------------------------------------------------------------------------------------------------------------
// y = A x for the rows [0, rows); LANES threads per row, shuffle reduction in a fixed order.  If `partial` is given,
// block b also writes partial[b] = sum over its rows of w[row] * y[row] (w = x for the CG's p.Ap).
template <int LANES>
__global__ void __launch_bounds__(RED_BLOCK)
    k_spmv(long long rows, const long long* __restrict__ rowptr, const int* __restrict__ colidx,
           const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y,
           const double* __restrict__ w, double* __restrict__ partial)
{
  constexpr int ROWS_PER_BLOCK = RED_BLOCK / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  double acc_dot = 0.;
  for (long long row0 = (long long)blockIdx.x * ROWS_PER_BLOCK; row0 < rows;
       row0 += (long long)gridDim.x * ROWS_PER_BLOCK) {
    const long long row = row0 + sub;
    double s = 0.;
    if (row < rows) {
      const long long b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
      for (long long p = b + lane; p < e; p += LANES)
        s += __ldg(values + p) * __ldg(x + __ldg(colidx + p));
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1)
      s += __shfl_xor_sync(0xffffffffu, s, o);
    if (row < rows && lane == 0) {
      y[row] = s;
      if (w)
        acc_dot += w[row] * s;
    }
  }
  if (partial) {
    const double bs = block_sum(acc_dot);
    if (threadIdx.x == 0)
      partial[blockIdx.x] = bs;
  }
}

__global__ void k_extract_inverse_diagonal(long long rows, const long long* __restrict__ rowptr,
                                           const int* __restrict__ colidx, const double* __restrict__ values,
                                           double* __restrict__ dinv, int use_diag)
{
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows)
    return;
  double d = 1.;
  if (use_diag) {
    const long long p = csr_find(colidx, __ldg(rowptr + row), __ldg(rowptr + row + 1), (int)row);
    if (p >= 0 && values[p] != 0.)
      d = 1. / values[p];
  }
  dinv[row] = d;
}

// ---- conjugate gradients ------------------------------------------------------------------------------------------
// state: x, r, z, p, q and the partial-sum arrays ppq (p.q), prz[2] (r.z of the current / previous iteration), prr.
// No scalar is ever written by more than one block or read while it is written: alpha and beta are recomputed by every
// block from the partial arrays in the same fixed order.

// r = b - A x (q holds A x), z = dinv r, p = z; prz[cur], prr
__global__ void __launch_bounds__(RED_BLOCK)
    k_cg_init(long long n, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
              double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ prz,
              double* __restrict__ prr)
{
  double rz = 0., rr = 0.;
  for (long long i = (long long)blockIdx.x * RED_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * RED_BLOCK) {
    const double ri = b[i] - q[i];
    const double zi = dinv[i] * ri;
    r[i] = ri;
    z[i] = zi;
    p[i] = zi;
    rz += ri * zi;
    rr += ri * ri;
  }
  rz = block_sum(rz);
  rr = block_sum(rr);
  if (threadIdx.x == 0) {
    prz[blockIdx.x] = rz;
    prr[blockIdx.x] = rr;
  }
}

// alpha = rz / pq; x += alpha p; r -= alpha q; z = dinv r; prz_new, prr
__global__ void __launch_bounds__(RED_BLOCK)
    k_cg_update_xr(long long n, int nb_spmv, const double* __restrict__ ppq, int nb, const double* __restrict__ prz_cur,
                   const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ dinv,
                   double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                   double* __restrict__ prz_new, double* __restrict__ prr)
{
  const double pq = sum_partials(ppq, nb_spmv);
  const double rz_cur = sum_partials(prz_cur, nb);
  const double alpha = pq != 0. ? rz_cur / pq : 0.;
  double rz = 0., rr = 0.;
  for (long long i = (long long)blockIdx.x * RED_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * RED_BLOCK) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * q[i];
    const double zi = dinv[i] * ri;
    r[i] = ri;
    z[i] = zi;
    rz += ri * zi;
    rr += ri * ri;
  }
  rz = block_sum(rz);
  rr = block_sum(rr);
  if (threadIdx.x == 0) {
    prz_new[blockIdx.x] = rz;
    prr[blockIdx.x] = rr;
  }
}

// beta = rz_new / rz_old; p = z + beta p; block 0 also records |r|^2 of this iteration
__global__ void __launch_bounds__(RED_BLOCK)
    k_cg_update_p(long long n, int nb, const double* __restrict__ prz_new, const double* __restrict__ prz_old,
                  const double* __restrict__ prr, const double* __restrict__ z, double* __restrict__ p,
                  double* __restrict__ rr_history)
{
  const double rz_new = sum_partials(prz_new, nb);
  const double rz_old = sum_partials(prz_old, nb);
  const double beta = rz_old != 0. ? rz_new / rz_old : 0.;
  if (blockIdx.x == 0) {
    const double rr = sum_partials(prr, nb);
    if (threadIdx.x == 0)
      *rr_history = rr;
  }
  for (long long i = (long long)blockIdx.x * RED_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * RED_BLOCK)
    p[i] = z[i] + beta * p[i];
}

// ---- BiCGStab (right-preconditioned) ----------------------------------------------------------------------------------
// scalar state S (device): [0] rho, [1] alpha, [2] omega, [3] rho_new, [4] beta, [5] |r|^2
enum
{
  S_RHO = 0,
  S_ALPHA,
  S_OMEGA,
  S_RHO_NEW,
  S_BETA,
  S_RR,
  S_COUNT
};

// up to two dot products in one pass: partial_a[b] = sum a1 a2, partial_b[b] = sum b1 b2
__global__ void __launch_bounds__(RED_BLOCK)
    k_dot2(long long n, const double* __restrict__ a1, const double* __restrict__ a2, const double* __restrict__ b1,
           const double* __restrict__ b2, double* __restrict__ partial_a, double* __restrict__ partial_b)
{
  double sa = 0., sb = 0.;
  for (long long i = (long long)blockIdx.x * RED_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * RED_BLOCK) {
    sa += a1[i] * a2[i];
    if (b1)
      sb += b1[i] * b2[i];
  }
  sa = block_sum(sa);
  sb = block_sum(sb);
  if (threadIdx.x == 0) {
    partial_a[blockIdx.x] = sa;
    if (b1)
      partial_b[blockIdx.x] = sb;
  }
}

enum
{
  OP_INIT = 0, // rho = alpha = omega = 1, rr = sum a
  OP_BETA,     // rho_new = sum a; beta = (rho_new / rho) (alpha / omega); rho = rho_new
  OP_ALPHA,    // alpha = rho / sum a
  OP_OMEGA,    // omega = sum a / sum b  (t.s / t.t)
  OP_RR        // rr = sum a, recorded into history
};

__global__ void __launch_bounds__(RED_BLOCK)
    k_bicg_scalar(int op, const double* __restrict__ pa, const double* __restrict__ pb, int nb, double* __restrict__ S,
                  double* __restrict__ history)
{
  const double a = sum_partials(pa, nb);
  const double b = pb ? sum_partials(pb, nb) : 0.;
  if (threadIdx.x != 0)
    return;
  switch (op) {
    case OP_INIT:
      S[S_RHO] = S[S_ALPHA] = S[S_OMEGA] = 1.;
      S[S_RR] = a;
      break;
    case OP_BETA: {
      const double rho = S[S_RHO], omega = S[S_OMEGA];
      S[S_RHO_NEW] = a;
      S[S_BETA] = (rho != 0. && omega != 0.) ? (a / rho) * (S[S_ALPHA] / omega) : 0.;
      S[S_RHO] = a;
      break;
    }
    case OP_ALPHA:
      S[S_ALPHA] = a != 0. ? S[S_RHO] / a : 0.;
      break;
    case OP_OMEGA:
      S[S_OMEGA] = b != 0. ? a / b : 0.;
      break;
    case OP_RR:
      S[S_RR] = a;
      if (history)
        *history = a;
      break;
  }
}

// p = r + beta (p - omega v); y = dinv p
__global__ void k_bicg_p(long long n, const double* __restrict__ S, const double* __restrict__ r,
                         const double* __restrict__ v, const double* __restrict__ dinv, double* __restrict__ p,
                         double* __restrict__ y)
{
  const double beta = S[S_BETA], omega = S[S_OMEGA];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double pi = r[i] + beta * (p[i] - omega * v[i]);
    p[i] = pi;
    y[i] = dinv[i] * pi;
  }
}

// s = r - alpha v; z = dinv s
__global__ void k_bicg_s(long long n, const double* __restrict__ S, const double* __restrict__ r,
                         const double* __restrict__ v, const double* __restrict__ dinv, double* __restrict__ s,
                         double* __restrict__ z)
{
  const double alpha = S[S_ALPHA];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double si = r[i] - alpha * v[i];
    s[i] = si;
    z[i] = dinv[i] * si;
  }
}

// x += alpha y + omega z; r = s - omega t
__global__ void k_bicg_xr(long long n, const double* __restrict__ S, const double* __restrict__ y,
                          const double* __restrict__ z, const double* __restrict__ s, const double* __restrict__ t,
                          double* __restrict__ x, double* __restrict__ r)
{
  const double alpha = S[S_ALPHA], omega = S[S_OMEGA];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] += alpha * y[i] + omega * z[i];
    r[i] = s[i] - omega * t[i];
  }
}

__global__ void k_residual(long long n, const double* __restrict__ b, const double* __restrict__ q,
                           double* __restrict__ r, double* __restrict__ rhat)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double ri = b[i] - q[i];
    r[i] = ri;
    rhat[i] = ri;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// n4 norms and interpolation
// ------------------------------------------------------------------------------------------------------------------
struct Apply2Params
{
  GridDev g;
  SpaceDev sp;
  FormDev form;
  FnDev f;
  int has_f;
  int has_u;
};

// BilinearForm::compute_locally (bilinear-form.hh:340-352) -> LocalElementIntegralBilinearForm::apply2
// (local/bilinear-forms/integrals.hh:97-134) with the one-function "bases" {e}, e = u_h - f
__global__ void __launch_bounds__(RED_BLOCK)
    k_bilinear_apply2(const Apply2Params* __restrict__ pp, const double* __restrict__ dofs, double* __restrict__ partial)
{
  const Apply2Params& P = *pp;
  const GridDev& g = P.g;
  const SpaceDev& sp = P.sp;
  const FormDev& F = P.form;
  const int D = g.d, n1 = sp.K + 1, m = F.m;
  const int my = D > 1 ? m : 1, mz = D > 2 ? m : 1;
  double acc = 0.;
  for (long long e = (long long)blockIdx.x * RED_BLOCK + threadIdx.x; e < g.ne; e += (long long)gridDim.x * RED_BLOCK) {
    long long idx[3];
    elem_coords(g, e, idx);
    double lower[3], ext[3], inv_ext[3];
    cell_geometry(g, idx, lower, ext);
    for (int k = 0; k < 3; ++k)
      inv_ext[k] = 1. / ext[k];
    const double ie = ext[0] * (D > 1 ? ext[1] : 1.) * (D > 2 ? ext[2] : 1.);
    double u[MAX_NLOC];
    if (P.has_u)
      for (int i = 0; i < sp.nloc; ++i)
        u[i] = dofs[global_index(g, sp, idx, i)];
    double local = 0.;
    for (int qz = 0; qz < mz; ++qz)
      for (int qy = 0; qy < my; ++qy)
        for (int qx = 0; qx < m; ++qx) {
          const int q[3] = {qx, qy, qz};
          double x[3] = {0., 0., 0.}, w = 1.;
          for (int k = 0; k < D; ++k) {
            x[k] = lower[k] + F.qx[q[k]] * ext[k];
            w *= F.qw[q[k]];
          }
          double val = 0., grad[3] = {0., 0., 0.};
          if (P.has_u) {
            for (int i = 0; i < sp.nloc; ++i) {
              const int a0 = i % n1, a1 = D > 1 ? (i / n1) % n1 : 0, a2 = D > 2 ? i / (n1 * n1) : 0;
              const double v0 = F.phi[qx][a0], v1 = D > 1 ? F.phi[qy][a1] : 1., v2 = D > 2 ? F.phi[qz][a2] : 1.;
              val += u[i] * (v0 * v1 * v2);
              grad[0] += u[i] * (inv_ext[0] * (F.dphi[qx][a0] * v1 * v2));
              if (D > 1)
                grad[1] += u[i] * (inv_ext[1] * (v0 * F.dphi[qy][a1] * v2));
              if (D > 2)
                grad[2] += u[i] * (inv_ext[2] * (v0 * v1 * F.dphi[qz][a2]));
            }
          }
          if (P.has_f) {
            double fg[3] = {0., 0., 0.};
            if (P.f.kind == GDTB_FN_QP_VALUE_GRAD) {
              const double* src = P.f.data + (e * (long long)(m * my * mz) + (qx + m * (qy + my * qz))) * (1 + D);
              val -= __ldg(src);
              for (int k = 0; k < D; ++k)
                fg[k] = __ldg(src + 1 + k);
            } else
              val -= fn_scalar(P.f, D, e, x);
            if (P.f.kind == GDTB_FN_BUILTIN)
              builtin_grad(P.f, D, x, fg);
            for (int k = 0; k < 3; ++k)
              grad[k] -= fg[k];
          }
          double v = 0.;
          const double xh[3] = {F.qx[qx], D > 1 ? F.qx[qy] : 0., D > 2 ? F.qx[qz] : 0.};
          const EvalPt pt = {qx + m * (qy + my * qz), idx, xh};
          for (int t = 0; t < F.n_terms; ++t) {
            if (F.terms[t].kind == GDTB_INT_LAPLACE) {
              double kap[9], kg[3] = {0., 0., 0.};
              fn_tensor(F.terms[t].diffusion, g, e, x, pt, kap);
              for (int r = 0; r < D; ++r)
                for (int c = 0; c < D; ++c)
                  kg[r] += kap[r * 3 + c] * grad[c];
              double s = 0.;
              for (int r = 0; r < D; ++r)
                s += kg[r] * grad[r];
              v += s; // laplace.hh:101
            } else
              v += (fn_scalar(F.terms[t].diffusion, g, e, x, pt) * val) * val; // product.hh:128
          }
          local += v * (ie * w); // integrals.hh:119,131
        }
    acc += F.scaling * local;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0)
    partial[blockIdx.x] = acc;
}

// DoFs = f at the Lagrange points; a DoF shared by several elements is written by the last element of the walk that
// contains it (the element with the highest index), like the reference's sequential overwrite
__global__ void k_lagrange_interpolate(const GridDev g, const SpaceDev sp, const FnDev f, double* __restrict__ dofs)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.ne * sp.nloc)
    return;
  const long long e = t / sp.nloc;
  const int i = int(t - e * sp.nloc);
  long long idx[3];
  elem_coords(g, e, idx);
  const int K = sp.K, n1 = K + 1;
  const int a[3] = {i % n1, g.d > 1 ? (i / n1) % n1 : 0, g.d > 2 ? i / (n1 * n1) : 0};
  if (sp.kind == GDTB_SPACE_CG)
    for (int k = 0; k < g.d; ++k)
      if (a[k] == K && idx[k] != g.n[k] - 1)
        return; // a later element owns this lattice point
  double lower[3], ext[3], x[3] = {0., 0., 0.};
  cell_geometry(g, idx, lower, ext);
  for (int k = 0; k < g.d; ++k)
    x[k] = lower[k] + (K > 0 ? double(a[k]) / double(K) : 0.5) * ext[k];
  dofs[global_index(g, sp, idx, i)] = fn_scalar(f, g.d, e, x);
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
struct CsrView
{
  long long rows;
  const long long* rowptr;
  const int* colidx;
  const double* values;
  double avg_row; // nnz / rows, picks the lanes per row
};

int launch_spmv(Launch& L, const CsrView& A, const double* x, double* y, const double* w, double* partial, int nb)
{
  if (A.rows == 0)
    return GDTB_OK;
  auto grid_for = [&](int lanes) {
    const long long rows_per_block = RED_BLOCK / lanes;
    long long blocks = (A.rows + rows_per_block - 1) / rows_per_block;
    if (partial)
      blocks = nb; // fixed geometry: the partial-sum layout (and with it the rounding) never changes
    return (unsigned)std::min<long long>(blocks, 1 << 30);
  };
  if (A.avg_row <= 6.)
    k_spmv<4><<<grid_for(4), RED_BLOCK, 0, L.stream>>>(A.rows, A.rowptr, A.colidx, A.values, x, y, w, partial);
  else if (A.avg_row <= 14.)
    k_spmv<8><<<grid_for(8), RED_BLOCK, 0, L.stream>>>(A.rows, A.rowptr, A.colidx, A.values, x, y, w, partial);
  else if (A.avg_row <= 48.)
    k_spmv<16><<<grid_for(16), RED_BLOCK, 0, L.stream>>>(A.rows, A.rowptr, A.colidx, A.values, x, y, w, partial);
  else
    k_spmv<32><<<grid_for(32), RED_BLOCK, 0, L.stream>>>(A.rows, A.rowptr, A.colidx, A.values, x, y, w, partial);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

int matop_pattern(gdtb_matop* op, const long long** rowptr, const int** colidx)
{
  if (op->pattern) {
    *rowptr = op->pattern->d_rowptr;
    *colidx = op->pattern->d_colidx;
    return GDTB_OK;
  }
  if (!op->d_own_rowptr) {
    // pattern-free operators exist for the CG Q1 and CG Q2 element stencils: materialise the closed-form pattern
    long long nnz = 0;
    GridDev whole = op->grid; // the pattern is the global one even when the operator holds a slab of its rows
    whole.layer_lo = 0;
    whole.layer_hi = whole.n[whole.d - 1];
    if (op->test.kind == GDTB_SPACE_DG)
      GDTB_TRY(pattern_structured_dg(op->ctx->launch, whole, op->test, &op->d_own_rowptr, &op->d_own_colidx, &nnz));
    else if (op->test.K == 2)
      GDTB_TRY(pattern_structured_cg_q2(op->ctx->launch, op->grid, op->test, &op->d_own_rowptr, &op->d_own_colidx, &nnz));
    else
      GDTB_TRY(pattern_structured_cg_q1(op->ctx->launch, op->grid, op->test, &op->d_own_rowptr, &op->d_own_colidx, &nnz));
  }
  *rowptr = op->d_own_rowptr;
  *colidx = op->d_own_colidx;
  return GDTB_OK;
}

} // namespace

namespace gdtb {
int internal_matop_pattern(gdtb_matop* op, const long long** rowptr, const int** colidx)
{
  return matop_pattern(op, rowptr, colidx);
}
} // namespace gdtb

namespace {

int matop_csr(gdtb_matop* op, CsrView& A)
{
  if (op->slab)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "apply / apply_inverse on a slab-partitioned operator is not supported");
  GDTB_TRY(matop_pattern(op, &A.rowptr, &A.colidx));
  A.rows = op->test.size;
  A.values = op->d_values;
  A.avg_row = A.rows > 0 ? double(op->nnz_local) / double(A.rows) : 0.;
  return GDTB_OK;
}

struct DeviceBuffer
{
  void* p = nullptr;
  ~DeviceBuffer()
  {
    cudaFree(p);
  }
  int alloc(size_t bytes)
  {
    if (cudaMalloc(&p, std::max<size_t>(bytes, 8)) != cudaSuccess)
      return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory");
    return GDTB_OK;
  }
  template <class T>
  T* as()
  {
    return static_cast<T*>(p);
  }
};

int solve_csr(gdtb_ctx* ctx, const CsrView& A, const double* d_b, double* d_x, const gdtb_solver_opts* opts_in,
              gdtb_solver_info* info)
{
  gdtb_solver_opts o;
  o.type = GDTB_SOLVER_CG;
  o.preconditioner = GDTB_PRECOND_JACOBI;
  o.max_iter = 0;
  o.check_every = 0;
  o.precision = 0.;
  if (opts_in)
    o = *opts_in;
  if (o.type != GDTB_SOLVER_CG && o.type != GDTB_SOLVER_BICGSTAB)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "apply_inverse: unknown solver type");
  if (o.preconditioner != GDTB_PRECOND_NONE && o.preconditioner != GDTB_PRECOND_JACOBI)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "apply_inverse: unknown preconditioner");
  const long long n = A.rows;
  const long long max_iter = o.max_iter > 0 ? o.max_iter : std::max<long long>(10 * n, 100);
  int check = o.check_every > 0 ? o.check_every : 25;
  if (o.type == GDTB_SOLVER_CG && (check & 1))
    ++check; // an even number of iterations per graph keeps the r.z ping-pong parity fixed
  const double precision = o.precision > 0. ? o.precision : 1e-10;
  Launch& L = ctx->launch;
  cudaStream_t st = L.stream;
  const int nb = (int)std::min<long long>(std::max<long long>((n + RED_BLOCK - 1) / RED_BLOCK, 1),
                                          std::min(MAX_PARTIALS, 4 * L.sm_count));
  const int n_vec = o.type == GDTB_SOLVER_CG ? 5 : 9;
  DeviceBuffer vecs, parts, hist, scal;
  GDTB_TRY(vecs.alloc(sizeof(double) * (size_t)n * n_vec));
  GDTB_TRY(parts.alloc(sizeof(double) * MAX_PARTIALS * 4));
  GDTB_TRY(hist.alloc(sizeof(double) * (size_t)check));
  GDTB_TRY(scal.alloc(sizeof(double) * S_COUNT));
  double* v = vecs.as<double>();
  double *dinv = v, *r = v + n, *w1 = v + 2 * n, *w2 = v + 3 * n, *w3 = v + 4 * n;
  double* P0 = parts.as<double>();
  double *P1 = P0 + MAX_PARTIALS, *P2 = P0 + 2 * MAX_PARTIALS, *P3 = P0 + 3 * MAX_PARTIALS;
  double* history = hist.as<double>();
  double* S = scal.as<double>();
  const unsigned vgrid = (unsigned)std::max<long long>(std::min<long long>((n + 255) / 256, 8LL * L.sm_count), 1);

  k_extract_inverse_diagonal<<<(unsigned)std::max<long long>((n + 255) / 256, 1), 256, 0, st>>>(
      n, A.rowptr, A.colidx, A.values, dinv, o.preconditioner == GDTB_PRECOND_JACOBI);
  L.count++;
  double rr0 = 0., rr = 0.;
  long long it = 0;
  bool converged = false;
  std::vector<double> h_hist((size_t)check);
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int status = GDTB_OK;

  if (o.type == GDTB_SOLVER_CG) {
    double *z = w1, *p = w2, *q = w3;
    double* prz[2] = {P1, P2};
    GDTB_TRY(launch_spmv(L, A, d_x, q, nullptr, nullptr, 0));
    k_cg_init<<<nb, RED_BLOCK, 0, st>>>(n, d_b, q, dinv, r, z, p, prz[0], P3);
    k_finish_sum<<<1, RED_BLOCK, 0, st>>>(P3, nb, S + S_RR);
    L.count += 2;
    GDTB_CUDA(cudaMemcpyAsync(&rr0, S + S_RR, sizeof(double), cudaMemcpyDeviceToHost, st));
    GDTB_CUDA(cudaStreamSynchronize(st));
    rr = rr0;
    converged = rr0 == 0.;
    if (!converged) {
      // one graph = `check` iterations
      GDTB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      long long launched = 0;
      const long long count_before = L.count;
      for (int k = 0; k < check && status == GDTB_OK; ++k) {
        const int cur = k & 1;
        status = launch_spmv(L, A, p, q, p, P0, nb);
        k_cg_update_xr<<<nb, RED_BLOCK, 0, st>>>(n, nb, P0, nb, prz[cur], p, q, dinv, d_x, r, z, prz[cur ^ 1], P3);
        k_cg_update_p<<<nb, RED_BLOCK, 0, st>>>(n, nb, prz[cur ^ 1], prz[cur], P3, z, p, history + k);
        launched += 3;
      }
      cudaError_t err = cudaStreamEndCapture(st, &graph);
      L.count = count_before; // captured launches are counted per replay below
      if (status == GDTB_OK && err != cudaSuccess)
        status = fail(GDTB_ERR_CUDA, std::string("CG graph capture: ") + cudaGetErrorString(err));
      if (status == GDTB_OK && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess)
        status = fail(GDTB_ERR_CUDA, "CG graph instantiation failed");
      while (status == GDTB_OK && !converged && it < max_iter) {
        if (cudaGraphLaunch(exec, st) != cudaSuccess) {
          status = fail(GDTB_ERR_CUDA, "CG graph launch failed");
          break;
        }
        L.count += launched;
        cudaMemcpyAsync(h_hist.data(), history, sizeof(double) * (size_t)check, cudaMemcpyDeviceToHost, st);
        cudaError_t e2 = cudaStreamSynchronize(st);
        if (e2 != cudaSuccess) {
          status = fail(GDTB_ERR_CUDA, std::string("CG: ") + cudaGetErrorString(e2));
          break;
        }
        // the residual history of the replay: first iteration that meets the tolerance (reported; the iterate has
        // moved on to the end of the replay and is at least as converged), stagnation = breakdown (rho or p.q became 0:
        // the update kernels then leave x and r untouched), fail fast instead of spinning until max_iter
        const double rr_before = rr;
        long long first_ok = -1;
        bool moved = false;
        for (int k = 0; k < check; ++k) {
          moved = moved || h_hist[(size_t)k] != rr_before;
          if (first_ok < 0 && std::sqrt(h_hist[(size_t)k]) <= precision * std::sqrt(rr0))
            first_ok = k;
        }
        rr = h_hist[(size_t)check - 1];
        if (first_ok >= 0 && rr == rr && std::sqrt(rr) <= precision * std::sqrt(rr0)) {
          it += first_ok + 1;
          converged = true;
          break;
        }
        it += check;
        if (rr == rr && !moved) {
          status = fail(GDTB_ERR_OPERATOR, "apply_inverse: CG broke down (the residual stopped changing: rho or p.q "
                                           "vanished) (XT::LA::Exceptions::linear_solver_failed)");
          break;
        }
        if (!(rr == rr)) {
          status = fail(GDTB_ERR_OPERATOR, "apply_inverse: CG broke down (NaN residual) -- is the matrix symmetric positive definite?");
          break;
        }
        converged = std::sqrt(rr) <= precision * std::sqrt(rr0);
      }
    }
  } else {
    double *rhat = w1, *p = w2, *vv = w3, *y = v + 5 * n, *s = v + 6 * n, *z = v + 7 * n, *t = v + 8 * n;
    GDTB_TRY(launch_spmv(L, A, d_x, vv, nullptr, nullptr, 0));
    k_residual<<<vgrid, 256, 0, st>>>(n, d_b, vv, r, rhat);
    k_dot2<<<nb, RED_BLOCK, 0, st>>>(n, r, r, nullptr, nullptr, P0, nullptr);
    k_bicg_scalar<<<1, RED_BLOCK, 0, st>>>(OP_INIT, P0, nullptr, nb, S, nullptr);
    L.count += 3;
    GDTB_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * (size_t)n, st));
    GDTB_CUDA(cudaMemsetAsync(vv, 0, sizeof(double) * (size_t)n, st));
    GDTB_CUDA(cudaMemcpyAsync(&rr0, S + S_RR, sizeof(double), cudaMemcpyDeviceToHost, st));
    GDTB_CUDA(cudaStreamSynchronize(st));
    rr = rr0;
    converged = rr0 == 0.;
    if (!converged) {
      GDTB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      long long launched = 0;
      const long long count_before = L.count;
      for (int k = 0; k < check && status == GDTB_OK; ++k) {
        k_dot2<<<nb, RED_BLOCK, 0, st>>>(n, rhat, r, nullptr, nullptr, P0, nullptr);
        k_bicg_scalar<<<1, RED_BLOCK, 0, st>>>(OP_BETA, P0, nullptr, nb, S, nullptr);
        k_bicg_p<<<vgrid, 256, 0, st>>>(n, S, r, vv, dinv, p, y);
        status = launch_spmv(L, A, y, vv, nullptr, nullptr, 0);
        k_dot2<<<nb, RED_BLOCK, 0, st>>>(n, rhat, vv, nullptr, nullptr, P0, nullptr);
        k_bicg_scalar<<<1, RED_BLOCK, 0, st>>>(OP_ALPHA, P0, nullptr, nb, S, nullptr);
        k_bicg_s<<<vgrid, 256, 0, st>>>(n, S, r, vv, dinv, s, z);
        if (status == GDTB_OK)
          status = launch_spmv(L, A, z, t, nullptr, nullptr, 0);
        k_dot2<<<nb, RED_BLOCK, 0, st>>>(n, t, s, t, t, P0, P1);
        k_bicg_scalar<<<1, RED_BLOCK, 0, st>>>(OP_OMEGA, P0, P1, nb, S, nullptr);
        k_bicg_xr<<<vgrid, 256, 0, st>>>(n, S, y, z, s, t, d_x, r);
        k_dot2<<<nb, RED_BLOCK, 0, st>>>(n, r, r, nullptr, nullptr, P0, nullptr);
        k_bicg_scalar<<<1, RED_BLOCK, 0, st>>>(OP_RR, P0, nullptr, nb, S, history + k);
        launched += 13;
      }
      L.count = count_before;
      cudaError_t err = cudaStreamEndCapture(st, &graph);
      if (status == GDTB_OK && err != cudaSuccess)
        status = fail(GDTB_ERR_CUDA, std::string("BiCGStab graph capture: ") + cudaGetErrorString(err));
      if (status == GDTB_OK && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess)
        status = fail(GDTB_ERR_CUDA, "BiCGStab graph instantiation failed");
      while (status == GDTB_OK && !converged && it < max_iter) {
        if (cudaGraphLaunch(exec, st) != cudaSuccess) {
          status = fail(GDTB_ERR_CUDA, "BiCGStab graph launch failed");
          break;
        }
        L.count += launched;
        cudaMemcpyAsync(h_hist.data(), history, sizeof(double) * (size_t)check, cudaMemcpyDeviceToHost, st);
        cudaError_t e2 = cudaStreamSynchronize(st);
        if (e2 != cudaSuccess) {
          status = fail(GDTB_ERR_CUDA, std::string("BiCGStab: ") + cudaGetErrorString(e2));
          break;
        }
        // the residual history of the replay: first iteration that meets the tolerance (reported; the iterate has
        // moved on to the end of the replay and is at least as converged), stagnation = breakdown (rho or p.q became 0:
        // the update kernels then leave x and r untouched), fail fast instead of spinning until max_iter
        const double rr_before = rr;
        long long first_ok = -1;
        bool moved = false;
        for (int k = 0; k < check; ++k) {
          moved = moved || h_hist[(size_t)k] != rr_before;
          if (first_ok < 0 && std::sqrt(h_hist[(size_t)k]) <= precision * std::sqrt(rr0))
            first_ok = k;
        }
        rr = h_hist[(size_t)check - 1];
        if (first_ok >= 0 && rr == rr && std::sqrt(rr) <= precision * std::sqrt(rr0)) {
          it += first_ok + 1;
          converged = true;
          break;
        }
        it += check;
        if (rr == rr && !moved) {
          status = fail(GDTB_ERR_OPERATOR, "apply_inverse: BiCGStab broke down (the residual stopped changing: rho or p.q "
                                           "vanished) (XT::LA::Exceptions::linear_solver_failed)");
          break;
        }
        if (!(rr == rr)) {
          status = fail(GDTB_ERR_OPERATOR, "apply_inverse: BiCGStab broke down (NaN residual)");
          break;
        }
        converged = std::sqrt(rr) <= precision * std::sqrt(rr0);
      }
    }
  }
  if (exec)
    cudaGraphExecDestroy(exec);
  if (graph)
    cudaGraphDestroy(graph);
  if (info) {
    info->iterations = (int32_t)std::min<long long>(it, 2147483647LL);
    info->converged = converged ? 1 : 0;
    info->initial_residual = std::sqrt(rr0);
    info->residual = std::sqrt(rr);
  }
  if (status != GDTB_OK)
    return status;
  if (!converged)
    return fail(GDTB_ERR_OPERATOR, "apply_inverse: the linear solver did not converge within max_iter iterations "
                                   "(XT::LA::Exceptions::linear_solver_failed)");
  return GDTB_OK;
}

} // namespace

// ======================================================================================================================
// C ABI
// ======================================================================================================================
extern "C" {

int gdtb_dirichlet_create(gdtb_ctx* ctx, const gdtb_space* space, uint32_t boundary_mask, gdtb_dirichlet** out)
{
  if (!space || !out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_dirichlet_create: NULL argument");
  GDTB_TRY(internal_check_ctx(ctx));
  auto dc = std::make_unique<gdtb_dirichlet>();
  dc->ctx = ctx;
  dc->grid = space->grid;
  dc->space = space->dev;
  dc->mask = boundary_mask;
  dc->n_dofs = 0;
  dc->d_dofs = nullptr;
  dc->d_flags = nullptr;
  const long long n = space->dev.size;
  Launch& L = ctx->launch;
  if (cudaMalloc(&dc->d_flags, (size_t)std::max<long long>(n, 1)) != cudaSuccess)
    return fail(GDTB_ERR_OUT_OF_MEMORY, "dirichlet: out of device memory");
  // (every exit below frees what has been allocated so far)
  if (cudaMemsetAsync(dc->d_flags, 0, (size_t)std::max<long long>(n, 1), L.stream) != cudaSuccess) {
    cudaFree(dc->d_flags);
    return fail(GDTB_ERR_CUDA, "dirichlet: cudaMemsetAsync failed");
  }
  // FV (P0) elements carry their only local key in the element interior: nothing lies on an intersection
  if (space->dev.kind != GDTB_SPACE_FV && space->dev.K > 0 && dc->grid.ne > 0) {
    k_dirichlet_flags<<<(unsigned)((dc->grid.ne + 255) / 256), 256, 0, L.stream>>>(dc->grid, dc->space, boundary_mask,
                                                                                   dc->d_flags);
    L.count++;
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
      cudaFree(dc->d_flags);
      return fail(GDTB_ERR_CUDA, std::string("k_dirichlet_flags: ") + cudaGetErrorString(err));
    }
  }
  // ascending DoF list (std::set order): stream compaction of the flags
  long long* d_count = nullptr;
  long long* d_list = nullptr;
  void* d_temp = nullptr;
  size_t temp_bytes = 0;
  thrust::counting_iterator<long long> ids(0);
  int status = GDTB_OK;
  if (cudaMalloc(&d_count, sizeof(long long)) != cudaSuccess
      || cudaMalloc(&d_list, sizeof(long long) * (size_t)std::max<long long>(n, 1)) != cudaSuccess)
    status = fail(GDTB_ERR_OUT_OF_MEMORY, "dirichlet: out of device memory");
  if (status == GDTB_OK) {
    cub::DeviceSelect::Flagged(nullptr, temp_bytes, ids, dc->d_flags, d_list, d_count, n, L.stream);
    if (cudaMalloc(&d_temp, std::max<size_t>(temp_bytes, 8)) != cudaSuccess)
      status = fail(GDTB_ERR_OUT_OF_MEMORY, "dirichlet: out of device memory");
  }
  if (status == GDTB_OK) {
    cub::DeviceSelect::Flagged(d_temp, temp_bytes, ids, dc->d_flags, d_list, d_count, n, L.stream);
    L.count++;
    long long count = 0;
    cudaMemcpyAsync(&count, d_count, sizeof(long long), cudaMemcpyDeviceToHost, L.stream);
    if (cudaStreamSynchronize(L.stream) != cudaSuccess)
      status = fail(GDTB_ERR_CUDA, "dirichlet: stream compaction failed");
    else {
      dc->n_dofs = count;
      if (cudaMalloc(&dc->d_dofs, sizeof(long long) * (size_t)std::max<long long>(count, 1)) != cudaSuccess)
        status = fail(GDTB_ERR_OUT_OF_MEMORY, "dirichlet: out of device memory");
      else
        cudaMemcpy(dc->d_dofs, d_list, sizeof(long long) * (size_t)count, cudaMemcpyDeviceToDevice);
    }
  }
  cudaFree(d_temp);
  cudaFree(d_list);
  cudaFree(d_count);
  if (status != GDTB_OK) {
    cudaFree(dc->d_flags);
    cudaFree(dc->d_dofs);
    return status;
  }
  *out = dc.release();
  return GDTB_OK;
}

int gdtb_dirichlet_destroy(gdtb_dirichlet* dc)
{
  if (!dc)
    return GDTB_OK;
  cudaFree(dc->d_dofs);
  cudaFree(dc->d_flags);
  delete dc;
  return GDTB_OK;
}

int64_t gdtb_dirichlet_size(const gdtb_dirichlet* dc)
{
  return dc ? dc->n_dofs : 0;
}

int gdtb_dirichlet_dofs_download(const gdtb_dirichlet* dc, int64_t* dofs)
{
  if (!dc || !dofs)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_dirichlet_dofs_download: NULL argument");
  static_assert(sizeof(int64_t) == sizeof(long long), "");
  GDTB_CUDA(cudaMemcpy(dofs, dc->d_dofs, sizeof(long long) * (size_t)dc->n_dofs, cudaMemcpyDeviceToHost));
  return GDTB_OK;
}

int gdtb_dirichlet_device(const gdtb_dirichlet* dc, const int64_t** d_dofs, const uint8_t** d_flags)
{
  if (!dc)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_dirichlet_device: NULL argument");
  if (d_dofs)
    *d_dofs = reinterpret_cast<const int64_t*>(dc->d_dofs);
  if (d_flags)
    *d_flags = dc->d_flags;
  return GDTB_OK;
}

static int dirichlet_apply_impl(gdtb_dirichlet* dc, long long rows, long long cols, long long row_begin,
                                long long row_end, const long long* rowptr, const int* colidx, long long value_offset,
                                double avg_row, double* d_values, double* d_vector, int only_clear,
                                int ensure_symmetry)
{
  gdtb_ctx* ctx = dc->ctx;
  Launch& L = ctx->launch;
  // the kernel below reuses the context's device error flag: report a pending out-of-pattern error of an earlier
  // asynchronous assembly first instead of clearing it
  if (ctx->error_flag_pending)
    GDTB_TRY(gdtb_ctx_synchronize(ctx));
  if (d_values) {
    // dirichlet-constraints.hh is only meaningful for square operators on the constrained space
    if (rows != dc->space.size || cols != dc->space.size)
      return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "dirichlet apply: the matrix does not live on the constrained space");
    const long long nrows = row_end - row_begin;
    if (nrows > 0) {
      GDTB_CUDA(cudaMemsetAsync(ctx->d_error_flag, 0, sizeof(int), L.stream));
      const int lanes = avg_row <= 6. ? 4 : (avg_row <= 14. ? 8 : (avg_row <= 48. ? 16 : 32));
      const long long threads = nrows * lanes;
      const unsigned grid = (unsigned)((threads + 255) / 256);
#define GDTB_DC_LAUNCH(LN)                                                                                            \
  k_dirichlet_apply<LN><<<grid, 256, 0, L.stream>>>(row_begin, row_end, rowptr, colidx, value_offset, d_values,       \
                                                    dc->d_flags, only_clear, ensure_symmetry, ctx->d_error_flag)
      if (lanes == 4)
        GDTB_DC_LAUNCH(4);
      else if (lanes == 8)
        GDTB_DC_LAUNCH(8);
      else if (lanes == 16)
        GDTB_DC_LAUNCH(16);
      else
        GDTB_DC_LAUNCH(32);
#undef GDTB_DC_LAUNCH
      L.count++;
      GDTB_CUDA(cudaGetLastError());
    }
  }
  if (d_vector && dc->n_dofs > 0) {
    k_dirichlet_vector<<<(unsigned)((dc->n_dofs + 255) / 256), 256, 0, L.stream>>>(dc->d_dofs, dc->n_dofs, row_begin,
                                                                                   row_end, d_vector);
    L.count++;
    GDTB_CUDA(cudaGetLastError());
  }
  int flag = 0;
  if (d_values) {
    GDTB_CUDA(cudaMemcpyAsync(&flag, ctx->d_error_flag, sizeof(int), cudaMemcpyDeviceToHost, L.stream));
  }
  GDTB_CUDA(cudaStreamSynchronize(L.stream));
  if (flag == 2)
    return fail(GDTB_ERR_OPERATOR, "dirichlet apply: unit_row on a row whose diagonal entry is not in the pattern");
  return GDTB_OK;
}

int gdtb_dirichlet_apply(gdtb_dirichlet* dc, gdtb_matop* op, gdtb_vecfun* fun, int only_clear, int ensure_symmetry)
{
  if (!dc)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_dirichlet_apply: NULL constraints");
  GDTB_TRY(internal_check_ctx(dc->ctx));
  if (op) {
    const long long* rowptr;
    const int* colidx;
    GDTB_TRY(matop_pattern(op, &rowptr, &colidx));
    const double avg = double(op->nnz_local) / double(std::max<long long>(op->row_end - op->row_begin, 1));
    GDTB_TRY(dirichlet_apply_impl(dc, op->test.size, op->ansatz.size, op->row_begin, op->row_end, rowptr, colidx,
                                  op->value_offset, avg, op->d_values, nullptr, only_clear, ensure_symmetry));
  }
  if (fun) {
    if (fun->space.size != dc->space.size)
      return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "dirichlet apply: the vector does not live on the constrained space");
    GDTB_TRY(dirichlet_apply_impl(dc, 0, 0, fun->row_begin, fun->row_end, nullptr, nullptr, 0, 0., nullptr, fun->d_vec,
                                  only_clear, ensure_symmetry));
  }
  return GDTB_OK;
}

int gdtb_dirichlet_apply_device(gdtb_dirichlet* dc, const gdtb_pattern* pattern, double* d_values, double* d_vector,
                                int only_clear, int ensure_symmetry)
{
  if (!dc || (d_values && !pattern))
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_dirichlet_apply_device: NULL argument");
  GDTB_TRY(internal_check_ctx(dc->ctx));
  const long long n = dc->space.size;
  if (d_values)
    return dirichlet_apply_impl(dc, pattern->rows, pattern->cols, 0, pattern->rows, pattern->d_rowptr,
                                pattern->d_colidx, 0, double(pattern->nnz) / double(std::max<long long>(pattern->rows, 1)),
                                d_values, d_vector, only_clear, ensure_symmetry);
  return dirichlet_apply_impl(dc, 0, 0, 0, n, nullptr, nullptr, 0, 0., nullptr, d_vector, only_clear, ensure_symmetry);
}

int gdtb_matop_pattern_device(gdtb_matop* op, const int64_t** d_rowptr, const int32_t** d_colidx)
{
  if (!op)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_pattern_device: NULL operator");
  GDTB_TRY(internal_check_ctx(op->ctx));
  const long long* rp;
  const int* ci;
  GDTB_TRY(matop_pattern(op, &rp, &ci));
  if (d_rowptr)
    *d_rowptr = reinterpret_cast<const int64_t*>(rp);
  if (d_colidx)
    *d_colidx = ci;
  return GDTB_OK;
}

int gdtb_matop_apply(gdtb_matop* op, const double* d_source, double* d_range)
{
  if (!op || !d_source || !d_range)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_apply: NULL argument");
  GDTB_TRY(internal_check_ctx(op->ctx));
  CsrView A;
  GDTB_TRY(matop_csr(op, A));
  GDTB_TRY(launch_spmv(op->ctx->launch, A, d_source, d_range, nullptr, nullptr, 0));
  GDTB_CUDA(cudaStreamSynchronize(op->ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_matop_apply_host(gdtb_matop* op, const double* source, double* range)
{
  if (!op || !source || !range)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_apply_host: NULL argument");
  GDTB_TRY(internal_check_ctx(op->ctx));
  DeviceBuffer s, r;
  GDTB_TRY(s.alloc(sizeof(double) * (size_t)op->ansatz.size));
  GDTB_TRY(r.alloc(sizeof(double) * (size_t)op->test.size));
  // uploads on the context's stream: a blocking copy on the legacy stream is not ordered against a non-blocking stream
  cudaStream_t st = op->ctx->launch.stream;
  GDTB_CUDA(cudaMemcpyAsync(s.p, source, sizeof(double) * (size_t)op->ansatz.size, cudaMemcpyHostToDevice, st));
  GDTB_TRY(gdtb_matop_apply(op, s.as<double>(), r.as<double>()));
  GDTB_CUDA(cudaMemcpyAsync(range, r.p, sizeof(double) * (size_t)op->test.size, cudaMemcpyDeviceToHost, st));
  GDTB_CUDA(cudaStreamSynchronize(st));
  return GDTB_OK;
}

int gdtb_csr_apply_device(gdtb_ctx* ctx, const gdtb_pattern* pattern, const double* d_values, const double* d_source,
                          double* d_range)
{
  if (!pattern || !d_values || !d_source || !d_range)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_csr_apply_device: NULL argument");
  GDTB_TRY(internal_check_ctx(ctx));
  CsrView A{pattern->rows, pattern->d_rowptr, pattern->d_colidx, d_values,
            double(pattern->nnz) / double(std::max<long long>(pattern->rows, 1))};
  GDTB_TRY(launch_spmv(ctx->launch, A, d_source, d_range, nullptr, nullptr, 0));
  GDTB_CUDA(cudaStreamSynchronize(ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_matop_apply_inverse(gdtb_matop* op, const double* d_rhs, double* d_x, const gdtb_solver_opts* opts,
                             gdtb_solver_info* info)
{
  if (!op || !d_rhs || !d_x)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_apply_inverse: NULL argument");
  GDTB_TRY(internal_check_ctx(op->ctx));
  if (op->test.size != op->ansatz.size)
    return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "apply_inverse needs a square operator");
  CsrView A;
  GDTB_TRY(matop_csr(op, A));
  return solve_csr(op->ctx, A, d_rhs, d_x, opts, info);
}

int gdtb_matop_apply_inverse_host(gdtb_matop* op, const double* rhs, double* x, const gdtb_solver_opts* opts,
                                  gdtb_solver_info* info)
{
  if (!op || !rhs || !x)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_apply_inverse_host: NULL argument");
  GDTB_TRY(internal_check_ctx(op->ctx));
  const size_t bytes = sizeof(double) * (size_t)op->test.size;
  DeviceBuffer b, sol;
  GDTB_TRY(b.alloc(bytes));
  GDTB_TRY(sol.alloc(bytes));
  cudaStream_t stream = op->ctx->launch.stream;
  GDTB_CUDA(cudaMemcpyAsync(b.p, rhs, bytes, cudaMemcpyHostToDevice, stream));
  GDTB_CUDA(cudaMemcpyAsync(sol.p, x, bytes, cudaMemcpyHostToDevice, stream));
  const int st = gdtb_matop_apply_inverse(op, b.as<double>(), sol.as<double>(), opts, info);
  GDTB_CUDA(cudaMemcpyAsync(x, sol.p, bytes, cudaMemcpyDeviceToHost, stream));
  GDTB_CUDA(cudaStreamSynchronize(stream));
  return st;
}

int gdtb_csr_apply_inverse_device(gdtb_ctx* ctx, const gdtb_pattern* pattern, const double* d_values,
                                  const double* d_rhs, double* d_x, const gdtb_solver_opts* opts,
                                  gdtb_solver_info* info)
{
  if (!pattern || !d_values || !d_rhs || !d_x)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_csr_apply_inverse_device: NULL argument");
  GDTB_TRY(internal_check_ctx(ctx));
  if (pattern->rows != pattern->cols)
    return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "apply_inverse needs a square matrix");
  CsrView A{pattern->rows, pattern->d_rowptr, pattern->d_colidx, d_values,
            double(pattern->nnz) / double(std::max<long long>(pattern->rows, 1))};
  return solve_csr(ctx, A, d_rhs, d_x, opts, info);
}

int gdtb_bilinear_form_apply2(gdtb_ctx* ctx, const gdtb_space* space, const double* d_dofs, const gdtb_function* f,
                              const gdtb_form* form, double* result)
{
  if (!space || !form || !result || (!d_dofs && !f))
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_bilinear_form_apply2: NULL argument");
  GDTB_TRY(internal_check_ctx(ctx));
  if (form->n_terms < 1 || form->n_terms > GDTB_MAX_TERMS)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "form: n_terms must be in [1, GDTB_MAX_TERMS]");
  if (f) {
    GDTB_TRY(internal_validate_function(*f, "apply2: f"));
    if (f->kind > GDTB_FN_BUILTIN && f->kind != GDTB_FN_QP_VALUE_GRAD)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "apply2: f must be a constant, per-element, analytic or value-and-gradient "
                                            "sampled function (a discrete function enters through the DoF vector)");
    if (f->kind == GDTB_FN_CONST_TENSOR || f->kind == GDTB_FN_ELEM_TENSOR)
      return fail(GDTB_ERR_INVALID_ARGUMENT, "apply2: f must be scalar");
  }
  const SpaceDev& sp = space->dev;
  const GridDev& g = space->grid;
  auto P = std::make_unique<Apply2Params>();
  std::memset(P.get(), 0, sizeof(Apply2Params));
  P->g = g;
  P->sp = sp;
  P->has_u = d_dofs ? 1 : 0;
  P->has_f = f ? 1 : 0;
  LoweredForm owner;
  gdtb_form fm = *form;
  gdtb_function ff;
  int status = GDTB_OK;
  if (f) {
    ff = *f;
    status = internal_lower_function(ctx, g, ff, owner);
    P->f = internal_to_dev(ff, &owner);
  }
  // order of the one-function basis: the discrete function's order, or the declared order of f if that is higher
  const int e_order = std::max(d_dofs ? sp.K : 0, f ? f->order : 0);
  int order = 0;
  FormDev& F = P->form;
  F.n_terms = fm.n_terms;
  F.scaling = fm.scaling;
  for (int t = 0; t < fm.n_terms && status == GDTB_OK; ++t) {
    gdtb_integrand& in = fm.terms[t];
    if (in.kind != GDTB_INT_LAPLACE && in.kind != GDTB_INT_PRODUCT) {
      status = fail(GDTB_ERR_INTEGRAND, "apply2: only element integrands (Laplace, product) are supported");
      break;
    }
    status = internal_validate_function(in.diffusion, "apply2: integrand.diffusion");
    if (status == GDTB_OK)
      status = internal_lower_function(ctx, g, in.diffusion, owner);
    order = std::max(order, in.diffusion.order + e_order + e_order); // laplace.hh:74-79, product.hh:89-100
    F.terms[t].kind = in.kind;
    F.terms[t].prefactor = in.prefactor;
    F.terms[t].diffusion = internal_to_dev(in.diffusion, &owner);
  }
  if (status == GDTB_OK) {
    F.m = gauss_points_for_order(order + fm.over_integrate);
    if (F.m > MAX_Q1D)
      status = fail(GDTB_ERR_NOT_IMPLEMENTED, "quadrature order too high (more than 8 Gauss points per direction)");
    else if (f && f->kind == GDTB_FN_QP_VALUE_GRAD && f->qp_per_element != ipow(F.m, g.d))
      status = fail(GDTB_ERR_SHAPES_DO_NOT_MATCH,
                    "apply2: f was not sampled for the rule of this form (gdtb_bilinear_form_quadrature_order)");
  }
  DeviceBuffer dP, partial, res;
  const int nb = (int)std::min<long long>(std::max<long long>((g.ne + RED_BLOCK - 1) / RED_BLOCK, 1),
                                          std::min(MAX_PARTIALS, 8 * ctx->launch.sm_count));
  if (status == GDTB_OK) {
    gauss_legendre_01(F.m, F.qx, F.qw);
    for (int q = 0; q < F.m; ++q)
      lagrange_1d(sp.K, F.qx[q], F.phi[q], F.dphi[q]);
    status = dP.alloc(sizeof(Apply2Params));
    if (status == GDTB_OK)
      status = partial.alloc(sizeof(double) * (size_t)nb);
    if (status == GDTB_OK)
      status = res.alloc(sizeof(double));
  }
  if (status == GDTB_OK) {
    cudaStream_t st = ctx->launch.stream;
    cudaMemcpyAsync(dP.p, P.get(), sizeof(Apply2Params), cudaMemcpyHostToDevice, st);
    k_bilinear_apply2<<<nb, RED_BLOCK, 0, st>>>(dP.as<Apply2Params>(), d_dofs, partial.as<double>());
    k_finish_sum<<<1, RED_BLOCK, 0, st>>>(partial.as<double>(), nb, res.as<double>());
    ctx->launch.count += 2;
    cudaMemcpyAsync(result, res.p, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaError_t err = cudaStreamSynchronize(st);
    if (err != cudaSuccess)
      status = fail(GDTB_ERR_CUDA, std::string("apply2: ") + cudaGetErrorString(err));
  }
  free_form(owner);
  return status;
}

int gdtb_bilinear_form_quadrature_order(const gdtb_space* space, int has_dofs, int f_order, const gdtb_form* form,
                                        int32_t* order)
{
  if (!space || !form || !order)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_bilinear_form_quadrature_order: NULL argument");
  if (form->n_terms < 1 || form->n_terms > GDTB_MAX_TERMS)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "form: n_terms must be in [1, GDTB_MAX_TERMS]");
  const int e_order = std::max(has_dofs ? space->dev.K : 0, std::max(f_order, 0));
  int o = 0;
  for (int t = 0; t < form->n_terms; ++t)
    o = std::max(o, form->terms[t].diffusion.order + e_order + e_order);
  *order = o + form->over_integrate;
  return GDTB_OK;
}

int gdtb_bilinear_form_apply2_host(gdtb_ctx* ctx, const gdtb_space* space, const double* dofs, const gdtb_function* f,
                                   const gdtb_form* form, double* result)
{
  if (!space)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_bilinear_form_apply2_host: NULL argument");
  GDTB_TRY(internal_check_ctx(ctx));
  DeviceBuffer d;
  if (dofs) {
    GDTB_TRY(d.alloc(sizeof(double) * (size_t)space->dev.size));
    GDTB_CUDA(cudaMemcpyAsync(d.p, dofs, sizeof(double) * (size_t)space->dev.size, cudaMemcpyHostToDevice,
                              ctx->launch.stream));
  }
  return gdtb_bilinear_form_apply2(ctx, space, dofs ? d.as<double>() : nullptr, f, form, result);
}

int gdtb_lagrange_interpolate(gdtb_ctx* ctx, const gdtb_space* space, const gdtb_function* f, double* d_dofs)
{
  if (!space || !f || !d_dofs)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_lagrange_interpolate: NULL argument");
  GDTB_TRY(internal_check_ctx(ctx));
  if (space->dev.kind == GDTB_SPACE_FV)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_lagrange_interpolate: use gdtb_fv_interpolate for finite-volume spaces");
  GDTB_TRY(internal_validate_function(*f, "interpolate: f"));
  if (f->kind > GDTB_FN_BUILTIN)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "interpolate: f must be a constant, per-element or analytic function");
  if (f->kind == GDTB_FN_CONST_TENSOR || f->kind == GDTB_FN_ELEM_TENSOR)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "interpolate: f must be scalar");
  LoweredForm owner;
  gdtb_function ff = *f;
  // (no early return below: the device clones held by `owner` are released on every path)
  int status = internal_lower_function(ctx, space->grid, ff, owner);
  const long long total = space->grid.ne * space->dev.nloc;
  if (status == GDTB_OK && total > 0) {
    k_lagrange_interpolate<<<(unsigned)((total + 255) / 256), 256, 0, ctx->launch.stream>>>(space->grid, space->dev,
                                                                                            internal_to_dev(ff), d_dofs);
    ctx->launch.count++;
    cudaError_t err = cudaStreamSynchronize(ctx->launch.stream);
    if (err != cudaSuccess)
      status = fail(GDTB_ERR_CUDA, std::string("interpolate: ") + cudaGetErrorString(err));
  }
  free_form(owner);
  return status;
}

int gdtb_lagrange_interpolate_host(gdtb_ctx* ctx, const gdtb_space* space, const gdtb_function* f, double* dofs)
{
  if (!space || !dofs)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_lagrange_interpolate_host: NULL argument");
  GDTB_TRY(internal_check_ctx(ctx));
  DeviceBuffer d;
  GDTB_TRY(d.alloc(sizeof(double) * (size_t)space->dev.size));
  GDTB_TRY(gdtb_lagrange_interpolate(ctx, space, f, d.as<double>()));
  GDTB_CUDA(cudaMemcpy(dofs, d.p, sizeof(double) * (size_t)space->dev.size, cudaMemcpyDeviceToHost));
  return GDTB_OK;
}

} // extern "C"

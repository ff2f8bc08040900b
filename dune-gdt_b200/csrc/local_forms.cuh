// dune-gdt_b200/csrc/local_forms.cuh -- device-side restatement of dune-gdt's local forms, one local-matrix ROW at a time.
//
// Shared by the coloured-scatter kernels (assemble_generic.cu) and the DG row-gather kernel (assemble_dg_gather.cu):
//   element_row  : LocalElementIntegralBilinearForm::apply2 (local/bilinear-forms/integrals.hh:97-134)
//                  + LocalLaplaceIntegrand / LocalElementProductIntegrand / sums
//                    (local/integrands/laplace.hh:81-102, product.hh:104-130, combined.hh:212-227)
//   coupling_row : LocalCouplingIntersectionIntegralBilinearForm::apply2 (integrals.hh:197-267)
//                  + InnerCoupling (laplace-ipdg.hh:107-186) / InnerPenalty (ipdg.hh:113-171) / sums (combined.hh:381-431)
//   boundary_row : LocalIntersectionIntegralBilinearForm::apply2 (integrals.hh:338-369)
//                  + DirichletCoupling (laplace-ipdg.hh:340-368) / BoundaryPenalty (ipdg.hh:254-282)
// Geometry (J^{-T} = diag(1/ext), integration element) and coefficients are evaluated per quadrature point in FP64
// from the cell's own corners, in the reference's operation order.
#pragma once

#include "common.cuh"

namespace gdtb {

template <int D, int K>
struct Loc
{
  static constexpr int N1 = K + 1;
  static constexpr int N = D == 1 ? N1 : (D == 2 ? N1 * N1 : N1 * N1 * N1);
  __host__ __device__ static constexpr int a(int i, int k)
  {
    return k == 0 ? i % N1 : (k == 1 ? (i / N1) % N1 : i / (N1 * N1));
  }
};

struct Tables
{
  double phi[MAX_Q1D][MAX_K + 1];
  double dphi[MAX_Q1D][MAX_K + 1];
  double phi_end[2][MAX_K + 1];
  double dphi_end[2][MAX_K + 1];
};

__device__ inline void load_tables(Tables& s, const FormDev& f)
{
  for (int t = threadIdx.x; t < MAX_Q1D * (MAX_K + 1); t += blockDim.x) {
    (&s.phi[0][0])[t] = (&f.phi[0][0])[t];
    (&s.dphi[0][0])[t] = (&f.dphi[0][0])[t];
  }
  for (int t = threadIdx.x; t < 2 * (MAX_K + 1); t += blockDim.x) {
    (&s.phi_end[0][0])[t] = (&f.phi_end[0][0])[t];
    (&s.dphi_end[0][0])[t] = (&f.dphi_end[0][0])[t];
  }
  __syncthreads();
}

// value and physical gradient of local basis function with tensor index (a0,a1,a2) from the 1D rows
// pv[k] / pd[k] (already selected for the current point); gradient = J^{-T} * reference gradient
// (spaces/basis/default.hh:167-174), J^{-T} = diag(1/ext).
template <int D>
__device__ inline void basis_at(const double* const* pv, const double* const* pd, const double* inv_ext, int a0,
                                int a1, int a2, double& val, double* grad)
{
  const double v0 = pv[0][a0];
  const double v1 = D > 1 ? pv[1][a1] : 1.;
  const double v2 = D > 2 ? pv[2][a2] : 1.;
  val = v0 * v1 * v2;
  grad[0] = inv_ext[0] * (pd[0][a0] * v1 * v2);
  grad[1] = D > 1 ? inv_ext[1] * (v0 * pd[1][a1] * v2) : 0.;
  grad[2] = D > 2 ? inv_ext[2] * (v0 * v1 * pd[2][a2]) : 0.;
}

template <int D>
__device__ inline double dotD(const double* a, const double* b)
{
  double s = a[0] * b[0];
  if (D > 1)
    s += a[1] * b[1];
  if (D > 2)
    s += a[2] * b[2];
  return s;
}

template <int D>
__device__ inline void matvecD(const double* T, const double* g, double* y)
{
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double s = 0.;
#pragma unroll
    for (int c = 0; c < D; ++c)
      s += T[r * 3 + c] * g[c];
    y[r] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// intersections
// ------------------------------------------------------------------------------------------------
struct FaceGeo
{
  double normal[3];
  double ie;
  double diameter;
};

template <int D>
__device__ inline FaceGeo make_face(const double* ext_in, int k, int s)
{
  FaceGeo f;
  f.normal[0] = f.normal[1] = f.normal[2] = 0.;
  f.normal[k] = s ? 1. : -1.;
  f.ie = 1.;
  double d2 = 0.;
#pragma unroll
  for (int j = 0; j < D; ++j)
    if (j != k) {
      f.ie *= ext_in[j];
      d2 += ext_in[j] * ext_in[j];
    }
  f.diameter = sqrt(d2);
  return f;
}

// default_intersection_diameter (ipdg.hh:27-38) or |I|
template <int D>
__device__ inline double intersection_h(const IntegrandDev& t, const FaceGeo& f, const double* ext_in,
                                        const double* ext_out, bool neighbor)
{
  if (t.hI_kind == GDTB_HI_VOLUME)
    return f.ie;
  if (D == 1)
    return neighbor ? 0.5 * (ext_in[0] + ext_out[0]) : ext_in[0];
  return f.diameter;
}

// select the 1D table rows for a face quadrature point: direction k is pinned to the end `side`,
// the other directions take the face rule's points in ascending axis order
template <int D>
__device__ inline void face_rows(const Tables& tab, const FormDev& f, int k, int side, int q1, int q2,
                                 const double** pv, const double** pd, double* xh, double& wq)
{
  int j = 0;
  wq = 1.;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    if (r == k) {
      pv[r] = tab.phi_end[side];
      pd[r] = tab.dphi_end[side];
      xh[r] = side;
    } else if (r < D) {
      const int q = j == 0 ? q1 : q2;
      ++j;
      pv[r] = tab.phi[q];
      pd[r] = tab.dphi[q];
      xh[r] = f.qx[q];
      wq *= f.qw[q];
    } else {
      pv[r] = tab.phi[0];
      pd[r] = tab.dphi[0];
      xh[r] = 0.;
    }
  }
}


// row i of the local element matrix of all summands of `f` on element idx, ADDED to acc[0 .. N)
template <int D, int K>
__device__ inline void element_row(const GridDev& g, const FormDev& f, const Tables& tab, const long long* idx, int i,
                                   double* __restrict__ acc)
{
  using L = Loc<D, K>;
  constexpr int N = L::N;
  const long long e = elem_index(g, idx);
  double lower[3], ext[3], inv_ext[3];
  cell_geometry(g, idx, lower, ext);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    inv_ext[k] = 1. / ext[k];
  const double ie = ext[0] * (D > 1 ? ext[1] : 1.) * (D > 2 ? ext[2] : 1.);
  const int ai0 = L::a(i, 0), ai1 = L::a(i, 1), ai2 = L::a(i, 2);
  const int m = f.m;
  const int my = D > 1 ? m : 1, mz = D > 2 ? m : 1;
  for (int qz = 0; qz < mz; ++qz)
    for (int qy = 0; qy < my; ++qy)
      for (int qx = 0; qx < m; ++qx) {
        const int q[3] = {qx, qy, qz};
        const double* pv[3];
        const double* pd[3];
        double x[3];
        double w = 1.;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          pv[k] = tab.phi[k < D ? q[k] : 0];
          pd[k] = tab.dphi[k < D ? q[k] : 0];
          x[k] = k < D ? lower[k] + f.qx[q[k]] * ext[k] : 0.;
          if (k < D)
            w *= f.qw[q[k]];
        }
        const double factor = ie * w; // integrals.hh:119
        double vi, gi[3];
        basis_at<D>(pv, pd, inv_ext, ai0, ai1, ai2, vi, gi);
        // coefficients of the summands at this point
        double kap[GDTB_MAX_TERMS][9];
        double wgt[GDTB_MAX_TERMS];
        const double xh[3] = {f.qx[qx], D > 1 ? f.qx[qy] : 0., D > 2 ? f.qx[qz] : 0.};
        const EvalPt pt = {qx + m * (qy + my * qz), idx, xh};
        for (int tt = 0; tt < f.n_terms; ++tt) {
          if (f.terms[tt].kind == GDTB_INT_LAPLACE)
            fn_tensor(f.terms[tt].diffusion, g, e, x, pt, kap[tt]);
          else
            wgt[tt] = fn_scalar(f.terms[tt].diffusion, g, e, x, pt);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double vj, gj[3];
          basis_at<D>(pv, pd, inv_ext, L::a(j, 0), L::a(j, 1), L::a(j, 2), vj, gj);
          double v = 0.;
          for (int tt = 0; tt < f.n_terms; ++tt) {
            if (f.terms[tt].kind == GDTB_INT_LAPLACE) {
              double kg[3];
              matvecD<D>(kap[tt], gj, kg); // (weight * ansatz_grad) . test_grad, laplace.hh:101
              v += dotD<D>(kg, gi);
            } else
              v += (wgt[tt] * vi) * vj; // product.hh:128
          }
          acc[j] += v * factor; // integrals.hh:131
        }
      }
}

// One row of the four coupling blocks of the inner (or periodic wrap) face between idx_in (the inside element, the
// one with the smaller index) and idx_out: row i of the inside element (row_inside: blocks in_in -> acc_a, in_out ->
// acc_b) or of the outside element (blocks out_in -> acc_a, out_out -> acc_b); acc_a are the columns of the inside
// element, acc_b those of the outside element.  The face is the (k, s) intersection of the inside element
// (s = 1: upper face, regular inner face; s = 0: lower face, periodic wrap).  Results are ADDED.
template <int D, int K>
__device__ inline void coupling_row(const GridDev& g, const FormDev& f, const Tables& tab, const long long* idx_in,
                                    const long long* idx_out, int k, int s, bool row_inside, int i,
                                    double* __restrict__ acc_a, double* __restrict__ acc_b)
{
  using L = Loc<D, K>;
  constexpr int N = L::N;
  const long long e_in = elem_index(g, idx_in), e_out = elem_index(g, idx_out);
  double lo_in[3], ext_in[3], lo_out[3], ext_out[3], inv_in[3], inv_out[3];
  cell_geometry(g, idx_in, lo_in, ext_in);
  cell_geometry(g, idx_out, lo_out, ext_out);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    inv_in[r] = 1. / ext_in[r];
    inv_out[r] = 1. / ext_out[r];
  }
  const FaceGeo face = make_face<D>(ext_in, k, s);
  const int m1 = D > 1 ? f.m : 1, m2 = D > 2 ? f.m : 1;
  for (int q2 = 0; q2 < m2; ++q2)
    for (int q1 = 0; q1 < m1; ++q1) {
      const double *pvi[3], *pdi[3], *pvo[3], *pdo[3];
      double xh_in[3], xh_out[3], x_in[3], x_out[3], wq, wq2;
      face_rows<D>(tab, f, k, s, q1, q2, pvi, pdi, xh_in, wq);
      face_rows<D>(tab, f, k, 1 - s, q1, q2, pvo, pdo, xh_out, wq2);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        x_in[r] = lo_in[r] + xh_in[r] * ext_in[r];
        x_out[r] = lo_out[r] + xh_out[r] * ext_out[r];
      }
      // this thread's test function (value + gradient on its own side)
      double vi_in, gi_in[3], vi_out, gi_out[3];
      basis_at<D>(pvi, pdi, inv_in, L::a(i, 0), L::a(i, 1), L::a(i, 2), vi_in, gi_in);
      basis_at<D>(pvo, pdo, inv_out, L::a(i, 0), L::a(i, 1), L::a(i, 2), vi_out, gi_out);
      double va[N], vb[N]; // summed integrand values of this point
#pragma unroll
      for (int j = 0; j < N; ++j)
        va[j] = vb[j] = 0.;
      for (int tt = 0; tt < f.n_terms; ++tt) {
        const IntegrandDev& in = f.terms[tt];
        double w_in[9], w_out[9], wn[3];
        const EvalPt p_in = {-1, idx_in, xh_in}, p_out = {-1, idx_out, xh_out};
        fn_tensor(in.weight, g, e_in, x_in, p_in, w_in);
        fn_tensor(in.weight, g, e_out, x_out, p_out, w_out);
        matvecD<D>(w_out, face.normal, wn);
        const double delta_plus = dotD<D>(face.normal, wn);
        matvecD<D>(w_in, face.normal, wn);
        const double delta_minus = dotD<D>(face.normal, wn);
        if (in.kind == GDTB_INT_IPDG_INNER_COUPLING) {
          double k_in[9], k_out[9], kg[3];
          fn_tensor(in.diffusion, g, e_in, x_in, p_in, k_in);
          fn_tensor(in.diffusion, g, e_out, x_out, p_out, k_out);
          const double weight_minus = delta_plus / (delta_plus + delta_minus);
          const double weight_plus = delta_minus / (delta_plus + delta_minus);
          const double sp_ = in.prefactor;
          matvecD<D>(k_in, gi_in, kg);
          const double fi_in = dotD<D>(kg, face.normal);
          matvecD<D>(k_out, gi_out, kg);
          const double fi_out = dotD<D>(kg, face.normal);
#pragma unroll
          for (int j = 0; j < N; ++j) {
            double vj_in, gj_in[3], vj_out, gj_out[3];
            basis_at<D>(pvi, pdi, inv_in, L::a(j, 0), L::a(j, 1), L::a(j, 2), vj_in, gj_in);
            basis_at<D>(pvo, pdo, inv_out, L::a(j, 0), L::a(j, 1), L::a(j, 2), vj_out, gj_out);
            matvecD<D>(k_in, gj_in, kg);
            const double fj_in = dotD<D>(kg, face.normal);
            matvecD<D>(k_out, gj_out, kg);
            const double fj_out = dotD<D>(kg, face.normal);
            if (row_inside) { // laplace-ipdg.hh:158-170
              va[j] += -1.0 * weight_minus * fj_in * vi_in;
              va[j] += -1.0 * sp_ * weight_minus * vj_in * fi_in;
              vb[j] += -1.0 * weight_plus * fj_out * vi_in;
              vb[j] += sp_ * weight_minus * vj_out * fi_in;
            } else { // laplace-ipdg.hh:172-185
              va[j] += weight_minus * fj_in * vi_out;
              va[j] += -1.0 * sp_ * weight_plus * vj_in * fi_out;
              vb[j] += weight_plus * fj_out * vi_out;
              vb[j] += sp_ * weight_plus * vj_out * fi_out;
            }
          }
        } else { // GDTB_INT_IPDG_INNER_PENALTY, ipdg.hh:149-170
          const double weight = (delta_plus * delta_minus) / (delta_plus + delta_minus);
          const double h = intersection_h<D>(in, face, ext_in, ext_out, true);
          const double penalty = (in.prefactor * weight) / h;
#pragma unroll
          for (int j = 0; j < N; ++j) {
            double vj_in, gj[3], vj_out;
            basis_at<D>(pvi, pdi, inv_in, L::a(j, 0), L::a(j, 1), L::a(j, 2), vj_in, gj);
            basis_at<D>(pvo, pdo, inv_out, L::a(j, 0), L::a(j, 1), L::a(j, 2), vj_out, gj);
            if (row_inside) {
              va[j] += penalty * vj_in * vi_in;
              vb[j] += -1.0 * penalty * vj_out * vi_in;
            } else {
              va[j] += -1.0 * penalty * vj_in * vi_out;
              vb[j] += penalty * vj_out * vi_out;
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < N; ++j) { // integrals.hh:254-265
        acc_a[j] += va[j] * face.ie * wq;
        acc_b[j] += vb[j] * face.ie * wq;
      }
    }
}

// row i of the boundary-face matrix of face (k, s) of element idx, ADDED to acc[0 .. N)
template <int D, int K>
__device__ inline void boundary_row(const GridDev& g, const FormDev& f, const Tables& tab, const long long* idx, int k,
                                    int s, int i, double* __restrict__ acc)
{
  using L = Loc<D, K>;
  constexpr int N = L::N;
  const long long e = elem_index(g, idx);
  double lo[3], ext[3], inv_ext[3];
  cell_geometry(g, idx, lo, ext);
#pragma unroll
  for (int r = 0; r < 3; ++r)
    inv_ext[r] = 1. / ext[r];
  const FaceGeo face = make_face<D>(ext, k, s);
  const int m1 = D > 1 ? f.m : 1, m2 = D > 2 ? f.m : 1;
  for (int q2 = 0; q2 < m2; ++q2)
    for (int q1 = 0; q1 < m1; ++q1) {
      const double *pv[3], *pd[3];
      double xh[3], x[3], wq;
      face_rows<D>(tab, f, k, s, q1, q2, pv, pd, xh, wq);
#pragma unroll
      for (int r = 0; r < 3; ++r)
        x[r] = lo[r] + xh[r] * ext[r];
      double vi, gi[3];
      basis_at<D>(pv, pd, inv_ext, L::a(i, 0), L::a(i, 1), L::a(i, 2), vi, gi);
      double va[N];
#pragma unroll
      for (int j = 0; j < N; ++j)
        va[j] = 0.;
      for (int tt = 0; tt < f.n_terms; ++tt) {
        const IntegrandDev& in = f.terms[tt];
        const EvalPt pt = {-1, idx, xh};
        if (in.kind == GDTB_INT_IPDG_DIRICHLET_COUPLING) { // laplace-ipdg.hh:362-367
          double kap[9], kg[3];
          fn_tensor(in.diffusion, g, e, x, pt, kap);
          matvecD<D>(kap, gi, kg);
          const double fi = dotD<D>(kg, face.normal);
#pragma unroll
          for (int j = 0; j < N; ++j) {
            double vj, gj[3];
            basis_at<D>(pv, pd, inv_ext, L::a(j, 0), L::a(j, 1), L::a(j, 2), vj, gj);
            matvecD<D>(kap, gj, kg);
            const double fj = dotD<D>(kg, face.normal);
            va[j] += -1.0 * fj * vi;
            va[j] += -1.0 * in.prefactor * vj * fi;
          }
        } else { // GDTB_INT_IPDG_BOUNDARY_PENALTY, ipdg.hh:276-281
          double w[9], wn[3];
          fn_tensor(in.weight, g, e, x, pt, w);
          matvecD<D>(w, face.normal, wn);
          const double h = intersection_h<D>(in, face, ext, ext, false);
          const double penalty = (in.prefactor * dotD<D>(face.normal, wn)) / h;
#pragma unroll
          for (int j = 0; j < N; ++j) {
            double vj, gj[3];
            basis_at<D>(pv, pd, inv_ext, L::a(j, 0), L::a(j, 1), L::a(j, 2), vj, gj);
            va[j] += penalty * vj * vi;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < N; ++j)
        acc[j] += va[j] * face.ie * wq; // integrals.hh:366-368
    }
}

} // namespace gdtb

// dune-gdt_b200/csrc/assemble_dg_fast.cu -- factorised DG row-gather kernels (order 1, constant / element-wise scalar
// coefficients): see assemble_dg_gather.cu for the work decomposition (one thread per DG row, rows of a work item =
// one contiguous CSR segment staged in shared memory and written by a TMA bulk store).
#include <cstdlib>

#include "dg_gather.cuh"

namespace gdtb {

namespace {

// ---- factorised path: DG order 1, constant / element-wise scalar coefficients ------------------------------------
// On an affine axis-aligned cell every integrand of the path is a product over the axes, and the Gauss rules are
// tensor rules, so the quadrature sums of the reference factorise EXACTLY into 1D reference tables of the form's own
// rule (m points): M1[a][b] = sum_q w_q phi_a phi_b, K1[a][b] = sum_q w_q phi_a' phi_b', and the end values
// phi_a(s), phi_a'(s) for the pinned axis of a face.  With per-axis cell data h_o, 1/h_o:
//   element Laplace   kappa sum_r (K1[i_r][j_r] / h_r) prod_{o != r} h_o M1[i_o][j_o]
//   element product   w prod_o h_o M1[i_o][j_o]
//   face normal to k  C[i_k][j_k] prod_{o != k} h_o M1[i_o][j_o],  C = the 2 x 2 matrix of the integrand's terms in
//                     the end values (laplace-ipdg.hh:149-185, 362-367; ipdg.hh:149-170, 276-281), coefficients per
//                     adjacent element.
// Same roles and signs as coupling_row / boundary_row of local_forms.cuh; results agree with the quadrature loops to
// rounding (the parity tests compare both paths against the oracle).
struct DgFastTab
{
  double M1[2][2], K1[2][2], pe[2][2], de[2][2];
  // all coefficients constant (CC kernels): the 2 x 2 face matrices of the form's terms are affine in the cell data
  //   mult = {1 / h_k(inside), 1 / h_k(outside), 1 / |I|, 1 / diam(I)}
  // fa / fb[s][i_k][j_k][m]: columns of the inside / outside element for a row of the inside (s = 1) or outside
  // (s = 0) element; boundary forms: fa[s] for the face with outer normal -+e_k, mult[0] = 1 / h_k.  F.scaling included.
  double fa[2][2][2][4], fb[2][2][2][4];
  double elap, emass; // element forms: sum of scaling * kappa over the Laplace terms / scaling * w over the products
};

// the CC tables of one form (role: 0 element, 1 coupling, 2 boundary); one thread
__device__ inline void dg_fast_tables_cc(const FormDev& F, int role, DgFastTab& T)
{
  for (int s = 0; s < 2; ++s)
    for (int ik = 0; ik < 2; ++ik)
      for (int jk = 0; jk < 2; ++jk)
        for (int m = 0; m < 4; ++m)
          T.fa[s][ik][jk][m] = T.fb[s][ik][jk][m] = 0.;
  T.elap = T.emass = 0.;
  for (int tt = 0; tt < F.n_terms; ++tt) {
    const IntegrandDev& in = F.terms[tt];
    const double sc = F.scaling;
    if (role == 0) {
      if (in.kind == GDTB_INT_LAPLACE)
        T.elap += sc * in.diffusion.c[0];
      else
        T.emass += sc * in.diffusion.c[0];
      continue;
    }
    const int slot = in.hI_kind == GDTB_HI_VOLUME ? 2 : 3;
    for (int ik = 0; ik < 2; ++ik)
      for (int jk = 0; jk < 2; ++jk) {
        if (role == 1) {
          const double dp = in.weight.c[0], dm = in.weight.c[0]; // delta_plus, delta_minus
          if (in.kind == GDTB_INT_IPDG_INNER_COUPLING) {
            const double c = in.diffusion.c[0], sp_ = in.prefactor;
            const double wm = dp / (dp + dm), wp = dm / (dp + dm);
            // s = 1: row of the inside element (laplace-ipdg.hh:158-170)
            T.fa[1][ik][jk][0] += sc * (-1.0 * wm * c * (T.de[1][jk] * T.pe[1][ik] + sp_ * T.pe[1][jk] * T.de[1][ik]));
            T.fb[1][ik][jk][1] += sc * (-1.0 * wp * c * T.de[0][jk] * T.pe[1][ik]);
            T.fb[1][ik][jk][0] += sc * (sp_ * wm * c * T.pe[0][jk] * T.de[1][ik]);
            // s = 0: row of the outside element (laplace-ipdg.hh:172-185)
            T.fa[0][ik][jk][0] += sc * (wm * c * T.de[1][jk] * T.pe[0][ik]);
            T.fa[0][ik][jk][1] += sc * (-1.0 * sp_ * wp * c * T.pe[1][jk] * T.de[0][ik]);
            T.fb[0][ik][jk][1] += sc * (wp * c * (T.de[0][jk] * T.pe[0][ik] + sp_ * T.pe[0][jk] * T.de[0][ik]));
          } else { // inner penalty (ipdg.hh:149-170): sigma (delta+ delta- / (delta+ + delta-)) / h
            const double pw = sc * in.prefactor * ((dp * dm) / (dp + dm));
            T.fa[1][ik][jk][slot] += pw * T.pe[1][jk] * T.pe[1][ik];
            T.fb[1][ik][jk][slot] += -1.0 * pw * T.pe[0][jk] * T.pe[1][ik];
            T.fa[0][ik][jk][slot] += -1.0 * pw * T.pe[1][jk] * T.pe[0][ik];
            T.fb[0][ik][jk][slot] += pw * T.pe[0][jk] * T.pe[0][ik];
          }
        } else {
          for (int s = 0; s < 2; ++s) {
            const double sg = s ? 1. : -1.;
            if (in.kind == GDTB_INT_IPDG_DIRICHLET_COUPLING) // laplace-ipdg.hh:362-367
              T.fa[s][ik][jk][0] += sc * (-1.0 * in.diffusion.c[0] * sg
                                          * (T.de[s][jk] * T.pe[s][ik] + in.prefactor * T.pe[s][jk] * T.de[s][ik]));
            else // boundary penalty (ipdg.hh:276-281): sigma (n . omega n) / h
              T.fa[s][ik][jk][slot] += sc * in.prefactor * in.weight.c[0] * T.pe[s][jk] * T.pe[s][ik];
          }
        }
      }
  }
}

// intersection_h of local_forms.cuh for axis-aligned faces: |I|, or the face diameter (1D: element lengths)
template <int D>
__device__ __forceinline__ double dg_face_h(const int hI_kind, const double* h, int k, double h_in, double h_out,
                                            bool neighbor)
{
  double ie = 1., d2 = 0.;
#pragma unroll
  for (int o = 0; o < D; ++o)
    if (o != k) {
      ie *= h[o];
      d2 += h[o] * h[o];
    }
  if (hI_kind == GDTB_HI_VOLUME)
    return ie;
  if (D == 1)
    return neighbor ? 0.5 * (h_in + h_out) : h_in;
  return sqrt(d2);
}

// coefficient of the SWIPDG descriptor (constant bank) or of the general form array
__device__ __forceinline__ double dg_sw(const DgGatherParams::SwFn& f, long long e)
{
  return f.data ? __ldg(f.data + e) : f.c;
}

// 1 / intersection_h for the CC tables: 1 / |I| (volume) or 1 / diameter (1D: element lengths) from the cell data
template <int D>
__device__ __forceinline__ double dg_inv_face(const double* h, const double* hinv, int k, double h_in, double h_out,
                                              bool neighbor, bool volume)
{
  if (D == 1)
    return volume ? 1. : (neighbor ? 1. / (0.5 * (h_in + h_out)) : 1. / h_in);
  if (volume || D == 2) { // 2D: the face is an interval, diameter == |I|
    double inv = 1.;
#pragma unroll
    for (int o = 0; o < D; ++o)
      if (o != k)
        inv *= hinv[o];
    return inv;
  }
  double d2 = 0.;
#pragma unroll
  for (int o = 0; o < D; ++o)
    if (o != k)
      d2 += h[o] * h[o];
  return 1. / sqrt(d2);
}

// block[j] += sc * c2[j_k] * prod_{o != k} tM[o][j_o]  (j = j_0 + 2 j_1 + 4 j_2)
template <int D>
__device__ __forceinline__ void dg_add_face_block(double* __restrict__ block, const double sc, const double* c2, int k,
                                                  const double (*tM)[2])
{
  constexpr int N = 1 << D;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double v = sc * c2[(j >> k) & 1];
#pragma unroll
    for (int o = 0; o < D; ++o)
      if (o != k)
        v *= tM[o][(j >> o) & 1];
    block[j] += v;
  }
}

// one block of a row (N = 2^D doubles, N even) into the stage: 16-byte stores when the stage is 16-byte aligned (block
// offsets are even), which halves the shared-memory wavefronts of the strided row layout
template <int N>
__device__ __forceinline__ void dg_store_block(double* __restrict__ blk, const double (&v)[N], bool aligned16)
{
  if (aligned16) {
#pragma unroll
    for (int j = 0; j < N; j += 2)
      *reinterpret_cast<double2*>(blk + j) = make_double2(v[j], v[j + 1]);
    return;
  }
#pragma unroll
  for (int j = 0; j < N; ++j)
    blk[j] = v[j];
}

// ONE: exactly one element, one coupling and one boundary form (the SWIPDG operator of the reference's drivers): the
// form loops have compile-time trip counts
// PER: some direction of the grid is periodic (wrap neighbours: dg_block_positions; the coupling forms whose filter
// includes the periodic intersections also run over the wrap faces)
template <int D, bool ACCUMULATE, bool CC, bool ONE, bool PER = false>
__global__ void __launch_bounds__(DGG_THREADS)
    k_dg_gather_fast(const __grid_constant__ DgGatherParams p, double* __restrict__ values, int stage_doubles, int nbuf)
{
  constexpr int N = 1 << D;
  const int n_elem = ONE ? 1 : p.n_elem, n_coup = ONE ? 1 : p.n_coup, n_bnd = ONE ? 1 : p.n_bnd;
  // element-wise coefficients + the SWIPDG form structure (DgGatherParams::swip): term counts and kinds are compile time
  constexpr bool SW = ONE && !CC;
  extern __shared__ __align__(16) double smem[];
  __shared__ DgFastTab tabs[DGG_MAX_FORMS];
  const GridDev& g = p.g;
  const int n_forms = n_elem + n_coup + n_bnd;
  if ((int)threadIdx.x < n_forms) {
    const FormDev& f = p.forms[threadIdx.x];
    DgFastTab& t = tabs[threadIdx.x];
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        double sm = 0., sk = 0.;
        for (int q = 0; q < f.m; ++q) {
          sm += f.qw[q] * f.phi[q][a] * f.phi[q][b];
          sk += f.qw[q] * f.dphi[q][a] * f.dphi[q][b];
        }
        t.M1[a][b] = sm;
        t.K1[a][b] = sk;
        t.pe[a][b] = f.phi_end[a][b];
        t.de[a][b] = f.dphi_end[a][b];
      }
    if (CC)
      dg_fast_tables_cc(f, (int)threadIdx.x < n_elem ? 0 : ((int)threadIdx.x < n_elem + n_coup ? 1 : 2), t);
  }
  __syncthreads();
  const FormDev* f_elem = p.forms;
  const FormDev* f_coup = p.forms + n_elem;
  const FormDev* f_bnd = p.forms + n_elem + n_coup;
  const DgFastTab* t_elem = tabs;
  const DgFastTab* t_coup = tabs + n_elem;
  const DgFastTab* t_bnd = tabs + n_elem + n_coup;
  constexpr int EPI = DGG_THREADS / N; // elements per item
  // the element range [e_begin, e_end) this process owns (a slab of element layers; the whole grid otherwise)
  const long long nitems = (p.e_end - p.e_begin + EPI - 1) / EPI;
  int buf = 0;

  for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const long long e0 = p.e_begin + item * EPI;
    const int ne_item = (int)min((long long)EPI, p.e_end - e0);
    int idx0[3], idx1[3];
    dg_decode<D>(p, (unsigned)e0, idx0);
    const long long start = (long long)N * N * dg_blocks_before<D>(g, e0, idx0) - p.value_offset;
    long long end;
    if (e0 + ne_item < g.ne) {
      dg_decode<D>(p, (unsigned)(e0 + ne_item), idx1);
      end = (long long)N * N * dg_blocks_before<D>(g, e0 + ne_item, idx1) - p.value_offset;
    } else {
      dg_decode<D>(p, (unsigned)(g.ne - 1), idx1);
      end = (long long)N * N * (dg_blocks_before<D>(g, g.ne - 1, idx1) + dg_nblocks<D>(g, idx1)) - p.value_offset;
    }
    const int seg = int(end - start);
    const int phase = int((reinterpret_cast<unsigned long long>(values + start) >> 3) & 1ULL);
    double* stage = smem + buf * stage_doubles + phase;

    const int le = threadIdx.x / N, i = threadIdx.x & (N - 1);
    int idx[3] = {0, 0, 0};
    int row_off = 0;
    if (le < ne_item) {
      dg_decode<D>(p, (unsigned)(e0 + le), idx);
      row_off = int((long long)N * N * dg_blocks_before<D>(g, e0 + le, idx) - p.value_offset - start)
                + i * dg_nblocks<D>(g, idx) * N;
    }
    // wait for the bulk store that last used this stage only now, after the item's index arithmetic (see the Q2 kernel)
    if (!ACCUMULATE && item != (long long)blockIdx.x) {
      if (threadIdx.x == 0) {
        if (nbuf == 1)
          dg_bulk_wait_read0();
        else
          dg_bulk_wait_read1();
      }
      __syncthreads();
    }
    if (le < ne_item) {
      const long long e = e0 + le;
      double* row = stage + row_off;
      long long estride[3] = {1, g.n[0], g.n[0] * g.n[1]};
      bool has_lo[D], has_hi[D];
      double h[D], hinv[D];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        has_lo[k] = idx[k] > 0 || (PER && dg_periodic(g, k));
        has_hi[k] = idx[k] < g.n[k] - 1 || (PER && dg_periodic(g, k));
        // per-axis tables of the grid (k_q1_axis_tables: entry i + 1 = h_k of cell i, 1 / h_k behind it)
        h[k] = __ldg(p.axis_tab[k] + idx[k] + 1);
        hinv[k] = __ldg(p.axis_tab[k] + p.axis_tab_inv + idx[k] + 1);
      }
      double self[N];
#pragma unroll
      for (int j = 0; j < N; ++j)
        self[j] = 0.;

      // ---- element forms --------------------------------------------------------------------------------------
      for (int f = 0; f < n_elem; ++f) {
        const FormDev& F = f_elem[f];
        const DgFastTab& T = t_elem[f];
        double tM[D][2], tK[D][2];
#pragma unroll
        for (int o = 0; o < D; ++o) {
          const int io = (i >> o) & 1;
          tM[o][0] = h[o] * T.M1[io][0];
          tM[o][1] = h[o] * T.M1[io][1];
          tK[o][0] = hinv[o] * T.K1[io][0];
          tK[o][1] = hinv[o] * T.K1[io][1];
        }
        // CC: the terms of a form share its tables, their constant coefficients are summed up front (two passes)
        for (int tt = 0; tt < (CC ? 2 : (SW ? 1 : F.n_terms)); ++tt) {
          const double c = CC ? (tt == 0 ? T.elap : T.emass)
                              : (SW ? p.sw.s_elem * dg_sw(p.sw.elem_kappa, e) : F.scaling * dg_coef(F.terms[tt].diffusion, e));
          if (CC && c == 0.)
            continue;
          if (CC ? tt == 0 : (SW || F.terms[tt].kind == GDTB_INT_LAPLACE)) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
              double sum = 0.;
#pragma unroll
              for (int r = 0; r < D; ++r) {
                double v = tK[r][(j >> r) & 1];
#pragma unroll
                for (int o = 0; o < D; ++o)
                  if (o != r)
                    v *= tM[o][(j >> o) & 1];
                sum += v;
              }
              self[j] = fma(c, sum, self[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < N; ++j) {
              double v = c;
#pragma unroll
              for (int o = 0; o < D; ++o)
                v *= tM[o][(j >> o) & 1];
              self[j] += v;
            }
          }
        }
      }

      // ---- faces ------------------------------------------------------------------------------------------------
      // position of the neighbour blocks in the row: z-, y-, x-, self, x+, y+, z+ (existing ones only)
      int pos_lo[D], pos_hi[D], pos_self;
      if (PER) {
        dg_block_positions<D>(g, idx, pos_lo, pos_hi, pos_self);
        pos_self *= N;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          pos_lo[k] *= N;
          pos_hi[k] *= N;
        }
      } else {
        int pos = 0;
#pragma unroll
        for (int k = D - 1; k >= 0; --k) {
          pos_lo[k] = pos;
          pos += has_lo[k] ? N : 0;
        }
        pos_self = pos;
        pos += N;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          pos_hi[k] = pos;
          pos += has_hi[k] ? N : 0;
        }
      }

#pragma unroll
      for (int k = 0; k < D; ++k) {
        const int ik = (i >> k) & 1;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const bool has = s ? has_hi[k] : has_lo[k];
          if (has) {
            // inner face: s == 0: this element is the outside one (inside = e - stride), s == 1: it is the inside one
            // A periodic wrap face is treated with the same geometric roles (inside = the cell below the face = the last
            // cell of the line, outside = the first): the IPDG forms are invariant under swapping inside / outside
            // together with the normal, so this equals the reference's walk (inside = smaller index) up to rounding.
            const bool wrap = PER && (s ? idx[k] == g.n[k] - 1 : idx[k] == 0);
            const int nb_k = wrap ? (s ? 0 : (int)g.n[k] - 1) : idx[k] + (s ? 1 : -1); // the neighbour cell along k
            const long long e_nb = e + (long long)(nb_k - idx[k]) * estride[k];
            const long long e_in = s ? e : e_nb, e_out = s ? e_nb : e;
            const double* nb_tab = p.axis_tab[k] + nb_k + 1;
            const double h_nb = __ldg(nb_tab), hinv_nb = __ldg(nb_tab + p.axis_tab_inv);
            const double h_in = s ? h[k] : h_nb, h_out = s ? h_nb : h[k];
            const double hinv_in = s ? hinv[k] : hinv_nb, hinv_out = s ? hinv_nb : hinv[k];
            double nbb[N];
#pragma unroll
            for (int j = 0; j < N; ++j)
              nbb[j] = 0.;
            for (int f = 0; f < n_coup; ++f) {
              if (PER && wrap && !((p.coup_on_periodic >> f) & 1u))
                continue; // ApplyOn::InnerIntersectionsOnce skips the periodic intersections
              const FormDev& F = f_coup[f];
              const DgFastTab& T = t_coup[f];
              double tM[D][2];
#pragma unroll
              for (int o = 0; o < D; ++o) {
                const int io = (i >> o) & 1;
                tM[o][0] = h[o] * T.M1[io][0];
                tM[o][1] = h[o] * T.M1[io][1];
              }
              if (CC) {
                const double mult[4] = {hinv_in, hinv_out, dg_inv_face<D>(h, hinv, k, h_in, h_out, true, true),
                                        dg_inv_face<D>(h, hinv, k, h_in, h_out, true, false)};
                double ca[2], cb[2];
#pragma unroll
                for (int jk = 0; jk < 2; ++jk) {
                  const double* A = T.fa[s][ik][jk];
                  const double* B = T.fb[s][ik][jk];
                  ca[jk] = fma(A[0], mult[0], fma(A[1], mult[1], fma(A[2], mult[2], A[3] * mult[3])));
                  cb[jk] = fma(B[0], mult[0], fma(B[1], mult[1], fma(B[2], mult[2], B[3] * mult[3])));
                }
                dg_add_face_block<D>(self, 1., s ? ca : cb, k, tM);
                dg_add_face_block<D>(nbb, 1., s ? cb : ca, k, tM);
                continue;
              }
              // test function on its own side: inside element -> upper end (1), outside element -> lower end (0)
              const double vi = s ? T.pe[1][ik] : T.pe[0][ik];
              const double gi = s ? T.de[1][ik] * hinv_in : T.de[0][ik] * hinv_out;
              double ca[2] = {0., 0.}, cb[2] = {0., 0.}; // columns of the inside / outside element
              // SWIPDG with kappa = omega from one array (the usual case): one load per cell of the face
              double sw_in = 0., sw_out = 0.;
              if (SW && p.sw.coup_same) {
                sw_in = dg_sw(p.sw.coup_kappa, e_in);
                sw_out = dg_sw(p.sw.coup_kappa, e_out);
              }
#pragma unroll
              for (int tt = 0; tt < (SW ? 2 : F.n_terms); ++tt) {
                const IntegrandDev& in = F.terms[tt];
                const DgGatherParams::SwFn& wfn = tt == 0 ? p.sw.coup_weight : p.sw.pen_weight;
                const double delta_plus = SW ? (p.sw.coup_same ? sw_out : dg_sw(wfn, e_out)) : dg_coef(in.weight, e_out);
                const double delta_minus = SW ? (p.sw.coup_same ? sw_in : dg_sw(wfn, e_in)) : dg_coef(in.weight, e_in);
                if (SW ? tt == 0 : in.kind == GDTB_INT_IPDG_INNER_COUPLING) {
                  const double k_in = SW ? (p.sw.coup_same ? sw_in : dg_sw(p.sw.coup_kappa, e_in)) : dg_coef(in.diffusion, e_in);
                  const double k_out = SW ? (p.sw.coup_same ? sw_out : dg_sw(p.sw.coup_kappa, e_out)) : dg_coef(in.diffusion, e_out);
                  // one reciprocal instead of two FP64 divisions (each ~30 instructions; agrees to an ulp)
                  const double rsum = __drcp_rn(delta_plus + delta_minus);
                  const double wm = delta_plus * rsum, wp = delta_minus * rsum;
                  const double sp_ = SW ? p.sw.coup_prefactor : in.prefactor;
                  const double fi = s ? k_in * gi : k_out * gi; // (kappa grad psi_i) . n on the test function's side
#pragma unroll
                  for (int jk = 0; jk < 2; ++jk) {
                    const double vj_in = T.pe[1][jk], vj_out = T.pe[0][jk];
                    const double fj_in = k_in * (T.de[1][jk] * hinv_in), fj_out = k_out * (T.de[0][jk] * hinv_out);
                    if (s) { // laplace-ipdg.hh:158-170 (in_in, in_out)
                      ca[jk] += -1.0 * wm * fj_in * vi;
                      ca[jk] += -1.0 * sp_ * wm * vj_in * fi;
                      cb[jk] += -1.0 * wp * fj_out * vi;
                      cb[jk] += sp_ * wm * vj_out * fi;
                    } else { // laplace-ipdg.hh:172-185 (out_in, out_out)
                      ca[jk] += wm * fj_in * vi;
                      ca[jk] += -1.0 * sp_ * wp * vj_in * fi;
                      cb[jk] += wp * fj_out * vi;
                      cb[jk] += sp_ * wp * vj_out * fi;
                    }
                  }
                } else { // GDTB_INT_IPDG_INNER_PENALTY, ipdg.hh:149-170
                  const double weight = (delta_plus * delta_minus) * __drcp_rn(delta_plus + delta_minus);
                  const double penalty =
                      SW ? (p.sw.pen_prefactor * weight)
                               * dg_inv_face<D>(h, hinv, k, h_in, h_out, true, p.sw.pen_hI == GDTB_HI_VOLUME)
                         : (in.prefactor * weight) * __drcp_rn(dg_face_h<D>(in.hI_kind, h, k, h_in, h_out, true));
#pragma unroll
                  for (int jk = 0; jk < 2; ++jk) {
                    const double vj_in = T.pe[1][jk], vj_out = T.pe[0][jk];
                    if (s) {
                      ca[jk] += penalty * vj_in * vi;
                      cb[jk] += -1.0 * penalty * vj_out * vi;
                    } else {
                      ca[jk] += -1.0 * penalty * vj_in * vi;
                      cb[jk] += penalty * vj_out * vi;
                    }
                  }
                }
              }
              // own columns: inside element -> ca, outside element -> cb
              const double sc_coup = SW ? p.sw.s_coup : F.scaling;
              dg_add_face_block<D>(self, sc_coup, s ? ca : cb, k, tM);
              dg_add_face_block<D>(nbb, sc_coup, s ? cb : ca, k, tM);
            }
            double* blk = row + (s ? pos_hi[k] : pos_lo[k]);
            dg_store_block<N>(blk, nbb, phase == 0);
          } else {
            // boundary face (k, s) with outer normal sg e_k
            const double sg = s ? 1. : -1.;
            for (int f = 0; f < n_bnd; ++f) {
              const FormDev& F = f_bnd[f];
              const DgFastTab& T = t_bnd[f];
              double tM[D][2];
#pragma unroll
              for (int o = 0; o < D; ++o) {
                const int io = (i >> o) & 1;
                tM[o][0] = h[o] * T.M1[io][0];
                tM[o][1] = h[o] * T.M1[io][1];
              }
              if (CC) {
                const double mult[4] = {hinv[k], 0., dg_inv_face<D>(h, hinv, k, h[k], h[k], false, true),
                                        dg_inv_face<D>(h, hinv, k, h[k], h[k], false, false)};
                double cc2[2];
#pragma unroll
                for (int jk = 0; jk < 2; ++jk) {
                  const double* A = T.fa[s][ik][jk];
                  cc2[jk] = fma(A[0], mult[0], fma(A[2], mult[2], A[3] * mult[3]));
                }
                dg_add_face_block<D>(self, 1., cc2, k, tM);
                continue;
              }
              const double vi = T.pe[s][ik], gi = sg * (T.de[s][ik] * hinv[k]);
              double ca[2] = {0., 0.};
#pragma unroll
              for (int tt = 0; tt < (SW ? 2 : F.n_terms); ++tt) {
                const IntegrandDev& in = F.terms[tt];
                if (SW ? tt == 0 : in.kind == GDTB_INT_IPDG_DIRICHLET_COUPLING) { // laplace-ipdg.hh:362-367
                  const double kap = SW ? dg_sw(p.sw.bnd_kappa, e) : dg_coef(in.diffusion, e);
                  const double fi = kap * gi;
#pragma unroll
                  for (int jk = 0; jk < 2; ++jk) {
                    const double vj = T.pe[s][jk], fj = kap * (sg * (T.de[s][jk] * hinv[k]));
                    ca[jk] += -1.0 * fj * vi;
                    ca[jk] += -1.0 * (SW ? p.sw.bnd_prefactor : in.prefactor) * vj * fi;
                  }
                } else { // GDTB_INT_IPDG_BOUNDARY_PENALTY, ipdg.hh:276-281
                  const double penalty =
                      SW ? (p.sw.bndpen_prefactor * dg_sw(p.sw.bnd_weight, e))
                               * dg_inv_face<D>(h, hinv, k, h[k], h[k], false, p.sw.bndpen_hI == GDTB_HI_VOLUME)
                         : (in.prefactor * dg_coef(in.weight, e)) * __drcp_rn(dg_face_h<D>(in.hI_kind, h, k, h[k], h[k], false));
#pragma unroll
                  for (int jk = 0; jk < 2; ++jk)
                    ca[jk] += penalty * T.pe[s][jk] * vi;
                }
              }
              dg_add_face_block<D>(self, SW ? p.sw.s_bnd : F.scaling, ca, k, tM);
            }
          }
        }
      }
      dg_store_block<N>(row + pos_self, self, phase == 0);
    }

    if (ACCUMULATE) {
      __syncthreads();
      for (int t = threadIdx.x; t < seg; t += blockDim.x)
        values[start + t] += stage[t];
      __syncthreads();
    } else {
      dg_fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        const int head = phase;
        const int body = (seg - head) & ~1;
        if (head)
          values[start] = stage[0];
        if (body > 0)
          dg_bulk_store_s2g(values + start + head, stage + head, (unsigned)(body * sizeof(double)));
        if (head + body < seg)
          values[start + head + body] = stage[head + body];
        dg_bulk_commit();
        // two stages: the store of this item overlaps the next item's arithmetic; one stage (more blocks per SM): the
        // stage must have been read out before the next item is written, other blocks fill the gap
      }
      buf = nbuf == 1 ? 0 : buf ^ 1;
    }
  }
  if (!ACCUMULATE && threadIdx.x == 0)
    dg_bulk_wait0();
}

template <int D>
int launch_dg_gather_fast(Launch& L, DgGatherParams& p, double* values, bool accumulate)
{
  const bool cc = p.fast == 2;
  constexpr int N = 1 << D;
  for (int k = 0; k < 2; ++k)
    p.magic[k] = p.g.n[k] > 1 ? ~0ULL / (unsigned long long)p.g.n[k] + 1 : 0;
  const int stage_doubles = ((DGG_THREADS * N * (2 * D + 1) + 2) + 1) & ~1;
  static const int nbuf_env = std::getenv("GDTB_DG_NBUF") ? std::atoi(std::getenv("GDTB_DG_NBUF")) : 0;
  const int nbuf = accumulate ? 1 : (nbuf_env == 1 || nbuf_env == 2 ? nbuf_env : 1); // measured: 0.62 ms vs 0.72 ms (C3)
  const size_t smem = (size_t)nbuf * stage_doubles * sizeof(double);
  const bool one = p.n_elem == 1 && p.n_coup == 1 && p.n_bnd == 1;
  auto kern = accumulate ? (cc ? k_dg_gather_fast<D, true, true, false> : k_dg_gather_fast<D, true, false, false>)
                         : (cc ? (one ? k_dg_gather_fast<D, false, true, true> : k_dg_gather_fast<D, false, true, false>)
                               : (one && p.swip ? k_dg_gather_fast<D, false, false, true> : k_dg_gather_fast<D, false, false, false>));
  if (p.g.periodic)
    kern = accumulate ? (cc ? k_dg_gather_fast<D, true, true, false, true> : k_dg_gather_fast<D, true, false, false, true>)
                      : (cc ? k_dg_gather_fast<D, false, true, false, true> : k_dg_gather_fast<D, false, false, false, true>);
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  GDTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DGG_THREADS, smem));
  if (per_sm < 1)
    return fail(GDTB_ERR_CUDA, "dg_gather: kernel does not fit on an SM");
  // shared-memory carve-out: what the resident blocks need (dynamic + static + 1 KB each), the rest of the 256 KB stays L1
  {
    cudaFuncAttributes fa;
    GDTB_CUDA(cudaFuncGetAttributes(&fa, kern));
    const size_t per_block = smem + fa.sharedSizeBytes + 1024;
    GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   (int)std::min<size_t>(100, ((size_t)per_sm * per_block * 100) / (228 * 1024) + 2)));
  }
  const long long nitems = ((p.e_end - p.e_begin) * N + DGG_THREADS - 1) / DGG_THREADS;
  long long grid = (long long)per_sm * L.sm_count;
  if (grid > nitems)
    grid = nitems;
  note_kernel(L, KF_DG_GATHER, reinterpret_cast<const void*>(kern));
  time_begin(L, KF_DG_GATHER);
  kern<<<(unsigned)grid, DGG_THREADS, smem, L.stream>>>(p, values, stage_doubles, nbuf);
  time_end(L, KF_DG_GATHER);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

} // namespace

int launch_dg_gather_fast_d(Launch& L, const DgGatherParams& p, double* values, bool accumulate)
{
  DgGatherParams q = p;
  switch (p.g.d) {
    case 1: return launch_dg_gather_fast<1>(L, q, values, accumulate);
    case 2: return launch_dg_gather_fast<2>(L, q, values, accumulate);
    default: return launch_dg_gather_fast<3>(L, q, values, accumulate);
  }
}

} // namespace gdtb

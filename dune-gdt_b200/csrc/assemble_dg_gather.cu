// dune-gdt_b200/csrc/assemble_dg_gather.cu -- owner-computes-rows assembly for discontinuous-Lagrange spaces
// (SWIPDG: element + inner-coupling + boundary forms, Stencil::element_and_intersection) on structured cube grids.
//
// The reference walks every inner face once and scatters the four coupling blocks into the rows of BOTH adjacent
// elements (LocalCouplingIntersectionBilinearFormAssembler::apply_local, local/assembler/bilinear-form-assemblers.hh:
// 238-278), plus the element and boundary-face forms (:110-128, :380-396): 4 n^2 + ... locked read-modify-writes per
// face.  A DG row belongs to exactly one element, so here one thread owns one row (element e, local test function i)
// and gathers everything the walk would have added to it: the rows of the out_in / out_out blocks of the faces on
// which e is the outside element (its lower faces), its element forms, the boundary forms of its boundary faces and
// the in_in / in_out rows of its upper faces.  Face quadratures are therefore evaluated from both sides (twice the
// flops of the face-once walk) but every matrix value is written exactly once, without atomics or colour passes, and
// in a fixed summation order.  The local forms themselves are the quadrature-faithful restatements of
// local_forms.cuh (coefficients evaluated per quadrature point).
//
// Layout: rows are consecutive per element and the blocks of a row are ordered by the neighbour's element index
// (z-, y-, x-, self, x+, y+, z+ on a non-periodic cube grid), so block positions are closed forms; rowptr is read once
// per row, colidx never.  A work item is a run of consecutive rows = one contiguous CSR segment, staged in shared
// memory and written by a TMA bulk store (double-buffered, persistent CTAs), like the CG gather kernels.
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "kernels.hpp"
#include "local_forms.cuh"

namespace gdtb {

namespace {

__device__ __forceinline__ void dg_fence_proxy_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void dg_bulk_store_s2g(double* gdst, const double* ssrc, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void dg_bulk_commit()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__device__ __forceinline__ void dg_bulk_wait_read1()
{
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

__device__ __forceinline__ void dg_bulk_wait_read0()
{
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ void dg_bulk_wait0()
{
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__device__ inline void load_tables_from(Tables& s, const FormDev* f)
{
  for (int t = threadIdx.x; t < MAX_Q1D * (MAX_K + 1); t += blockDim.x) {
    (&s.phi[0][0])[t] = (&f->phi[0][0])[t];
    (&s.dphi[0][0])[t] = (&f->dphi[0][0])[t];
  }
  for (int t = threadIdx.x; t < 2 * (MAX_K + 1); t += blockDim.x) {
    (&s.phi_end[0][0])[t] = (&f->phi_end[0][0])[t];
    (&s.dphi_end[0][0])[t] = (&f->dphi_end[0][0])[t];
  }
}

template <int N>
__device__ __forceinline__ void axpy_clear(double* __restrict__ y, double a, double* __restrict__ x)
{
#pragma unroll
  for (int j = 0; j < N; ++j) {
    y[j] += a * x[j];
    x[j] = 0.;
  }
}

template <int D, int K, bool ACCUMULATE>
__global__ void __launch_bounds__(DGG_THREADS)
    k_dg_gather(const __grid_constant__ DgGatherParams p, double* __restrict__ values, int stage_doubles)
{
  using L = Loc<D, K>;
  constexpr int N = L::N;
  extern __shared__ __align__(16) double smem[];
  __shared__ Tables tabs[DGG_MAX_FORMS];
  const GridDev& g = p.g;
  const int n_forms = p.n_elem + p.n_coup + p.n_bnd;
  for (int f = 0; f < n_forms; ++f)
    load_tables_from(tabs[f], p.forms + f);
  __syncthreads();
  const FormDev* f_elem = p.forms;
  const FormDev* f_coup = p.forms + p.n_elem;
  const FormDev* f_bnd = p.forms + p.n_elem + p.n_coup;
  const Tables* t_elem = tabs;
  const Tables* t_coup = tabs + p.n_elem;
  const Tables* t_bnd = tabs + p.n_elem + p.n_coup;
  // rows of the element range [e_begin, e_end) this process owns (a slab of element layers; the whole grid otherwise):
  // DG rows are element-owned, the neighbours only enter through their index, geometry and coefficients
  const long long row_first = p.e_begin * N, nrows_total = (p.e_end - p.e_begin) * N;
  const long long nitems = (nrows_total + DGG_THREADS - 1) / DGG_THREADS;
  int buf = 0;

  for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const long long r0 = row_first + item * DGG_THREADS;
    const int nr = (int)min((long long)DGG_THREADS, row_first + nrows_total - r0);
    const long long start = __ldg(p.rowptr + r0) - p.value_offset;
    const int seg = int(__ldg(p.rowptr + r0 + nr) - p.value_offset - start);
    const int phase = int((reinterpret_cast<unsigned long long>(values + start) >> 3) & 1ULL);
    double* stage = smem + buf * stage_doubles + phase;

    if ((int)threadIdx.x < nr) {
      const long long r = r0 + threadIdx.x;
      const long long e = r / N;
      const int i = int(r - e * N);
      long long idx[3];
      elem_coords(g, e, idx);
      double self[N], nb[2 * D][N], ta[N], tb[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        self[j] = ta[j] = tb[j] = 0.;
#pragma unroll
        for (int b = 0; b < 2 * D; ++b)
          nb[b][j] = 0.;
      }
      bool has_lo[D], has_hi[D];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        has_lo[k] = idx[k] > 0;
        has_hi[k] = idx[k] < g.n[k] - 1;
      }
      // faces on which this element is the OUTSIDE one (the lower neighbours were visited earlier by the walker):
      // rows of out_in (columns of the neighbour) and out_out (own columns)
#pragma unroll
      for (int k = D - 1; k >= 0; --k)
        if (has_lo[k]) {
          long long in[3] = {idx[0], idx[1], idx[2]};
          in[k] -= 1;
          for (int f = 0; f < p.n_coup; ++f) {
            coupling_row<D, K>(g, f_coup[f], t_coup[f], in, idx, k, 1, false, i, ta, tb);
            axpy_clear<N>(nb[k], f_coup[f].scaling, ta);
            axpy_clear<N>(self, f_coup[f].scaling, tb);
          }
        }
      // element forms
      for (int f = 0; f < p.n_elem; ++f) {
        element_row<D, K>(g, f_elem[f], t_elem[f], idx, i, ta);
        axpy_clear<N>(self, f_elem[f].scaling, ta);
      }
      // own intersections in order: boundary forms on boundary faces, coupling forms (as inside) on upper inner faces
#pragma unroll
      for (int k = 0; k < D; ++k) {
        if (!has_lo[k])
          for (int f = 0; f < p.n_bnd; ++f) {
            boundary_row<D, K>(g, f_bnd[f], t_bnd[f], idx, k, 0, i, ta);
            axpy_clear<N>(self, f_bnd[f].scaling, ta);
          }
        if (has_hi[k]) {
          long long out[3] = {idx[0], idx[1], idx[2]};
          out[k] += 1;
          for (int f = 0; f < p.n_coup; ++f) {
            coupling_row<D, K>(g, f_coup[f], t_coup[f], idx, out, k, 1, true, i, ta, tb);
            axpy_clear<N>(self, f_coup[f].scaling, ta);
            axpy_clear<N>(nb[D + k], f_coup[f].scaling, tb);
          }
        } else
          for (int f = 0; f < p.n_bnd; ++f) {
            boundary_row<D, K>(g, f_bnd[f], t_bnd[f], idx, k, 1, i, ta);
            axpy_clear<N>(self, f_bnd[f].scaling, ta);
          }
      }
      // the row in CSR order: blocks by ascending neighbour index
      double* row = stage + int(__ldg(p.rowptr + r) - p.value_offset - start);
#pragma unroll
      for (int k = D - 1; k >= 0; --k)
        if (has_lo[k]) {
#pragma unroll
          for (int j = 0; j < N; ++j)
            row[j] = nb[k][j];
          row += N;
        }
#pragma unroll
      for (int j = 0; j < N; ++j)
        row[j] = self[j];
      row += N;
#pragma unroll
      for (int k = 0; k < D; ++k)
        if (has_hi[k]) {
#pragma unroll
          for (int j = 0; j < N; ++j)
            row[j] = nb[D + k][j];
          row += N;
        }
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
        dg_bulk_wait_read1();
      }
      __syncthreads();
      buf ^= 1;
    }
  }
  if (!ACCUMULATE && threadIdx.x == 0)
    dg_bulk_wait0();
}

template <int D, int K>
int launch_dg_gather_dk(Launch& L, const DgGatherParams& p, double* values, bool accumulate)
{
  constexpr int N = Loc<D, K>::N;
  const int stage_doubles = ((DGG_THREADS * N * (2 * D + 1) + 2) + 1) & ~1;
  const size_t smem = (size_t)(accumulate ? 1 : 2) * stage_doubles * sizeof(double);
  auto kern = accumulate ? k_dg_gather<D, K, true> : k_dg_gather<D, K, false>;
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  GDTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DGG_THREADS, smem));
  if (per_sm < 1)
    return fail(GDTB_ERR_CUDA, "dg_gather: kernel does not fit on an SM");
  const long long nitems = ((p.e_end - p.e_begin) * N + DGG_THREADS - 1) / DGG_THREADS;
  long long grid = (long long)per_sm * L.sm_count;
  if (grid > nitems)
    grid = nitems;
  note_kernel(L, KF_DG_GATHER, reinterpret_cast<const void*>(kern));
  time_begin(L, KF_DG_GATHER);
  kern<<<(unsigned)grid, DGG_THREADS, smem, L.stream>>>(p, values, stage_doubles);
  time_end(L, KF_DG_GATHER);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

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

__device__ __forceinline__ double dg_coef(const FnDev& f, long long e)
{
  return f.kind == GDTB_FN_ELEM_SCALAR ? __ldg(f.data + e) : f.c[0];
}

__device__ __forceinline__ double dg_ext(const GridDev& g, int k, int i)
{
  const double lower = __dadd_rn(g.lo[k], __dmul_rn(double(i), g.h[k]));
  const double upper = __dadd_rn(g.lo[k], __dmul_rn(double(i + 1), g.h[k]));
  return __dsub_rn(upper, lower);
}

// blocks (element + existing neighbours) of all elements before e in the element_and_intersection pattern
template <int D>
__host__ __device__ __forceinline__ long long dg_blocks_before(const GridDev& g, const long long e, const int* idx)
{
  const long long nx = g.n[0];
  long long P = e;
  const long long m = D > 1 ? (long long)idx[1] + (D > 2 ? g.n[1] * idx[2] : 0) : 0; // complete x-lines before e
  P += (e - m - (idx[0] > 0 ? 1 : 0)) + (e - m);                                       // lower / upper x neighbours
  if (D > 1) {
    const long long z = D > 2 ? idx[2] : 0;
    const long long y0 = z * nx + (idx[1] > 0 ? nx : idx[0]);                  // elements before e with y == 0
    const long long y1 = z * nx + (idx[1] == g.n[1] - 1 ? (long long)idx[0] : 0); // ... with y == n_y - 1
    P += (e - y0) + (e - y1);
  }
  if (D > 2) {
    const long long plane = nx * g.n[1];
    P += (e - min(e, plane)) + (e - max(0LL, e - (g.n[2] - 1) * plane));
  }
  return P;
}

template <int D>
__host__ __device__ __forceinline__ int dg_nblocks(const GridDev& g, const int* idx)
{
  int nb = 1;
#pragma unroll
  for (int k = 0; k < D; ++k)
    nb += (idx[k] > 0 ? 1 : 0) + (idx[k] < g.n[k] - 1 ? 1 : 0);
  return nb;
}

template <int D>
__device__ __forceinline__ void dg_decode(const DgGatherParams& p, const unsigned e, int* idx)
{
  const GridDev& g = p.g;
  const unsigned nx = (unsigned)g.n[0];
  const unsigned t1 = D > 1 ? (nx == 1 ? e : (unsigned)__umul64hi((unsigned long long)e, p.magic[0])) : 0;
  idx[0] = int(e - t1 * nx);
  idx[1] = idx[2] = 0;
  if (D == 2)
    idx[1] = (int)t1;
  if (D == 3) {
    const unsigned ny = (unsigned)g.n[1];
    const unsigned t2 = ny == 1 ? t1 : (unsigned)__umul64hi((unsigned long long)t1, p.magic[1]);
    idx[1] = int(t1 - t2 * ny);
    idx[2] = (int)t2;
  }
}

// intersection_h of local_forms.cuh for axis-aligned faces: |I|, or the face diameter (1D: element lengths)
template <int D>
__device__ __forceinline__ double dg_face_h(const IntegrandDev& t, const double* h, int k, double h_in, double h_out,
                                            bool neighbor)
{
  double ie = 1., d2 = 0.;
#pragma unroll
  for (int o = 0; o < D; ++o)
    if (o != k) {
      ie *= h[o];
      d2 += h[o] * h[o];
    }
  if (t.hI_kind == GDTB_HI_VOLUME)
    return ie;
  if (D == 1)
    return neighbor ? 0.5 * (h_in + h_out) : h_in;
  return sqrt(d2);
}

// 1 / intersection_h for the CC tables: 1 / |I| (volume) or 1 / diameter (1D: element lengths) from the cell data
template <int D>
__device__ __forceinline__ double dg_inv_face(const double* h, const double* hinv, int k, double h_in, double h_out,
                                              bool neighbor, bool volume)
{
  if (D == 1)
    return volume ? 1. : (neighbor ? 1. / (0.5 * (h_in + h_out)) : 1. / h_in);
  double inv = 1., d2 = 0.;
#pragma unroll
  for (int o = 0; o < D; ++o)
    if (o != k) {
      inv *= hinv[o];
      d2 += h[o] * h[o];
    }
  if (volume || D == 2)
    return inv; // 2D: the face is an interval, diameter == |I|
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

template <int D, bool ACCUMULATE, bool CC>
__global__ void __launch_bounds__(DGG_THREADS)
    k_dg_gather_fast(const __grid_constant__ DgGatherParams p, double* __restrict__ values, int stage_doubles, int nbuf)
{
  constexpr int N = 1 << D;
  extern __shared__ __align__(16) double smem[];
  __shared__ DgFastTab tabs[DGG_MAX_FORMS];
  const GridDev& g = p.g;
  const int n_forms = p.n_elem + p.n_coup + p.n_bnd;
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
      dg_fast_tables_cc(f, (int)threadIdx.x < p.n_elem ? 0 : ((int)threadIdx.x < p.n_elem + p.n_coup ? 1 : 2), t);
  }
  __syncthreads();
  const FormDev* f_elem = p.forms;
  const FormDev* f_coup = p.forms + p.n_elem;
  const FormDev* f_bnd = p.forms + p.n_elem + p.n_coup;
  const DgFastTab* t_elem = tabs;
  const DgFastTab* t_coup = tabs + p.n_elem;
  const DgFastTab* t_bnd = tabs + p.n_elem + p.n_coup;
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
    if (le < ne_item) {
      const long long e = e0 + le;
      int idx[3];
      dg_decode<D>(p, (unsigned)e, idx);
      const int nblocks = dg_nblocks<D>(g, idx);
      double* row = stage + int((long long)N * N * dg_blocks_before<D>(g, e, idx) - p.value_offset - start) + i * nblocks * N;
      long long estride[3] = {1, g.n[0], g.n[0] * g.n[1]};
      bool has_lo[D], has_hi[D];
      double h[D], hinv[D];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        has_lo[k] = idx[k] > 0;
        has_hi[k] = idx[k] < g.n[k] - 1;
        h[k] = dg_ext(g, k, idx[k]);
        hinv[k] = __drcp_rn(h[k]);
      }
      double self[N];
#pragma unroll
      for (int j = 0; j < N; ++j)
        self[j] = 0.;

      // ---- element forms --------------------------------------------------------------------------------------
      for (int f = 0; f < p.n_elem; ++f) {
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
        for (int tt = 0; tt < (CC ? 2 : F.n_terms); ++tt) {
          const double c = CC ? (tt == 0 ? T.elap : T.emass) : F.scaling * dg_coef(F.terms[tt].diffusion, e);
          if (CC && c == 0.)
            continue;
          if (CC ? tt == 0 : F.terms[tt].kind == GDTB_INT_LAPLACE) {
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
      int pos = 0;
      int pos_lo[D], pos_hi[D];
#pragma unroll
      for (int k = D - 1; k >= 0; --k) {
        pos_lo[k] = pos;
        pos += has_lo[k] ? N : 0;
      }
      const int pos_self = pos;
      pos += N;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        pos_hi[k] = pos;
        pos += has_hi[k] ? N : 0;
      }

#pragma unroll
      for (int k = 0; k < D; ++k) {
        const int ik = (i >> k) & 1;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const bool has = s ? has_hi[k] : has_lo[k];
          if (has) {
            // inner face: s == 0: this element is the outside one (inside = e - stride), s == 1: it is the inside one
            const long long e_in = s ? e : e - estride[k], e_out = s ? e + estride[k] : e;
            const double h_in = s ? h[k] : dg_ext(g, k, idx[k] - 1), h_out = s ? dg_ext(g, k, idx[k] + 1) : h[k];
            const double hinv_in = s ? hinv[k] : __drcp_rn(h_in), hinv_out = s ? __drcp_rn(h_out) : hinv[k];
            double nbb[N];
#pragma unroll
            for (int j = 0; j < N; ++j)
              nbb[j] = 0.;
            for (int f = 0; f < p.n_coup; ++f) {
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
              for (int tt = 0; tt < F.n_terms; ++tt) {
                const IntegrandDev& in = F.terms[tt];
                const double delta_plus = dg_coef(in.weight, e_out), delta_minus = dg_coef(in.weight, e_in);
                if (in.kind == GDTB_INT_IPDG_INNER_COUPLING) {
                  const double k_in = dg_coef(in.diffusion, e_in), k_out = dg_coef(in.diffusion, e_out);
                  const double wm = delta_plus / (delta_plus + delta_minus), wp = delta_minus / (delta_plus + delta_minus);
                  const double sp_ = in.prefactor;
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
                  const double weight = (delta_plus * delta_minus) / (delta_plus + delta_minus);
                  const double penalty = (in.prefactor * weight) / dg_face_h<D>(in, h, k, h_in, h_out, true);
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
              dg_add_face_block<D>(self, F.scaling, s ? ca : cb, k, tM);
              dg_add_face_block<D>(nbb, F.scaling, s ? cb : ca, k, tM);
            }
            double* blk = row + (s ? pos_hi[k] : pos_lo[k]);
#pragma unroll
            for (int j = 0; j < N; ++j)
              blk[j] = nbb[j];
          } else {
            // boundary face (k, s) with outer normal sg e_k
            const double sg = s ? 1. : -1.;
            for (int f = 0; f < p.n_bnd; ++f) {
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
              for (int tt = 0; tt < F.n_terms; ++tt) {
                const IntegrandDev& in = F.terms[tt];
                if (in.kind == GDTB_INT_IPDG_DIRICHLET_COUPLING) { // laplace-ipdg.hh:362-367
                  const double kap = dg_coef(in.diffusion, e);
                  const double fi = kap * gi;
#pragma unroll
                  for (int jk = 0; jk < 2; ++jk) {
                    const double vj = T.pe[s][jk], fj = kap * (sg * (T.de[s][jk] * hinv[k]));
                    ca[jk] += -1.0 * fj * vi;
                    ca[jk] += -1.0 * in.prefactor * vj * fi;
                  }
                } else { // GDTB_INT_IPDG_BOUNDARY_PENALTY, ipdg.hh:276-281
                  const double penalty = (in.prefactor * dg_coef(in.weight, e)) / dg_face_h<D>(in, h, k, h[k], h[k], false);
#pragma unroll
                  for (int jk = 0; jk < 2; ++jk)
                    ca[jk] += penalty * T.pe[s][jk] * vi;
                }
              }
              dg_add_face_block<D>(self, F.scaling, ca, k, tM);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < N; ++j)
        row[pos_self + j] = self[j];
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
        if (nbuf == 1)
          dg_bulk_wait_read0();
        else
          dg_bulk_wait_read1();
      }
      __syncthreads();
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
  auto kern = accumulate ? (cc ? k_dg_gather_fast<D, true, true> : k_dg_gather_fast<D, true, false>)
                         : (cc ? k_dg_gather_fast<D, false, true> : k_dg_gather_fast<D, false, false>);
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  GDTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DGG_THREADS, smem));
  if (per_sm < 1)
    return fail(GDTB_ERR_CUDA, "dg_gather: kernel does not fit on an SM");
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

bool dg_gather_supported(int d, int K)
{
  return (K == 1 && d >= 1 && d <= 3) || (K == 2 && d >= 1 && d <= 2);
}

bool dg_gather_fast_supported(const GridDev& g, int K)
{
  return K == 1 && g.d >= 1 && g.d <= 3 && g.ne < (1LL << 31) / (1 << g.d);
}

int launch_dg_gather(Launch& L, const DgGatherParams& p, double* values, bool accumulate)
{
  if (p.fast) {
    DgGatherParams q = p;
    switch (p.g.d) {
      case 1: return launch_dg_gather_fast<1>(L, q, values, accumulate);
      case 2: return launch_dg_gather_fast<2>(L, q, values, accumulate);
      default: return launch_dg_gather_fast<3>(L, q, values, accumulate);
    }
  }
  switch (p.g.d * 10 + p.sp.K) {
    case 11: return launch_dg_gather_dk<1, 1>(L, p, values, accumulate);
    case 21: return launch_dg_gather_dk<2, 1>(L, p, values, accumulate);
    case 31: return launch_dg_gather_dk<3, 1>(L, p, values, accumulate);
    case 12: return launch_dg_gather_dk<1, 2>(L, p, values, accumulate);
    case 22: return launch_dg_gather_dk<2, 2>(L, p, values, accumulate);
    default: return fail(GDTB_ERR_NOT_IMPLEMENTED, "dg_gather: unsupported (dimension, order)");
  }
}

// ---- closed-form sparsity pattern of the DG element_and_intersection stencil (non-periodic grid) --------------------
// One thread per row (element e, local DoF i): blocks of nloc columns per existing neighbour in ascending element index
// (z-, y-, x-, e, x+, y+, z+), row starts from dg_blocks_before.
namespace {

template <int D>
__global__ void __launch_bounds__(128) k_dg_pattern(const __grid_constant__ DgGatherParams p, int nloc,
                                                     long long* __restrict__ rowptr, int* __restrict__ colidx)
{
  const GridDev& g = p.g;
  const long long rows = g.ne * nloc;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= rows; r += (long long)gridDim.x * blockDim.x) {
    const long long e = r == rows ? g.ne - 1 : r / nloc;
    const int i = r == rows ? nloc : int(r - e * nloc);
    int idx[3];
    dg_decode<D>(p, (unsigned)e, idx);
    const int nb = dg_nblocks<D>(g, idx);
    const long long start = (long long)nloc * nloc * dg_blocks_before<D>(g, e, idx) + (long long)i * nb * nloc;
    rowptr[r] = start;
    if (r == rows)
      continue;
    int* out = colidx + start;
    const long long stride[3] = {1, g.n[0], g.n[0] * g.n[1]};
#pragma unroll
    for (int k = D - 1; k >= 0; --k)
      if (idx[k] > 0)
        for (int j = 0; j < nloc; ++j)
          *out++ = (int)((e - stride[k]) * nloc + j);
    for (int j = 0; j < nloc; ++j)
      *out++ = (int)(e * nloc + j);
#pragma unroll
    for (int k = 0; k < D; ++k)
      if (idx[k] < g.n[k] - 1)
        for (int j = 0; j < nloc; ++j)
          *out++ = (int)((e + stride[k]) * nloc + j);
  }
}

} // namespace

int pattern_structured_dg(Launch& L, const GridDev& g, const SpaceDev& sp, long long** d_rowptr, int** d_colidx,
                          long long* nnz_out)
{
  if (g.periodic)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "structured DG pattern: non-periodic grids only");
  if (sp.size >= (1LL << 31) || g.ne >= (1LL << 31))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "structured DG pattern: more than 2^31 degrees of freedom");
  const int nloc = sp.nloc;
  long long blocks = g.ne; // element blocks + one block per (element, existing neighbour)
  for (int k = 0; k < g.d; ++k)
    blocks += 2 * (g.ne / g.n[k]) * (g.n[k] - 1);
  const long long nnz = blocks * nloc * nloc;
  DgGatherParams p;
  std::memset(&p, 0, sizeof(p));
  p.g = g;
  p.sp = sp;
  for (int k = 0; k < 2; ++k)
    p.magic[k] = g.n[k] > 1 ? ~0ULL / (unsigned long long)g.n[k] + 1 : 0;
  long long* rowptr = nullptr;
  int* colidx = nullptr;
  if (cudaMalloc(&rowptr, sizeof(long long) * (size_t)(sp.size + 1)) != cudaSuccess
      || cudaMalloc(&colidx, sizeof(int) * (size_t)nnz) != cudaSuccess) {
    cudaFree(rowptr);
    return fail(GDTB_ERR_OUT_OF_MEMORY, "pattern: out of device memory");
  }
  const unsigned grid = (unsigned)std::min<long long>((sp.size + 1 + 127) / 128, (long long)L.sm_count * 64);
  switch (g.d) {
    case 1: k_dg_pattern<1><<<grid, 128, 0, L.stream>>>(p, nloc, rowptr, colidx); break;
    case 2: k_dg_pattern<2><<<grid, 128, 0, L.stream>>>(p, nloc, rowptr, colidx); break;
    default: k_dg_pattern<3><<<grid, 128, 0, L.stream>>>(p, nloc, rowptr, colidx); break;
  }
  L.count++;
  cudaError_t err = cudaStreamSynchronize(L.stream);
  if (err != cudaSuccess) {
    cudaFree(rowptr);
    cudaFree(colidx);
    return fail(GDTB_ERR_CUDA, std::string("k_dg_pattern: ") + cudaGetErrorString(err));
  }
  *d_rowptr = rowptr;
  *d_colidx = colidx;
  *nnz_out = nnz;
  return GDTB_OK;
}

template <int D>
static void dg_host_rowptr_d(const GridDev& g, int nloc, long long* rowptr)
{
  long long r = 0;
  int idx[3] = {0, 0, 0};
  for (long long e = 0; e < g.ne; ++e) {
    idx[0] = int(e % g.n[0]);
    idx[1] = D > 1 ? int((e / g.n[0]) % g.n[1]) : 0;
    idx[2] = D > 2 ? int(e / (g.n[0] * g.n[1])) : 0;
    const long long base = (long long)nloc * nloc * dg_blocks_before<D>(g, e, idx);
    const int nb = dg_nblocks<D>(g, idx);
    for (int i = 0; i < nloc; ++i)
      rowptr[r++] = base + (long long)i * nb * nloc;
    if (e == g.ne - 1)
      rowptr[r] = base + (long long)nloc * nb * nloc;
  }
}

long long dg_value_offset(const GridDev& g, const SpaceDev& sp, long long e)
{
  const long long nn = (long long)sp.nloc * sp.nloc;
  const long long ee = std::min(e, g.ne - 1);
  int idx[3] = {int(ee % g.n[0]), g.d > 1 ? int((ee / g.n[0]) % g.n[1]) : 0, g.d > 2 ? int(ee / (g.n[0] * g.n[1])) : 0};
  long long blocks;
  int nb;
  switch (g.d) {
    case 1: blocks = dg_blocks_before<1>(g, ee, idx); nb = dg_nblocks<1>(g, idx); break;
    case 2: blocks = dg_blocks_before<2>(g, ee, idx); nb = dg_nblocks<2>(g, idx); break;
    default: blocks = dg_blocks_before<3>(g, ee, idx); nb = dg_nblocks<3>(g, idx); break;
  }
  return nn * (e >= g.ne ? blocks + nb : blocks);
}

int dg_host_rowptr(const GridDev& g, const SpaceDev& sp, long long* rowptr)
{
  switch (g.d) {
    case 1: dg_host_rowptr_d<1>(g, sp.nloc, rowptr); break;
    case 2: dg_host_rowptr_d<2>(g, sp.nloc, rowptr); break;
    default: dg_host_rowptr_d<3>(g, sp.nloc, rowptr); break;
  }
  return GDTB_OK;
}

} // namespace gdtb

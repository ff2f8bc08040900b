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

#include "dg_gather.cuh"
#include "local_forms.cuh"

namespace gdtb {

namespace {

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
  if (p.fast)
    return launch_dg_gather_fast_d(L, p, values, accumulate);
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
    int lo[D], hi[D], self;
    dg_block_positions<D>(g, idx, lo, hi, self);
    for (int j = 0; j < nloc; ++j)
      out[self * nloc + j] = (int)(e * nloc + j);
#pragma unroll
    for (int k = 0; k < D; ++k) {
      // neighbour cells along k (periodic wrap included)
      const long long e_lo = e + (idx[k] > 0 ? -stride[k] : (g.n[k] - 1) * stride[k]);
      const long long e_hi = e + (idx[k] < g.n[k] - 1 ? stride[k] : -(g.n[k] - 1) * stride[k]);
      if (lo[k] >= 0)
        for (int j = 0; j < nloc; ++j)
          out[lo[k] * nloc + j] = (int)(e_lo * nloc + j);
      if (hi[k] >= 0)
        for (int j = 0; j < nloc; ++j)
          out[hi[k] * nloc + j] = (int)(e_hi * nloc + j);
    }
  }
}

} // namespace

int pattern_structured_dg(Launch& L, const GridDev& g, const SpaceDev& sp, long long** d_rowptr, int** d_colidx,
                          long long* nnz_out)
{
  if (!dg_closed_form_grid(g))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "structured DG pattern: periodic directions need at least three cells");
  if (sp.size >= (1LL << 31) || g.ne >= (1LL << 31))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "structured DG pattern: more than 2^31 degrees of freedom");
  const int nloc = sp.nloc;
  long long blocks = g.ne; // element blocks + one block per (element, existing neighbour)
  for (int k = 0; k < g.d; ++k)
    blocks += dg_periodic(g, k) ? 2 * g.ne : 2 * (g.ne / g.n[k]) * (g.n[k] - 1);
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

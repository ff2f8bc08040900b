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
  const long long nrows_total = g.ne * N;
  const long long nitems = (nrows_total + DGG_THREADS - 1) / DGG_THREADS;
  int buf = 0;

  for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const long long r0 = item * DGG_THREADS;
    const int nr = (int)min((long long)DGG_THREADS, nrows_total - r0);
    const long long start = __ldg(p.rowptr + r0);
    const int seg = int(__ldg(p.rowptr + r0 + nr) - start);
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
      double* row = stage + int(__ldg(p.rowptr + r) - start);
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
  const long long nitems = (p.g.ne * N + DGG_THREADS - 1) / DGG_THREADS;
  long long grid = (long long)per_sm * L.sm_count;
  if (grid > nitems)
    grid = nitems;
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

int launch_dg_gather(Launch& L, const DgGatherParams& p, double* values, bool accumulate)
{
  switch (p.g.d * 10 + p.sp.K) {
    case 11: return launch_dg_gather_dk<1, 1>(L, p, values, accumulate);
    case 21: return launch_dg_gather_dk<2, 1>(L, p, values, accumulate);
    case 31: return launch_dg_gather_dk<3, 1>(L, p, values, accumulate);
    case 12: return launch_dg_gather_dk<1, 2>(L, p, values, accumulate);
    case 22: return launch_dg_gather_dk<2, 2>(L, p, values, accumulate);
    default: return fail(GDTB_ERR_NOT_IMPLEMENTED, "dg_gather: unsupported (dimension, order)");
  }
}

} // namespace gdtb

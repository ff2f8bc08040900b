// dune-gdt_b200/csrc/dg_gather.cuh -- pieces shared by the DG row-gather kernels (assemble_dg_gather.cu: the
// quadrature-faithful kernel and the closed-form pattern; assemble_dg_fast.cu: the factorised kernels): TMA bulk-store
// helpers, the closed-form block positions of the element_and_intersection pattern and the element-index decode.
#pragma once
#include "common.cuh"
#include "kernels.hpp"

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

// A periodic direction k (n_k >= 3: both neighbours exist and differ) gives every element two neighbours along k: the
// wrap neighbour of the first cell is the last one (larger index: its block comes after the regular upper neighbour's),
// the wrap neighbour of the last cell the first one (smaller index: before the regular lower neighbour's).
__host__ __device__ __forceinline__ bool dg_periodic(const GridDev& g, int k)
{
  return (g.periodic >> k) & 1;
}

// blocks (element + existing neighbours) of all elements before e in the element_and_intersection pattern
template <int D>
__host__ __device__ __forceinline__ long long dg_blocks_before(const GridDev& g, const long long e, const int* idx)
{
  const long long nx = g.n[0];
  long long P = e;
  const long long m = D > 1 ? (long long)idx[1] + (D > 2 ? g.n[1] * idx[2] : 0) : 0; // complete x-lines before e
  if (dg_periodic(g, 0))
    P += 2 * e;
  else
    P += (e - m - (idx[0] > 0 ? 1 : 0)) + (e - m); // lower / upper x neighbours
  if (D > 1) {
    if (dg_periodic(g, 1))
      P += 2 * e;
    else {
      const long long z = D > 2 ? idx[2] : 0;
      const long long y0 = z * nx + (idx[1] > 0 ? nx : idx[0]);                  // elements before e with y == 0
      const long long y1 = z * nx + (idx[1] == g.n[1] - 1 ? (long long)idx[0] : 0); // ... with y == n_y - 1
      P += (e - y0) + (e - y1);
    }
  }
  if (D > 2) {
    if (dg_periodic(g, 2))
      P += 2 * e;
    else {
      const long long plane = nx * g.n[1];
      P += (e - min(e, plane)) + (e - max(0LL, e - (g.n[2] - 1) * plane));
    }
  }
  return P;
}

template <int D>
__host__ __device__ __forceinline__ int dg_nblocks(const GridDev& g, const int* idx)
{
  int nb = 1;
#pragma unroll
  for (int k = 0; k < D; ++k)
    nb += dg_periodic(g, k) ? 2 : (idx[k] > 0 ? 1 : 0) + (idx[k] < g.n[k] - 1 ? 1 : 0);
  return nb;
}

// Positions (in blocks) of the neighbour blocks inside a row, blocks in ascending order of the neighbour's element
// index: for k = D-1 .. 0 [upper wrap neighbour][lower neighbour], the element itself, for k = 0 .. D-1
// [upper neighbour][lower wrap neighbour].  lo[k] / hi[k] = -1 where there is no neighbour; returns the block count.
template <int D>
__host__ __device__ __forceinline__ int dg_block_positions(const GridDev& g, const int* idx, int* lo, int* hi, int& self)
{
  int pos = 0;
#pragma unroll
  for (int k = D - 1; k >= 0; --k) {
    lo[k] = hi[k] = -1;
    if (dg_periodic(g, k) && idx[k] == g.n[k] - 1)
      hi[k] = pos++;
    if (idx[k] > 0)
      lo[k] = pos++;
  }
  self = pos++;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    if (idx[k] < g.n[k] - 1)
      hi[k] = pos++;
    if (dg_periodic(g, k) && idx[k] == 0)
      lo[k] = pos++;
  }
  return pos;
}

// the closed forms above cover a grid whose periodic directions all have at least three cells
inline bool dg_closed_form_grid(const GridDev& g)
{
  for (int k = 0; k < g.d; ++k)
    if (dg_periodic(g, k) && g.n[k] < 3)
      return false;
  return true;
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

} // namespace

// factorised kernels (assemble_dg_fast.cu)
int launch_dg_gather_fast_d(Launch& L, const DgGatherParams& p, double* values, bool accumulate);

} // namespace gdtb

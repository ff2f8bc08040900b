// dune-gdt_b200/csrc/q2_layout.cuh -- CSR layout arithmetic of the continuous-Lagrange Q2 element stencil on the MCMG
// lattice numbering (closed-form row starts, coupling boxes, column groups) and the TMA bulk-store helpers, shared by
// assemble_q2_gather.cu (constant / element-wise coefficients, per-row kernels) and assemble_q2_qp.cu (coefficients per
// quadrature point, x-fused kernel).  See the header of assemble_q2_gather.cu for the lattice view.
#pragma once

#include "common.cuh"
#include "kernels.hpp"

namespace gdtb {
namespace {

__device__ __forceinline__ void q2_fence_proxy_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void q2_bulk_store_s2g(double* gdst, const double* ssrc, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void q2_bulk_commit()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__device__ __forceinline__ void q2_bulk_wait_read1()
{
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

__device__ __forceinline__ void q2_bulk_wait_read0()
{
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ void q2_bulk_wait0()
{
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void q2_cp_async16(void* sdst, const void* gsrc)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc)
               : "memory");
}

__device__ __forceinline__ void q2_cp_async_commit()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void q2_cp_async_wait_all()
{
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// read-only 16-byte load that the compiler may not sink towards its first use (the work-item record of the NEXT item is
// requested a whole item ahead)
__device__ __forceinline__ int4 q2_ldg_int4_here(const int4* ptr)
{
  int4 v;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}

__device__ __forceinline__ void q2_prefetch_l1(const void* ptr)
{
  asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
}

__device__ __forceinline__ double q2_cell_extent(double lo, double h, int i)
{
  const double lower = __dadd_rn(lo, __dmul_rn(double(i), h));
  const double upper = __dadd_rn(lo, __dmul_rn(double(i + 1), h));
  return __dsub_rn(upper, lower);
}

#ifndef Q2G_THREADS_N
#define Q2G_THREADS_N 256
#endif
constexpr int Q2G_THREADS = Q2G_THREADS_N;

// Closed-form CSR row starts.  Along one axis with N elements a row at lattice coordinate p = 2 c + S couples to
// L(c) lattice points: S = 1: 3 (never clipped); S = 0: 5 - 2 [c == 0] - 2 [c == N].  The rows of a row group are
// lexicographic (x fastest) and a row holds Lx Ly Ll entries, so with the 1D prefix sums PL(c) = sum_{c' < c} L(c')
// and totals T the entries before row (cx, cy, cl) of its group number
//   Tx Ty PLl(cl) + Ll(cl) (Tx PLy(cy) + Ly(cy) PLx(cx)).
// rowptr is never read (it is what the pattern builder produces for this space; the parity tests compare both).
struct Q2AxisLen
{
  int L;        // entries along the axis for this row
  long long PL; // entries along the axis of the rows before it
};
__host__ __device__ __forceinline__ Q2AxisLen q2_axis_len(int S, int c, int N)
{
  Q2AxisLen r;
  if (S) {
    r.L = 3;
    r.PL = 3LL * c;
  } else {
    r.L = 5 - (c == 0 ? 2 : 0) - (c == N ? 2 : 0);
    r.PL = 5LL * c - (c > 0 ? 2 : 0);
  }
  return r;
}
__host__ __device__ __forceinline__ long long q2_axis_total(int S, long long N)
{
  return S ? 3 * N : 5 * N + 1;
}

// n / d for a run-time constant d through its magic m = floor(2^64 / d) + 1 (exact for all 32-bit n; d = 1 has no magic)
__host__ __device__ __forceinline__ unsigned q2_div(const unsigned n, const unsigned d, const unsigned long long m)
{
#ifdef __CUDA_ARCH__
  return d == 1 ? n : (unsigned)__umul64hi((unsigned long long)n, m);
#else
  (void)m;
  return n / d;
#endif
}

// decode the lexicographic row index inside a row group into element-lattice coordinates
template <int D>
__host__ __device__ __forceinline__ void q2_decode(const Q2RowGroup& rg, const unsigned lex, int& cx, int& cy, int& cl)
{
  const unsigned t1 = q2_div(lex, rg.ex, rg.mex);
  cx = int(lex - t1 * rg.ex);
  if (D == 3) {
    const unsigned t2 = q2_div(t1, rg.ey, rg.mey);
    cy = int(t1 - t2 * rg.ey);
    cl = (int)t2;
  } else {
    cy = 0;
    cl = (int)t1;
  }
}

// number of matrix entries of the row group that precede row (cx, cy, cl); the part inside one layer of the last axis
// stays below 2^32 (25 (N + 1)^2 entries)
template <int D>
__host__ __device__ __forceinline__ long long q2_row_offset(const GridDev& g, const Q2RowGroup& rg, const int cx, const int cy,
                                                   const int cl)
{
  const int s = rg.s;
  const Q2AxisLen X = q2_axis_len(s & 1, cx, (int)g.n[0]);
  if (D == 3) {
    const Q2AxisLen Y = q2_axis_len((s >> 1) & 1, cy, (int)g.n[1]);
    const Q2AxisLen Z = q2_axis_len((s >> 2) & 1, cl, (int)g.n[2]);
    const unsigned inner = rg.Tx * (unsigned)Y.PL + (unsigned)Y.L * (unsigned)X.PL;
    return rg.TxTy * Z.PL + (long long)((unsigned long long)(unsigned)Z.L * inner);
  }
  const Q2AxisLen Y = q2_axis_len((s >> 1) & 1, cl, (int)g.n[1]);
  return (long long)rg.Tx * Y.PL + (long long)((unsigned long long)(unsigned)Y.L * (unsigned)X.PL);
}

// position of component comp (K: 0..4, M: 5..9) of the lattice point p = 2 c + S in the component-major x table of the
// sum-factorised kernels (assemble_q2_gather.cu, q2_load_x)
__host__ __device__ __forceinline__ long long q2_x_index(const int S, const int comp, const int c, const long long Nx)
{
  return (long long)(S * 10 + comp) * (Nx + 1) + c;
}

// per-axis description of a row's coupling box for parity S (compile time): the row's lattice coordinate is
// p = 2 c + S; box offsets a = 0 .. A-1 mean q = p - R + a with R = S ? 1 : 2, A = S ? 3 : 5
template <int S>
struct AxisBox
{
  static constexpr int R = S ? 1 : 2;
  static constexpr int A = S ? 3 : 5;
  static constexpr int NE = S ? 1 : 2; // elements along this axis that contain p
  // parity of q for offset a (p = S mod 2)
  __host__ __device__ static constexpr int parity(int a)
  {
    return (S + R + a) & 1;
  }
  // element candidate o: lattice offset of its first node relative to the box start: S: 0; !S: 2 o
  __host__ __device__ static constexpr int first(int o)
  {
    return S ? 0 : 2 * o;
  }
  // local 1D index (0, 1, 2 = left, middle, right node) of p in element candidate o
  __host__ __device__ static constexpr int local(int o)
  {
    return S ? 1 : 2 - 2 * o;
  }
};

struct AxisRuntime
{
  int idx[5];    // index of q among the box values of its own parity
  bool valid[5]; // q inside the lattice
  int n[2];      // number of even / odd values in the (clipped) box
  double ha[2], hb[2]; // h and 1/h of the candidate elements (0 if outside the grid)
  int e[2];      // element index of the candidates
};

template <int S>
__device__ __forceinline__ void axis_setup(AxisRuntime& ax, int c, int N, double lo, double h)
{
  using B = AxisBox<S>;
  const int p = 2 * c + S;
  const int blo = max(0, p - B::R), bhi = min(2 * N, p + B::R);
  // counts of even / odd values in [blo, bhi]
  const int first_even = blo + (blo & 1), first_odd = blo + 1 - (blo & 1);
  ax.n[0] = first_even <= bhi ? (bhi - first_even) / 2 + 1 : 0;
  ax.n[1] = first_odd <= bhi ? (bhi - first_odd) / 2 + 1 : 0;
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    ax.idx[a] = 0;
    ax.valid[a] = false;
    if (a < B::A) {
      const int q = p - B::R + a;
      ax.valid[a] = q >= 0 && q <= 2 * N;
      ax.idx[a] = (q - (B::parity(a) ? first_odd : first_even)) >> 1;
    }
  }
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    ax.e[o] = 0;
    ax.ha[o] = ax.hb[o] = 0.;
    if (o < B::NE) {
      const int e = S ? c : c - 1 + o;
      ax.e[o] = e;
      if (e >= 0 && e < N) {
        const double ext = q2_cell_extent(lo, h, e);
        ax.ha[o] = ext;
        ax.hb[o] = __drcp_rn(ext);
      }
    }
  }
}

// group order of the column groups = ascending global index: codim ascending, shift bitset ascending
__device__ __forceinline__ constexpr int q2_group_order(int D, int rank)
{
  return D == 3 ? (rank == 0 ? 7 : rank == 1 ? 3 : rank == 2 ? 5 : rank == 3 ? 6 : rank == 4 ? 1 : rank == 5 ? 2
                                                                                        : rank == 6 ? 4
                                                                                                    : 0)
                : (rank == 0 ? 3 : rank == 1 ? 1 : rank == 2 ? 2 : 0);
}

} // namespace
} // namespace gdtb

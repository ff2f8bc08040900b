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

// entries along x of the rows before row c of a lattice line (q2_axis_len(S, c, N).PL as an int)
__device__ __forceinline__ int q2_xpl(const int S, const int c)
{
  return S ? 3 * c : 5 * c - (c > 0 ? 2 : 0);
}

// start of the lattice line (cy, cl) of a row group relative to the group's first value, and the entries w a row of the
// line holds per entry along x: the row (cx, cy, cl) starts at line + w * q2_xpl(cx)  (q2_row_offset, regrouped)
template <int D>
__device__ __forceinline__ void q2_line(const GridDev& g, const Q2RowGroup& rg, const int cy, const int cl, long long& line,
                                        int& w)
{
  const int s = rg.s;
  if (D == 3) {
    const Q2AxisLen Y = q2_axis_len((s >> 1) & 1, cy, (int)g.n[1]);
    const Q2AxisLen Z = q2_axis_len((s >> 2) & 1, cl, (int)g.n[2]);
    w = Y.L * Z.L;
    line = rg.TxTy * Z.PL + (long long)((unsigned long long)(unsigned)Z.L * (rg.Tx * (unsigned)Y.PL));
  } else {
    const Q2AxisLen Y = q2_axis_len((s >> 1) & 1, cl, (int)g.n[1]);
    w = Y.L;
    line = (long long)rg.Tx * Y.PL;
  }
}

// number of box offsets a in [lo, hi] with a & 1 == par
__device__ __forceinline__ int q2_count_par(int lo, int hi, int par)
{
  const int first = lo + ((lo ^ par) & 1);
  return first <= hi ? ((hi - first) >> 1) + 1 : 0;
}

// writes the (a_y, a_x) plane of a row into its CSR positions (closed forms; interior rows: compile-time offsets)
template <int D, int SX, int SY, int SL>
__device__ __forceinline__ void q2_scatter_plane(const GridDev& g, const int px, const int py, const int pl, const int slot,
                                                 const double (&acc)[D == 3 ? AxisBox<SY>::A : 1][AxisBox<SX>::A],
                                                 double* __restrict__ row)
{
  using BX = AxisBox<SX>;
  using BY = AxisBox<SY>;
  using BL = AxisBox<SL>;
  constexpr int last = D - 1;
  const int Nx = (int)g.n[0], Ny = D == 3 ? (int)g.n[1] : 1, Nl = (int)g.n[last];
  constexpr int AX = BX::A, AY = D == 3 ? BY::A : 1, AL = BL::A;
  // parity of the lattice point at box offset a is a & 1 (S + R = 2 for both parities); the column groups of a row
  // come in ascending global index (codim ascending, shift ascending), each lexicographic with x fastest
  const bool interior = px >= BX::R && px + BX::R <= 2 * Nx && (D == 2 || (py >= BY::R && py + BY::R <= 2 * Ny))
                        && pl >= BL::R && pl + BL::R <= 2 * Nl;
  const int parl = slot & 1;
  if (interior) {
    // unclipped box: every count and group start is a compile-time number, only the plane is a run-time choice
    constexpr int NX[2] = {(AX + 1) / 2, AX / 2}, NY[2] = {D == 3 ? (AY + 1) / 2 : 1, D == 3 ? AY / 2 : 0};
    constexpr int NL[2] = {(AL + 1) / 2, AL / 2};
    int start_[2][1 << (D - 1)]; // [plane parity][sx | sy << 1]
    {
      int running = 0;
#pragma unroll
      for (int r = 0; r < (1 << D); ++r) {
        const int s = q2_group_order(D, r);
        const int sx = s & 1, sy = D == 3 ? (s >> 1) & 1 : 0, sl = (s >> (D - 1)) & 1;
        start_[sl][s & ((1 << (D - 1)) - 1)] = running;
        running += NX[sx] * (D == 3 ? NY[sy] : 1) * NL[sl];
      }
    }
    const int idx_l = slot >> 1;
    double* plane[1 << (D - 1)];
#pragma unroll
    for (int sxy = 0; sxy < (1 << (D - 1)); ++sxy) {
      const int sx = sxy & 1, sy = D == 3 ? sxy >> 1 : 0;
      plane[sxy] = row + (parl ? start_[1][sxy] : start_[0][sxy]) + idx_l * (NX[sx] * (D == 3 ? NY[sy] : 1));
    }
#pragma unroll
    for (int a = 0; a < AY; ++a)
#pragma unroll
      for (int b = 0; b < AX; ++b)
        plane[(b & 1) | (D == 3 ? (a & 1) << 1 : 0)][(a >> 1) * NX[b & 1] + (b >> 1)] = acc[a][b];
    return;
  }

  // clipped box at the grid boundary: offsets [lo, hi] are inside the lattice
  const int xlo = max(0, BX::R - px), xhi = BX::A - 1 - max(0, px + BX::R - 2 * Nx);
  const int ylo = D == 3 ? max(0, BY::R - py) : 0, yhi = D == 3 ? BY::A - 1 - max(0, py + BY::R - 2 * Ny) : 0;
  const int llo = max(0, BL::R - pl), lhi = BL::A - 1 - max(0, pl + BL::R - 2 * Nl);
  const int nx0 = q2_count_par(xlo, xhi, 0), nx1 = q2_count_par(xlo, xhi, 1);
  const int ny0 = D == 3 ? q2_count_par(ylo, yhi, 0) : 1, ny1 = D == 3 ? q2_count_par(ylo, yhi, 1) : 0;
  const int nl0 = q2_count_par(llo, lhi, 0), nl1 = q2_count_par(llo, lhi, 1);
  const int idx_l = (slot - llo) >> 1;
  int base[1 << (D - 1)];
  {
    int running = 0;
#pragma unroll
    for (int r = 0; r < (1 << D); ++r) {
      const int s = q2_group_order(D, r);
      const int sx = s & 1, sy = D == 3 ? (s >> 1) & 1 : 0, sl = (s >> (D - 1)) & 1;
      if (sl == parl)
        base[s & ((1 << (D - 1)) - 1)] = running;
      running += (sx ? nx1 : nx0) * (D == 3 ? (sy ? ny1 : ny0) : 1) * (sl ? nl1 : nl0);
    }
  }
  // entries of equal parity along x are consecutive, idx_x(b) = (b >> 1) - shift_par
  const int shx0 = (xlo + 1) >> 1, shx1 = xlo >> 1;
#pragma unroll
  for (int a = 0; a < AY; ++a) {
    if (D == 3 && (a < ylo || a > yhi))
      continue;
    const int pary = a & 1;
    const int line = D == 3 ? idx_l * (pary ? ny1 : ny0) + ((a - ylo) >> 1) : idx_l;
    double* r0 = row + base[D == 3 ? (pary << 1) : 0] + line * nx0 - shx0;
    double* r1 = row + base[D == 3 ? (1 | (pary << 1)) : 1] + line * nx1 - shx1;
#pragma unroll
    for (int b = 0; b < AX; ++b) {
      if (b < xlo || b > xhi)
        continue;
      if (b & 1)
        r1[b >> 1] = acc[a][b];
      else
        r0[b >> 1] = acc[a][b];
    }
  }
}

} // namespace
} // namespace gdtb

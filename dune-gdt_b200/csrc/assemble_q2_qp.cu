// dune-gdt_b200/csrc/assemble_q2_qp.cu -- continuous-Lagrange Q2 row gather in 3D for ONE integrand whose coefficient
// varies inside the cells: LocalLaplaceIntegrand with kappa(x) I / LocalElementProductIntegrand with w(x)
// (local/integrands/laplace.hh:81-102, product.hh:104-130) inside LocalElementIntegralBilinearForm::apply2
// (local/bilinear-forms/integrals.hh:97-134), the coefficient given as one value per quadrature point of the form's rule
// (GDTB_FN_QP_SCALAR, or any grid function sampled by k_sample_function).
//
// Work decomposition ("x-fused").  A lattice line (p_y, p_l) of Q2 DoFs alternates vertex-type (p_x = 2 c) and mid-type
// (p_x = 2 c + 1) points; in the MCMG numbering they belong to two row groups (parity patterns s and s | 1), each
// contiguous along x.  One warp handles 31 consecutive x-elements c of one line and ONE lattice plane of the last axis
// (the warps of a CTA are the 5 / 3 planes of the coupling box); lane c evaluates, for every element (c, e_y, e_l) that
// touches the line, the rows i_x = 0, 1, 2 of the local matrix restricted to that plane:
//   stage l:  A^t[q_y][q_x]   = sum_ql kappa(q) PT^t[q_l][i_l][j_l]                      t in {MM, KK}   (shared by the 3 rows)
//   stage y:  B^c[j_y][q_x]   = sum_qy A^t[q_y][q_x] PT^t'[q_y][i_y][j_y]                3 term combinations (shared)
//   stage x:  L[i_x][j_y][j_x] = sum_qx (w_x B^1 PT^KK + (w_y B^2 + w_l B^3) PT^MM)[q_x][i_x][j_x]
// i.e. 108 FMAs per row and plane where the per-row kernel (assemble_q2_gather.cu, SF = 3) spends 252.  Rows i_x = 0, 1
// are the lane's own vertex / mid row; row i_x = 2 belongs to the vertex 2 c + 2 and travels to lane c + 1 by a warp
// shuffle (lane 0 of every warp is the halo element below the chunk).  Every matrix entry is still written exactly
// once, in a fixed summation order (deterministic, no atomics); the two CSR segments of a chunk (vertex rows, mid rows)
// are staged in shared memory and leave the SM as TMA bulk stores.
//
// Coefficient stream.  The <= 4 element lines a chunk touches are contiguous runs of the [element][q] array: one thread
// issues one TMA bulk load per line into shared memory (cp.async.bulk.shared::cluster.global + mbarrier complete_tx),
// all warps read the samples from there.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "kernels.hpp"
#include "q2_layout.cuh"

namespace gdtb {

namespace {

constexpr int XF_ELEMS = 31;   // x-elements per chunk (lane 0 is the halo element below them)
constexpr int XF_WARPS = 5;    // planes of the coupling box of a vertex-type last-axis coordinate
constexpr int XF_THREADS = 32 * XF_WARPS;
constexpr int XF_SEG0_MAX = XF_ELEMS * 125 + 2; // vertex rows: up to 5^3 entries
constexpr int XF_SEG1_MAX = XF_ELEMS * 75 + 2;  // mid rows: up to 3 * 5^2 entries
constexpr int XF_STAGE0 = (XF_SEG0_MAX + 1) & ~1;
constexpr int XF_STAGE = XF_STAGE0 + ((XF_SEG1_MAX + 1) & ~1);

struct XfParams
{
  GridDev g;
  CgQpGroup G;
  Q2RowGroup rg[8];        // by parity pattern s
  long long item_begin[5]; // items of the line kinds k = s_y | s_l << 1, prefix sums
  int lines_y[4];
  long long cl_lo[4], cl_hi[4]; // owned lattice layers c_l along the last axis for the line kind
  int chunks;
  unsigned long long m_chunks, m_lines[4]; // division magics floor(2^64 / d) + 1 of chunks / lines_y (q2_div)
  int line_cap;      // doubles of one staged element line
  int use_tma;       // coefficient array 16-byte aligned
  long long coef_e_end; // one past the last element index the coefficient array holds
  long long coef_e_begin;
};

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra DONE_%=;\n"
               "bra WAIT_%=;\n"
               "DONE_%=:\n"
               "}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void bulk_load_g2s(double* sdst, const double* gsrc, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// h and 1 / h of the element candidates of the lattice coordinate 2 c + S along one axis (axis_setup without the CSR part)
template <int S>
__device__ __forceinline__ void xf_axis_geometry(const int c, const int N, const double lo, const double h, double (&ha)[2],
                                                 double (&hb)[2])
{
  using B = AxisBox<S>;
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    ha[o] = hb[o] = 0.;
    if (o < B::NE) {
      const int e = S ? c : c - 1 + o;
      if (e >= 0 && e < N) {
        const double ext = q2_cell_extent(lo, h, e);
        ha[o] = ext;
        hb[o] = __drcp_rn(ext);
      }
    }
  }
}

template <int M, int KIND, int SY, int SL>
__device__ __forceinline__ void xf_line(const XfParams& p, const int cy, const int cl, const int c0, const int slot,
                                        const double* __restrict__ coef_s, const int (&line_shift)[4],
                                        double* __restrict__ stage0, double* __restrict__ stage1, const int w_line)
{
  using BY = AxisBox<SY>;
  using BL = AxisBox<SL>;
  constexpr int NQ = M * M * M;
  constexpr int AYN = BY::A;
  const GridDev& g = p.g;
  const CgQpGroup& G = p.G;
  const int Nx = (int)g.n[0], Ny = (int)g.n[1], Nl = (int)g.n[2];
  const int lane = threadIdx.x & 31;
  const int c = c0 - 1 + lane;
  if (slot >= BL::A)
    return;
  // geometry of the element candidates along y and the last axis (h and 1 / h, zero outside the grid); the plane must lie
  // inside the lattice
  double ay_ha[2], ay_hb[2], al_ha[2], al_hb[2];
  xf_axis_geometry<SY>(cy, Ny, g.lo[1], g.h[1], ay_ha, ay_hb);
  xf_axis_geometry<SL>(cl, Nl, g.lo[2], g.h[2], al_ha, al_hb);
  const int ql = 2 * cl + SL - BL::R + slot;
  const bool valid_l = ql >= 0 && ql <= 2 * Nl;
  if (!valid_l)
    return;
  const bool elem_ok = c >= 0 && c < Nx;
  double hax = 0., hbx = 0.;
  if (elem_ok) {
    hax = q2_cell_extent(g.lo[0], g.h[0], c);
    hbx = __drcp_rn(hax);
  }
  double accV[AYN][5], accM[AYN][3];
#pragma unroll
  for (int a = 0; a < AYN; ++a) {
#pragma unroll
    for (int b = 0; b < 5; ++b)
      accV[a][b] = 0.;
#pragma unroll
    for (int b = 0; b < 3; ++b)
      accM[a][b] = 0.;
  }

#pragma unroll
  for (int ol = 0; ol < BL::NE; ++ol) {
    const int jl = slot - BL::first(ol);
    if (jl < 0 || jl > 2)
      continue; // the element candidate does not contain the plane (uniform over the warp)
    const int il = BL::local(ol);
    const double hal = al_ha[ol], hbl = al_hb[ol];
    if (hal == 0.)
      continue; // outside the grid
    double plM[M], plK[M];
#pragma unroll
    for (int q = 0; q < M; ++q) {
      plM[q] = jl == 0 ? G.pt[QPT_MM][q][il][0] : (jl == 1 ? G.pt[QPT_MM][q][il][1] : G.pt[QPT_MM][q][il][2]);
      plK[q] = jl == 0 ? G.pt[QPT_KK][q][il][0] : (jl == 1 ? G.pt[QPT_KK][q][il][1] : G.pt[QPT_KK][q][il][2]);
    }
#pragma unroll
    for (int oy = 0; oy < BY::NE; ++oy) {
      const double hay = ay_ha[oy], hby = ay_hb[oy];
      if (hay == 0.)
        continue;
      const int iy = BY::local(oy);
      double send[3][3];
#pragma unroll
      for (int jy = 0; jy < 3; ++jy)
#pragma unroll
        for (int jx = 0; jx < 3; ++jx)
          send[jy][jx] = 0.;
      if (elem_ok) {
        const double* kq = coef_s + (ol * BY::NE + oy) * p.line_cap + (lane + line_shift[ol * BY::NE + oy]) * NQ;
        const double ie = hax * hay * hal; // integrals.hh:119
        // ---- stage l ----------------------------------------------------------------------------------------
        double AM[M][M], AK[M][M];
#pragma unroll
        for (int qy = 0; qy < M; ++qy)
#pragma unroll
          for (int qx = 0; qx < M; ++qx) {
            double am = 0., ak = 0.;
#pragma unroll
            for (int ql = 0; ql < M; ++ql) {
              const double k = kq[qx + M * (qy + M * ql)];
              am = fma(k, plM[ql], am);
              if (KIND != Q1G_MASS)
                ak = fma(k, plK[ql], ak);
            }
            AM[qy][qx] = am;
            AK[qy][qx] = ak;
          }
        // ---- stage y: BK is contracted with PT^KK along x, BM with PT^MM ------------------------------------------
        const double wx = G.scale * (ie * (hbx * hbx)), wy = G.scale * (ie * (hby * hby)), wl = G.scale * (ie * (hbl * hbl));
        double BK[3][M], BM[3][M];
#pragma unroll
        for (int jy = 0; jy < 3; ++jy)
#pragma unroll
          for (int qx = 0; qx < M; ++qx) {
            double b1 = 0., b2 = 0., b3 = 0.;
#pragma unroll
            for (int qy = 0; qy < M; ++qy) {
              b1 = fma(AM[qy][qx], G.pt[QPT_MM][qy][iy][jy], b1);
              if (KIND != Q1G_MASS) {
                b2 = fma(AM[qy][qx], G.pt[QPT_KK][qy][iy][jy], b2);
                b3 = fma(AK[qy][qx], G.pt[QPT_MM][qy][iy][jy], b3);
              }
            }
            if (KIND == Q1G_MASS) {
              BM[jy][qx] = (G.scale * ie) * b1;
              BK[jy][qx] = 0.;
            } else {
              BK[jy][qx] = wx * b1;
              BM[jy][qx] = fma(wy, b2, wl * b3);
            }
          }
        // ---- stage x: the three rows i_x of the element ----------------------------------------------------------
#pragma unroll
        for (int ix = 0; ix < 3; ++ix)
#pragma unroll
          for (int jy = 0; jy < 3; ++jy)
#pragma unroll
            for (int jx = 0; jx < 3; ++jx) {
              double v = 0.;
#pragma unroll
              for (int qx = 0; qx < M; ++qx) {
                v = fma(BM[jy][qx], G.pt[QPT_MM][qx][ix][jx], v);
                if (KIND != Q1G_MASS)
                  v = fma(BK[jy][qx], G.pt[QPT_KK][qx][ix][jx], v);
              }
              if (ix == 0)
                accV[BY::first(oy) + jy][2 + jx] += v;
              else if (ix == 1)
                accM[BY::first(oy) + jy][jx] += v;
              else
                send[jy][jx] = v;
            }
      }
      // row i_x = 2 of element c is a row of the vertex 2 c + 2: hand it to the lane above
#pragma unroll
      for (int jy = 0; jy < 3; ++jy)
#pragma unroll
        for (int jx = 0; jx < 3; ++jx) {
          const double r = __shfl_up_sync(0xffffffffu, send[jy][jx], 1);
          if (lane > 0)
            accV[BY::first(oy) + jy][jx] += r;
        }
    }
  }

  // ---- the lane's rows: vertex 2 c (c <= N_x) and mid point 2 c + 1 (c < N_x); lane 0 is the halo ----------------
  if (lane == 0 || c > Nx)
    return;
  // scatter through the closed-form positions shared with the per-row kernels (interior rows: compile-time offsets)
  {
    double* row = stage0 + w_line * (q2_xpl(0, c) - q2_xpl(0, c0));
    q2_scatter_plane<3, 0, SY, SL>(g, 2 * c, 2 * cy + SY, 2 * cl + SL, slot, accV, row);
  }
  if (c < Nx) {
    double* row = stage1 + w_line * (3 * (c - c0));
    q2_scatter_plane<3, 1, SY, SL>(g, 2 * c + 1, 2 * cy + SY, 2 * cl + SL, slot, accM, row);
  }
}

// register budget: 2 blocks x 5 warps per SM.  __launch_bounds__(160, 2) makes ptxas stop at 168 registers (it rounds
// the block up to 6 warps); XF_MAXNREG states the budget directly (200 x 160 x 2 = 64 000 registers).
#ifdef XF_MAXNREG
#define XF_KERNEL_ATTR __maxnreg__(XF_MAXNREG)
#else
#define XF_KERNEL_ATTR __launch_bounds__(XF_THREADS, 2)
#endif

template <int M, int KIND, bool ACCUMULATE>
__global__ void XF_KERNEL_ATTR
    k_q2_qp_xfused(const __grid_constant__ XfParams p, double* __restrict__ values)
{
  constexpr int NQ = M * M * M;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) unsigned long long bar;
  double* stage = smem;
  double* coef_s = smem + XF_STAGE;
  const GridDev& g = p.g;
  const int Nx = (int)g.n[0], Ny = (int)g.n[1], Nl = (int)g.n[2];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned parity = 0;
  for (long long item = blockIdx.x; item < p.item_begin[4]; item += gridDim.x) {
    // ---- decode (uniform): line kind, line, chunk ----------------------------------------------------------
    int k = 0;
#pragma unroll
    for (int j = 1; j < 4; ++j)
      if (item >= p.item_begin[j])
        k = j;
    const int sy = k & 1, sl = k >> 1;
    // (32-bit divisions by run-time constants through their magics: the item count is checked at launch)
    const unsigned t0 = (unsigned)(item - p.item_begin[k]);
    const unsigned t1 = q2_div(t0, (unsigned)p.chunks, p.m_chunks);
    const int chunk = int(t0 - t1 * (unsigned)p.chunks);
    const unsigned t2 = q2_div(t1, (unsigned)p.lines_y[k], p.m_lines[k]);
    const int cy = int(t1 - t2 * (unsigned)p.lines_y[k]);
    const int cl = int(p.cl_lo[k] + t2);
    const int c0 = chunk * XF_ELEMS;
    // ---- coefficient stream: the element lines (o_y, o_l) of the chunk, elements c0 - 1 .. c0 + 30 ---------------
    const int ney = sy ? 1 : 2, nel = sl ? 1 : 2;
    const int cb = max(c0 - 1, 0), ce = min(c0 + XF_ELEMS, Nx);
    int line_shift[4] = {0, 0, 0, 0}; // element c of line L sits at slot (lane + line_shift[L]) of its staged line
    bool tma[4] = {false, false, false, false};
    long long e_first[4] = {0, 0, 0, 0};
    unsigned total_bytes = 0;
#pragma unroll
    for (int L = 0; L < 4; ++L) {
      const int ol = L / ney, oy = L - ol * ney;
      if (L >= ney * nel)
        continue;
      const int ey = sy ? cy : cy - 1 + oy, el = sl ? cl : cl - 1 + ol;
      if (ey < 0 || ey >= Ny || el < 0 || el >= Nl)
        continue;
      const long long ef = ((long long)el * Ny + ey) * Nx + cb;
      const long long e_al = ef & ~1LL;
      long long cnt = (ce - cb) + (ef - e_al);
      cnt += cnt & 1;
      e_first[L] = ef;
      // lane of element c: c - c0 + 1; staged slot: c - cb + (ef - e_al)
      line_shift[L] = int(ef - e_al) - (cb - (c0 - 1));
      tma[L] = p.use_tma && e_al >= p.coef_e_begin && e_al + cnt <= p.coef_e_end;
      if (tma[L])
        total_bytes += (unsigned)(cnt * NQ * sizeof(double));
    }
    if (threadIdx.x == 0 && total_bytes > 0) {
      mbar_expect_tx(&bar, total_bytes);
#pragma unroll
      for (int L = 0; L < 4; ++L)
        if (tma[L]) {
          const long long e_al = e_first[L] & ~1LL;
          long long cnt = (ce - cb) + (e_first[L] - e_al);
          cnt += cnt & 1;
          bulk_load_g2s(coef_s + L * p.line_cap, p.G.coef + e_al * NQ, (unsigned)(cnt * NQ * sizeof(double)), &bar);
        }
    }
    // lines that cannot take the bulk path (unaligned array, array end): plain coalesced loads
#pragma unroll
    for (int L = 0; L < 4; ++L) {
      const int ol = L / ney, oy = L - ol * ney;
      if (L >= ney * nel || tma[L])
        continue;
      const int ey = sy ? cy : cy - 1 + oy, el = sl ? cl : cl - 1 + ol;
      if (ey < 0 || ey >= Ny || el < 0 || el >= Nl)
        continue;
      const int shift = int(e_first[L] - (e_first[L] & ~1LL));
      const double* src = p.G.coef + e_first[L] * NQ;
      double* dst = coef_s + L * p.line_cap + shift * NQ;
      for (int i = threadIdx.x; i < (ce - cb) * NQ; i += blockDim.x)
        dst[i] = __ldg(src + i);
    }
    // (the CSR segments of the chunk are worked out HERE, while the bulk loads issued above are in flight)
    const int s0 = (sy << 1) | (sl << 2);
    const Q2RowGroup& rg0 = p.rg[s0];
    const Q2RowGroup& rg1 = p.rg[s0 | 1];
    // the two CSR segments of the chunk: vertex rows c0 .. c0 + n0 - 1, mid rows c0 .. c0 + n1 - 1
    const int n0 = min(XF_ELEMS, Nx + 1 - c0), n1 = min(XF_ELEMS, Nx - c0);
    // the vertex rows and the mid rows of the chunk are consecutive along x on the line (c_y, c_l) of their groups: a row
    // starts at line + w * q2_xpl(c), w = L_y L_l entries per entry along x (q2_line); a segment that ends at the line end
    // ends where the next line starts
    long long line0, line1;
    int w0, w1;
    q2_line<3>(g, rg0, cy, cl, line0, w0);
    q2_line<3>(g, rg1, cy, cl, line1, w1);
    const long long off0a = line0 + w0 * q2_xpl(0, c0);
    const long long off0b = c0 + n0 < (int)rg0.ex ? line0 + w0 * q2_xpl(0, c0 + n0) : line0 + (long long)w0 * (int)rg0.Tx;
    long long off1a = 0, off1b = 0;
    if (n1 > 0) {
      off1a = line1 + w1 * q2_xpl(1, c0);
      off1b = c0 + n1 < (int)rg1.ex ? line1 + w1 * q2_xpl(1, c0 + n1) : line1 + (long long)w1 * (int)rg1.Tx;
    }
    const long long start0 = rg0.value_begin + off0a, start1 = rg1.value_begin + off1a;
    const int seg0 = int(off0b - off0a), seg1 = int(off1b - off1a);
    const int phase0 = int((reinterpret_cast<unsigned long long>(values + start0) >> 3) & 1ULL);
    const int phase1 = int((reinterpret_cast<unsigned long long>(values + start1) >> 3) & 1ULL);
    double* stage0 = stage + phase0;
    double* stage1 = stage + XF_STAGE0 + phase1;

    if (total_bytes > 0) {
      mbar_wait(&bar, parity);
      parity ^= 1;
    }
    // the output stage must have been read out by the previous item's bulk store: waited for here, after this item's
    // index arithmetic and coefficient loads, so that the drain overlaps them
    if (!ACCUMULATE && item != (long long)blockIdx.x && threadIdx.x == 0)
      q2_bulk_wait_read0();
    __syncthreads();

    switch (k) {
      case 0: xf_line<M, KIND, 0, 0>(p, cy, cl, c0, warp, coef_s, line_shift, stage0, stage1, w0); break;
      case 1: xf_line<M, KIND, 1, 0>(p, cy, cl, c0, warp, coef_s, line_shift, stage0, stage1, w0); break;
      case 2: xf_line<M, KIND, 0, 1>(p, cy, cl, c0, warp, coef_s, line_shift, stage0, stage1, w0); break;
      default: xf_line<M, KIND, 1, 1>(p, cy, cl, c0, warp, coef_s, line_shift, stage0, stage1, w0); break;
    }

    if (ACCUMULATE) {
      __syncthreads();
      for (int i = threadIdx.x; i < seg0; i += blockDim.x)
        values[start0 + i] += stage0[i];
      for (int i = threadIdx.x; i < seg1; i += blockDim.x)
        values[start1 + i] += stage1[i];
      __syncthreads();
    } else {
      q2_fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        {
          const int head = phase0, body = (seg0 - head) & ~1;
          if (head)
            values[start0] = stage0[0];
          if (body > 0)
            q2_bulk_store_s2g(values + start0 + head, stage0 + head, (unsigned)(body * sizeof(double)));
          if (head + body < seg0)
            values[start0 + head + body] = stage0[head + body];
        }
        if (seg1 > 0) {
          const int head = phase1, body = (seg1 - head) & ~1;
          if (head)
            values[start1] = stage1[0];
          if (body > 0)
            q2_bulk_store_s2g(values + start1 + head, stage1 + head, (unsigned)(body * sizeof(double)));
          if (head + body < seg1)
            values[start1 + head + body] = stage1[head + body];
        }
        q2_bulk_commit();
      }
    }
  }
  if (!ACCUMULATE && threadIdx.x == 0)
    q2_bulk_wait0();
}

using XfKernel = void (*)(const XfParams, double*);

template <int M>
XfKernel xf_kernel_m(int kind, bool accumulate)
{
  if (kind == Q1G_MASS)
    return accumulate ? k_q2_qp_xfused<M, Q1G_MASS, true> : k_q2_qp_xfused<M, Q1G_MASS, false>;
  return accumulate ? k_q2_qp_xfused<M, Q1G_LAPLACE_SCALAR, true> : k_q2_qp_xfused<M, Q1G_LAPLACE_SCALAR, false>;
}

} // namespace

bool q2_qp_xfused_supported(int d, int m, int kind)
{
  return d == 3 && (m == 2 || m == 3) && (kind == Q1G_LAPLACE_SCALAR || kind == Q1G_MASS)
         && !std::getenv("GDTB_Q2_QP_NO_XFUSED");
}

int launch_q2_qp_xfused(Launch& L, const GridDev& g, const CgQpGroup& group, const SpaceDev& sp, long long coef_e_begin,
                        long long coef_e_end, double* values, bool accumulate)
{
  if (!q2_qp_xfused_supported(g.d, group.m, group.kind))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "q2_qp_xfused: 3D, 2 or 3 Gauss points per direction, scalar coefficients");
  if (sp.size >= (1LL << 31))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "q2_gather: more than 2^31 degrees of freedom");
  for (int k = 0; k < 3; ++k)
    if (g.n[k] > 5000)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "q2_gather: more than 5000 elements along one axis");
  XfParams p;
  std::memset(&p, 0, sizeof(p));
  p.g = g;
  p.G = group;
  Q2SlabRange ranges[8];
  q2_slab_ranges(g, sp, ranges);
  {
    int r = 0;
    for (int c = 0; c <= 3; ++c)
      for (int s = 0; s < 8; ++s) {
        int pc = 0;
        for (int k = 0; k < 3; ++k)
          pc += (s >> k) & 1;
        if (pc != 3 - c)
          continue;
        Q2RowGroup& rg = p.rg[s];
        rg.s = s;
        rg.rows = 1;
        for (int k = 0; k < 3; ++k)
          rg.rows *= ((s >> k) & 1) ? g.n[k] : g.n[k] + 1;
        rg.row_begin = sp.cg.codim_offset[c] + sp.cg.group_offset[s];
        rg.ex = (unsigned)((s & 1) ? g.n[0] : g.n[0] + 1);
        rg.ey = (unsigned)((s & 2) ? g.n[1] : g.n[1] + 1);
        rg.mex = rg.ex > 1 ? ~0ULL / rg.ex + 1 : 0;
        rg.mey = rg.ey > 1 ? ~0ULL / rg.ey + 1 : 0;
        rg.Tx = (unsigned)q2_axis_total(s & 1, g.n[0]);
        rg.TxTy = (long long)rg.Tx * q2_axis_total((s >> 1) & 1, g.n[1]);
        const long long per_layer = (long long)rg.ex * rg.ey;
        rg.lex_begin = ranges[r].row_begin - rg.row_begin;
        rg.lex_end = ranges[r].row_end - rg.row_begin;
        const int SL = (s >> 2) & 1;
        const long long off_begin = rg.TxTy * q2_axis_len(SL, (int)(rg.lex_begin / per_layer), (int)g.n[2]).PL;
        rg.off_end = off_begin + ranges[r].count;
        rg.value_begin = ranges[r].local_offset - off_begin;
        ++r;
      }
  }
  p.chunks = (int)((g.n[0] + 1 + XF_ELEMS - 1) / XF_ELEMS);
  long long items = 0;
  for (int k = 0; k < 4; ++k) {
    const int sy = k & 1, sl = k >> 1;
    const Q2RowGroup& rg = p.rg[(sy << 1) | (sl << 2)];
    const long long per_layer = (long long)rg.ex * rg.ey;
    p.lines_y[k] = (int)(sy ? g.n[1] : g.n[1] + 1);
    p.cl_lo[k] = rg.lex_begin / per_layer;
    p.cl_hi[k] = rg.lex_end / per_layer;
    p.item_begin[k] = items;
    items += (p.cl_hi[k] - p.cl_lo[k]) * p.lines_y[k] * p.chunks;
  }
  p.item_begin[4] = items;
  if (items <= 0)
    return GDTB_OK;
  if (items >= (1LL << 31))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "q2_qp_xfused: more than 2^31 work items");
  p.m_chunks = p.chunks > 1 ? ~0ULL / (unsigned long long)p.chunks + 1 : 0;
  for (int k = 0; k < 4; ++k)
    p.m_lines[k] = p.lines_y[k] > 1 ? ~0ULL / (unsigned long long)p.lines_y[k] + 1 : 0;
  const int nq = group.m * group.m * group.m;
  p.line_cap = (XF_ELEMS + 3) * nq;
  p.line_cap += p.line_cap & 1;
  p.use_tma = (reinterpret_cast<uintptr_t>(group.coef) & 15) == 0 && !std::getenv("GDTB_Q2_QP_NO_TMA");
  p.coef_e_begin = coef_e_begin;
  p.coef_e_end = coef_e_end;
  const size_t smem = sizeof(double) * ((size_t)XF_STAGE + 4 * (size_t)p.line_cap);
  XfKernel kern = group.m == 2 ? xf_kernel_m<2>(group.kind, accumulate) : xf_kernel_m<3>(group.kind, accumulate);
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // two resident blocks need 2 x (smem + 1 KB) of the 228 KB: the rest stays L1 (register spills of the unrolled
  // arithmetic and the table loads live there) instead of the maximal shared-memory carve-out
  static const int carve_env = std::getenv("GDTB_XF_CARVEOUT") ? std::atoi(std::getenv("GDTB_XF_CARVEOUT")) : 0;
  int carve = carve_env > 0 ? carve_env : (int)((2 * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 1;
  carve = std::min(100, std::max(carve, 1));
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  int per_sm = 0;
  GDTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, XF_THREADS, smem));
  if (per_sm < 1)
    return fail(GDTB_ERR_CUDA, "q2_qp_xfused: kernel does not fit on an SM");
  const long long grid = std::min<long long>((long long)per_sm * L.sm_count, items);
  note_kernel(L, KF_Q2_GATHER, reinterpret_cast<const void*>(kern));
  time_begin(L, KF_Q2_GATHER);
  kern<<<(unsigned)grid, XF_THREADS, smem, L.stream>>>(p, values);
  time_end(L, KF_Q2_GATHER);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

} // namespace gdtb

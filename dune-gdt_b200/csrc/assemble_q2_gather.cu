// dune-gdt_b200/csrc/assemble_q2_gather.cu -- owner-computes-rows ("row gather") assembly for continuous-Lagrange
// Q2 spaces on axis-aligned structured grids (2D / 3D), element forms with element-wise constant coefficients.
//
// Replaces the same chain as assemble_q1_gather.cu (LocalElementBilinearFormAssembler::apply_local,
// local/assembler/bilinear-form-assemblers.hh:110-128; LocalElementIntegralBilinearForm::apply2,
// local/bilinear-forms/integrals.hh:97-134; LocalLaplaceIntegrand / LocalElementProductIntegrand::evaluate,
// local/integrands/laplace.hh:81-102, product.hh:104-130; add_to_entry [EXT]) for order 2.
//
// Lattice view.  The Q2 DoFs of an N_x x N_y x N_z cube grid are the points p of the lattice [0, 2 N_k]^d; p_k odd
// means "inside an element along axis k".  The parity pattern s(p) is the YaspGrid shift bitset of the sub-entity the
// DoF sits on, and the MCMG-based ContinuousMapper (spaces/mapper/continuous.hh:117-150) numbers the DoFs
// [cells | faces | edges | vertices] (codim ascending), each codim by ascending shift bitset, each group
// lexicographically (x fastest) -- common.cuh::cg_global_index.  Row p couples to the lattice box |q_k - p_k| <= 1
// (p_k odd) or <= 2 (p_k even), clipped to the grid: 27 / 45 / 75 / 125 entries for cell / face / edge / vertex rows.
// Sorted by global index the entries of a row are grouped by the parity pattern of q in the same group order and are
// lexicographic inside a group, so the CSR position of every entry is a closed form: colidx is never read, rowptr only
// once per work item.
//
// Sum factorisation.  On an affine axis-aligned cell with an element-wise constant coefficient the quadrature sum of
// the reference factorises exactly into 1D reference tables (the form's own Gauss rule):
//   L_e[i][j] = c_e sum_r |det J_e| / h_r^2  prod_k T^{(r,k)}[i_k][j_k],  T^{(r,k)} = K1 (k == r) or M1 (k != r),
//   K1[a][b] = sum_q w_q phi_a'(x_q) phi_b'(x_q),  M1[a][b] = sum_q w_q phi_a(x_q) phi_b(x_q)   (3 x 3 each)
// (mass: c_e |det J_e| prod_k M1).  A thread owns one row p and one value q_last of the last axis ("plane"): it sums,
// over the <= 2^d elements that contain both p and the plane, the products built up axis by axis (x innermost) into
// <= 5 x 5 accumulators -- no atomics, fixed order, every value written once.  Geometry (h_k = upper - lower,
// 1 / h_k, |det J|) and coefficients are evaluated per element on the device in FP64; elements outside the grid get
// zero factors.  Rows of one work item are consecutive in CSR: the segment is staged in shared memory and leaves the
// SM as one TMA bulk store, double-buffered, persistent CTAs (as in assemble_q1_gather.cu).
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "kernels.hpp"
#include "q2_layout.cuh"

// Single-group sum-factorised kernel: resident blocks per SM the register allocation aims for.  Measured on C5
// (ms per assembly): 2 blocks x 2 stages 1.90, 3 blocks x 1 stage 1.85, 4 blocks (64 registers, 150 B of spills) x 1 stage
// 1.76.  The other variants are register-bound at 2 blocks and keep two stages.
#ifndef Q2G_MIN_BLOCKS
#define Q2G_MIN_BLOCKS 3
#endif

#ifndef Q2G_PE_MIN_BLOCKS
#define Q2G_PE_MIN_BLOCKS 2
#endif
#ifndef Q2G_SLOT_MAJOR
#define Q2G_SLOT_MAJOR 1
#endif

namespace gdtb {

namespace {

// ---- coefficients per quadrature point (kernels.hpp, CgQpGroup): sum factorisation over the tensor rule -------------
// A[qy][qx] = sum_ql kq[qx, qy, ql] * pl[ql]: the plane's factor PT[t_l][ql][i_l][j_l] contracted first
template <int D, int M>
__device__ __forceinline__ void q2qp_contract_last(const double (&kq)[D == 3 ? M * M * M : M * M], const double (&pl)[M],
                                                   double (&A)[D == 3 ? M : 1][M])
{
#pragma unroll
  for (int qy = 0; qy < (D == 3 ? M : 1); ++qy)
#pragma unroll
    for (int qx = 0; qx < M; ++qx) {
      double a = 0.;
#pragma unroll
      for (int ql = 0; ql < M; ++ql)
        a = fma(kq[D == 3 ? qx + M * (qy + M * ql) : qx + M * ql], pl[ql], a);
      A[qy][qx] = a;
    }
}

// blk[jy][jx] += w * sum_{qy, qx} A[qy][qx] PT[ty][qy][iy][jy] PT[tx][qx][ix][jx]
template <int D, int M>
__device__ __forceinline__ void q2qp_plane_term(const CgQpGroup& G, const double (&A)[D == 3 ? M : 1][M], const int ty,
                                                const int tx, const int iy, const int ix, const double w,
                                                double (&blk)[D == 3 ? 3 : 1][3])
{
  if constexpr (D == 3) {
    double B[3][M]; // [jy][qx]
#pragma unroll
    for (int jy = 0; jy < 3; ++jy)
#pragma unroll
      for (int qx = 0; qx < M; ++qx) {
        double b = 0.;
#pragma unroll
        for (int qy = 0; qy < M; ++qy)
          b = fma(A[qy][qx], G.pt[ty][qy][iy][jy], b);
        B[jy][qx] = b;
      }
#pragma unroll
    for (int jy = 0; jy < 3; ++jy)
#pragma unroll
      for (int jx = 0; jx < 3; ++jx) {
        double c = 0.;
#pragma unroll
        for (int qx = 0; qx < M; ++qx)
          c = fma(B[jy][qx], G.pt[tx][qx][ix][jx], c);
        blk[jy][jx] = fma(w, c, blk[jy][jx]);
      }
  } else {
#pragma unroll
    for (int jx = 0; jx < 3; ++jx) {
      double c = 0.;
#pragma unroll
      for (int qx = 0; qx < M; ++qx)
        c = fma(A[0][qx], G.pt[tx][qx][ix][jx], c);
      blk[0][jx] = fma(w, c, blk[0][jx]);
    }
  }
}

// the factors PT[t][ql][il][jl] of the plane (jl is a run-time value of the thread: select, do not index)
template <int M>
__device__ __forceinline__ void q2qp_plane_factors(const CgQpGroup& G, const int t, const int il, const int jl, double (&pl)[M])
{
#pragma unroll
  for (int ql = 0; ql < M; ++ql)
    pl[ql] = jl == 0 ? G.pt[t][ql][il][0] : (jl == 1 ? G.pt[t][ql][il][1] : G.pt[t][ql][il][2]);
}

// One (row, plane): D == 3: SX, SY in-thread axes, SL = parity of the last (plane) axis; D == 2: SX in-thread, SL = y.
// M == 0: element-wise constant coefficients (1D reference tables); M > 0: one integrand of kind KIND with a coefficient
// value / tensor per quadrature point, M Gauss points per direction.
template <int D, int SX, int SY, int SL, int M = 0, int KIND = 0>
__device__ __forceinline__ void q2_row_plane(const Q2GatherParams& p, const int cx, const int cy, const int cl,
                                             const int slot, double* __restrict__ row)
{
  using BX = AxisBox<SX>;
  using BY = AxisBox<SY>;
  using BL = AxisBox<SL>;
  const GridDev& g = p.g;
  constexpr int last = D - 1;
  const int Nx = (int)g.n[0], Ny = D == 3 ? (int)g.n[1] : 1, Nl = (int)g.n[last];

  AxisRuntime ax, ay, al;
  axis_setup<SX>(ax, cx, Nx, g.lo[0], g.h[0]);
  if (D == 3)
    axis_setup<SY>(ay, cy, Ny, g.lo[1], g.h[1]);
  else {
    ay.n[0] = 1;
    ay.n[1] = 0;
    ay.idx[0] = 0;
    ay.valid[0] = true;
    ay.ha[0] = ay.hb[0] = 1.;
    ay.e[0] = 0;
  }
  axis_setup<SL>(al, cl, Nl, g.lo[last], g.h[last]);
  // the plane: offset `slot` along the last axis
  if (slot >= BL::A)
    return;
  int pl = 0, idx_l = 0;
  bool valid_l = false;
#pragma unroll
  for (int a = 0; a < BL::A; ++a)
    if (a == slot) {
      pl = BL::parity(a);
      idx_l = al.idx[a];
      valid_l = al.valid[a];
    }
  if (!valid_l)
    return;
  // start of each column group inside the row: groups in ascending global index order, sizes prod_k n_k[parity];
  // bit layout of a parity pattern: bit 0 = x, bit 1 = y (3D) / last (2D), bit 2 = last (3D).  Only the groups
  // with the plane's parity along the last axis are needed: base[px | py << 1]
  int base[1 << (D - 1)];
  {
    int all[1 << D];
    int running = 0;
#pragma unroll
    for (int r = 0; r < (1 << D); ++r) {
      const int s = q2_group_order(D, r);
      all[s] = running;
      const int sx = s & 1, sy = D == 3 ? (s >> 1) & 1 : 0, sl = (s >> (D - 1)) & 1;
      running += ax.n[sx] * (D == 3 ? ay.n[sy] : 1) * al.n[sl];
    }
#pragma unroll
    for (int sxy = 0; sxy < (1 << (D - 1)); ++sxy)
      base[sxy] = pl ? all[sxy | (1 << (D - 1))] : all[sxy];
  }

  constexpr int AY = D == 3 ? BY::A : 1, NEY = D == 3 ? BY::NE : 1;
  double acc[AY][BX::A];
#pragma unroll
  for (int a = 0; a < AY; ++a)
#pragma unroll
    for (int b = 0; b < BX::A; ++b)
      acc[a][b] = 0.;

#pragma unroll
  for (int ol = 0; ol < BL::NE; ++ol) {
    // does element candidate ol (along the last axis) contain the plane?  local index of the plane in it
    const int jl = slot - BL::first(ol);
    if (jl < 0 || jl > 2)
      continue;
    const int il = BL::local(ol);
#pragma unroll
    for (int oy = 0; oy < NEY; ++oy)
#pragma unroll
      for (int ox = 0; ox < BX::NE; ++ox) {
        const double hax = ax.ha[ox], hbx = ax.hb[ox];
        const double hay = D == 3 ? ay.ha[oy] : 1., hby = D == 3 ? ay.hb[oy] : 1.;
        const double hal = al.ha[ol], hbl = al.hb[ol];
        const double ie = hax * hay * hal; // |det J_e| (0 for an element outside the grid)
        const bool valid = ie != 0.;
        const long long e = (long long)ax.e[ox] + (long long)Nx * ((D == 3 ? ay.e[oy] : al.e[ol]) + (D == 3 ? (long long)Ny * al.e[ol] : 0));
        const int ix = BX::local(ox), iy = D == 3 ? BY::local(oy) : 0;
        if constexpr (M > 0) {
          if (!valid)
            continue;
          const CgQpGroup& G = p.qp;
          constexpr int NQ = D == 3 ? M * M * M : M * M;
          double blk[D == 3 ? 3 : 1][3];
#pragma unroll
          for (int jy = 0; jy < (D == 3 ? 3 : 1); ++jy)
#pragma unroll
            for (int jx = 0; jx < 3; ++jx)
              blk[jy][jx] = 0.;
          double kq[NQ], pl[M], A[D == 3 ? M : 1][M];
          const double hb_[3] = {hbx, D == 3 ? hby : hbl, hbl}; // 1 / h along x, y (2D: last), last
          if constexpr (KIND == Q1G_LAPLACE_TENSOR) {
            const double* src = G.coef + e * (long long)(NQ * D * D);
#pragma unroll 1
            for (int rc = 0; rc < D * D; ++rc) {
              const int r = rc / D, c = rc - r * D;
#pragma unroll
              for (int q = 0; q < NQ; ++q)
                kq[q] = __ldg(src + q * (D * D) + rc);
              // axis k takes the test derivative if k == r and the ansatz derivative if k == c (laplace.hh:98-101)
              const int tx = 0 == r ? (0 == c ? QPT_KK : QPT_KM) : (0 == c ? QPT_MK : QPT_MM);
              const int ty = 1 == r ? (1 == c ? QPT_KK : QPT_KM) : (1 == c ? QPT_MK : QPT_MM);
              const int tl = last == r ? (last == c ? QPT_KK : QPT_KM) : (last == c ? QPT_MK : QPT_MM);
              q2qp_plane_factors<M>(G, tl, il, jl, pl);
              q2qp_contract_last<D, M>(kq, pl, A);
              const double br = r == 0 ? hb_[0] : (r == 1 ? hb_[1] : hb_[2]), bc = c == 0 ? hb_[0] : (c == 1 ? hb_[1] : hb_[2]);
              q2qp_plane_term<D, M>(G, A, ty, tx, iy, ix, G.scale * (ie * (br * bc)), blk);
            }
          } else {
            const double* src = G.coef + e * (long long)NQ;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
              kq[q] = __ldg(src + q);
            q2qp_plane_factors<M>(G, QPT_MM, il, jl, pl);
            q2qp_contract_last<D, M>(kq, pl, A);
            if constexpr (KIND == Q1G_MASS)
              q2qp_plane_term<D, M>(G, A, QPT_MM, QPT_MM, iy, ix, G.scale * ie, blk);
            else {
              // kappa = c(x) I: r = x and (3D) r = y share the mass factor along the last axis, r = last has its own
              q2qp_plane_term<D, M>(G, A, QPT_MM, QPT_KK, iy, ix, G.scale * (ie * (hbx * hbx)), blk);
              if constexpr (D == 3)
                q2qp_plane_term<D, M>(G, A, QPT_KK, QPT_MM, iy, ix, G.scale * (ie * (hby * hby)), blk);
              q2qp_plane_factors<M>(G, QPT_KK, il, jl, pl);
              q2qp_contract_last<D, M>(kq, pl, A);
              q2qp_plane_term<D, M>(G, A, QPT_MM, QPT_MM, iy, ix, G.scale * (ie * (hbl * hbl)), blk);
            }
          }
#pragma unroll
          for (int jy = 0; jy < (D == 3 ? 3 : 1); ++jy)
#pragma unroll
            for (int jx = 0; jx < 3; ++jx)
              acc[D == 3 ? BY::first(oy) + jy : 0][BX::first(ox) + jx] += blk[jy][jx];
          continue;
        }
#pragma unroll 1
        for (int gi = 0; gi < p.n_groups; ++gi) {
          const Q2Group& G = p.group[gi];
          double cf = G.scale;
          if (G.coef_elem)
            cf *= valid ? __ldg(G.coef + e) : 0.;
          if (G.kind == Q1G_MASS) {
            // c |det J| M1 x M1 x M1
            const double cl_ = cf * ie * G.TM[il][jl];
#pragma unroll
            for (int jy = 0; jy < (D == 3 ? 3 : 1); ++jy) {
              const double cyv = D == 3 ? cl_ * G.TM[iy][jy] : cl_;
#pragma unroll
              for (int jx = 0; jx < 3; ++jx)
                acc[D == 3 ? BY::first(oy) + jy : 0][BX::first(ox) + jx] =
                    fma(cyv, G.TM[ix][jx], acc[D == 3 ? BY::first(oy) + jy : 0][BX::first(ox) + jx]);
            }
          } else {
            // Laplace with kappa = c I: sum_r c |det J| / h_r^2 (K1 along r, M1 along the other axes)
#pragma unroll
            for (int r = 0; r < D; ++r) {
              double w;
              if (r == 0)
                w = cf * (hbx * hay * hal);
              else if (r == last)
                w = cf * (hax * hay * hbl);
              else
                w = cf * (hax * hby * hal);
              const double cl_ = w * (r == last ? G.TK[il][jl] : G.TM[il][jl]);
#pragma unroll
              for (int jy = 0; jy < (D == 3 ? 3 : 1); ++jy) {
                const double cyv = D == 3 ? cl_ * ((D == 3 && r == 1) ? G.TK[iy][jy] : G.TM[iy][jy]) : cl_;
#pragma unroll
                for (int jx = 0; jx < 3; ++jx)
                  acc[D == 3 ? BY::first(oy) + jy : 0][BX::first(ox) + jx] =
                      fma(cyv, r == 0 ? G.TK[ix][jx] : G.TM[ix][jx], acc[D == 3 ? BY::first(oy) + jy : 0][BX::first(ox) + jx]);
              }
            }
          }
        }
      }
  }

  // scatter the plane into the row (CSR order by closed-form position)
#pragma unroll
  for (int a = 0; a < AY; ++a) {
    if (D == 3 && !ay.valid[a])
      continue;
    const int py = D == 3 ? BY::parity(a) : 0;
#pragma unroll
    for (int b = 0; b < BX::A; ++b) {
      if (!ax.valid[b])
        continue;
      const int px = BX::parity(b);
      int pos;
      if (D == 3)
        pos = base[px | (py << 1)] + (idx_l * ay.n[py] + ay.idx[a]) * ax.n[px] + ax.idx[b];
      else
        pos = base[px] + idx_l * ax.n[px] + ax.idx[b];
      row[pos] = acc[a][b];
    }
  }
}

// ---- sum-factorised variant (every coefficient constant) ------------------------------------------------------
// With c_e = c the sum over the elements around p factorises too: along every axis the <= 2 elements that contain p
// contribute the 1D row vectors K[a] = sum_e K1[i_e][a - first_e] / h_e and M[a] = sum_e M1[i_e][a - first_e] h_e
// (a = box offset), and the row is   c sum_r prod_k (k == r ? K^(k) : M^(k))   (mass: c prod_k M^(k)).  The vectors
// depend on one lattice coordinate only; k_q2_axis_tables evaluates them once per launch (geometry per element in
// FP64 as above), the row kernel loads 22 numbers and spends 2 FMA-class instructions per matrix entry:
//   row[a_l][a_y][a_x] = MY[a_y] * P[a_x] + Q[a_y] * MX[a_x],  P = c (ML KX + KL MX),  Q = c ML KY.
__global__ void __launch_bounds__(128) k_q2_axis_tables(const __grid_constant__ Q2GatherParams p, double* __restrict__ tab)
{
  const GridDev& g = p.g;
  const int D = g.d;
  long long total = 0;
  for (int k = 0; k < D; ++k)
    total += 2 * g.n[k] + 1;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total * p.n_groups;
       t += (long long)gridDim.x * blockDim.x) {
    const int gi = int(t / total);
    long long r = t - gi * total;
    int k = 0;
    while (r >= 2 * g.n[k] + 1) {
      r -= 2 * g.n[k] + 1;
      ++k;
    }
    const int pt = (int)r, S = pt & 1, c = pt >> 1, N = (int)g.n[k];
    const Q2Group& G = p.group[gi];
    double K[5] = {0., 0., 0., 0., 0.}, M[5] = {0., 0., 0., 0., 0.};
    for (int o = 0; o < (S ? 1 : 2); ++o) {
      const int e = S ? c : c - 1 + o;
      if (e < 0 || e >= N)
        continue;
      const double ext = q2_cell_extent(g.lo[k], g.h[k], e);
      const double inv = __drcp_rn(ext);
      const int first = S ? 0 : 2 * o, il = S ? 1 : 2 - 2 * o;
      for (int j = 0; j < 3; ++j) {
        K[first + j] = fma(inv, G.TK[il][j], K[first + j]);
        M[first + j] = fma(ext, G.TM[il][j], M[first + j]);
      }
    }
    if (k == 0) { // component-major along x (q2_x_index)
      double* out = tab + gi * p.sf_group_stride + p.sf_axis_off[0];
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        out[q2_x_index(S, a, c, g.n[0])] = K[a];
        out[q2_x_index(S, 5 + a, c, g.n[0])] = M[a];
      }
    } else {
      double* out = tab + gi * p.sf_group_stride + p.sf_axis_off[k] + 10LL * pt;
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        out[a] = K[a];
        out[5 + a] = M[a];
      }
    }
  }
}

// Per-ELEMENT 1D factors for a single integrand whose coefficient is one value per element (Q2GatherParams::sf == 2): for
// the lattice point p = 2 c + S of axis k and its element candidate o (S = 1: the element c; S = 0: c - 1 and c) the row
// vectors K_o[j] = K1[i_o][j] / h_e, M_o[j] = M1[i_o][j] h_e (j = local column index in that element, zero for an
// element outside the grid): 12 doubles per point, [o][K0 K1 K2 M0 M1 M2]; x component-major (q2_xpe_index).
__host__ __device__ __forceinline__ long long q2_xpe_index(const int S, const int comp, const int c, const long long Nx)
{
  return (long long)(S * 12 + comp) * (Nx + 1) + c;
}

template <int D>
__global__ void __launch_bounds__(128) k_q2_axis_tables_pe(const __grid_constant__ Q2GatherParams p, double* __restrict__ tab)
{
  const GridDev& g = p.g;
  long long total = 0;
  for (int k = 0; k < D; ++k)
    total += 2 * g.n[k] + 1;
  const Q2Group& G = p.group[0];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    int k = 0;
    while (r >= 2 * g.n[k] + 1) {
      r -= 2 * g.n[k] + 1;
      ++k;
    }
    const int pt = (int)r, S = pt & 1, c = pt >> 1, N = (int)g.n[k];
    double v[12];
#pragma unroll
    for (int i = 0; i < 12; ++i)
      v[i] = 0.;
    for (int o = 0; o < (S ? 1 : 2); ++o) {
      const int e = S ? c : c - 1 + o;
      if (e < 0 || e >= N)
        continue;
      const double ext = q2_cell_extent(g.lo[k], g.h[k], e);
      const double inv = __drcp_rn(ext);
      const int il = S ? 1 : 2 - 2 * o;
      for (int j = 0; j < 3; ++j) {
        v[6 * o + j] = inv * G.TK[il][j];
        v[6 * o + 3 + j] = ext * G.TM[il][j];
      }
    }
    if (k == 0) {
      double* out = tab + p.sf_axis_off[0];
#pragma unroll
      for (int i = 0; i < 12; ++i)
        out[q2_xpe_index(S, i, c, g.n[0])] = v[i];
    } else {
      double* out = tab + p.sf_axis_off[k] + 12LL * pt;
#pragma unroll
      for (int i = 0; i < 12; ++i)
        out[i] = v[i];
    }
  }
}

template <int A>
__device__ __forceinline__ void q2_load_axis(const double* __restrict__ t, double (&K)[A], double (&M)[A])
{
  // 10 doubles per lattice point, 16-byte aligned pairs: K0 K1 | K2 K3 | K4 M0 | M1 M2 | M3 M4
  const double2* q = reinterpret_cast<const double2*>(t);
  const double2 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = __ldg(q + 3);
  K[0] = v0.x;
  K[1] = v0.y;
  K[2] = v1.x;
  M[0] = v2.y;
  M[1] = v3.x;
  M[2] = v3.y;
  if (A == 5) {
    const double2 v4 = __ldg(q + 4);
    K[A - 2] = v1.y;
    K[A - 1] = v2.x;
    M[A - 2] = v4.x;
    M[A - 1] = v4.y;
  }
}

// The x axis keeps its table component-major ("structure of arrays"): tab_x[(S * 10 + comp) * (N_x + 1) + c] for the
// lattice point p = 2 c + S.  With the slot-major thread layout the lanes of a warp hold consecutive c, so each of the
// 2 A loads of a warp is one contiguous run (2 shared-memory/L1 wavefronts) instead of 32 entries 80 bytes apart
// (round-2 profile: the strided x loads were two thirds of the LSU wavefronts and the top stall, long scoreboard 6.2).
template <int A>
__device__ __forceinline__ void q2_load_x(const double* __restrict__ t, const int S, const int c, const long long Nx,
                                          double (&K)[A], double (&M)[A])
{
  const double* q = t + q2_x_index(S, 0, c, Nx);
  const long long stride = Nx + 1;
#pragma unroll
  for (int a = 0; a < A; ++a) {
    K[a] = __ldg(q + a * stride);
    M[a] = __ldg(q + (5 + a) * stride);
  }
}

// one group's contribution to the (a_y, a_x) plane of a row; FIRST: assign instead of accumulate
template <int D, int AX, int AY, bool FIRST>
__device__ __forceinline__ void q2_sf_group(const Q2GatherParams& p, const int gi, const int px, const int py, const int pl,
                                            const int slot, double (&acc)[AY][AX])
{
  constexpr int last = D - 1;
  const Q2Group& G = p.group[gi];
  const double* tab = p.sf_tab + gi * p.sf_group_stride;
  double KX[AX], MX[AX], KY[AY], MY[AY];
  q2_load_x<AX>(tab + p.sf_axis_off[0], px & 1, px >> 1, p.g.n[0], KX, MX);
  if (D == 3)
    q2_load_axis<AY>(tab + p.sf_axis_off[1] + 10LL * py, KY, MY);
  const double* tl = tab + p.sf_axis_off[last] + 10LL * pl;
  const double ML = G.scale * __ldg(tl + 5 + slot);
  if (G.kind == Q1G_MASS) {
#pragma unroll
    for (int a = 0; a < AY; ++a) {
      const double m = D == 3 ? ML * MY[a] : ML;
#pragma unroll
      for (int b = 0; b < AX; ++b)
        acc[a][b] = FIRST ? m * MX[b] : fma(m, MX[b], acc[a][b]);
    }
  } else {
    const double KL = G.scale * __ldg(tl + slot);
    double P[AX];
#pragma unroll
    for (int b = 0; b < AX; ++b)
      P[b] = fma(ML, KX[b], KL * MX[b]);
#pragma unroll
    for (int a = 0; a < AY; ++a) {
      const double q = D == 3 ? ML * KY[a] : 0.;
#pragma unroll
      for (int b = 0; b < AX; ++b) {
        if (D == 3)
          acc[a][b] = FIRST ? fma(MY[a], P[b], q * MX[b]) : fma(MY[a], P[b], fma(q, MX[b], acc[a][b]));
        else
          acc[a][b] = FIRST ? P[b] : acc[a][b] + P[b];
      }
    }
  }
}

template <int D, int SX, int SY, int SL, bool NG1>
__device__ __forceinline__ void q2_row_plane_sf(const Q2GatherParams& p, const int cx, const int cy, const int cl,
                                                const int slot, double* __restrict__ row)
{
  using BX = AxisBox<SX>;
  using BY = AxisBox<SY>;
  using BL = AxisBox<SL>;
  const GridDev& g = p.g;
  constexpr int last = D - 1;
  if (slot >= BL::A)
    return;
  const int Nl = (int)g.n[last];
  const int px = 2 * cx + SX, py = 2 * cy + SY, pl = 2 * cl + SL;
  constexpr int AX = BX::A, AY = D == 3 ? BY::A : 1;
  // the plane must lie inside the lattice
  const int ql = pl - BL::R + slot;
  if (ql < 0 || ql > 2 * Nl)
    return;

  double acc[AY][AX];
  q2_sf_group<D, AX, AY, true>(p, 0, px, py, pl, slot, acc);
  if (!NG1) { // one group (the common case) lets the compiler sink the arithmetic to the stores: fewer live registers
#pragma unroll 1
    for (int gi = 1; gi < p.n_groups; ++gi)
      q2_sf_group<D, AX, AY, false>(p, gi, px, py, pl, slot, acc);
  }

  q2_scatter_plane<D, SX, SY, SL>(g, px, py, pl, slot, acc, row);
}

// Single integrand with ONE coefficient value per element (sf == 2): the plane of a row is
//   sum_{o_l, o_y} [ (ML MY[j_y]) PX + (ML KY[j_y] + KL MY[j_y]) QX ],  PX = sum_{o_x} kappa_e KX_o, QX = sum_{o_x} kappa_e MX_o
// (mass: (ML MY[j_y]) QX) over the element candidates that contain both the row's lattice point and the plane -- the
// per-element factors come from k_q2_axis_tables_pe, the coefficient of an element is read once per (o_x, o_y, o_l);
// about 100-200 FMAs per thread instead of the ~1 400 instructions of the generic per-element loop (q2_row_plane)
template <int D, int SX, int SY, int SL>
__device__ __forceinline__ void q2_row_plane_pe(const Q2GatherParams& p, const int cx, const int cy, const int cl,
                                                const int slot, double* __restrict__ row)
{
  using BX = AxisBox<SX>;
  using BY = AxisBox<SY>;
  using BL = AxisBox<SL>;
  const GridDev& g = p.g;
  constexpr int last = D - 1;
  if (slot >= BL::A)
    return;
  const int Nx = (int)g.n[0], Ny = D == 3 ? (int)g.n[1] : 1, Nl = (int)g.n[last];
  const int px = 2 * cx + SX, py = 2 * cy + SY, pl = 2 * cl + SL;
  constexpr int AX = BX::A, AY = D == 3 ? BY::A : 1, NEX = BX::NE, NEY = D == 3 ? BY::NE : 1;
  const int ql = pl - BL::R + slot;
  if (ql < 0 || ql > 2 * Nl)
    return;
  const Q2Group& G = p.group[0];
  const double* tab = p.sf_tab;
  const bool mass = G.kind == Q1G_MASS;

  double KX[NEX][3], MX[NEX][3], KY[NEY][3], MY[NEY][3];
  {
    const double* q = tab + p.sf_axis_off[0] + q2_xpe_index(SX, 0, cx, Nx);
    const long long stride = Nx + 1;
#pragma unroll
    for (int o = 0; o < NEX; ++o)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        KX[o][j] = __ldg(q + (6 * o + j) * stride);
        MX[o][j] = __ldg(q + (6 * o + 3 + j) * stride);
      }
  }
  if (D == 3) {
    const double* q = tab + p.sf_axis_off[1] + 12LL * py;
#pragma unroll
    for (int o = 0; o < NEY; ++o)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        KY[o][j] = __ldg(q + 6 * o + j);
        MY[o][j] = __ldg(q + 6 * o + 3 + j);
      }
  }
  // the coefficients of the <= 8 elements first: all requests are in flight before the first one is needed
  double kap[BL::NE][NEY][NEX];
#pragma unroll
  for (int ol = 0; ol < BL::NE; ++ol) {
    const int jl = slot - BL::first(ol);
    const int el = SL ? cl : cl - 1 + ol;
    const bool vl = jl >= 0 && jl <= 2 && el >= 0 && el < Nl;
#pragma unroll
    for (int oy = 0; oy < NEY; ++oy) {
      const int ey = D == 3 ? (SY ? cy : cy - 1 + oy) : 0;
      const bool vy = D < 3 || (ey >= 0 && ey < Ny);
#pragma unroll
      for (int ox = 0; ox < NEX; ++ox) {
        const int ex = SX ? cx : cx - 1 + ox;
        const bool valid = vl && vy && ex >= 0 && ex < Nx;
        const long long e = (long long)ex + (long long)Nx * (D == 3 ? ey + (long long)Ny * el : el);
        kap[ol][oy][ox] = valid ? __ldg(G.coef + e) : 0.;
      }
    }
  }
  double acc[AY][AX];
#pragma unroll
  for (int a = 0; a < AY; ++a)
#pragma unroll
    for (int b = 0; b < AX; ++b)
      acc[a][b] = 0.;

#pragma unroll
  for (int ol = 0; ol < BL::NE; ++ol) {
    const int jl = slot - BL::first(ol);
    if (jl < 0 || jl > 2)
      continue; // the element candidate does not contain the plane
    const int el = SL ? cl : cl - 1 + ol;
    if (el < 0 || el >= Nl)
      continue;
    const double* tl = tab + p.sf_axis_off[last] + 12LL * pl + 6 * ol;
    const double KL = G.scale * __ldg(tl + jl), ML = G.scale * __ldg(tl + 3 + jl);
#pragma unroll
    for (int oy = 0; oy < NEY; ++oy) {
      double PX[AX], QX[AX];
#pragma unroll
      for (int b = 0; b < AX; ++b)
        PX[b] = QX[b] = 0.;
#pragma unroll
      for (int ox = 0; ox < NEX; ++ox) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          PX[BX::first(ox) + j] = fma(kap[ol][oy][ox], KX[ox][j], PX[BX::first(ox) + j]);
          QX[BX::first(ox) + j] = fma(kap[ol][oy][ox], MX[ox][j], QX[BX::first(ox) + j]);
        }
      }
#pragma unroll
      for (int jy = 0; jy < (D == 3 ? 3 : 1); ++jy) {
        const int a = D == 3 ? BY::first(oy) + jy : 0;
        const double my = D == 3 ? MY[oy][jy] : 1., ky = D == 3 ? KY[oy][jy] : 0.;
        if (mass) {
          const double c = ML * my;
#pragma unroll
          for (int b = 0; b < AX; ++b)
            acc[a][b] = fma(c, QX[b], acc[a][b]);
        } else {
          const double cA = ML * my, cB = fma(ML, ky, KL * my);
#pragma unroll
          for (int b = 0; b < AX; ++b)
            acc[a][b] = fma(cA, PX[b], fma(cB, QX[b], acc[a][b]));
        }
      }
    }
  }
  q2_scatter_plane<D, SX, SY, SL>(g, px, py, pl, slot, acc, row);
}

template <int SF, int D, int SX, int SY, int SL, int M = 0, int KIND = 0>
__device__ __forceinline__ void q2_dispatch(const Q2GatherParams& p, const int cx, const int cy, const int cl,
                                            const int slot, double* __restrict__ row)
{
  if constexpr (SF == 3)
    q2_row_plane<D, SX, SY, SL, M, KIND>(p, cx, cy, cl, slot, row);
  else if (SF == 4)
    q2_row_plane_pe<D, SX, SY, SL>(p, cx, cy, cl, slot, row);
  else if (SF == 2)
    q2_row_plane_sf<D, SX, SY, SL, true>(p, cx, cy, cl, slot, row);
  else if (SF == 1)
    q2_row_plane_sf<D, SX, SY, SL, false>(p, cx, cy, cl, slot, row);
  else
    q2_row_plane<D, SX, SY, SL>(p, cx, cy, cl, slot, row);
}

static_assert(Q2G_THREADS / 3 < 128, "the work-item records pack the rows of an item into 7 bits");
// work-item records (Q2GatherParams::items): what the LN bookkeeping of k_q2_gather computes per item, once per grid / slab
template <int D>
__global__ void __launch_bounds__(128) k_q2_items(const __grid_constant__ Q2GatherParams p, int4* __restrict__ recs)
{
  const GridDev& g = p.g;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= p.n_items)
    return;
  int gi = 0;
  for (int k = 1; k < p.n_rowgroups; ++k)
    if (item >= p.rg[k].item_begin)
      gi = k;
  const Q2RowGroup& rg = p.rg[gi];
  const int s = rg.s;
  const bool five = !((s >> (D - 1)) & 1);
  const int rpi = five ? Q2G_THREADS / 5 : Q2G_THREADS / 3;
  const unsigned lrow0 = unsigned(rg.lex_begin) + unsigned(item - rg.item_begin) * rpi;
  const int nrows = (int)min((long long)rpi, rg.lex_end - lrow0);
  int ux, uy, ul;
  q2_decode<D>(rg, lrow0, ux, uy, ul);
  const int ex = (int)rg.ex;
  const int n0 = min(nrows, ex - ux);
  int y1 = uy + 1, l1 = ul, wrap = 0;
  if (D == 2 || y1 == (int)rg.ey) {
    y1 = 0;
    l1 = ul + 1;
    wrap = 1;
  }
  long long line0, line1;
  int w0, w1;
  q2_line<D>(g, rg, uy, ul, line0, w0);
  q2_line<D>(g, rg, y1, l1, line1, w1);
  const int sx = s & 1;
  const long long off0 = line0 + w0 * q2_xpl(sx, ux);
  const int rest = nrows - n0;
  const long long off1 = rest > 0           ? line1 + w1 * q2_xpl(sx, rest)
                         : ux + nrows < ex ? line0 + w0 * q2_xpl(sx, ux + nrows)
                                           : line0 + (long long)w0 * (int)rg.Tx;
  const long long start = rg.value_begin + off0;
  int4 a, b;
  a.x = (int)(unsigned)(start & 0xffffffffLL);
  a.y = (int)(unsigned)((unsigned long long)start >> 32);
  a.z = int(off1 - off0);
  a.w = s | (nrows << 3) | (n0 << 10) | (w0 << 17) | ((rest > 0 ? w1 : 0) << 22) | (wrap << 27);
  b.x = ux;
  b.y = uy;
  b.z = ul;
  b.w = rest > 0 ? int(line1 - off0) : 0;
  recs[2 * item] = a;
  recs[2 * item + 1] = b;
}

// SF: 0 per-element coefficients, 1 sum-factorised (constant coefficients), 2 sum-factorised with a single group,
// 3 one integrand of kind KIND with a coefficient per quadrature point (M Gauss points per direction), 4 one integrand
// with one coefficient value per element (per-element 1D factor tables, q2_row_plane_pe)
// LN: every lattice line of every row group is at least as long as a work item (checked at launch), so an item touches
// at most two lines: row starts come from two uniform line records + one multiply-add per thread instead of a decode
// (two divisions) and a full q2_row_offset per thread, and the end of the segment from the same records.
// LN == 2: the uniform part (row group, first row, segment, line records) comes from the work-item records of k_q2_items.
template <int D, bool ACCUMULATE, int SF, int M = 0, int KIND = 0, int LN = 0>
__global__ void __launch_bounds__(Q2G_THREADS, SF == 2 ? Q2G_MIN_BLOCKS : (SF == 4 ? Q2G_PE_MIN_BLOCKS : (SF == 3 ? 1 : 2)))
    k_q2_gather(const __grid_constant__ Q2GatherParams p, double* __restrict__ values, int stage_doubles, int nbuf)
{
  extern __shared__ __align__(16) double smem[];
  const GridDev& g = p.g;
  int buf = 0;
  int gi = 0;
#if Q2G_SLOT_MAJOR
  // slot-major: the lanes of a warp hold CONSECUTIVE ROWS of one plane, so one store instruction of the warp hits rows
  // whose starts differ by the (odd) row length -- 16 distinct 8-byte banks per half warp instead of the 5-slots-per-row
  // pattern (2.1 x the minimal shared-memory wavefronts in the round-2 profile); surplus threads (slot >= slots) idle
  const int slot5 = threadIdx.x / (Q2G_THREADS / 5), lr5 = threadIdx.x - slot5 * (Q2G_THREADS / 5);
  const int slot3 = threadIdx.x / (Q2G_THREADS / 3), lr3 = threadIdx.x - slot3 * (Q2G_THREADS / 3);
#else
  const int lr5 = threadIdx.x / 5, slot5 = threadIdx.x - 5 * lr5;
  const int lr3 = threadIdx.x / 3, slot3 = threadIdx.x - 3 * lr3;
#endif
  // LN == 2: the record of the NEXT item is requested one iteration ahead
  const int4* const recs = reinterpret_cast<const int4*>(p.items);
  int4 nra = make_int4(0, 0, 0, 0), nrb = nra;
  if (LN == 2 && (long long)blockIdx.x < p.n_items) {
    nra = q2_ldg_int4_here(recs + 2 * blockIdx.x);
    nrb = q2_ldg_int4_here(recs + 2 * blockIdx.x + 1);
  }
  for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    // ---- per item (uniform): row group, rows, CSR segment ---------------------------------------------------
    int s, nrows, ux, uy, ul, n0 = 0, y1 = 0, l1 = 0;
    int cx = 0, cy = 0, cl = 0, row_off = 0;
    long long start;
    int seg;
    bool five;
    int lr, slot;
    bool active;
    if (LN == 2) {
      const int4 ra = nra, rb = nrb;
      if (item + gridDim.x < p.n_items) {
        nra = q2_ldg_int4_here(recs + 2 * (item + gridDim.x));
        nrb = q2_ldg_int4_here(recs + 2 * (item + gridDim.x) + 1);
      }
      start = (long long)(((unsigned long long)(unsigned)ra.y << 32) | (unsigned long long)(unsigned)ra.x);
      seg = ra.z;
      const int pk = ra.w; // s | nrows << 3 | n0 << 10 | w0 << 17 | w1 << 22 | wrap << 27
      s = pk & 7;
      nrows = (pk >> 3) & 127;
      n0 = (pk >> 10) & 127;
      const int w0 = (pk >> 17) & 31, w1 = (pk >> 22) & 31, wrap = (pk >> 27) & 1;
      ux = rb.x;
      uy = rb.y;
      ul = rb.z;
      y1 = wrap ? 0 : uy + 1;
      l1 = ul + wrap;
      five = !((s >> (D - 1)) & 1);
      lr = five ? lr5 : lr3;
      slot = five ? slot5 : slot3;
      active = lr < nrows && slot < (five ? 5 : 3);
      const int sx = s & 1;
      const bool second = lr >= n0;
      cx = second ? lr - n0 : ux + lr;
      cy = second ? y1 : uy;
      cl = second ? l1 : ul;
      row_off = second ? rb.w + w1 * q2_xpl(sx, cx) : w0 * (q2_xpl(sx, cx) - q2_xpl(sx, ux));
    } else {
    if (LN) { // items ascend: the row group only moves forward
      while (gi + 1 < p.n_rowgroups && (int)item >= (int)p.rg[gi + 1].item_begin)
        ++gi;
    } else {
      gi = 0;
#pragma unroll
      for (int k = 1; k < 8; ++k)
        if (k < p.n_rowgroups && (int)item >= (int)p.rg[k].item_begin)
          gi = k;
    }
    const Q2RowGroup& rg = p.rg[gi];
    s = rg.s;
    five = !((s >> (D - 1)) & 1);
    const int rpi = five ? Q2G_THREADS / 5 : Q2G_THREADS / 3;
    const unsigned lrow0 = unsigned(rg.lex_begin) + unsigned(item - rg.item_begin) * rpi; // first row of the item
    nrows = (int)min((long long)rpi, rg.lex_end - lrow0);
    q2_decode<D>(rg, lrow0, ux, uy, ul);
    lr = five ? lr5 : lr3;
    slot = five ? slot5 : slot3;
    active = lr < nrows && slot < (five ? 5 : 3);
    if (LN) {
      const int ex = (int)rg.ex;
      n0 = min(nrows, ex - ux); // rows of the item on its first line
      y1 = uy + 1;              // the line after it
      l1 = ul;
      if (D == 2 || y1 == (int)rg.ey) {
        y1 = 0;
        l1 = ul + 1;
      }
      long long line0, line1;
      int w0, w1;
      q2_line<D>(g, rg, uy, ul, line0, w0);
      q2_line<D>(g, rg, y1, l1, line1, w1);
      const int sx = s & 1;
      const long long off0 = line0 + w0 * q2_xpl(sx, ux);
      const int rest = nrows - n0;
      // end of the segment: inside the first line, exactly at its end (= start of the next line), or on the second line
      const long long off1 = rest > 0           ? line1 + w1 * q2_xpl(sx, rest)
                             : ux + nrows < ex ? line0 + w0 * q2_xpl(sx, ux + nrows)
                                               : line0 + (long long)w0 * (int)rg.Tx;
      start = rg.value_begin + off0;
      seg = int(off1 - off0);
      const bool second = lr >= n0;
      cx = second ? lr - n0 : ux + lr;
      cy = second ? y1 : uy;
      cl = second ? l1 : ul;
      row_off = (second ? int(line1 - off0) : int(line0 - off0)) + (second ? w1 : w0) * q2_xpl(sx, cx);
    } else {
      int vx, vy, vl;
      const long long off0 = q2_row_offset<D>(g, rg, ux, uy, ul);
      long long off1 = rg.off_end;
      if ((long long)lrow0 + nrows < rg.lex_end) {
        q2_decode<D>(rg, lrow0 + nrows, vx, vy, vl);
        off1 = q2_row_offset<D>(g, rg, vx, vy, vl);
      }
      start = rg.value_begin + off0;
      seg = int(off1 - off0);
      if (active) {
        q2_decode<D>(rg, lrow0 + lr, cx, cy, cl);
        row_off = int(q2_row_offset<D>(g, rg, cx, cy, cl) - off0);
      }
    }
    }
    const int phase = int((reinterpret_cast<unsigned long long>(values + start) >> 3) & 1ULL);
    double* stage = smem + buf * stage_doubles + phase;

    // the stage about to be written must have been read out by the bulk store that used it last: waiting HERE (after
    // this item's index arithmetic) instead of right after the store lets the drain overlap the bookkeeping
    if (!ACCUMULATE && item != (long long)blockIdx.x) {
      if (threadIdx.x == 0) {
        if (nbuf == 1)
          q2_bulk_wait_read0();
        else
          q2_bulk_wait_read1();
      }
      __syncthreads();
    }
    if (active) {
      double* row = stage + row_off;
      if (D == 3) {
        switch (s) {
          case 0: q2_dispatch<SF, 3, 0, 0, 0, M, KIND>(p, cx, cy, cl, slot, row); break;
          case 1: q2_dispatch<SF, 3, 1, 0, 0, M, KIND>(p, cx, cy, cl, slot, row); break;
          case 2: q2_dispatch<SF, 3, 0, 1, 0, M, KIND>(p, cx, cy, cl, slot, row); break;
          case 3: q2_dispatch<SF, 3, 1, 1, 0, M, KIND>(p, cx, cy, cl, slot, row); break;
          case 4: q2_dispatch<SF, 3, 0, 0, 1, M, KIND>(p, cx, cy, cl, slot, row); break;
          case 5: q2_dispatch<SF, 3, 1, 0, 1, M, KIND>(p, cx, cy, cl, slot, row); break;
          case 6: q2_dispatch<SF, 3, 0, 1, 1, M, KIND>(p, cx, cy, cl, slot, row); break;
          default: q2_dispatch<SF, 3, 1, 1, 1, M, KIND>(p, cx, cy, cl, slot, row); break;
        }
      } else {
        switch (s) {
          case 0: q2_dispatch<SF, 2, 0, 0, 0, M, KIND>(p, cx, cy, cl, slot, row); break;
          case 1: q2_dispatch<SF, 2, 1, 0, 0, M, KIND>(p, cx, cy, cl, slot, row); break;
          case 2: q2_dispatch<SF, 2, 0, 0, 1, M, KIND>(p, cx, cy, cl, slot, row); break;
          default: q2_dispatch<SF, 2, 1, 0, 1, M, KIND>(p, cx, cy, cl, slot, row); break;
        }
      }
    }

    if (ACCUMULATE) {
      __syncthreads();
      for (int i = threadIdx.x; i < seg; i += blockDim.x)
        values[start + i] += stage[i];
      __syncthreads();
    } else {
      q2_fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        const int head = phase;
        const int body = (seg - head) & ~1;
        if (head)
          values[start] = stage[0];
        if (body > 0)
          q2_bulk_store_s2g(values + start + head, stage + head, (unsigned)(body * sizeof(double)));
        if (head + body < seg)
          values[start + head + body] = stage[head + body];
        q2_bulk_commit();
      }
      buf = nbuf == 1 ? 0 : buf ^ 1;
    }
  }
  if (!ACCUMULATE && threadIdx.x == 0)
    q2_bulk_wait0();
}

} // namespace

int q2_slab_ranges(const GridDev& g, const SpaceDev& sp, Q2SlabRange* out)
{
  const int d = g.d;
  const long long Nl = g.n[d - 1];
  const long long zb = g.layer_lo, ze = g.layer_hi;
  int r = 0;
  long long global_value = 0, local = 0;
  for (int c = 0; c <= d; ++c)
    for (int s = 0; s < (1 << d); ++s) {
      int pc = 0;
      for (int k = 0; k < d; ++k)
        pc += (s >> k) & 1;
      if (pc != d - c)
        continue;
      long long per_layer = 1, layer_entries = 1, total = 1;
      for (int k = 0; k < d - 1; ++k) {
        per_layer *= ((s >> k) & 1) ? g.n[k] : g.n[k] + 1;
        layer_entries *= q2_axis_total((s >> k) & 1, g.n[k]);
      }
      const int SL = (s >> (d - 1)) & 1;
      total = layer_entries * q2_axis_total(SL, Nl);
      // lattice layers p_l = 2 c_l + SL of the element layers [zb, ze); the last slab also owns the top (even) layer
      const long long cl_lo = zb, cl_hi = SL ? ze : (ze == Nl ? Nl + 1 : ze);
      const long long layers_in_group = SL ? Nl : Nl + 1;
      const long long off_lo = layer_entries * q2_axis_len(SL, (int)cl_lo, (int)Nl).PL;
      const long long off_hi = cl_hi == layers_in_group ? total : layer_entries * q2_axis_len(SL, (int)cl_hi, (int)Nl).PL;
      Q2SlabRange& R = out[r++];
      const long long group_row_begin = sp.cg.codim_offset[c] + sp.cg.group_offset[s];
      R.row_begin = group_row_begin + cl_lo * per_layer;
      R.row_end = group_row_begin + cl_hi * per_layer;
      R.value_offset = global_value + off_lo;
      R.local_offset = local;
      R.count = off_hi - off_lo;
      local += R.count;
      global_value += total;
    }
  return r;
}

long long q2_pe_table_doubles(const GridDev& g)
{
  long long total = 24 * (g.n[0] + 1);
  for (int k = 1; k < g.d; ++k)
    total += 12 * (2 * g.n[k] + 1);
  return total;
}

long long q2_sf_table_doubles(const GridDev& g)
{
  long long total = 20 * (g.n[0] + 1); // x: component-major, both parities padded to N_x + 1 points (q2_x_index)
  for (int k = 1; k < g.d; ++k)
    total += 10 * (2 * g.n[k] + 1);
  return total;
}

namespace {

using Q2Kernel = void (*)(const Q2GatherParams, double*, int, int);

template <int D, int M>
Q2Kernel q2_qp_kernel_dm(int kind, bool accumulate)
{
  switch (kind) {
    case Q1G_LAPLACE_SCALAR:
      return accumulate ? k_q2_gather<D, true, 3, M, Q1G_LAPLACE_SCALAR> : k_q2_gather<D, false, 3, M, Q1G_LAPLACE_SCALAR>;
    case Q1G_MASS: return accumulate ? k_q2_gather<D, true, 3, M, Q1G_MASS> : k_q2_gather<D, false, 3, M, Q1G_MASS>;
    default:
      return accumulate ? k_q2_gather<D, true, 3, M, Q1G_LAPLACE_TENSOR> : k_q2_gather<D, false, 3, M, Q1G_LAPLACE_TENSOR>;
  }
}

template <int D>
Q2Kernel q2_qp_kernel_d(int m, int kind, bool accumulate)
{
  switch (m) {
    case 2: return q2_qp_kernel_dm<D, 2>(kind, accumulate);
    case 3: return q2_qp_kernel_dm<D, 3>(kind, accumulate);
    case 4: return q2_qp_kernel_dm<D, 4>(kind, accumulate);
    default: return nullptr;
  }
}

int launch_q2_impl(Launch& L, Q2GatherParams& p, const SpaceDev& sp, double* values, bool accumulate, bool qp);
void q2_setup_rowgroups(Q2GatherParams& p, const SpaceDev& sp);
bool q2_lines_apply(const Q2GatherParams& p);

} // namespace

bool q2_qp_supported(int d, int m, int kind)
{
  (void)kind;
  return (d == 2 && m >= 2 && m <= 4) || (d == 3 && m >= 2 && m <= 3);
}

long long q2_item_count(const GridDev& g, const SpaceDev& sp)
{
  if ((g.d != 2 && g.d != 3) || sp.size >= (1LL << 31))
    return 0;
  Q2GatherParams p;
  std::memset(&p, 0, sizeof(p));
  p.g = g;
  q2_setup_rowgroups(p, sp);
  return q2_lines_apply(p) ? p.n_items : 0;
}

int launch_q2_gather(Launch& L, Q2GatherParams& p, const SpaceDev& sp, double* values, bool accumulate)
{
  return launch_q2_impl(L, p, sp, values, accumulate, false);
}

int launch_q2_gather_qp(Launch& L, Q2GatherParams& p, const CgQpGroup& group, const SpaceDev& sp, double* values,
                        bool accumulate)
{
  p.qp = group;
  p.sf = 0;
  p.n_groups = 0;
  return launch_q2_impl(L, p, sp, values, accumulate, true);
}

namespace {

// row groups in ascending global index order (codim ascending, shift bitset ascending), restricted to the slab, and the
// work items of every group
void q2_setup_rowgroups(Q2GatherParams& p, const SpaceDev& sp)
{
  const GridDev& g = p.g;
  const int d = g.d;
  Q2SlabRange ranges[8];
  p.n_rowgroups = q2_slab_ranges(g, sp, ranges);
  p.n_items = 0;
  {
    int r = 0;
    for (int c = 0; c <= d; ++c)
      for (int s = 0; s < (1 << d); ++s) {
        int pc = 0;
        for (int k = 0; k < d; ++k)
          pc += (s >> k) & 1;
        if (pc != d - c)
          continue;
        Q2RowGroup& rg = p.rg[r];
        rg.s = s;
        rg.rows = 1;
        for (int k = 0; k < d; ++k)
          rg.rows *= ((s >> k) & 1) ? g.n[k] : g.n[k] + 1;
        rg.row_begin = sp.cg.codim_offset[c] + sp.cg.group_offset[s];
        rg.ex = (unsigned)((s & 1) ? g.n[0] : g.n[0] + 1);
        rg.ey = d == 3 ? (unsigned)((s & 2) ? g.n[1] : g.n[1] + 1) : 1u;
        rg.mex = rg.ex > 1 ? ~0ULL / rg.ex + 1 : 0;
        rg.mey = rg.ey > 1 ? ~0ULL / rg.ey + 1 : 0;
        rg.Tx = (unsigned)q2_axis_total(s & 1, g.n[0]);
        rg.TxTy = (long long)rg.Tx * (d == 3 ? q2_axis_total((s >> 1) & 1, g.n[1]) : 1);
        // owned rows: whole lattice layers along the last axis
        const long long per_layer = (long long)rg.ex * rg.ey;
        rg.lex_begin = ranges[r].row_begin - rg.row_begin;
        rg.lex_end = ranges[r].row_end - rg.row_begin;
        const int SL = (s >> (d - 1)) & 1;
        const long long layer_entries = d == 3 ? rg.TxTy : (long long)rg.Tx;
        const long long off_begin = layer_entries * q2_axis_len(SL, (int)(rg.lex_begin / per_layer), (int)g.n[d - 1]).PL;
        rg.off_end = off_begin + ranges[r].count;
        rg.value_begin = ranges[r].local_offset - off_begin;
        const int slots = SL ? 3 : 5;
        const int rpi = Q2G_THREADS / slots;
        rg.item_begin = p.n_items;
        p.n_items += (rg.lex_end - rg.lex_begin + rpi - 1) / rpi;
        ++r;
      }
  }
}

// the line-based bookkeeping needs every lattice line of every (non-empty) row group to hold a work item's rows
bool q2_lines_apply(const Q2GatherParams& p)
{
  for (int r = 0; r < p.n_rowgroups; ++r)
    if (p.rg[r].lex_end > p.rg[r].lex_begin && (int)p.rg[r].ex < Q2G_THREADS / 3)
      return false;
  return true;
}

template <int SF>
Q2Kernel q2_pick(int d, bool accumulate, int ln)
{
  if (accumulate)
    return d == 3 ? k_q2_gather<3, true, SF> : k_q2_gather<2, true, SF>;
  if (ln == 2)
    return d == 3 ? k_q2_gather<3, false, SF, 0, 0, 2> : k_q2_gather<2, false, SF, 0, 0, 2>;
  if (ln == 1)
    return d == 3 ? k_q2_gather<3, false, SF, 0, 0, 1> : k_q2_gather<2, false, SF, 0, 0, 1>;
  return d == 3 ? k_q2_gather<3, false, SF> : k_q2_gather<2, false, SF>;
}

int launch_q2_impl(Launch& L, Q2GatherParams& p, const SpaceDev& sp, double* values, bool accumulate, bool qp)
{
  const GridDev& g = p.g;
  const int d = g.d;
  if (d != 2 && d != 3)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "q2_gather: 2D and 3D grids only");
  if (sp.size >= (1LL << 31))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "q2_gather: more than 2^31 degrees of freedom");
  for (int k = 0; k < d; ++k)
    if (g.n[k] > 5000)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "q2_gather: more than 5000 elements along one axis");
  q2_setup_rowgroups(p, sp);
  int max_row = 1;
  for (int k = 0; k < d; ++k)
    max_row *= 5;
  // stage: the longest segment is rows_per_item rows of the longest row kind that uses that slot count
  // (5 slots: rows up to 5^d entries, 51 rows; 3 slots: rows up to 3 * 5^(d-1), 85 rows)
  const int seg5 = (Q2G_THREADS / 5) * max_row, seg3 = (Q2G_THREADS / 3) * (max_row / 5 * 3);
  const int stage_doubles = ((std::max(seg5, seg3) + 2) + 1) & ~1;
  static const int nbuf_env = std::getenv("GDTB_Q2_NBUF") ? std::atoi(std::getenv("GDTB_Q2_NBUF")) : 0;
  const bool single_group = p.sf && p.n_groups == 1; // sf == 2 (one per-element integrand) included
  const int nbuf = accumulate ? 1 : (nbuf_env == 1 || nbuf_env == 2 ? nbuf_env : (single_group ? 1 : 2));
  const size_t smem = (size_t)nbuf * stage_doubles * sizeof(double);
  // line-based bookkeeping (LN) needs every lattice line of every row group to hold at least one work item's rows
  static const bool ln_off = std::getenv("GDTB_Q2_NO_LN") != nullptr;
  static const bool items_off = std::getenv("GDTB_Q2_NO_ITEMS") != nullptr;
  int ln = !ln_off && q2_lines_apply(p) ? 1 : 0;
  if (ln && p.items && !items_off && !accumulate && !qp)
    ln = 2; // work-item records
  Q2Kernel kern = q2_pick<0>(d, accumulate, ln);
  if (qp) {
    if (!q2_qp_supported(d, p.qp.m, p.qp.kind))
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "q2_gather_qp: unsupported number of Gauss points per direction");
    kern = d == 3 ? q2_qp_kernel_d<3>(p.qp.m, p.qp.kind, accumulate) : q2_qp_kernel_d<2>(p.qp.m, p.qp.kind, accumulate);
  } else if (p.sf == 2) {
    kern = q2_pick<4>(d, accumulate, ln);
    long long off = 0;
    for (int k = 0; k < d; ++k) {
      p.sf_axis_off[k] = off;
      off += k == 0 ? 24 * (g.n[0] + 1) : 12 * (2 * g.n[k] + 1);
    }
    p.sf_group_stride = off;
  } else if (p.sf) {
    kern = p.n_groups == 1 ? q2_pick<2>(d, accumulate, ln) : q2_pick<1>(d, accumulate, ln);
    long long off = 0;
    for (int k = 0; k < d; ++k) {
      p.sf_axis_off[k] = off;
      off += k == 0 ? 20 * (g.n[0] + 1) : 10 * (2 * g.n[k] + 1);
    }
    p.sf_group_stride = off;
  }
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  GDTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Q2G_THREADS, smem));
  if (per_sm < 1)
    return fail(GDTB_ERR_CUDA, "q2_gather: kernel does not fit on an SM");
  // shared-memory carve-out: what the resident blocks need (dynamic + static + 1 KB each), the rest of the 256 KB stays L1
  {
    cudaFuncAttributes fa;
    GDTB_CUDA(cudaFuncGetAttributes(&fa, kern));
    const size_t per_block = smem + fa.sharedSizeBytes + 1024;
    GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   (int)std::min<size_t>(100, ((size_t)per_sm * per_block * 100) / (228 * 1024) + 2)));
  }
  long long grid = (long long)per_sm * L.sm_count;
  if (grid > p.n_items)
    grid = p.n_items;
  note_kernel(L, KF_Q2_GATHER, reinterpret_cast<const void*>(kern));
  time_begin(L, KF_Q2_GATHER);
  if (p.sf == 2) {
    const long long work = p.sf_group_stride / 12;
    if (d == 3)
      k_q2_axis_tables_pe<3><<<(unsigned)((work + 127) / 128), 128, 0, L.stream>>>(p, const_cast<double*>(p.sf_tab));
    else
      k_q2_axis_tables_pe<2><<<(unsigned)((work + 127) / 128), 128, 0, L.stream>>>(p, const_cast<double*>(p.sf_tab));
    L.count++;
  } else if (p.sf) {
    const long long work = p.sf_group_stride / 10 * p.n_groups;
    k_q2_axis_tables<<<(unsigned)((work + 127) / 128), 128, 0, L.stream>>>(p, const_cast<double*>(p.sf_tab));
    L.count++;
  }
  if (ln == 2 && !p.items_ready) {
    if (d == 3)
      k_q2_items<3><<<(unsigned)((p.n_items + 127) / 128), 128, 0, L.stream>>>(p, reinterpret_cast<int4*>(p.items));
    else
      k_q2_items<2><<<(unsigned)((p.n_items + 127) / 128), 128, 0, L.stream>>>(p, reinterpret_cast<int4*>(p.items));
    L.count++;
    p.items_ready = 1; // tells the caller that the buffer now holds this grid's / slab's records
  } else if (ln != 2)
    p.items_ready = 0;
  kern<<<(unsigned)grid, Q2G_THREADS, smem, L.stream>>>(p, values, stage_doubles, nbuf);
  time_end(L, KF_Q2_GATHER);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

} // namespace

// ---- closed-form sparsity pattern of the CG Q2 element stencil ------------------------------------------------------
// One thread per row: the columns of a row are the lattice points of its clipped coupling box, grouped by parity pattern
// in ascending global index (codim ascending, shift ascending) and lexicographic inside a group -- written in that
// order they are the sorted, duplicate-free row the sort-and-unique builder produces (tests compare both with the oracle).
namespace {

struct Q2PatternParams
{
  GridDev g;
  int n_rowgroups;
  Q2RowGroup rg[8];
  long long group_row_begin[8]; // by parity pattern s
};

template <int D>
__global__ void __launch_bounds__(128) k_q2_pattern(const __grid_constant__ Q2PatternParams p, long long rows,
                                                     long long* __restrict__ rowptr, int* __restrict__ colidx,
                                                     long long nnz)
{
  const GridDev& g = p.g;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= rows; r += (long long)gridDim.x * blockDim.x) {
    if (r == rows) {
      rowptr[rows] = nnz;
      continue;
    }
    int gi = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k)
      if (k < p.n_rowgroups && r >= p.rg[k].row_begin)
        gi = k;
    const Q2RowGroup& rg = p.rg[gi];
    const int s = rg.s;
    int c[3];
    q2_decode<D>(rg, (unsigned)(r - rg.row_begin), c[0], c[1], c[2]);
    const long long start = rg.value_begin + q2_row_offset<D>(g, rg, c[0], c[1], c[2]);
    rowptr[r] = start;
    // clipped coupling box per axis in lattice coordinates
    int qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0};
    const int cc[3] = {c[0], D == 3 ? c[1] : c[2], c[2]}; // axis order x, y (2D: last), z
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const int S = (s >> k) & 1, R = S ? 1 : 2, pk = 2 * cc[k] + S;
      qlo[k] = max(0, pk - R);
      qhi[k] = min(2 * (int)g.n[k], pk + R);
    }
    int* out = colidx + start;
    for (int rank = 0; rank < (1 << D); ++rank) {
      const int sq = q2_group_order(D, rank);
      // first lattice value of the group's parity inside the box, per axis
      int f[3] = {0, 0, 0};
      bool empty = false;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const int par = (sq >> k) & 1;
        f[k] = qlo[k] + ((qlo[k] ^ par) & 1);
        empty = empty || f[k] > qhi[k];
      }
      if (empty)
        continue;
      const long long ex = (sq & 1) ? g.n[0] : g.n[0] + 1;
      const long long ey = D == 3 ? ((sq & 2) ? g.n[1] : g.n[1] + 1) : ((sq & 2) ? g.n[1] : g.n[1] + 1);
      const long long base = p.group_row_begin[sq];
      if (D == 3) {
        for (int qz = f[2]; qz <= qhi[2]; qz += 2)
          for (int qy = f[1]; qy <= qhi[1]; qy += 2)
            for (int qx = f[0]; qx <= qhi[0]; qx += 2)
              *out++ = (int)(base + (qx >> 1) + ex * ((qy >> 1) + ey * (long long)(qz >> 1)));
      } else {
        for (int qy = f[1]; qy <= qhi[1]; qy += 2)
          for (int qx = f[0]; qx <= qhi[0]; qx += 2)
            *out++ = (int)(base + (qx >> 1) + ex * (long long)(qy >> 1));
      }
    }
  }
}

} // namespace

int pattern_structured_cg_q2(Launch& L, const GridDev& g, const SpaceDev& sp, long long** d_rowptr, int** d_colidx,
                             long long* nnz_out)
{
  const int d = g.d;
  if (d != 2 && d != 3)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "structured Q2 pattern: 2D and 3D grids only");
  // reuse the row-group setup of the gather kernel on the whole grid
  Q2GatherParams gp;
  std::memset(&gp, 0, sizeof(gp));
  gp.g = g;
  gp.g.layer_lo = 0;
  gp.g.layer_hi = g.n[d - 1];
  Q2PatternParams p;
  std::memset(&p, 0, sizeof(p));
  p.g = gp.g;
  Q2SlabRange ranges[8];
  p.n_rowgroups = q2_slab_ranges(gp.g, sp, ranges);
  long long nnz = 0;
  {
    int r = 0;
    for (int c = 0; c <= d; ++c)
      for (int s = 0; s < (1 << d); ++s) {
        int pc = 0;
        for (int k = 0; k < d; ++k)
          pc += (s >> k) & 1;
        if (pc != d - c)
          continue;
        Q2RowGroup& rg = p.rg[r];
        rg.s = s;
        rg.row_begin = ranges[r].row_begin;
        rg.rows = ranges[r].row_end - ranges[r].row_begin;
        rg.ex = (unsigned)((s & 1) ? g.n[0] : g.n[0] + 1);
        rg.ey = d == 3 ? (unsigned)((s & 2) ? g.n[1] : g.n[1] + 1) : 1u;
        rg.mex = rg.ex > 1 ? ~0ULL / rg.ex + 1 : 0;
        rg.mey = rg.ey > 1 ? ~0ULL / rg.ey + 1 : 0;
        rg.Tx = (unsigned)q2_axis_total(s & 1, g.n[0]);
        rg.TxTy = (long long)rg.Tx * (d == 3 ? q2_axis_total((s >> 1) & 1, g.n[1]) : 1);
        rg.value_begin = ranges[r].value_offset;
        p.group_row_begin[s] = ranges[r].row_begin;
        nnz += ranges[r].count;
        ++r;
      }
  }
  if (sp.size >= (1LL << 31))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "structured Q2 pattern: more than 2^31 degrees of freedom");
  long long* rowptr = nullptr;
  int* colidx = nullptr;
  if (cudaMalloc(&rowptr, sizeof(long long) * (size_t)(sp.size + 1)) != cudaSuccess
      || cudaMalloc(&colidx, sizeof(int) * (size_t)nnz) != cudaSuccess) {
    cudaFree(rowptr);
    return fail(GDTB_ERR_OUT_OF_MEMORY, "pattern: out of device memory");
  }
  const unsigned grid = (unsigned)std::min<long long>((sp.size + 1 + 127) / 128, (long long)L.sm_count * 64);
  if (d == 3)
    k_q2_pattern<3><<<grid, 128, 0, L.stream>>>(p, sp.size, rowptr, colidx, nnz);
  else
    k_q2_pattern<2><<<grid, 128, 0, L.stream>>>(p, sp.size, rowptr, colidx, nnz);
  L.count++;
  cudaError_t err = cudaStreamSynchronize(L.stream);
  if (err != cudaSuccess) {
    cudaFree(rowptr);
    cudaFree(colidx);
    return fail(GDTB_ERR_CUDA, std::string("k_q2_pattern: ") + cudaGetErrorString(err));
  }
  *d_rowptr = rowptr;
  *d_colidx = colidx;
  *nnz_out = nnz;
  return GDTB_OK;
}

int q2_host_rowptr(const GridDev& g0, const SpaceDev& sp, long long* rowptr)
{
  GridDev g = g0;
  const int d = g.d;
  g.layer_lo = 0;
  g.layer_hi = g.n[d - 1];
  Q2SlabRange ranges[8];
  q2_slab_ranges(g, sp, ranges);
  int r = 0;
  long long row = 0;
  for (int c = 0; c <= d; ++c)
    for (int s = 0; s < (1 << d); ++s) {
      int pc = 0;
      for (int k = 0; k < d; ++k)
        pc += (s >> k) & 1;
      if (pc != d - c)
        continue;
      Q2RowGroup rg;
      std::memset(&rg, 0, sizeof(rg));
      rg.s = s;
      rg.ex = (unsigned)((s & 1) ? g.n[0] : g.n[0] + 1);
      rg.ey = d == 3 ? (unsigned)((s & 2) ? g.n[1] : g.n[1] + 1) : 1u;
      rg.Tx = (unsigned)q2_axis_total(s & 1, g.n[0]);
      rg.TxTy = (long long)rg.Tx * (d == 3 ? q2_axis_total((s >> 1) & 1, g.n[1]) : 1);
      if (ranges[r].row_begin != row)
        return fail(GDTB_ERR_SPACE, "q2_host_rowptr: row groups are not contiguous");
      for (long long lex = 0; lex < ranges[r].row_end - ranges[r].row_begin; ++lex) {
        int cx, cy, cl;
        if (d == 3) {
          q2_decode<3>(rg, (unsigned)lex, cx, cy, cl);
          rowptr[row++] = ranges[r].value_offset + q2_row_offset<3>(g, rg, cx, cy, cl);
        } else {
          q2_decode<2>(rg, (unsigned)lex, cx, cy, cl);
          rowptr[row++] = ranges[r].value_offset + q2_row_offset<2>(g, rg, cx, cy, cl);
        }
      }
      if (r == (1 << d) - 1)
        rowptr[row] = ranges[r].value_offset + ranges[r].count;
      ++r;
    }
  return row == sp.size ? GDTB_OK : fail(GDTB_ERR_SPACE, "q2_host_rowptr: space size mismatch");
}

} // namespace gdtb

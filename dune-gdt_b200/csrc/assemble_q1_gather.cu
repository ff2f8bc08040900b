// dune-gdt_b200/csrc/assemble_q1_gather.cu -- owner-computes-rows ("row gather") assembly for
// continuous-Lagrange Q1 spaces on axis-aligned structured grids.
//
// Replaces, for element forms whose integrands are sums of LocalLaplaceIntegrand / LocalElementProductIntegrand
// with element-wise constant coefficients, the whole chain
//   LocalElementBilinearFormAssembler::apply_local  (local/assembler/bilinear-form-assemblers.hh:110-128)
//   LocalElementIntegralBilinearForm::apply2        (local/bilinear-forms/integrals.hh:97-134)
//   LocalLaplaceIntegrand / LocalElementProductIntegrand::evaluate (laplace.hh:81-102, product.hh:104-130)
//   MatrixType::add_to_entry [EXT]
// and the matching functional chain (functional-assemblers.hh:77-86, local/functionals/integrals.hh:72-98).
//
// Formulation.  On an axis-aligned cell with an element-constant coefficient c_e the local matrix is
// c_e * Lref, where Lref ("reference tensor") is the quadrature sum evaluated once for the cell shape with
// the form's own Gauss rule (host side, capi.cu).  Instead of scattering 8x8 local matrices (read-modify-write,
// 2^d colour passes, 5x the compulsory traffic -- SURVEY.md section 8d) every CSR row is produced exactly once by
// the thread that owns its vertex: A[v][v+delta] = sum_{o in {0,1}^d} c_{e(v,o)} * Lref[i(o)][j(o,delta)],
// i.e. at most 2^d * 2^d = 64 FMAs per row in 3D.  The CSR position of every entry is a closed form of the
// vertex coordinates (tensor-product stencil), so neither rowptr nor colidx is read.
//
// Data movement.  One work item = one x-line chunk of vertices; its CSR values form ONE contiguous segment.
// Rows are staged in shared memory in CSR order and leave the SM as a single TMA bulk store
// (cp.async.bulk.global.shared::cta, SASS UBLKCP) -- HBM sees only full-line sequential writes.  The kernel is
// persistent-strided over work items with several CTAs per SM so that one CTA's store drains while the
// others compute.  Deterministic (no atomics, fixed summation order).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.hpp"

// resident blocks per SM the register allocation aims for, and the number of shared-memory stages (see the kernel)
#ifndef Q1G_MIN_BLOCKS
#define Q1G_MIN_BLOCKS 2
#endif
#ifndef Q1G_DEFAULT_NBUF
#define Q1G_DEFAULT_NBUF 2
#endif

namespace gdtb {

namespace {

__device__ __forceinline__ void fence_proxy_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void bulk_store_s2g(double* gdst, const double* ssrc, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void bulk_commit()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__device__ __forceinline__ void bulk_wait_read0()
{
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ void bulk_wait_read1()
{
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

__device__ __forceinline__ void bulk_wait0()
{
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// number of stencil columns before vertex index i along one axis with N cells: sum_{i' < i} n(i'),
// n(i') = 2 at the two ends, 3 inside
__host__ __device__ __forceinline__ int S_axis(int i, int N)
{
  return i == 0 ? 0 : (i > N ? 3 * N + 1 : 3 * i - 1);
}

__host__ __device__ __forceinline__ int n_axis(int i, int N)
{
  return 1 + (i > 0 ? 1 : 0) + (i < N ? 1 : 0);
}

template <int D>
struct P3
{
  static constexpr int value = D == 1 ? 3 : (D == 2 ? 9 : 27);
};

// stencil index of the column offset (dx, dy, dz) in {-1,0,1}^D: lexicographic in (dz, dy, dx) = CSR order
template <int D>
__device__ __forceinline__ constexpr int delta_index(int dx, int dy, int dz)
{
  return (dx + 1) + (D > 1 ? 3 * (dy + 1) : 0) + (D > 2 ? 9 * (dz + 1) : 0);
}

// closed-form CSR row pointer of vertex (ix, iy, iz) of the CG-Q1 element stencil (tensor-product pattern)
template <int D>
__host__ __device__ __forceinline__ long long q1_rowptr(int ix, int iy, int iz, int Nx, int Ny, int Nz)
{
  if (D == 1)
    return S_axis(ix, Nx);
  const long long Wx = 3LL * Nx + 1;
  if (D == 2)
    return S_axis(iy, Ny) * Wx + (long long)(n_axis(iy, Ny) * S_axis(ix, Nx));
  const long long Wy = 3LL * Ny + 1;
  return S_axis(iz, Nz) * Wy * Wx
         + n_axis(iz, Nz) * (S_axis(iy, Ny) * Wx + (long long)(n_axis(iy, Ny) * S_axis(ix, Nx)));
}

// extent of cell i along one axis exactly as YaspGrid<EquidistantOffsetCoordinates> hands it to the geometry:
// upper - lower with lower = origin + i * h, upper = origin + (i + 1) * h (no FMA contraction, so that the ulp
// noise of the extents is the one of the CPU path)
__device__ __forceinline__ double cell_extent(double lo, double h, int i)
{
  const double lower = __dadd_rn(lo, __dmul_rn(double(i), h));
  const double upper = __dadd_rn(lo, __dmul_rn(double(i + 1), h));
  return __dsub_rn(upper, lower);
}

// exact v / d for v < 2^31 by one widening multiply: magic = ceil(2^(31+l) / d), l = ceil(log2 d) (kernels.hpp)
__device__ __forceinline__ unsigned fast_div(unsigned v, const FastDiv& fd)
{
  return (unsigned)(((unsigned long long)v * fd.magic) >> fd.shift);
}

// vertex index -> (ix, iy, iz), x fastest; one past the last vertex decodes to i_last = N_last + 1, others 0
template <int D>
__device__ __forceinline__ void q1_decode(unsigned v, const FastDiv& dx, const FastDiv& dy, int& ix, int& iy, int& iz)
{
  ix = (int)v;
  iy = 0;
  iz = 0;
  if (D > 1) {
    const unsigned t = fast_div(v, dx);
    ix = int(v - t * dx.d);
    iy = (int)t;
    if (D > 2) {
      const unsigned u = fast_div(t, dy);
      iy = int(t - u * dy.d);
      iz = (int)u;
    }
  }
}

// read-only 16-byte load that stays where it is written (the record of the NEXT item is requested an item ahead)
__device__ __forceinline__ int4 q1_ldg_int4_here(const int4* ptr)
{
  int4 v;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}

// work-item records of k_q1_gather<..., PREF> (3D, ROWS rows per item, every lattice line at least ROWS vertices long):
// the uniform per-item bookkeeping (first vertex, CSR segment, the two lines an item touches) once per grid / slab
__global__ void __launch_bounds__(128) k_q1_items(const Q1GatherParams p, long long nrows, int nitems, int rows_per_item,
                                                   int4* __restrict__ recs)
{
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= nitems)
    return;
  const GridDev& g = p.g;
  const int Nx = (int)g.n[0], Ny = (int)g.n[1], Nz = (int)g.n[2];
  const long long l0 = (long long)item * rows_per_item;
  const int nr = (int)min((long long)rows_per_item, nrows - l0);
  const unsigned r0 = (unsigned)(p.row_offset + l0);
  int bx, by, bz, cx, cy, cz;
  q1_decode<3>(r0, p.div_vx, p.div_vy, bx, by, bz);
  q1_decode<3>(r0 + nr, p.div_vx, p.div_vy, cx, cy, cz);
  const long long gstart = q1_rowptr<3>(bx, by, bz, Nx, Ny, Nz);
  const long long gend = q1_rowptr<3>(cx, cy, cz, Nx, Ny, Nz);
  int y1 = by + 1, z1 = bz;
  if (y1 > Ny) {
    y1 = 0;
    z1 = bz + 1;
  }
  const bool two = bx + nr > Nx + 1;
  const int w0 = n_axis(by, Ny) * n_axis(bz, Nz);
  const int w1 = two ? n_axis(y1, Ny) * n_axis(z1, Nz) : 0;
  const long long start = gstart - p.value_offset;
  int4 a, b;
  a.x = (int)(unsigned)(start & 0xffffffffLL);
  a.y = (int)(unsigned)((unsigned long long)start >> 32);
  a.z = int(gend - gstart);
  a.w = nr | (w0 << 9) | (w1 << 13);
  b.x = bx;
  b.y = by;
  b.z = bz;
  b.w = two ? int(q1_rowptr<3>(0, y1, z1, Nx, Ny, Nz) - gstart) : 0;
  recs[2 * item] = a;
  recs[2 * item + 1] = b;
}

// per-qp kernel, kappa(x) I in 3D: the direction terms share their sum-factorisation stages (1) or run one by one (0)
#ifndef Q1G_QP_SHARED_STAGES
#define Q1G_QP_SHARED_STAGES 1
#endif
// per-pair factorised arithmetic of the one-kappa-per-cell Laplace rows (see the kernel); 0: cell-by-cell sum (A/B)
#ifndef Q1G_PF3
#define Q1G_PF3 1
#endif
// the eight coefficients of a vertex through cp.async slots requested one item ahead (1) or plain loads (0, default: the
// slot machinery costs more issue slots and registers than the latency it hides once the arithmetic is pair-factorised)
#ifndef Q1G_PE_SLOTS
#define Q1G_PE_SLOTS 0
#endif
constexpr int Q1G_ROWS = 256; // rows (vertices) per work item = threads per CTA
#ifndef Q1G_ROWS_PREF
#define Q1G_ROWS_PREF 224 // the same with the coefficient prefetch slots (k_q1_gather<..., PREF>)
#endif
static_assert(Q1G_ROWS_PREF < 512, "the work-item records pack the rows of an item into 9 bits");

// bounded wait for a counter another GPU raises in this GPU's memory (system-scope acquire); gives up after ~1 s
__device__ __forceinline__ void q1_wait_counter(const int* counter, int expect, int* timeout_flag)
{
  for (long long spin = 0; spin < (1LL << 23); ++spin) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v - expect >= 0)
      return;
    __nanosleep(100);
  }
  atomicExch(timeout_flag, 1);
}

// Contribution of ONE element (offset (ox, oy, oz) around the vertex) to the vertex' row: the row i = (2^D-1) ^ o of
// its local matrix L_e = sum_g scale_g * coefficient_g(e) * sum_{r,c} (1/h_r)(1/h_c) |det J_e| M_g[r][c] is added to
// the stencil accumulators.  P0 / P1 receive the ansatz vertices with s_last = 0 / 1 (D == 3: the two z-planes the
// element touches, index 3 (oy + sy) + ox + sx; D < 3: both point to the full accumulator array, index delta_index).
// a[k] = h_k(e) and b[k] = 1 / h_k(e), both zeroed for a cell outside the grid / slab: such an element contributes
// exact zeros and no branch is needed.
template <int D, int NG, int KIND0, bool CELLDATA>
__device__ __forceinline__ void q1_add_element(const Q1GatherParams& p, int ox, int oy, int oz, const double (&a)[3],
                                               const double (&b)[3], bool valid, long long e, double* __restrict__ P0,
                                               double* __restrict__ P1)
{
  constexpr int NO = 1 << D;
  const int o = ox + 2 * oy + 4 * oz;
  const double ie = a[0] * (D > 1 ? a[1] : 1.) * (D > 2 ? a[2] : 1.); // |det J_e| (integrals.hh:119)
#pragma unroll
  for (int gi = 0; gi < NG; ++gi) {
    const Q1Group& G = p.group[gi];
    const int kind = KIND0 >= 0 ? KIND0 : G.kind;
    double cf = G.scale;
    if (CELLDATA && G.coef_elem && kind != Q1G_LAPLACE_TENSOR)
      cf *= valid ? __ldg(G.coef + e) : 0.;
    if (kind == Q1G_LAPLACE_SCALAR) {
      // kappa = c I: L_e = c sum_r |det J| / h_r^2 M[r][r]
      double w[3];
      w[0] = cf * (b[0] * (D > 1 ? a[1] : 1.) * (D > 2 ? a[2] : 1.));
      if (D > 1)
        w[1] = cf * (a[0] * b[1] * (D > 2 ? a[2] : 1.));
      if (D > 2)
        w[2] = cf * (a[0] * a[1] * b[2]);
#pragma unroll
      for (int s = 0; s < NO; ++s) {
        const int sx = s & 1, sy = (s >> 1) & 1, sz = (s >> 2) & 1;
        double* P = (D == 3 ? sz : (D == 2 ? 0 : 0)) ? P1 : P0;
        const int t = D == 3 ? 3 * (oy + sy) + ox + sx : delta_index<D>(ox - 1 + sx, oy - 1 + sy, 0);
#pragma unroll
        for (int r = 0; r < D; ++r)
          P[t] = fma(w[r], G.M[r * 3 + r][o][s], P[t]);
      }
    } else if (kind == Q1G_MASS) {
      const double w = cf * ie;
#pragma unroll
      for (int s = 0; s < NO; ++s) {
        const int sx = s & 1, sy = (s >> 1) & 1, sz = (s >> 2) & 1;
        double* P = (D == 3 && sz) ? P1 : P0;
        const int t = D == 3 ? 3 * (oy + sy) + ox + sx : delta_index<D>(ox - 1 + sx, oy - 1 + sy, 0);
        P[t] = fma(w, G.M[0][o][s], P[t]);
      }
    } else {
      double w[9];
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const double kap = (CELLDATA && G.coef_elem) ? (valid ? __ldg(G.coef + e * (D * D) + r * D + c) : 0.)
                                                       : G.kappa[r * 3 + c];
          w[r * 3 + c] = cf * kap * (b[r] * b[c]) * ie;
        }
#pragma unroll
      for (int s = 0; s < NO; ++s) {
        const int sx = s & 1, sy = (s >> 1) & 1, sz = (s >> 2) & 1;
        double* P = (D == 3 && sz) ? P1 : P0;
        const int t = D == 3 ? 3 * (oy + sy) + ox + sx : delta_index<D>(ox - 1 + sx, oy - 1 + sy, 0);
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
          for (int c = 0; c < D; ++c)
            P[t] = fma(w[r * 3 + c], G.M[r * 3 + c][o][s], P[t]);
      }
    }
  }
}

// single Laplace integrand with kappa = c I in 3D (the headline configuration): the three direction weights
// w[r] = c_e |det J_e| / h_r^2 are handed in, so that the caller can share the per-axis products between the 8 cells
__device__ __forceinline__ void q1_add_element_lap3(const Q1Group& G, int ox, int oy, int oz, const double (&w)[3],
                                                    double* __restrict__ P0, double* __restrict__ P1)
{
  const int o = ox + 2 * oy + 4 * oz;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int sx = s & 1, sy = (s >> 1) & 1, sz = (s >> 2) & 1;
    double* P = sz ? P1 : P0;
    const int t = 3 * (oy + sy) + ox + sx;
#pragma unroll
    for (int r = 0; r < 3; ++r)
      P[t] = fma(w[r], G.M[r * 3 + r][o][s], P[t]);
  }
}

// writes the (dy, dx) entries of one z-plane (D == 3) / of the whole row (D < 3) in CSR order; returns the count
template <int NP>
__device__ __forceinline__ int q1_store_plane(double* __restrict__ row, const double* __restrict__ P, bool full,
                                              bool cx0, bool cx1, bool cy0, bool cy1)
{
  if (full) {
#pragma unroll
    for (int k = 0; k < NP; ++k)
      row[k] = P[k];
    return NP;
  }
  int pos = 0;
#pragma unroll
  for (int dy = 0; dy < NP / 3; ++dy) {
    if (NP > 3 && (dy == 0 ? !cy0 : (dy == 2 ? !cy1 : false)))
      continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      if (dx == 0 ? !cx0 : (dx == 2 ? !cx1 : false))
        continue;
      row[pos++] = P[3 * dy + dx];
    }
  }
  return pos;
}

// Work item = Q1G_ROWS consecutive rows (vertices in mapper order).  Consecutive rows are consecutive in CSR, so the
// item's values form ONE contiguous segment [rowptr(r0), rowptr(r0 + Q1G_ROWS)): it is staged in shared memory in CSR
// order and leaves the SM as one TMA bulk store.  Thread t owns row r0 + t: it evaluates the geometry of the 2^D
// cells around its vertex, forms the row of each cell's local matrix that belongs to the vertex and sums the
// contributions per stencil column in a fixed order (deterministic, no atomics).
// CELLDATA = false: no per-element coefficient / source arrays and no per-cell right-hand-side terms are in play
// (constant coefficients, separable analytic source): the element index is never formed.
// PREF (3D, one Laplace integrand with one kappa per element): the eight coefficients a vertex needs are fetched one
// item ahead by cp.async into private shared-memory slots, so their DRAM latency hides behind the previous item's
// arithmetic; ROWS is lowered to make room for the slots next to the two stages of two resident blocks.
template <int D, int NG, int KIND0, bool ACCUMULATE, bool CELLDATA, bool P2P = false, int ROWS = Q1G_ROWS, bool PREF = false>
__global__ void __launch_bounds__(ROWS, (D == 3 && NG == 1 && KIND0 == Q1G_LAPLACE_SCALAR && !CELLDATA) ? Q1G_MIN_BLOCKS : 2)
    k_q1_gather(const __grid_constant__ Q1GatherParams p, double* __restrict__ values, double* __restrict__ rhs,
                long long nrows, int nitems, int stage_doubles, int nbuf)
{
  constexpr int NO = 1 << D; // elements around a vertex
  extern __shared__ __align__(16) double smem[];
  const GridDev& g = p.g;
  const int Nx = (int)g.n[0], Ny = D > 1 ? (int)g.n[1] : 1, Nz = D > 2 ? (int)g.n[2] : 1;
  // element range along the last direction (owner-computes slab + ghost layer)
  const int elo = (int)p.elem_lo, ehi = (int)p.elem_hi;
  const bool want_values = NG > 0 && values != nullptr;
  int buf = 0;
  // PREF: slot (o, t) = kappa of the cell with offset o around the vertex of thread t (0 for a cell outside the grid /
  // slab), behind the stage buffers
  double* const slots = smem + (size_t)nbuf * stage_doubles;
  // (PREF) vertex of thread t of an item whose first vertex is (bx, by, bz): the item's rows are consecutive along x and
  // a lattice line holds at least ROWS vertices (checked at launch), so at most one line break lies inside the item
  auto item_vertex = [&](int bx, int by, int bz, int& ix, int& iy, int& iz) -> bool {
    ix = bx + (int)threadIdx.x;
    iy = by;
    iz = bz;
    const bool second = ix > Nx;
    if (second) {
      ix -= Nx + 1;
      iy = by + 1;
      if (iy > Ny) {
        iy = 0;
        iz = bz + 1;
      }
    }
    return second;
  };
  auto prefetch_coef = [&](const int4& rb_, int nr_) {
    if ((int)threadIdx.x < nr_) {
      int ix, iy, iz;
      item_vertex(rb_.x, rb_.y, rb_.z, ix, iy, iz);
      const long long e0 = (long long)(ix - 1) + (long long)Nx * ((iy - 1) + (long long)Ny * (iz - 1));
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int cx_ = ix - 1 + (o & 1), cy_ = iy - 1 + ((o >> 1) & 1), cz_ = iz - 1 + (o >> 2);
        const bool valid = cx_ >= 0 && cx_ < Nx && cy_ >= 0 && cy_ < Ny && cz_ >= elo && cz_ < ehi;
        double* dst = slots + o * ROWS + threadIdx.x;
        if (valid) {
          const double* src = p.group[0].coef + (e0 + (o & 1) + (long long)Nx * (((o >> 1) & 1) + (long long)Ny * (o >> 2)));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
                       : "memory");
        } else
          *dst = 0.;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ... and two items ahead of that one thread per cell line asks the TMA unit to pull the line's window into L2
  // (cp.async.bulk.prefetch.L2): the DRAM reads of the coefficient stream then arrive as a few long bursts instead of
  // many 256-byte requests scattered between the writes (every read burst costs a bus turn-around in a write-only stream)
  auto prefetch_lines_l2 = [&](int item_) {
    if (item_ >= nitems || threadIdx.x >= 4)
      return;
    int ix, iy, iz;
    q1_decode<D>((unsigned)(p.row_offset + (long long)item_ * ROWS), p.div_vx, p.div_vy, ix, iy, iz);
    const long long ne = (long long)Nx * Ny * Nz;
    const long long first = (long long)(ix - 1) + (long long)Nx * ((iy - 1 + (int)(threadIdx.x & 1)) + (long long)Ny * (iz - 1 + (int)(threadIdx.x >> 1)));
    long long lo = max(first, 0LL), hi = min(first + ROWS + 2, ne);
    if (hi - lo < 4)
      return;
    const double* src = p.group[0].coef + lo;
    if (reinterpret_cast<unsigned long long>(src) & 8ULL) { // 16-byte alignment of the bulk prefetch
      ++src;
      ++lo;
    }
    const unsigned bytes = (unsigned)(((hi - lo) & ~1LL) * sizeof(double));
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
  };
  // (PREF) work-item records (k_q1_items): {start lo, start hi, seg, nr | w0 << 9 | w1 << 13}, {bx, by, bz, rel1}; the
  // record of the next item is requested one item ahead
  const int4* const recs = reinterpret_cast<const int4*>(p.items);
  int4 nra = make_int4(0, 0, 0, 0), nrb = nra;
  if (PREF) {
    if ((int)blockIdx.x < nitems) {
      nra = q1_ldg_int4_here(recs + 2 * blockIdx.x);
      nrb = q1_ldg_int4_here(recs + 2 * blockIdx.x + 1);
    }
    if (Q1G_PE_SLOTS)
      prefetch_coef(nrb, nra.w & 511);
    prefetch_lines_l2(blockIdx.x + (int)gridDim.x);
    prefetch_lines_l2(blockIdx.x + 2 * (int)gridDim.x);
  }

  for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
    // peer-memory halo: highest rows first (the top layer is what the neighbour above waits for), the bottom layer --
    // which waits for the neighbour below -- last
    const int item = P2P ? nitems - 1 - it : it;
    // ---- per item (uniform over the CTA): the CSR segment -------------------------------------------------
    const long long l0 = (long long)item * ROWS; // first local row
    int nr = (int)min((long long)ROWS, nrows - l0);
    const unsigned r0 = (unsigned)(p.row_offset + l0); // first global row (= vertex index)
    int4 ra = nra, rb = nrb;
    if (PREF) {
      nr = ra.w & 511;
      if (it + (int)gridDim.x < nitems) {
        nra = q1_ldg_int4_here(recs + 2 * (it + (int)gridDim.x));
        nrb = q1_ldg_int4_here(recs + 2 * (it + (int)gridDim.x) + 1);
      } else
        nra.w = 0; // no rows: the coefficient request below is empty
    }
    const Q1HaloP2p& H = p.halo;
    const long long top_row = nrows - H.layer_rows;
    const bool recv_item = P2P && H.has_lower && l0 < H.layer_rows;
    const bool send_item = P2P && H.has_upper && l0 + nr > top_row;
    if (P2P && (recv_item || send_item)) {
      if (threadIdx.x == 0) {
        if (recv_item)
          q1_wait_counter(H.my_flags + 0, H.expect_data, H.my_flags + 2); // the lower neighbour's layer has arrived
        if (send_item)
          q1_wait_counter(H.my_flags + 1, H.expect_ack, H.my_flags + 2);  // the neighbour has consumed this buffer's previous content
      }
      __syncthreads();
    }
    long long start = 0;
    int seg = 0, phase = 0;
    unsigned start32 = 0;
    double* stage = smem;
    if (PREF) {
      start = (long long)(((unsigned long long)(unsigned)ra.y << 32) | (unsigned long long)(unsigned)ra.x);
      seg = ra.z;
      phase = int((reinterpret_cast<unsigned long long>(values + start) >> 3) & 1ULL);
      stage = smem + buf * stage_doubles + phase;
    } else if (want_values) {
      int bx, by, bz, cx, cy, cz;
      q1_decode<D>(r0, p.div_vx, p.div_vy, bx, by, bz);
      q1_decode<D>(r0 + nr, p.div_vx, p.div_vy, cx, cy, cz);
      const long long gstart = q1_rowptr<D>(bx, by, bz, Nx, Ny, Nz);
      // one past the last vertex decodes to (0, 0, Nz + 1) [(0, Ny + 1) in 2D, Nx + 1 in 1D]: S_axis saturates
      const long long gend = q1_rowptr<D>(cx, cy, cz, Nx, Ny, Nz);
      seg = int(gend - gstart);
      start = gstart - p.value_offset;
      start32 = (unsigned)gstart;
      // keep the shared-memory and global 16-byte phases equal for the bulk copy
      phase = int((reinterpret_cast<unsigned long long>(values + start) >> 3) & 1ULL);
      stage = smem + buf * stage_doubles + phase;
    }

    // ---- per vertex --------------------------------------------------------------------------------------
    double c8[8];
    if (PREF && Q1G_PE_SLOTS) {
      // this item's coefficients arrived while the previous item was computed; the next item's are requested now
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int o = 0; o < 8; ++o)
        c8[o] = slots[o * ROWS + threadIdx.x];
      prefetch_coef(nrb, nra.w & 511);
    }
    if (PREF)
      prefetch_lines_l2(it + 3 * (int)gridDim.x);
    if ((int)threadIdx.x < nr) {
      int ix, iy, iz;
      bool second_line = false;
      if (PREF)
        second_line = item_vertex(rb.x, rb.y, rb.z, ix, iy, iz);
      else
        q1_decode<D>(r0 + threadIdx.x, p.div_vx, p.div_vy, ix, iy, iz);
      const int il[3] = {ix, iy, iz};
      const int Nl[3] = {Nx, Ny, Nz};
      // geometry of the two cells per axis around the vertex: ha[k][o] = h_k, hb[k][o] = 1 / h_k (J^{-T} = diag(1/h_k),
      // spaces/basis/default.hh:167-174; integration element = prod h_k) from the grid's per-axis tables
      // (k_q1_axis_tables: entry i + 1 belongs to cell i, the entries of the cells -1 and N_k are zero, so cells
      // outside the grid contribute exact zeros); along the last axis the slab's element range applies on top
      double ha[3][2], hb[3][2];
      bool vk[3][2];
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          ha[k][o] = 1.;
          hb[k][o] = 1.;
          vk[k][o] = o == 0;
          if (k < D) {
            const int c = il[k] - 1 + o;
            const double* tab = p.axis_tab[k] + (c + 1);
            ha[k][o] = __ldg(tab);
            hb[k][o] = __ldg(tab + p.axis_tab_inv);
            vk[k][o] = c >= 0 && c < Nl[k];
            if (k == D - 1) {
              vk[k][o] = c >= elo && c < ehi;
              ha[k][o] = vk[k][o] ? ha[k][o] : 0.;
              hb[k][o] = vk[k][o] ? hb[k][o] : 0.;
            }
          }
        }
      // element index of offset o = 0 (may be out of range; only dereferenced when valid)
      const long long e0 =
          CELLDATA ? (long long)(ix - 1) + (long long)Nx * ((D > 1 ? iy - 1 : 0) + (long long)Ny * (D > 2 ? iz - 1 : 0))
                   : 0;
      const bool cx0 = ix > 0, cx1 = ix < Nx;
      const bool cy0 = D > 1 && iy > 0, cy1 = D > 1 && iy < Ny, cz0 = D > 2 && iz > 0, cz1 = D > 2 && iz < Nz;
      const bool full_xy = cx0 && cx1 && (D < 2 || (cy0 && cy1));
      // wrap-around 32-bit arithmetic is exact for the (small) difference of two row pointers
      double* row = stage;
      if (PREF) {
        // row start from the item's line records: w = n(iy) n(iz) entries per entry along x
        const int w0 = (ra.w >> 9) & 15, w1 = (ra.w >> 13) & 15;
        row += second_line ? rb.w + w1 * S_axis(ix, Nx) : w0 * (S_axis(ix, Nx) - S_axis(rb.x, Nx));
      } else if (want_values)
        row += (unsigned)q1_rowptr<D>(ix, iy, iz, Nx, Ny, Nz) - start32;
      double bsum = 0.;

      if (D == 3) {
        // two passes over the element layers below / above the vertex; the z-plane of the stencil that is complete
        // after a pass is written out at once, so only two planes (18 values) are live at any time
        double P[3][9];
#pragma unroll
        for (int k = 0; k < 9; ++k)
          P[0][k] = P[1][k] = P[2][k] = 0.;
        constexpr bool LAP3 = NG == 1 && KIND0 == Q1G_LAPLACE_SCALAR;
        // products shared by the cells: |det J| / h_r^2 = (1/h_r) prod_{k != r} h_k, the x factor carries the scaling
        double fx_a[2], fx_b[2], yz_aa[2][2], yz_ba[2][2], yz_ab[2][2];
        if (LAP3) {
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            fx_a[o] = p.group[0].scale * ha[0][o];
            fx_b[o] = p.group[0].scale * hb[0][o];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              yz_aa[o][q] = ha[1][o] * ha[2][q];
              yz_ba[o][q] = hb[1][o] * ha[2][q];
              yz_ab[o][q] = ha[1][o] * hb[2][q];
            }
          }
        }
        // Constant kappa = c I (the headline configuration): the sum over the 8 cells factorises as well.  Per axis
        // the row couples to the offsets d = -1, 0, +1 through K^k[d] = sum_cells K1[.][.] / h_k and
        // M^k[d] = sum_cells M1[.][.] h_k (1D tables of the form's rule; the vertex is node 1 of the lower and node 0 of
        // the upper cell) and the stencil is c (K^x M^y M^z + M^x K^y M^z + M^x M^y K^z): 3 FP64 instructions per entry
        // instead of 8.  Cells outside the grid / slab carry h = 1/h = 0 and drop out.
        // (the PREF instantiation is only launched with one kappa per element: no sum-factorised stencil in it)
        const bool sf3 = !PREF && LAP3 && want_values && !(CELLDATA && p.group[0].coef_elem) && !p.no_sf3;
        if (sf3) {
          const Q1Group& G = p.group[0];
          double Kv[3][3], Mv[3][3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double sc = k == 0 ? G.scale : 1.;
            Kv[k][0] = sc * (hb[k][0] * G.K1[1][0]);
            Kv[k][1] = sc * fma(hb[k][0], G.K1[1][1], hb[k][1] * G.K1[0][0]);
            Kv[k][2] = sc * (hb[k][1] * G.K1[0][1]);
            Mv[k][0] = sc * (ha[k][0] * G.M1[1][0]);
            Mv[k][1] = sc * fma(ha[k][0], G.M1[1][1], ha[k][1] * G.M1[0][0]);
            Mv[k][2] = sc * (ha[k][1] * G.M1[0][1]);
          }
#pragma unroll
          for (int dz = 0; dz < 3; ++dz) {
            if ((dz == 0 && !cz0) || (dz == 2 && !cz1))
              continue;
            double Pl[9];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const double A = Mv[1][dy] * Mv[2][dz];
              const double B = fma(Kv[1][dy], Mv[2][dz], Mv[1][dy] * Kv[2][dz]);
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
                Pl[3 * dy + dx] = fma(Kv[0][dx], A, Mv[0][dx] * B);
            }
            row += q1_store_plane<9>(row, Pl, full_xy, cx0, cx1, cy0, cy1);
          }
        }
        // One Laplace integrand, kappa per cell (or constant with the sum-factorised stencil switched off): every cell's
        // row is kappa_e sum_r prod_k T^(r,k) with the 1D factors KK[k][o][s] = K1[1 - o][s] / h_k(o), MM[k][o][s] =
        // M1[1 - o][s] h_k(o) (s = local column, stencil offset o + s), so the two cells of an x-pair are combined first,
        //   PX[b] = sum_ox kappa KKx[ox][b - ox],  QX[b] = sum_ox kappa MMx[ox][b - ox],
        // and a (cell-pair, s_y, s_z) block is (MMy MMz) PX + (KKy MMz + MMy KKz) QX: 44 FP64 instructions per pair with
        // register operands instead of 60 + 48 constant-bank loads per pair of the cell-by-cell sum (q1_add_element_lap3)
        const bool pf3 = Q1G_PF3 && PREF && LAP3 && want_values && !sf3;
        double KKy[2][2], MMy[2][2], KKz[2][2], MMz[2][2], KKx[2][2], MMx[2][2];
        if (pf3) {
          const Q1Group& G = p.group[0];
#pragma unroll
          for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              KKx[o][q] = fx_b[o] * G.K1[1 - o][q];
              MMx[o][q] = fx_a[o] * G.M1[1 - o][q];
              KKy[o][q] = hb[1][o] * G.K1[1 - o][q];
              MMy[o][q] = ha[1][o] * G.M1[1 - o][q];
              KKz[o][q] = hb[2][o] * G.K1[1 - o][q];
              MMz[o][q] = ha[2][o] * G.M1[1 - o][q];
            }
        }
#pragma unroll
        for (int oz = 0; oz < 2; ++oz) {
          if (sf3 && (!CELLDATA || (!p.rhs_has_const && !p.rhs_has_elem)))
            break;
          if (pf3) {
#pragma unroll
            for (int oy = 0; oy < 2; ++oy) {
              double k0 = 1., k1 = 1.;
              if (CELLDATA && p.group[0].coef_elem) {
                if (PREF && Q1G_PE_SLOTS) {
                  k0 = c8[2 * oy + 4 * oz];
                  k1 = c8[1 + 2 * oy + 4 * oz];
                } else {
                  const long long e = e0 + (long long)Nx * (oy + (long long)Ny * oz);
                  const bool vyz = vk[1][oy] && vk[2][oz];
                  k0 = vyz && vk[0][0] ? __ldg(p.group[0].coef + e) : 0.;
                  k1 = vyz && vk[0][1] ? __ldg(p.group[0].coef + e + 1) : 0.;
                }
              }
              const double PX[3] = {k0 * KKx[0][0], fma(k0, KKx[0][1], k1 * KKx[1][0]), k1 * KKx[1][1]};
              const double QX[3] = {k0 * MMx[0][0], fma(k0, MMx[0][1], k1 * MMx[1][0]), k1 * MMx[1][1]};
#pragma unroll
              for (int sz = 0; sz < 2; ++sz)
#pragma unroll
                for (int sy = 0; sy < 2; ++sy) {
                  const double cA = MMy[oy][sy] * MMz[oz][sz];
                  const double cB = fma(KKy[oy][sy], MMz[oz][sz], MMy[oy][sy] * KKz[oz][sz]);
                  double* Pp = P[oz + sz] + 3 * (oy + sy);
#pragma unroll
                  for (int bb = 0; bb < 3; ++bb)
                    Pp[bb] = fma(cA, PX[bb], fma(cB, QX[bb], Pp[bb]));
                }
            }
          }
#pragma unroll
          for (int oxy = 0; oxy < 4; ++oxy) {
            const int ox = oxy & 1, oy = oxy >> 1;
            const double a[3] = {ha[0][ox], ha[1][oy], ha[2][oz]};
            const double b[3] = {hb[0][ox], hb[1][oy], hb[2][oz]};
            const bool valid = vk[0][ox] && vk[1][oy] && vk[2][oz];
            const long long e = CELLDATA ? e0 + ox + (long long)Nx * (oy + (long long)Ny * oz) : 0;
            if (want_values && !sf3 && !pf3) {
              if (LAP3) {
                // scheduling fence: keeps the weights of the 8 cells from being formed all at once (register pressure)
                double t0 = yz_aa[oy][oz];
                asm volatile("" : "+d"(t0));
                double w[3] = {fx_b[ox] * t0, fx_a[ox] * yz_ba[oy][oz], fx_a[ox] * yz_ab[oy][oz]};
                if (CELLDATA && p.group[0].coef_elem) {
                  const double c = PREF && Q1G_PE_SLOTS ? c8[ox + 2 * oy + 4 * oz] : (valid ? __ldg(p.group[0].coef + e) : 0.);
                  w[0] *= c;
                  w[1] *= c;
                  w[2] *= c;
                }
                q1_add_element_lap3(p.group[0], ox, oy, oz, w, P[oz], P[oz + 1]);
              } else
                q1_add_element<D, NG, KIND0, CELLDATA>(p, ox, oy, oz, a, b, valid, e, P[oz], P[oz + 1]);
            }
            if (!CELLDATA || (!p.rhs_has_const && !p.rhs_has_elem))
              continue;
            const double ie = a[0] * a[1] * a[2];
            if (p.rhs_has_const)
              bsum = fma(ie, p.rhs_S_const[ox + 2 * oy + 4 * oz], bsum);
            if (p.rhs_has_elem)
              bsum = fma(ie * p.rhs_S_elem[ox + 2 * oy + 4 * oz], valid ? __ldg(p.rhs_elem + e) : 0., bsum);
          }
          if (want_values && !sf3) {
            if (oz == 0) {
              if (cz0)
                row += q1_store_plane<9>(row, P[0], full_xy, cx0, cx1, cy0, cy1);
            } else {
              row += q1_store_plane<9>(row, P[1], full_xy, cx0, cx1, cy0, cy1);
              if (cz1)
                q1_store_plane<9>(row, P[2], full_xy, cx0, cx1, cy0, cy1);
            }
          }
        }
      } else {
        constexpr int ND = P3<D>::value;
        double P[ND];
#pragma unroll
        for (int k = 0; k < ND; ++k)
          P[k] = 0.;
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          const int ox = o & 1, oy = (o >> 1) & 1;
          const double a[3] = {ha[0][ox], ha[1][oy], 1.};
          const double b[3] = {hb[0][ox], hb[1][oy], 1.};
          const bool valid = vk[0][ox] && vk[1][oy];
          const long long e = CELLDATA ? e0 + ox + (long long)Nx * oy : 0;
          if (want_values)
            q1_add_element<D, NG, KIND0, CELLDATA>(p, ox, oy, 0, a, b, valid, e, P, P);
          if (!CELLDATA)
            continue;
          const double ie = a[0] * a[1];
          if (p.rhs_has_const)
            bsum = fma(ie, p.rhs_S_const[o], bsum);
          if (p.rhs_has_elem)
            bsum = fma(ie * p.rhs_S_elem[o], valid ? __ldg(p.rhs_elem + e) : 0., bsum);
        }
        if (want_values)
          q1_store_plane<ND>(row, P, full_xy, cx0, cx1, cy0, cy1);
      }

      if (p.has_rhs && rhs) {
        if (p.rhs_has_sep) {
          double t2 = p.rhs_sep_scale * __ldg(p.rhs_sep_tab + ix);
          if (D > 1)
            t2 *= __ldg(p.rhs_sep_tab + p.rhs_sep_stride + iy);
          if (D > 2)
            t2 *= __ldg(p.rhs_sep_tab + 2 * p.rhs_sep_stride + iz);
          bsum += t2;
        }
        const long long r = l0 + threadIdx.x;
        if (P2P && recv_item && r < H.layer_rows)
          bsum += __ldcg(H.recv_rhs + r); // partial sum of the interface row from the slab below
        if (P2P && send_item && r >= top_row)
          H.peer_rhs[r - top_row] = bsum; // partial sum of a row the slab above owns
        if (ACCUMULATE)
          rhs[r] += bsum;
        else
          rhs[r] = bsum;
      }
    }

    if (P2P && want_values && (recv_item || send_item)) {
      __syncthreads(); // the item's rows are complete in shared memory
      const long long top_value = (long long)p.halo.layer_values; // values of one interface layer
      if (recv_item) {
        // add the partial rows that arrived from below (bottom layer = local values [0, layer_values))
        const int n = (int)min((long long)seg, top_value - start);
        for (int i = threadIdx.x; i < n; i += blockDim.x)
          stage[i] += __ldcg(H.recv_values + start + i);
      }
      if (send_item) {
        // the top layer's partial rows go to the owner: local values [nnz_local - layer_values, nnz_local)
        const long long first = (long long)(p.halo_top_value_start);
        const int skip = (int)max(0LL, first - start);
        for (int i = threadIdx.x + skip; i < seg; i += blockDim.x)
          H.peer_values[start + i - first] = stage[i];
      }
    }
    if (P2P && send_item)
      __threadfence_system(); // peer stores (values and right-hand side) before the counter

    if (want_values) {
      if (ACCUMULATE) {
        __syncthreads();
        for (int i = threadIdx.x; i < seg; i += blockDim.x)
          values[start + i] += stage[i];
        __syncthreads();
      } else {
        // double-buffered TMA bulk store: the store of this item drains while the next one is computed
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
          // 16-byte aligned middle part by one bulk copy, at most one odd double at either end by plain stores
          const int head = phase;
          const int body = (seg - head) & ~1;
          if (head)
            values[start] = stage[0];
          if (body > 0)
            bulk_store_s2g(values + start + head, stage + head, (unsigned)(body * sizeof(double)));
          if (head + body < seg)
            values[start + head + body] = stage[head + body];
          bulk_commit();
          if (nbuf == 1)
            bulk_wait_read0(); // one stage: it must have been read out before the next item is written
          else
            bulk_wait_read1(); // the buffer written two items ago is free again
        }
        __syncthreads();
        buf = nbuf == 1 ? 0 : buf ^ 1;
      }
    }
    if (P2P && (recv_item || send_item)) {
      if (!want_values || ACCUMULATE)
        __syncthreads(); // every thread's peer stores / receive-buffer reads are done (the value path synchronised above)
      if (threadIdx.x == 0) {
        if (send_item)
          atomicAdd_system(H.peer_flags + 0, 1); // one more item of this step's interface layer is in place
        if (recv_item)
          atomicAdd_system(H.lower_flags + 1, 1); // this item no longer needs the receive buffer
      }
    }
  }
  if (!ACCUMULATE && want_values && threadIdx.x == 0)
    bulk_wait0();
}

// Per-axis geometry tables of the (tensor-product) grid, built once per grid on the device: for axis k and cell i
// tab[k][i + 1] = h_k(i) = upper - lower with YaspGrid's coordinates origin + i h (no FMA contraction), and
// tab[k][inv + i + 1] = 1 / h_k(i); the two border entries (cells -1 and N_k) are zero.
__global__ void k_q1_axis_tables(const GridDev g, double* __restrict__ t0, double* __restrict__ t1,
                                 double* __restrict__ t2, long long inv)
{
  double* tabs[3] = {t0, t1, t2};
  for (int k = 0; k < g.d; ++k)
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < g.n[k] + 2;
         j += (long long)gridDim.x * blockDim.x) {
      const long long i = j - 1;
      double h = 0., ih = 0.;
      if (i >= 0 && i < g.n[k]) {
        h = cell_extent(g.lo[k], g.h[k], (int)i);
        ih = __drcp_rn(h);
      }
      tabs[k][j] = h;
      tabs[k][inv + j] = ih;
    }
}

// Separable right-hand side tables: for f(x) = p0 * prod_k F_k(x_k),
//   B_k[i] = sum over the (valid) cells e in {i-1, i} of  ext_e * sum_q w_q phi_a(xi_q) F_k(lower_e + xi_q * ext_e),
// a = local index of vertex i in cell e.  One block, strided over (axis, vertex).
__global__ void k_q1_rhs_tables(const GridDev g, long long elem_lo, long long elem_hi, const FnDev f, int m,
                                const double* __restrict__ qx,
                                const double* __restrict__ qw, const double* __restrict__ phi,
                                double* __restrict__ tab, long long stride)
{
  const int last = g.d - 1;
  for (int k = 0; k < g.d; ++k) {
    const long long elo = (k == last) ? elem_lo : 0;
    const long long ehi = (k == last) ? elem_hi : g.n[k];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i <= g.n[k];
         i += (long long)gridDim.x * blockDim.x) {
      double s = 0.;
      for (int o = 0; o < 2; ++o) {
        const long long e = i - 1 + o;
        if (e < elo || e >= ehi)
          continue;
        const int a = 1 - o;
        const double lower = __dadd_rn(g.lo[k], __dmul_rn(double(e), g.h[k]));
        const double upper = __dadd_rn(g.lo[k], __dmul_rn(double(e + 1), g.h[k]));
        const double ext = __dsub_rn(upper, lower);
        double sc = 0.;
        for (int q = 0; q < m; ++q) {
          const double x = lower + qx[q] * ext;
          double F = 1.;
          if (f.builtin == GDTB_BUILTIN_COS_PRODUCT)
            F = cos(f.p[1] * x);
          else if (f.builtin == GDTB_BUILTIN_GAUSSIAN && k == 0) {
            const double t = x - f.p[0];
            F = exp(-(t * t) / (2. * (f.p[1] * f.p[1])));
          } else if (f.builtin == GDTB_BUILTIN_INDICATOR && k == 0)
            F = (f.p[0] <= x && x <= f.p[1]) ? 1. : 0.;
          sc += qw[q] * phi[q * 2 + a] * F;
        }
        s += sc * ext; // this axis' factor of the integration element prod_k ext_k (local/functionals/integrals.hh:96)
      }
      tab[k * stride + i] = s;
    }
  }
}

} // namespace

int launch_q1_axis_tables(Launch& L, const GridDev& g, double* const* tabs, long long inv)
{
  k_q1_axis_tables<<<8, 256, 0, L.stream>>>(g, tabs[0], tabs[1], tabs[2], inv);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

int launch_q1_rhs_tables(Launch& L, const GridDev& g, long long elem_lo, long long elem_hi, const FnDev& f, int m,
                         const double* qx, const double* qw, const double* phi, double* tab, long long stride)
{
  k_q1_rhs_tables<<<8, 256, 0, L.stream>>>(g, elem_lo, elem_hi, f, m, qx, qw, phi, tab, stride);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

template <int D, int NG, int KIND0, bool CELLDATA>
static int launch_q1_gather_dnc(Launch& L, const Q1GatherParams& p, double* values, double* rhs, bool accumulate)
{
  const GridDev& g = p.g;
  long long layer_rows = 1;
  for (int k = 0; k < D - 1; ++k)
    layer_rows *= g.n[k] + 1;
  long long total_rows = layer_rows * (g.n[D - 1] + 1);
  if (total_rows >= (1LL << 31) - Q1G_ROWS)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "q1_gather: more than 2^31 vertices");
  const long long nrows = (p.row_hi - p.row_lo) * layer_rows;
  if (nrows <= 0)
    return GDTB_OK;
  // two stage buffers (double-buffered bulk store), each padded for the 16-byte phase shift
  const bool with_values = NG > 0 && values;
  static const bool no_sf3 = std::getenv("GDTB_Q1_NO_SF3") != nullptr;
  const_cast<Q1GatherParams&>(p).no_sf3 = no_sf3 ? 1 : 0;
  if (accumulate && p.halo_p2p)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "q1_gather: the peer-memory halo works in overwrite mode");
  // one kappa per element (3D Laplace): coefficients prefetched one item ahead (see the kernel)
  static const bool no_pref = std::getenv("GDTB_Q1_NO_PREFETCH") != nullptr;
  constexpr bool CAN_PREF = D == 3 && NG == 1 && KIND0 == Q1G_LAPLACE_SCALAR && CELLDATA;
  // (the PREF kernel takes its per-item bookkeeping from work-item records: needs the caller's buffer and lattice lines
  // of at least one work item's rows)
  const bool pref = CAN_PREF && with_values && p.group[0].coef_elem && !accumulate && !p.halo_p2p && !no_pref && p.items
                    && g.n[0] + 1 >= Q1G_ROWS_PREF;
  const int rows_per_item = pref ? Q1G_ROWS_PREF : Q1G_ROWS;
  const long long nitems = (nrows + rows_per_item - 1) / rows_per_item;
  // two stage buffers (double-buffered bulk store), each padded for the 16-byte phase shift
  const int stage_doubles = ((rows_per_item * P3<D>::value + 2) + 1) & ~1;
  static const int nbuf_env = std::getenv("GDTB_Q1_NBUF") ? std::atoi(std::getenv("GDTB_Q1_NBUF")) : 0;
  const int nbuf = accumulate ? 1 : (nbuf_env == 1 || nbuf_env == 2 ? nbuf_env : Q1G_DEFAULT_NBUF);
  const size_t smem =
      with_values ? ((size_t)nbuf * stage_doubles + (pref && Q1G_PE_SLOTS ? 8 * Q1G_ROWS_PREF : 0)) * sizeof(double) : 16;
  auto kern = accumulate ? k_q1_gather<D, NG, KIND0, true, CELLDATA, false>
                         : (p.halo_p2p ? k_q1_gather<D, NG, KIND0, false, CELLDATA, true>
                                       : k_q1_gather<D, NG, KIND0, false, CELLDATA, false>);
  if constexpr (CAN_PREF)
    if (pref)
      kern = k_q1_gather<D, NG, KIND0, false, CELLDATA, false, Q1G_ROWS_PREF, true>;
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  GDTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, rows_per_item, smem));
  if (per_sm < 1)
    return fail(GDTB_ERR_CUDA, "q1_gather: kernel does not fit on an SM");
  // shared-memory carve-out: what the resident blocks need (dynamic + static + 1 KB each), the rest of the 256 KB stays L1
  {
    cudaFuncAttributes fa;
    GDTB_CUDA(cudaFuncGetAttributes(&fa, kern));
    const size_t per_block = smem + fa.sharedSizeBytes + 1024;
    GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   (int)std::min<size_t>(100, ((size_t)per_sm * per_block * 100) / (228 * 1024) + 2)));
  }
  long long grid = (long long)per_sm * L.sm_count;
  if (grid > nitems)
    grid = nitems;
  note_kernel(L, KF_Q1_GATHER, reinterpret_cast<const void*>(kern));
  time_begin(L, KF_Q1_GATHER);
  if (pref && !p.items_ready) {
    k_q1_items<<<(unsigned)((nitems + 127) / 128), 128, 0, L.stream>>>(p, nrows, (int)nitems, rows_per_item,
                                                                       reinterpret_cast<int4*>(p.items));
    L.count++;
  }
  const_cast<Q1GatherParams&>(p).items_ready = pref ? 1 : 0;
  kern<<<(unsigned)grid, rows_per_item, smem, L.stream>>>(p, values, rhs, nrows, (int)nitems, stage_doubles, nbuf);
  time_end(L, KF_Q1_GATHER);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

template <int D, int NG, int KIND0>
static int launch_q1_gather_dn(Launch& L, const Q1GatherParams& p, double* values, double* rhs, bool accumulate)
{
  bool celldata = rhs && p.has_rhs && (p.rhs_has_const || p.rhs_has_elem);
  for (int g = 0; g < (values ? p.n_groups : 0); ++g)
    celldata = celldata || p.group[g].coef_elem;
  // The variant without the per-cell code paths measures SLOWER on B200 and is opt-in for experiments only.  Before the
  // sum-factorised path: 0.788 vs 0.719 ms on C2.  With it (the lean kernel then is the pure sum-factorised one, 106
  // registers): 0.612 ms at 2 blocks x 2 stages, 0.579 at 2 x 1, 0.604 at 3 blocks (80 registers) x 1, 0.640 at 4 blocks
  // (64 registers) x 1 -- against 0.565 ms for the default build.
  static const bool allow_lean = std::getenv("GDTB_Q1_LEAN") != nullptr;
  celldata = celldata || !allow_lean;
  return celldata ? launch_q1_gather_dnc<D, NG, KIND0, true>(L, p, values, rhs, accumulate)
                  : launch_q1_gather_dnc<D, NG, KIND0, false>(L, p, values, rhs, accumulate);
}

template <int D>
static int launch_q1_gather_d(Launch& L, const Q1GatherParams& p, double* values, double* rhs, bool accumulate)
{
  switch (values ? p.n_groups : 0) {
    case 0: return launch_q1_gather_dn<D, 0, -1>(L, p, nullptr, rhs, accumulate);
    case 1:
      // a single integrand: its kind is a compile-time constant of the kernel
      switch (p.group[0].kind) {
        case Q1G_LAPLACE_SCALAR: return launch_q1_gather_dn<D, 1, Q1G_LAPLACE_SCALAR>(L, p, values, rhs, accumulate);
        case Q1G_MASS: return launch_q1_gather_dn<D, 1, Q1G_MASS>(L, p, values, rhs, accumulate);
        default: return launch_q1_gather_dn<D, 1, Q1G_LAPLACE_TENSOR>(L, p, values, rhs, accumulate);
      }
    case 2: return launch_q1_gather_dn<D, 2, -1>(L, p, values, rhs, accumulate);
    case 3: return launch_q1_gather_dn<D, 3, -1>(L, p, values, rhs, accumulate);
    default: return fail(GDTB_ERR_INVALID_ARGUMENT, "q1_gather: too many integrand groups");
  }
}

long long q1_pref_item_capacity(const GridDev& g, long long row_lo, long long row_hi)
{
  if (g.d != 3 || g.n[0] + 1 < Q1G_ROWS_PREF)
    return 0;
  const long long nrows = (row_hi - row_lo) * (g.n[0] + 1) * (g.n[1] + 1);
  return nrows > 0 ? (nrows + Q1G_ROWS_PREF - 1) / Q1G_ROWS_PREF : 0;
}

int launch_q1_gather(Launch& L, const Q1GatherParams& p, double* values, double* rhs, bool accumulate)
{
  switch (p.g.d) {
    case 1: return launch_q1_gather_d<1>(L, p, values, rhs, accumulate);
    case 2: return launch_q1_gather_d<2>(L, p, values, rhs, accumulate);
    case 3: return launch_q1_gather_d<3>(L, p, values, rhs, accumulate);
    default: return fail(GDTB_ERR_INVALID_ARGUMENT, "q1_gather: dimension must be 1, 2 or 3");
  }
}

// ==================================================================================================================
// Coefficients that vary inside a cell (one value / tensor per quadrature point): k_q1_gather_qp
// ==================================================================================================================
namespace {

template <int D, int M>
struct QpCount
{
  static constexpr int value = D == 1 ? M : (D == 2 ? M * M : M * M * M);
};

// out[s] += w * sum_q kq[q] prod_k PT[t_k][q_k][i_k][s_k]: one (r, c) term of one element's local-matrix row, by sum
// factorisation over the tensor rule (last axis first).  i_k = local index of the vertex in the element, s = ansatz
// vertex, t_k = point-table type along axis k; all of them compile-time constants after unrolling, so the tables are
// read as constant-bank operands straight from the kernel parameters.
template <int D, int M>
__device__ __forceinline__ void q1qp_term(const CgQpGroup& G, const double (&kq)[QpCount<D, M>::value], const int tx,
                                          const int ty, const int tz, const int ix, const int iy, const int iz,
                                          const double w, double (&out)[1 << D])
{
  if constexpr (D == 3) {
    double A[M][M][2];
#pragma unroll
    for (int qx = 0; qx < M; ++qx)
#pragma unroll
      for (int qy = 0; qy < M; ++qy)
#pragma unroll
        for (int sz = 0; sz < 2; ++sz) {
          double a = 0.;
#pragma unroll
          for (int qz = 0; qz < M; ++qz)
            a = fma(kq[qx + M * (qy + M * qz)], G.pt[tz][qz][iz][sz], a);
          A[qx][qy][sz] = a;
        }
    double B[M][2][2];
#pragma unroll
    for (int qx = 0; qx < M; ++qx)
#pragma unroll
      for (int sy = 0; sy < 2; ++sy)
#pragma unroll
        for (int sz = 0; sz < 2; ++sz) {
          double b = 0.;
#pragma unroll
          for (int qy = 0; qy < M; ++qy)
            b = fma(A[qx][qy][sz], G.pt[ty][qy][iy][sy], b);
          B[qx][sy][sz] = b;
        }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int sx = s & 1, sy = (s >> 1) & 1, sz = s >> 2;
      double c = 0.;
#pragma unroll
      for (int qx = 0; qx < M; ++qx)
        c = fma(B[qx][sy][sz], G.pt[tx][qx][ix][sx], c);
      out[s] = fma(w, c, out[s]);
    }
  } else if constexpr (D == 2) {
    double A[M][2];
#pragma unroll
    for (int qx = 0; qx < M; ++qx)
#pragma unroll
      for (int sy = 0; sy < 2; ++sy) {
        double a = 0.;
#pragma unroll
        for (int qy = 0; qy < M; ++qy)
          a = fma(kq[qx + M * qy], G.pt[ty][qy][iy][sy], a);
        A[qx][sy] = a;
      }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int sx = s & 1, sy = s >> 1;
      double c = 0.;
#pragma unroll
      for (int qx = 0; qx < M; ++qx)
        c = fma(A[qx][sy], G.pt[tx][qx][ix][sx], c);
      out[s] = fma(w, c, out[s]);
    }
  } else {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      double c = 0.;
#pragma unroll
      for (int qx = 0; qx < M; ++qx)
        c = fma(kq[qx], G.pt[tx][qx][ix][s], c);
      out[s] = fma(w, c, out[s]);
    }
  }
}

// the NQ coefficient samples of element e (scalar kinds)
template <int NQ>
__device__ __forceinline__ void q1qp_load(const CgQpGroup& G, const long long e, double (&kq)[NQ])
{
  // the NQ samples of an element are contiguous; neighbouring lanes read neighbouring elements, i.e. addresses 8 NQ
  // bytes apart: every load instruction of a warp touches 32 NQ / 16 lines whatever its width, so the widest load
  // (256-bit, sm_100) cuts the LSU wavefronts of this stream 4 x (round-2 profile: LSU data pipe 66 % busy, 47 % of the
  // warp-state samples waiting for these loads)
  const double* src = G.coef + e * (long long)NQ;
  const unsigned long long base = reinterpret_cast<unsigned long long>(G.coef);
  if (NQ % 4 == 0 && (base & 31ULL) == 0) {
#pragma unroll
    for (int q = 0; q < NQ / 4; ++q)
      ldg256(src + 4 * q, kq[4 * q], kq[4 * q + 1], kq[4 * q + 2], kq[4 * q + 3]);
  } else if (NQ % 2 == 0 && (base & 15ULL) == 0) {
#pragma unroll
    for (int q = 0; q < NQ / 2; ++q) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(src) + q);
      kq[2 * q] = v.x;
      kq[2 * q + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      kq[q] = __ldg(src + q);
  }
}

// row i = (2^D - 1) ^ o of the local matrix of element e (offset o around the vertex): all (r, c) terms of the integrand;
// scalar kinds take the element's samples from the caller (kq_in, requested one element ahead)
template <int D, int M, int KIND>
__device__ __forceinline__ void q1qp_element(const CgQpGroup& G, const long long e, const int ox, const int oy,
                                             const int oz, const double (&a)[3], const double (&b)[3],
                                             const double (&kq_in)[QpCount<D, M>::value], double (&out)[1 << D])
{
  constexpr int NQ = QpCount<D, M>::value;
  const int ix = 1 - ox, iy = 1 - oy, iz = 1 - oz;
  const double ie = a[0] * (D > 1 ? a[1] : 1.) * (D > 2 ? a[2] : 1.); // integrals.hh:119
  if (KIND == Q1G_LAPLACE_TENSOR) {
    double kq[NQ];
    const double* src = G.coef + e * (long long)(NQ * D * D);
    // the (r, c) loop stays rolled (the point-table type becomes a run-time index into the constant bank): D * D unrolled
    // copies of the contraction per element would not fit the instruction cache
#pragma unroll 1
    for (int rc = 0; rc < D * D; ++rc) {
      const int r = rc / D, c = rc - r * D;
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        kq[q] = __ldg(src + q * (D * D) + rc);
      // (kappa grad phi_j) . grad psi_i: test derivative along r, ansatz derivative along c (laplace.hh:98-101)
      const int t0 = 0 == r ? (0 == c ? QPT_KK : QPT_KM) : (0 == c ? QPT_MK : QPT_MM);
      const int t1 = 1 == r ? (1 == c ? QPT_KK : QPT_KM) : (1 == c ? QPT_MK : QPT_MM);
      const int t2 = 2 == r ? (2 == c ? QPT_KK : QPT_KM) : (2 == c ? QPT_MK : QPT_MM);
      const double br = r == 0 ? b[0] : (r == 1 ? b[1] : b[2]), bc = c == 0 ? b[0] : (c == 1 ? b[1] : b[2]);
      q1qp_term<D, M>(G, kq, t0, t1, t2, ix, iy, iz, G.scale * (ie * (br * bc)), out);
    }
  } else {
    if (KIND == Q1G_MASS)
      q1qp_term<D, M>(G, kq_in, QPT_MM, QPT_MM, QPT_MM, ix, iy, iz, G.scale * ie, out);
    else if (D == 3 && Q1G_QP_SHARED_STAGES) {
      // kappa(x) I in 3D: the three direction terms share their stages (as the x-fused Q2 kernel does) --
      //   stage z: AM / AK = kappa contracted with MM_z / KK_z;  stage y: B1 = AM MM_y, B2 = AM KK_y, B3 = AK MM_y;
      //   stage x: (w_x B1) KK_x + (w_y B2 + w_z B3) MM_x  -- 144 instead of 168 FP64 instructions per row and cell
      const double wx = G.scale * (ie * (b[0] * b[0])), wy = G.scale * (ie * (b[1] * b[1])), wz = G.scale * (ie * (b[2] * b[2]));
      double BK[M][2][2], BM[M][2][2];
#pragma unroll
      for (int qx = 0; qx < M; ++qx) {
        double AM[M][2], AK[M][2];
#pragma unroll
        for (int qy = 0; qy < M; ++qy)
#pragma unroll
          for (int sz = 0; sz < 2; ++sz) {
            double am = 0., ak = 0.;
#pragma unroll
            for (int qz = 0; qz < M; ++qz) {
              const double k = kq_in[qx + M * (qy + M * qz)];
              am = fma(k, G.pt[QPT_MM][qz][iz][sz], am);
              ak = fma(k, G.pt[QPT_KK][qz][iz][sz], ak);
            }
            AM[qy][sz] = am;
            AK[qy][sz] = ak;
          }
#pragma unroll
        for (int sy = 0; sy < 2; ++sy)
#pragma unroll
          for (int sz = 0; sz < 2; ++sz) {
            double b1 = 0., b2 = 0., b3 = 0.;
#pragma unroll
            for (int qy = 0; qy < M; ++qy) {
              b1 = fma(AM[qy][sz], G.pt[QPT_MM][qy][iy][sy], b1);
              b2 = fma(AM[qy][sz], G.pt[QPT_KK][qy][iy][sy], b2);
              b3 = fma(AK[qy][sz], G.pt[QPT_MM][qy][iy][sy], b3);
            }
            BK[qx][sy][sz] = wx * b1;
            BM[qx][sy][sz] = fma(wy, b2, wz * b3);
          }
      }
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int sx = s & 1, sy = (s >> 1) & 1, sz = s >> 2;
        double c = out[s];
#pragma unroll
        for (int qx = 0; qx < M; ++qx) {
          c = fma(BK[qx][sy][sz], G.pt[QPT_KK][qx][ix][sx], c);
          c = fma(BM[qx][sy][sz], G.pt[QPT_MM][qx][ix][sx], c);
        }
        out[s] = c;
      }
    } else {
#pragma unroll
      for (int r = 0; r < D; ++r)
        q1qp_term<D, M>(G, kq_in, r == 0 ? QPT_KK : QPT_MM, r == 1 ? QPT_KK : QPT_MM, r == 2 ? QPT_KK : QPT_MM, ix, iy, iz,
                        G.scale * (ie * (b[r] * b[r])), out);
    }
  }
}

// Same work decomposition and data movement as k_q1_gather (one thread per vertex row, Q1G_ROWS rows per item, the
// item's CSR segment staged in shared memory and written by one TMA bulk store); the per-element arithmetic is the
// sum-factorised quadrature loop above.
template <int D, int M, int KIND, bool ACCUMULATE>
// 2 blocks per SM (128 registers) for up to 2^3 samples per element: measured 2.04 ms vs 2.21 ms with one block (C2 per-qp)
__global__ void __launch_bounds__(Q1G_ROWS, (D == 3 && M >= 3) ? 1 : 2)
    k_q1_gather_qp(const __grid_constant__ Q1QpParams p, double* __restrict__ values, long long nrows, int nitems,
                   int stage_doubles, int nbuf)
{
  constexpr int NO = 1 << D;
  extern __shared__ __align__(16) double smem[];
  const GridDev& g = p.g;
  const CgQpGroup& G = p.group;
  const int Nx = (int)g.n[0], Ny = D > 1 ? (int)g.n[1] : 1, Nz = D > 2 ? (int)g.n[2] : 1;
  const int elo = (int)p.elem_lo, ehi = (int)p.elem_hi;
  int buf = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const long long l0 = (long long)item * Q1G_ROWS;
    const int nr = (int)min((long long)Q1G_ROWS, nrows - l0);
    const unsigned r0 = (unsigned)(p.row_offset + l0);
    int bx, by, bz, cx, cy, cz;
    q1_decode<D>(r0, p.div_vx, p.div_vy, bx, by, bz);
    q1_decode<D>(r0 + nr, p.div_vx, p.div_vy, cx, cy, cz);
    const long long gstart = q1_rowptr<D>(bx, by, bz, Nx, Ny, Nz);
    const long long gend = q1_rowptr<D>(cx, cy, cz, Nx, Ny, Nz);
    const int seg = int(gend - gstart);
    const long long start = gstart - p.value_offset;
    const unsigned start32 = (unsigned)gstart;
    const int phase = int((reinterpret_cast<unsigned long long>(values + start) >> 3) & 1ULL);
    double* stage = smem + buf * stage_doubles + phase;

    if ((int)threadIdx.x < nr) {
      int ix, iy, iz;
      q1_decode<D>(r0 + threadIdx.x, p.div_vx, p.div_vy, ix, iy, iz);
      const int il[3] = {ix, iy, iz};
      const int Nl[3] = {Nx, Ny, Nz};
      double ha[3][2], hb[3][2];
      bool vk[3][2];
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          ha[k][o] = 1.;
          hb[k][o] = 1.;
          vk[k][o] = o == 0;
          if (k < D) {
            const int c = il[k] - 1 + o;
            const double* tab = p.axis_tab[k] + (c + 1);
            ha[k][o] = __ldg(tab);
            hb[k][o] = __ldg(tab + p.axis_tab_inv);
            vk[k][o] = c >= 0 && c < Nl[k];
            if (k == D - 1)
              vk[k][o] = c >= elo && c < ehi;
          }
        }
      const long long e0 =
          (long long)(ix - 1) + (long long)Nx * ((D > 1 ? iy - 1 : 0) + (long long)Ny * (D > 2 ? iz - 1 : 0));
      const bool cx0 = ix > 0, cx1 = ix < Nx;
      const bool cy0 = D > 1 && iy > 0, cy1 = D > 1 && iy < Ny, cz0 = D > 2 && iz > 0, cz1 = D > 2 && iz < Nz;
      const bool full_xy = cx0 && cx1 && (D < 2 || (cy0 && cy1));
      double* row = stage + ((unsigned)q1_rowptr<D>(ix, iy, iz, Nx, Ny, Nz) - start32);
      if constexpr (D == 3) {
        double P[3][9];
#pragma unroll
        for (int k = 0; k < 9; ++k)
          P[0][k] = P[1][k] = P[2][k] = 0.;
        // coefficient samples are requested one element ahead (cells outside the grid / slab read the thread's first
        // valid cell instead: the request does not depend on the branch below and can be issued early)
        constexpr int NQ = QpCount<D, M>::value;
        constexpr bool AHEAD = KIND != Q1G_LAPLACE_TENSOR && NQ <= 8;
        const long long e_safe = e0 + (vk[0][0] ? 0 : 1) + (long long)Nx * ((vk[1][0] ? 0 : 1) + (long long)Ny * (vk[2][0] ? 0 : 1));
        double kq_next[NQ];
        if (AHEAD)
          q1qp_load<NQ>(G, vk[0][0] && vk[1][0] && vk[2][0] ? e0 : e_safe, kq_next);
#pragma unroll
        for (int oz = 0; oz < 2; ++oz) {
#pragma unroll
          for (int oxy = 0; oxy < 4; ++oxy) {
            const int ox = oxy & 1, oy = oxy >> 1;
            double kq[NQ];
            if (KIND != Q1G_LAPLACE_TENSOR && !AHEAD) {
              if (vk[0][ox] && vk[1][oy] && vk[2][oz])
                q1qp_load<NQ>(G, e0 + ox + (long long)Nx * (oy + (long long)Ny * oz), kq);
            }
            if (AHEAD) {
#pragma unroll
              for (int q = 0; q < NQ; ++q)
                kq[q] = kq_next[q];
              if (oz * 4 + oxy < 7) {
                const int n = oz * 4 + oxy + 1, nx = n & 1, ny = (n >> 1) & 1, nz = n >> 2;
                const long long en = e0 + nx + (long long)Nx * (ny + (long long)Ny * nz);
                q1qp_load<NQ>(G, vk[0][nx] && vk[1][ny] && vk[2][nz] ? en : e_safe, kq_next);
              }
            }
            if (!(vk[0][ox] && vk[1][oy] && vk[2][oz]))
              continue;
            const double a[3] = {ha[0][ox], ha[1][oy], ha[2][oz]};
            const double b[3] = {hb[0][ox], hb[1][oy], hb[2][oz]};
            const long long e = e0 + ox + (long long)Nx * (oy + (long long)Ny * oz);
            double out[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
            q1qp_element<D, M, KIND>(G, e, ox, oy, oz, a, b, kq, out);
#pragma unroll
            for (int s = 0; s < 8; ++s) {
              const int sx = s & 1, sy = (s >> 1) & 1, sz = s >> 2;
              P[oz + sz][3 * (oy + sy) + ox + sx] += out[s];
            }
          }
          if (oz == 0) {
            if (cz0)
              row += q1_store_plane<9>(row, P[0], full_xy, cx0, cx1, cy0, cy1);
          } else {
            row += q1_store_plane<9>(row, P[1], full_xy, cx0, cx1, cy0, cy1);
            if (cz1)
              q1_store_plane<9>(row, P[2], full_xy, cx0, cx1, cy0, cy1);
          }
        }
      } else {
        constexpr int ND = P3<D>::value;
        double P[ND];
#pragma unroll
        for (int k = 0; k < ND; ++k)
          P[k] = 0.;
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          const int ox = o & 1, oy = (o >> 1) & 1;
          if (!(vk[0][ox] && vk[1][oy]))
            continue;
          const double a[3] = {ha[0][ox], ha[1][oy], 1.};
          const double b[3] = {hb[0][ox], hb[1][oy], 1.};
          const long long e = e0 + ox + (long long)Nx * oy;
          double out[NO];
#pragma unroll
          for (int s = 0; s < NO; ++s)
            out[s] = 0.;
          double kq[QpCount<D, M>::value];
          if (KIND != Q1G_LAPLACE_TENSOR)
            q1qp_load<QpCount<D, M>::value>(G, e, kq);
          q1qp_element<D, M, KIND>(G, e, ox, oy, 0, a, b, kq, out);
#pragma unroll
          for (int s = 0; s < NO; ++s)
            P[delta_index<D>(ox - 1 + (s & 1), oy - 1 + ((s >> 1) & 1), 0)] += out[s];
        }
        q1_store_plane<ND>(row, P, full_xy, cx0, cx1, cy0, cy1);
      }
    }

    if (ACCUMULATE) {
      __syncthreads();
      for (int i = threadIdx.x; i < seg; i += blockDim.x)
        values[start + i] += stage[i];
      __syncthreads();
    } else {
      fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        const int head = phase;
        const int body = (seg - head) & ~1;
        if (head)
          values[start] = stage[0];
        if (body > 0)
          bulk_store_s2g(values + start + head, stage + head, (unsigned)(body * sizeof(double)));
        if (head + body < seg)
          values[start + head + body] = stage[head + body];
        bulk_commit();
        if (nbuf == 1)
          bulk_wait_read0();
        else
          bulk_wait_read1();
      }
      __syncthreads();
      buf = nbuf == 1 ? 0 : buf ^ 1;
    }
  }
  if (!ACCUMULATE && threadIdx.x == 0)
    bulk_wait0();
}

template <int D, int M, int KIND>
int launch_q1_qp_dmk(Launch& L, const Q1QpParams& p, double* values, bool accumulate)
{
  const GridDev& g = p.g;
  long long layer_rows = 1;
  for (int k = 0; k < D - 1; ++k)
    layer_rows *= g.n[k] + 1;
  if (layer_rows * (g.n[D - 1] + 1) >= (1LL << 31) - Q1G_ROWS)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "q1_gather: more than 2^31 vertices");
  const long long nrows = (p.row_hi - p.row_lo) * layer_rows;
  const long long nitems = (nrows + Q1G_ROWS - 1) / Q1G_ROWS;
  if (nitems <= 0)
    return GDTB_OK;
  const int stage_doubles = ((Q1G_ROWS * P3<D>::value + 2) + 1) & ~1;
  const int nbuf = accumulate ? 1 : 2;
  const size_t smem = (size_t)nbuf * stage_doubles * sizeof(double);
  auto kern = accumulate ? k_q1_gather_qp<D, M, KIND, true> : k_q1_gather_qp<D, M, KIND, false>;
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  GDTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Q1G_ROWS, smem));
  if (per_sm < 1)
    return fail(GDTB_ERR_CUDA, "q1_gather_qp: kernel does not fit on an SM");
  // shared-memory carve-out: what the resident blocks need (dynamic + static + 1 KB each), the rest of the 256 KB stays L1
  {
    cudaFuncAttributes fa;
    GDTB_CUDA(cudaFuncGetAttributes(&fa, kern));
    const size_t per_block = smem + fa.sharedSizeBytes + 1024;
    GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   (int)std::min<size_t>(100, ((size_t)per_sm * per_block * 100) / (228 * 1024) + 2)));
  }
  long long grid = std::min<long long>((long long)per_sm * L.sm_count, nitems);
  note_kernel(L, KF_Q1_GATHER, reinterpret_cast<const void*>(kern));
  time_begin(L, KF_Q1_GATHER);
  kern<<<(unsigned)grid, Q1G_ROWS, smem, L.stream>>>(p, values, nrows, (int)nitems, stage_doubles, nbuf);
  time_end(L, KF_Q1_GATHER);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

template <int D, int M>
int launch_q1_qp_dm(Launch& L, const Q1QpParams& p, double* values, bool accumulate)
{
  switch (p.group.kind) {
    case Q1G_LAPLACE_SCALAR: return launch_q1_qp_dmk<D, M, Q1G_LAPLACE_SCALAR>(L, p, values, accumulate);
    case Q1G_MASS: return launch_q1_qp_dmk<D, M, Q1G_MASS>(L, p, values, accumulate);
    default: return launch_q1_qp_dmk<D, M, Q1G_LAPLACE_TENSOR>(L, p, values, accumulate);
  }
}

template <int D>
int launch_q1_qp_d(Launch& L, const Q1QpParams& p, double* values, bool accumulate)
{
  switch (p.group.m) {
    case 1: return launch_q1_qp_dm<D, 1>(L, p, values, accumulate);
    case 2: return launch_q1_qp_dm<D, 2>(L, p, values, accumulate);
    case 3: return launch_q1_qp_dm<D, 3>(L, p, values, accumulate);
    default: return fail(GDTB_ERR_NOT_IMPLEMENTED, "q1_gather_qp: 1 to 3 Gauss points per direction");
  }
}

} // namespace

bool q1_qp_supported(int d, int m, int kind)
{
  (void)kind;
  return d >= 1 && d <= 3 && m >= 1 && m <= 3;
}

int launch_q1_gather_qp(Launch& L, const Q1QpParams& p, double* values, bool accumulate)
{
  switch (p.g.d) {
    case 1: return launch_q1_qp_d<1>(L, p, values, accumulate);
    case 2: return launch_q1_qp_d<2>(L, p, values, accumulate);
    case 3: return launch_q1_qp_d<3>(L, p, values, accumulate);
    default: return fail(GDTB_ERR_INVALID_ARGUMENT, "q1_gather_qp: dimension must be 1, 2 or 3");
  }
}

int q1_halo_items(long long layer_rows, long long layers, bool top)
{
  const long long nrows = (layers + 1) * layer_rows; // the slab's own vertex layers plus the interface layer on top
  const long long nitems = (nrows + Q1G_ROWS - 1) / Q1G_ROWS;
  if (!top)
    return (int)std::min(nitems, (layer_rows + Q1G_ROWS - 1) / Q1G_ROWS); // items with a row below layer_rows
  return (int)(nitems - (nrows - layer_rows) / Q1G_ROWS);                  // items with a row in the last layer
}

int q1_host_rowptr(const GridDev& g, const SpaceDev& sp, long long* rowptr)
{
  const int Nx = (int)g.n[0], Ny = g.d > 1 ? (int)g.n[1] : 1, Nz = g.d > 2 ? (int)g.n[2] : 1;
  long long r = 0;
  for (int iz = 0; iz <= (g.d > 2 ? Nz : 0); ++iz)
    for (int iy = 0; iy <= (g.d > 1 ? Ny : 0); ++iy)
      for (int ix = 0; ix <= Nx; ++ix)
        rowptr[r++] = g.d == 1 ? q1_rowptr<1>(ix, 0, 0, Nx, Ny, Nz)
                               : (g.d == 2 ? q1_rowptr<2>(ix, iy, 0, Nx, Ny, Nz) : q1_rowptr<3>(ix, iy, iz, Nx, Ny, Nz));
  if (r != sp.size)
    return fail(GDTB_ERR_SPACE, "q1_host_rowptr: space size mismatch");
  // one past the last vertex decodes to i_last = N_last + 1 (the S_axis of the last axis saturates)
  rowptr[r] = g.d == 1 ? q1_rowptr<1>(Nx + 1, 0, 0, Nx, Ny, Nz)
                       : (g.d == 2 ? q1_rowptr<2>(0, Ny + 1, 0, Nx, Ny, Nz) : q1_rowptr<3>(0, 0, Nz + 1, Nx, Ny, Nz));
  return GDTB_OK;
}

} // namespace gdtb

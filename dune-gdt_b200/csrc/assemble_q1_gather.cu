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
#include "common.cuh"
#include "kernels.hpp"

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
__device__ __forceinline__ long long S_axis(long long i, long long N)
{
  return i == 0 ? 0 : (i > N ? 3 * N + 1 : 3 * i - 1);
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

template <int D, bool ACCUMULATE>
__global__ void k_q1_gather(const __grid_constant__ Q1GatherParams p, double* __restrict__ values,
                            double* __restrict__ rhs, int chunk, int nchunks, long long nitems, int stage_doubles)
{
  constexpr int NO = 1 << D;       // elements around a vertex
  constexpr int ND = P3<D>::value; // stencil size
  extern __shared__ __align__(16) double smem[];
  const GridDev& g = p.g;
  const int Nx = (int)g.n[0], Ny = D > 1 ? (int)g.n[1] : 1, Nz = D > 2 ? (int)g.n[2] : 1;
  const long long Wx = 3LL * Nx + 1, Wy = D > 1 ? 3LL * Ny + 1 : 1;
  const int lines_y = D == 3 ? Ny + 1 : 1;
  // element range along the last direction (owner-computes slab + ghost layer), vertex range in x
  const int elo = (int)p.elem_lo, ehi = (int)p.elem_hi;
  const int x_lo = D == 1 ? (int)p.row_lo : 0, x_hi = D == 1 ? (int)p.row_hi : Nx + 1;
  int buf = 0;

  for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    // ---- per line (uniform over the CTA) ----------------------------------------------------------------
    const int ch = int(item % nchunks);
    const long long line = item / nchunks;
    int iy = 0, iz = 0;
    if (D == 3) {
      iy = int(line % lines_y);
      iz = int(p.row_lo + line / lines_y);
    } else if (D == 2)
      iy = int(p.row_lo + line);
    const int x0 = x_lo + ch * chunk;
    const int x1 = min(x0 + chunk, x_hi);
    // element validity along y and z: e_k = i_k - 1 + o_k
    bool vy[2] = {true, true}, vz[2] = {true, true};
    if (D == 2) {
      vy[0] = iy - 1 >= elo && iy - 1 < ehi;
      vy[1] = iy >= elo && iy < ehi;
    } else if (D == 3) {
      vy[0] = iy >= 1;
      vy[1] = iy < Ny;
      vz[0] = iz - 1 >= elo && iz - 1 < ehi;
      vz[1] = iz >= elo && iz < ehi;
    }
    // columns that exist in y and z
    const bool cy0 = D > 1 && iy > 0, cy1 = D > 1 && iy < Ny, cz0 = D > 2 && iz > 0, cz1 = D > 2 && iz < Nz;
    const int ny = D > 1 ? 1 + cy0 + cy1 : 1, nz = D > 2 ? 1 + cz0 + cz1 : 1;
    const int c = ny * nz;
    const bool full_yz = (D < 2 || ny == 3) && (D < 3 || nz == 3);
    const int Sx0 = x0 == 0 ? 0 : 3 * x0 - 1;
    const int Sx1 = x1 > Nx ? 3 * Nx + 1 : 3 * x1 - 1;
    long long start;
    if (D == 3)
      start = S_axis(iz, Nz) * Wy * Wx + (long long)nz * (S_axis(iy, Ny) * Wx + (long long)ny * Sx0);
    else if (D == 2)
      start = S_axis(iy, Ny) * Wx + (long long)ny * Sx0;
    else
      start = Sx0;
    start -= p.value_offset;
    const int seg = c * (Sx1 - Sx0);
    // keep the shared-memory and global 16-byte phases equal for the bulk copy
    const int phase = values ? int((reinterpret_cast<unsigned long long>(values + start) >> 3) & 1ULL) : 0;
    double* stage = smem + buf * stage_doubles + phase;

    // ---- per vertex ----------------------------------------------------------------------------------------
    const int ix = x0 + (int)threadIdx.x;
    if (ix < x1) {
      bool vx[2];
      if (D == 1) {
        vx[0] = ix - 1 >= elo && ix - 1 < ehi;
        vx[1] = ix >= elo && ix < ehi;
      } else {
        vx[0] = ix >= 1;
        vx[1] = ix < Nx;
      }
      double acc[ND];
#pragma unroll
      for (int dlt = 0; dlt < ND; ++dlt)
        acc[dlt] = 0.;
      double b = 0.;
      // element index of offset o = 0 (may be out of range; only dereferenced when valid)
      const long long e0 = (long long)(ix - 1) + (long long)Nx * ((D > 1 ? iy - 1 : 0) + (long long)Ny * (D > 2 ? iz - 1 : 0));
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        const int ox = o & 1, oy = (o >> 1) & 1, oz = (o >> 2) & 1;
        const bool valid = vx[ox] && vy[oy] && vz[oz];
        if (valid) {
          const long long e = e0 + ox + (long long)Nx * (oy + (long long)Ny * oz);
          if (p.has_const) {
#pragma unroll
            for (int s = 0; s < NO; ++s)
              acc[delta_index<D>(ox - 1 + (s & 1), oy - 1 + ((s >> 1) & 1), oz - 1 + ((s >> 2) & 1))] +=
                  p.T_const[o][s];
          }
          for (int chn = 0; chn < p.n_elem; ++chn) {
            const double cf = __ldg(p.coef[chn] + e);
#pragma unroll
            for (int s = 0; s < NO; ++s) {
              const int dlt = delta_index<D>(ox - 1 + (s & 1), oy - 1 + ((s >> 1) & 1), oz - 1 + ((s >> 2) & 1));
              acc[dlt] = fma(cf, p.T_elem[chn][o][s], acc[dlt]);
            }
          }
          if (p.rhs_has_const)
            b += p.rhs_const;
          if (p.rhs_has_elem)
            b = fma(p.rhs_elem_scale, __ldg(p.rhs_elem + e), b);
        }
      }
      // rows in CSR order: (dz, dy, dx) ascending over the columns that exist
      if (values) {
        const int Sx = ix == 0 ? 0 : 3 * ix - 1;
        double* row = stage + c * (Sx - Sx0);
        if (full_yz && vx[0] && vx[1] && D > 1) {
#pragma unroll
          for (int k = 0; k < ND; ++k)
            row[k] = acc[k];
        } else {
          int pos = 0;
#pragma unroll
          for (int dz = -1; dz <= 1; ++dz) {
            if (D < 3 ? dz != 0 : (dz < 0 ? !cz0 : (dz > 0 ? !cz1 : false)))
              continue;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
              if (D < 2 ? dy != 0 : (dy < 0 ? !cy0 : (dy > 0 ? !cy1 : false)))
                continue;
#pragma unroll
              for (int dx = -1; dx <= 1; ++dx) {
                if (dx < 0 ? ix == 0 : (dx > 0 ? ix == Nx : false))
                  continue;
                row[pos++] = acc[delta_index<D>(dx, dy, dz)];
              }
            }
          }
        }
      }
      if (p.has_rhs && rhs) {
        if (p.rhs_has_sep) {
          double t = p.rhs_sep_scale * __ldg(p.rhs_sep_tab + ix);
          if (D > 1)
            t *= __ldg(p.rhs_sep_tab + p.rhs_sep_stride + iy);
          if (D > 2)
            t *= __ldg(p.rhs_sep_tab + 2 * p.rhs_sep_stride + iz);
          b += t;
        }
        long long r = ix;
        if (D > 1)
          r += (long long)(Nx + 1) * iy;
        if (D > 2)
          r += (long long)(Nx + 1) * (Ny + 1) * iz;
        r -= p.row_offset;
        if (ACCUMULATE)
          rhs[r] += b;
        else
          rhs[r] = b;
      }
    }

    if (values) {
      if (ACCUMULATE) {
        __syncthreads();
        for (int i = threadIdx.x; i < seg; i += blockDim.x)
          values[start + i] += stage[i];
        __syncthreads();
      } else {
        // double-buffered TMA bulk store: the store of this line drains while the next line is computed
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
          bulk_wait_read1(); // the buffer written two lines ago is free again
        }
        __syncthreads();
        buf ^= 1;
      }
    }
  }
  if (!ACCUMULATE && values && threadIdx.x == 0)
    bulk_wait0();
}

// Separable right-hand side tables: for f(x) = p0 * prod_k F_k(x_k),
//   B_k[i] = sum over the (valid) cells e in {i-1, i} of  sum_q w_q phi_a(xi_q) F_k(lower_e + xi_q * ext_e),
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
        const double lower = g.lo[k] + double(e) * g.h[k];
        const double upper = g.lo[k] + double(e + 1) * g.h[k];
        const double ext = upper - lower;
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
          s += qw[q] * phi[q * 2 + a] * F;
        }
      }
      tab[k * stride + i] = s;
    }
  }
}

} // namespace

int launch_q1_rhs_tables(Launch& L, const GridDev& g, long long elem_lo, long long elem_hi, const FnDev& f, int m,
                         const double* qx, const double* qw, const double* phi, double* tab, long long stride)
{
  k_q1_rhs_tables<<<8, 256, 0, L.stream>>>(g, elem_lo, elem_hi, f, m, qx, qw, phi, tab, stride);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

template <int D>
static int launch_q1_gather_d(Launch& L, const Q1GatherParams& p, double* values, double* rhs, bool accumulate)
{
  const GridDev& g = p.g;
  const long long nvx = D == 1 ? p.row_hi - p.row_lo : g.n[0] + 1;
  const int max_chunk = 512;
  const int nchunks = int((nvx + max_chunk - 1) / max_chunk);
  const int chunk = int((nvx + nchunks - 1) / nchunks);
  const int threads = ((chunk + 31) / 32) * 32;
  long long nlines = 1;
  if (D == 3)
    nlines = (g.n[1] + 1) * (p.row_hi - p.row_lo);
  else if (D == 2)
    nlines = p.row_hi - p.row_lo;
  const long long nitems = nlines * nchunks;
  // two stage buffers (double-buffered bulk store), each padded for the 16-byte phase shift
  const int stage_doubles = ((chunk * P3<D>::value + 2) + 1) & ~1;
  const size_t smem = values ? (size_t)(accumulate ? 1 : 2) * stage_doubles * sizeof(double) : 16;
  auto kern = accumulate ? k_q1_gather<D, true> : k_q1_gather<D, false>;
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  GDTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1)
    return fail(GDTB_ERR_CUDA, "q1_gather: kernel does not fit on an SM");
  long long grid = (long long)per_sm * L.sm_count;
  if (grid > nitems)
    grid = nitems;
  time_begin(L, KF_Q1_GATHER);
  kern<<<(unsigned)grid, threads, smem, L.stream>>>(p, values, rhs, chunk, nchunks, nitems, stage_doubles);
  time_end(L, KF_Q1_GATHER);
  L.count++;
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
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

} // namespace gdtb

// dune-gdt_b200/csrc/fv_tma.cu -- AdvectionFvOperator::apply (operators/advection-fv.hh:66-83, local/operators/
// advection-fv.hh:127-153, local/numerical-fluxes/upwind.hh:61-73, lax-friedrichs.hh:60-88) for a scalar conservation law
// on a 2D grid, with the source vector staged in shared memory by TMA bulk loads.
//
// Same cell-gather formulation and the same arithmetic as k_fv_march (fv.cu): a thread owns two adjacent cells in x and
// marches along y, the flux through the upper face of row j is reused as the lower-face flux of row j + 1, every cell sums
// (G_up - G_low) / ext_k over its axes and is written once (16-byte stores).  What differs is how u reaches the SM:
//   * a work unit is a strip of FVT_W = 512 columns times a run of rows; the CTA keeps a ring of FVT_NG groups of FVT_G
//     rows (each row: the strip's 512 cells + 2 halo cells either side) in shared memory;
//   * a loader warp issues one cp.async.bulk.shared::cluster.global per row (4 KB, + 16-byte copies for the periodic wrap
//     columns of the first / last strip) and arms the group's "full" mbarrier with the byte count (complete_tx);
//   * the eight marching warps wait on the group's mbarrier phase, read the next row and the x-neighbours from shared
//     memory, and hand a group back through its "empty" mbarrier when its rows are consumed (no block-wide barrier).
// So up to FVT_NG - 1 groups (20 rows = 82 KB per CTA, two CTAs per SM) are in flight ahead of the arithmetic whatever the
// occupancy -- the memory-level parallelism no longer depends on how many warps are resident (k_fv_march: long-scoreboard
// stall 14 per issue).  Units are sized so that their number is a multiple of the resident CTAs.
#include <algorithm>
#include <cstdint>

#include "common.cuh"
#include "fv_tma.hpp"

namespace gdtb {

namespace {

constexpr int FVT_W = 512;          // cells per strip row
constexpr int FVT_T = FVT_W / 2;    // threads per CTA (two cells each)
constexpr int FVT_G = 4;            // rows per group = per mbarrier
constexpr int FVT_NG = 6;           // groups in the ring
constexpr int FVT_ROWD = FVT_W + 4; // doubles per staged row: two halo cells either side (16-byte granularity)

__device__ __forceinline__ void fvt_mbar_init(unsigned long long* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void fvt_mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void fvt_mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "FVT_WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra FVT_DONE_%=;\n"
               "bra FVT_WAIT_%=;\n"
               "FVT_DONE_%=:\n"
               "}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void fvt_bulk_load(double* sdst, const double* gsrc, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// numerical flux along +e_k between the cells L (lower coordinate) and U: see fv.cu::flux_plus for why one formula serves
// inner and periodic wrap faces
template <int NUMFLUX, int KIND>
__device__ __forceinline__ double fvt_flux(const FvParams& p, const int k, const double uL, const double uU)
{
  if (NUMFLUX == GDTB_NUMFLUX_UPWIND) {
    if (KIND == GDTB_FLUX_LINEAR) {
      const double a = p.flux.p[k];
      return a * (a > 0. ? uL : uU);
    }
    const double w = (uL + uU) > 0. ? uL : uU;
    return 0.5 * w * w;
  }
  if (KIND == GDTB_FLUX_LINEAR)
    return (p.flux.p[k] * uL + p.flux.p[k] * uU) * 0.5 + (uL - uU) * (0.5 * p.lf_lambda_linear);
  return (0.5 * uL * uL + 0.5 * uU * uU) * 0.5 + (uL - uU) * (0.5 * fmax(fabs(uL), fabs(uU)));
}

__device__ __forceinline__ void fvt_mbar_arrive(unsigned long long* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// Warp-specialised: warps 0 .. 7 march (two cells per thread), warp 8 is the loader (one lane issues the bulk copies).
// full[slot]: the loader arms it with the byte count, the TMA unit completes it; empty[slot]: the eight marching warps
// arrive when they are done with the group's rows, the loader waits for it before refilling the slot.  No block-wide
// barrier inside the loop.
template <int NUMFLUX, int KIND>
__global__ void __launch_bounds__(FVT_T + 32, 2)
    k_fv_tma(const __grid_constant__ FvParams p, const double* __restrict__ u, double* __restrict__ out, const int strips,
             const int chunks, const int rows_per_chunk)
{
  extern __shared__ __align__(128) double ring[]; // FVT_NG * FVT_G rows of FVT_ROWD doubles
  __shared__ __align__(8) unsigned long long full[FVT_NG], empty[FVT_NG];
  const int n0 = (int)p.g.n[0], n1 = (int)p.g.n[1];
  const bool per0 = (p.g.periodic & 1) && n0 > 1, per1 = (p.g.periodic & 2) && n1 > 1;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < FVT_NG; ++s) {
      fvt_mbar_init(&full[s], 1);
      fvt_mbar_init(&empty[s], FVT_T / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool loader = threadIdx.x >= FVT_T;
  // running group counter of this CTA (identical in the loader and the marching warps): slot = index % FVT_NG, use
  // number = index / FVT_NG
  int slot = 0, use = 0;
  auto advance = [&]() {
    if (++slot == FVT_NG) {
      slot = 0;
      ++use;
    }
  };

  for (int unit = blockIdx.x; unit < strips * chunks; unit += gridDim.x) {
    const int strip = unit % strips, chunk = unit / strips;
    const int x0 = strip * FVT_W;
    const int j0 = chunk * rows_per_chunk, j1 = min(j0 + rows_per_chunk, n1);
    // staged rows: local r = 0 .. nrows - 1 stands for grid row j0 - 1 + r (the row below the run and the row above it)
    const int nrows = j1 - j0 + 2;
    const int ngroups = (nrows + FVT_G - 1) / FVT_G;

    if (loader) {
      if (threadIdx.x == FVT_T) {
        const bool left_inner = strip != 0, right_inner = strip != strips - 1;
        const unsigned row_bytes = FVT_W * 8u + ((left_inner || per0) ? 16u : 0u) + ((right_inner || per0) ? 16u : 0u);
        for (int gg = 0; gg < ngroups; ++gg) {
          if (use > 0)
            fvt_mbar_wait(&empty[slot], unsigned((use - 1) & 1)); // the marching warps are done with the slot's last content
          // rows of the group that exist (a run at a non-periodic domain edge has no row beyond the grid)
          int ra = gg * FVT_G, rb = min(ra + FVT_G, nrows);
          int present = 0;
          for (int r = ra; r < rb; ++r) {
            const int jr = j0 - 1 + r;
            present += (jr >= 0 && jr < n1) || per1 ? 1 : 0;
          }
          fvt_mbar_expect_tx(&full[slot], (unsigned)present * row_bytes);
          for (int r = ra; r < rb; ++r) {
            int jr = j0 - 1 + r;
            if (jr < 0 || jr >= n1) {
              if (!per1)
                continue;
              jr = jr < 0 ? jr + n1 : jr - n1;
            }
            double* dst = ring + (size_t)(slot * FVT_G + (r - ra)) * FVT_ROWD;
            const double* src = u + (long long)jr * n0 + x0;
            // [left halo (2 cells) | FVT_W cells | right halo (2 cells)]; inner strip boundaries are contiguous in memory
            if (left_inner && right_inner)
              fvt_bulk_load(dst, src - 2, FVT_ROWD * 8u, &full[slot]);
            else {
              const unsigned main_bytes = FVT_W * 8u + (left_inner ? 16u : 0u) + (right_inner ? 16u : 0u);
              fvt_bulk_load(dst + (left_inner ? 0 : 2), src - (left_inner ? 2 : 0), main_bytes, &full[slot]);
              if (!left_inner && per0)
                fvt_bulk_load(dst, u + (long long)jr * n0 + n0 - 2, 16u, &full[slot]);
              if (!right_inner && per0)
                fvt_bulk_load(dst + 2 + FVT_W, u + (long long)jr * n0, 16u, &full[slot]);
            }
          }
          advance();
        }
      } else {
        for (int gg = 0; gg < ngroups; ++gg)
          advance();
      }
      continue;
    }

    // ---- marching warps ------------------------------------------------------------------------------------------
    const int lane = threadIdx.x & 31;
    int rslot = slot; // next group to hand back to the loader
    int pos = slot * FVT_G; // ring row of local row 0 (rows of consecutive groups are consecutive in the ring)
    auto next_pos = [](const int q) { return q + 1 == FVT_NG * FVT_G ? 0 : q + 1; };
    auto wait_next_group = [&]() {
      fvt_mbar_wait(&full[slot], unsigned(use & 1));
      advance();
    };
    auto release_group = [&]() {
      __syncwarp();
      if (lane == 0)
        fvt_mbar_arrive(&empty[rslot]);
      if (++rslot == FVT_NG)
        rslot = 0;
    };
    const int cl = 2 + 2 * (int)threadIdx.x; // position of the thread's first cell inside a staged row
    const int ix = x0 + 2 * (int)threadIdx.x;
    const bool x_lo_edge = ix == 0, x_hi_edge = ix + 1 == n0 - 1;
    const bool x_lo = !x_lo_edge || per0, x_hi = !x_hi_edge || per0;
    const double2 rx = __ldg(reinterpret_cast<const double2*>(p.inv_ext[0] + ix));
    const double cxl = x_lo ? rx.x : 0., cxh = x_hi ? rx.y : 0.;

    wait_next_group(); // group 0: rows 0 .. FVT_G - 1
    int waited = 1, released = 0;
    double uc0, uc1, G_low0 = 0., G_low1 = 0.;
    {
      const double2 b = *reinterpret_cast<const double2*>(ring + (size_t)pos * FVT_ROWD + cl);
      pos = next_pos(pos);
      const double2 c = *reinterpret_cast<const double2*>(ring + (size_t)pos * FVT_ROWD + cl);
      uc0 = c.x, uc1 = c.y;
      if (j0 > 0 || per1) { // lower face of the first row of the run
        G_low0 = fvt_flux<NUMFLUX, KIND>(p, 1, b.x, uc0);
        G_low1 = fvt_flux<NUMFLUX, KIND>(p, 1, b.y, uc1);
      }
    }
    const double* prl = p.inv_ext[1] + j0;
    double* po = out + (long long)j0 * n0 + ix;
    for (int r = 1; r <= nrows - 2; ++r) {
      if (((r + 1) & (FVT_G - 1)) == 0) { // row r + 1 is the first of a new group
        wait_next_group();
        ++waited;
      }
      const double* row = ring + (size_t)pos * FVT_ROWD + cl;
      pos = next_pos(pos);
      const double xl = x_lo ? row[-1] : uc0, xr = x_hi ? row[2] : uc1;
      const bool has_up = j0 + r - 1 < n1 - 1 || per1;
      double un0 = uc0, un1 = uc1;
      if (has_up) {
        const double2 n = *reinterpret_cast<const double2*>(ring + (size_t)pos * FVT_ROWD + cl);
        un0 = n.x, un1 = n.y;
      }
      const double rl = __ldg(prl);
      ++prl;
      const double gx0 = fvt_flux<NUMFLUX, KIND>(p, 0, xl, uc0);
      const double gx1 = fvt_flux<NUMFLUX, KIND>(p, 0, uc0, uc1);
      const double gx2 = fvt_flux<NUMFLUX, KIND>(p, 0, uc1, xr);
      const double G_up0 = has_up ? fvt_flux<NUMFLUX, KIND>(p, 1, uc0, un0) : 0.;
      const double G_up1 = has_up ? fvt_flux<NUMFLUX, KIND>(p, 1, uc1, un1) : 0.;
      // same order of operations as k_fv_march (C = 2)
      double acc0 = gx1 * rx.x - gx0 * cxl;
      double acc1 = gx2 * cxh - gx1 * rx.y;
      acc0 += (G_up0 - G_low0) * rl;
      acc1 += (G_up1 - G_low1) * rl;
      const double res0 = p.euler ? uc0 - acc0 * p.dt : acc0;
      const double res1 = p.euler ? uc1 - acc1 * p.dt : acc1;
      *reinterpret_cast<double2*>(po) = make_double2(res0, res1);
      po += n0;
      G_low0 = G_up0, G_low1 = G_up1;
      uc0 = un0, uc1 = un1;
      if ((r & (FVT_G - 1)) == FVT_G - 1) { // rows <= r are consumed: the group of row r goes back to the loader
        release_group();
        ++released;
      }
    }
    // the groups of the run's tail (all of them have been waited for: waited == ngroups)
    for (; released < ngroups; ++released)
      release_group();
    (void)waited;
  }
}

} // namespace

bool fv_tma_eligible(const FvParams& p, const double* u, const double* out)
{
  const GridDev& g = p.g;
  return g.d == 2 && g.n[0] % FVT_W == 0 && g.n[1] >= 2 && !p.ghosted && !p.p2p && p.n_stage < 0 && p.out_mode == 0
         && (p.bnd_ext_mask | p.bnd_nf_mask) == 0 && p.apply_lo == 0 && p.apply_hi == g.n[1]
         && ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(p.inv_ext[0])) & 15) == 0
         && g.n[0] * g.n[1] < (1LL << 31);
}

int launch_fv_tma(Launch& L, const FvParams& p, const double* u, double* out)
{
  const GridDev& g = p.g;
  const int strips = (int)(g.n[0] / FVT_W);
  const size_t smem = sizeof(double) * (size_t)FVT_NG * FVT_G * FVT_ROWD;
  auto kern = p.flux.numflux == GDTB_NUMFLUX_LAX_FRIEDRICHS
                  ? (p.flux.kind == GDTB_FLUX_BURGERS ? k_fv_tma<GDTB_NUMFLUX_LAX_FRIEDRICHS, GDTB_FLUX_BURGERS>
                                                      : k_fv_tma<GDTB_NUMFLUX_LAX_FRIEDRICHS, GDTB_FLUX_LINEAR>)
                  : (p.flux.kind == GDTB_FLUX_BURGERS ? k_fv_tma<GDTB_NUMFLUX_UPWIND, GDTB_FLUX_BURGERS>
                                                      : k_fv_tma<GDTB_NUMFLUX_UPWIND, GDTB_FLUX_LINEAR>);
  GDTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  GDTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FVT_T + 32, smem));
  if (per_sm < 1)
    return fail(GDTB_ERR_CUDA, "fv_tma: kernel does not fit on an SM");
  per_sm = std::min(per_sm, 2);
  const long long resident = (long long)per_sm * L.sm_count;
  // units = strips x chunks: a multiple of the resident CTAs with runs of about 56 rows (halo re-read 2 / 56)
  long long chunks = std::max<long long>(1, (g.n[1] + 55) / 56);
  const long long want_units = ((strips * chunks + resident - 1) / resident) * resident;
  chunks = std::max<long long>(1, std::min<long long>(g.n[1] / 2, want_units / strips));
  const int rows_per_chunk = (int)((g.n[1] + chunks - 1) / chunks);
  chunks = (g.n[1] + rows_per_chunk - 1) / rows_per_chunk;
  const long long units = strips * chunks;
  const unsigned grid = (unsigned)std::min<long long>(resident, units);
  kern<<<grid, FVT_T + 32, smem, L.stream>>>(p, u, out, strips, (int)chunks, rows_per_chunk);
  GDTB_CUDA(cudaGetLastError());
  return GDTB_OK;
}

} // namespace gdtb

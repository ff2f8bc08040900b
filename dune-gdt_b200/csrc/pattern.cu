// dune-gdt_b200/csrc/pattern.cu -- sparsity patterns on the device.
//
// Replaces make_{element,intersection,element_and_intersection}_sparsity_pattern
// (dune/gdt/tools/sparsity-pattern.hh:34-144): the reference inserts (row, col) pairs per element (and per
// neighbour) into a vector-of-vectors with a linear duplicate search and sorts the rows afterwards.  Here:
//   pattern_sort_unique      : every (entity, ii, jj) emits one 64-bit key row << 32 | col, the keys are radix
//                              sorted, made unique and cut into CSR rows -- works for every space/stencil.
//   pattern_structured_cg_q1 : closed-form tensor-product stencil of CG Q1 on a cube grid (one thread per row).
// Both produce rowptr int64 / colidx int32 with ascending unique columns per row, i.e. exactly the container
// XT::LA::SparsityPatternDefault ends up as after sort() [EXT].
#include <algorithm>

#include <cub/cub.cuh>

#include "common.cuh"
#include "kernels.hpp"

namespace gdtb {

namespace {

// slots per element: own block (element stencils) + one block per face neighbour (intersection stencils)
__global__ void k_emit_keys(const GridDev g, const SpaceDev test, const SpaceDev ansatz, int stencil,
                            unsigned long long* __restrict__ keys, long long n_elements, int slots)
{
  const int nt = test.nloc, na = ansatz.nloc;
  const long long per_elem = (long long)slots * nt * na;
  const long long total = n_elements * per_elem;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long e = t / per_elem;
    int r = int(t % per_elem);
    const int slot = r / (nt * na);
    r -= slot * nt * na;
    const int ii = r / na, jj = r % na;
    long long idx[3], nb[3];
    elem_coords(g, e, idx);
    const long long row = global_index(g, test, idx, ii);
    long long col;
    bool valid = true;
    const bool own = stencil != GDTB_STENCIL_INTERSECTION && slot == 0;
    if (own)
      col = global_index(g, ansatz, idx, jj);
    else {
      const int face = stencil == GDTB_STENCIL_INTERSECTION ? slot : slot - 1;
      bool boundary;
      valid = face_neighbor(g, idx, face / 2, face % 2, nb, &boundary);
      col = valid ? global_index(g, ansatz, nb, jj) : 0;
    }
    // invalid slots repeat the largest key so that they vanish in the unique pass (dropped at the end)
    keys[t] = valid ? ((unsigned long long)row << 32) | (unsigned long long)(unsigned)col : ~0ULL;
  }
}

__global__ void k_rowptr_from_keys(const unsigned long long* __restrict__ keys, long long nkeys, long long rows,
                                   long long* __restrict__ rowptr)
{
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= rows;
       r += (long long)gridDim.x * blockDim.x) {
    // first key >= r << 32
    const unsigned long long target = (unsigned long long)r << 32;
    long long lo = 0, hi = nkeys;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (keys[mid] < target)
        lo = mid + 1;
      else
        hi = mid;
    }
    rowptr[r] = lo;
  }
}

__global__ void k_cols_from_keys(const unsigned long long* __restrict__ keys, long long nnz, int* __restrict__ colidx)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
       i += (long long)gridDim.x * blockDim.x)
    colidx[i] = (int)(unsigned)(keys[i] & 0xffffffffULL);
}

__device__ __forceinline__ long long S_axis(long long i, long long N)
{
  return i == 0 ? 0 : (i > N ? 3 * N + 1 : 3 * i - 1);
}

// CG Q1 element stencil: row of vertex (ix,iy,iz) = all vertices (ix+dx, iy+dy, iz+dz) inside the grid,
// ascending in the vertex index = lexicographic in (dz, dy, dx)
__global__ void k_structured_cg_q1(const GridDev g, long long rows, long long* __restrict__ rowptr,
                                   int* __restrict__ colidx)
{
  const int d = g.d;
  const long long Nx = g.n[0], Ny = d > 1 ? g.n[1] : 0, Nz = d > 2 ? g.n[2] : 0;
  const long long Vx = Nx + 1, Vy = d > 1 ? Ny + 1 : 1;
  const long long Wx = 3 * Nx + 1, Wy = d > 1 ? 3 * Ny + 1 : 1, Wz = d > 2 ? 3 * Nz + 1 : 1;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= rows;
       r += (long long)gridDim.x * blockDim.x) {
    if (r == rows) {
      rowptr[r] = Wx * Wy * Wz;
      continue;
    }
    const long long ix = r % Vx, iy = (r / Vx) % Vy, iz = r / (Vx * Vy);
    const int ny = d > 1 ? ((iy == 0 || iy == Ny) ? 2 : 3) : 1;
    const int nz = d > 2 ? ((iz == 0 || iz == Nz) ? 2 : 3) : 1;
    long long pos = (d > 2 ? S_axis(iz, Nz) : 0) * Wy * Wx + nz * ((d > 1 ? S_axis(iy, Ny) : 0) * Wx + ny * S_axis(ix, Nx));
    rowptr[r] = pos;
    for (int dz = -1; dz <= 1; ++dz) {
      if (d < 3 ? dz != 0 : (iz + dz < 0 || iz + dz > Nz))
        continue;
      for (int dy = -1; dy <= 1; ++dy) {
        if (d < 2 ? dy != 0 : (iy + dy < 0 || iy + dy > Ny))
          continue;
        for (int dx = -1; dx <= 1; ++dx) {
          if (ix + dx < 0 || ix + dx > Nx)
            continue;
          colidx[pos++] = (int)((ix + dx) + Vx * ((iy + dy) + Vy * (iz + dz)));
        }
      }
    }
  }
}

} // namespace

int pattern_sort_unique(Launch& L, const GridDev& g, const SpaceDev& test, const SpaceDev& ansatz, int stencil,
                        long long** d_rowptr, int** d_colidx, long long* nnz_out)
{
  const int slots = stencil == GDTB_STENCIL_ELEMENT ? 1 : (stencil == GDTB_STENCIL_INTERSECTION ? 2 * g.d : 1 + 2 * g.d);
  const long long nkeys = g.ne * slots * test.nloc * ansatz.nloc;
  if (test.size >= (1LL << 31) || ansatz.size >= (1LL << 31))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "pattern: more than 2^31 DoFs per space is not supported (int32 colidx)");
  unsigned long long *keys_a = nullptr, *keys_b = nullptr, *uniq = nullptr;
  long long* d_num = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() {
    cudaFree(keys_a);
    cudaFree(keys_b);
    cudaFree(uniq);
    cudaFree(d_num);
    cudaFree(tmp);
  };
#define PAT_CUDA(call)                                                                                                 \
  do {                                                                                                                 \
    cudaError_t err__ = (call);                                                                                        \
    if (err__ != cudaSuccess) {                                                                                        \
      cleanup();                                                                                                       \
      return fail(err__ == cudaErrorMemoryAllocation ? GDTB_ERR_OUT_OF_MEMORY : GDTB_ERR_CUDA,                         \
                  std::string(#call) + ": " + cudaGetErrorString(err__));                                              \
    }                                                                                                                  \
  } while (0)
  PAT_CUDA(cudaMalloc(&keys_a, sizeof(unsigned long long) * (size_t)nkeys));
  PAT_CUDA(cudaMalloc(&keys_b, sizeof(unsigned long long) * (size_t)nkeys));
  PAT_CUDA(cudaMalloc(&d_num, sizeof(long long)));
  {
    const int block = 256;
    const long long want = (nkeys + block - 1) / block;
    const unsigned grid = (unsigned)std::min<long long>(want, (long long)L.sm_count * 32);
    k_emit_keys<<<grid, block, 0, L.stream>>>(g, test, ansatz, stencil, keys_a, g.ne, slots);
    L.count++;
    PAT_CUDA(cudaGetLastError());
  }
  // radix sort over the significant bits only
  int row_bits = 1;
  while ((1LL << row_bits) < test.size + 1)
    ++row_bits;
  const int end_bit = 64; // invalid keys are all ones; keep them last
  (void)row_bits;
  cub::DoubleBuffer<unsigned long long> buf(keys_a, keys_b);
  size_t tmp_bytes = 0;
  PAT_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, buf, nkeys, 0, end_bit, L.stream));
  PAT_CUDA(cudaMalloc(&tmp, tmp_bytes));
  PAT_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, buf, nkeys, 0, end_bit, L.stream));
  L.count += 4;
  cudaFree(tmp);
  tmp = nullptr;
  unsigned long long* sorted = buf.Current();
  unsigned long long* other = buf.Alternate();
  size_t tmp2 = 0;
  PAT_CUDA(cub::DeviceSelect::Unique(nullptr, tmp2, sorted, other, d_num, nkeys, L.stream));
  PAT_CUDA(cudaMalloc(&tmp, tmp2));
  PAT_CUDA(cub::DeviceSelect::Unique(tmp, tmp2, sorted, other, d_num, nkeys, L.stream));
  L.count += 2;
  long long num = 0;
  PAT_CUDA(cudaMemcpyAsync(&num, d_num, sizeof(long long), cudaMemcpyDeviceToHost, L.stream));
  PAT_CUDA(cudaStreamSynchronize(L.stream));
  // drop the sentinel of invalid slots (largest key, at most one after unique)
  unsigned long long lastkey = 0;
  if (num > 0) {
    PAT_CUDA(cudaMemcpy(&lastkey, other + (num - 1), sizeof(lastkey), cudaMemcpyDeviceToHost));
    if (lastkey == ~0ULL)
      --num;
  }
  long long* rowptr = nullptr;
  int* colidx = nullptr;
  PAT_CUDA(cudaMalloc(&rowptr, sizeof(long long) * (size_t)(test.size + 1)));
  if (cudaMalloc(&colidx, sizeof(int) * (size_t)std::max<long long>(num, 1)) != cudaSuccess) {
    cudaFree(rowptr);
    cleanup();
    return fail(GDTB_ERR_OUT_OF_MEMORY, "pattern: out of device memory for colidx");
  }
  {
    const int block = 256;
    unsigned grid = (unsigned)std::min<long long>((test.size + 1 + block - 1) / block, (long long)L.sm_count * 32);
    k_rowptr_from_keys<<<grid, block, 0, L.stream>>>(other, num, test.size, rowptr);
    grid = (unsigned)std::min<long long>((num + block - 1) / block + 1, (long long)L.sm_count * 32);
    k_cols_from_keys<<<grid, block, 0, L.stream>>>(other, num, colidx);
    L.count += 2;
    cudaError_t err = cudaStreamSynchronize(L.stream);
    if (err != cudaSuccess) {
      cudaFree(rowptr);
      cudaFree(colidx);
      cleanup();
      return fail(GDTB_ERR_CUDA, std::string("pattern kernels: ") + cudaGetErrorString(err));
    }
  }
  cleanup();
#undef PAT_CUDA
  *d_rowptr = rowptr;
  *d_colidx = colidx;
  *nnz_out = num;
  return GDTB_OK;
}

int pattern_structured_cg_q1(Launch& L, const GridDev& g, const SpaceDev& sp, long long** d_rowptr, int** d_colidx,
                             long long* nnz_out)
{
  long long nnz = 1;
  for (int k = 0; k < g.d; ++k)
    nnz *= 3 * g.n[k] + 1;
  long long* rowptr = nullptr;
  int* colidx = nullptr;
  if (cudaMalloc(&rowptr, sizeof(long long) * (size_t)(sp.size + 1)) != cudaSuccess
      || cudaMalloc(&colidx, sizeof(int) * (size_t)nnz) != cudaSuccess) {
    cudaFree(rowptr);
    return fail(GDTB_ERR_OUT_OF_MEMORY, "pattern: out of device memory");
  }
  const int block = 256;
  const unsigned grid = (unsigned)std::min<long long>((sp.size + 1 + block - 1) / block, (long long)L.sm_count * 64);
  k_structured_cg_q1<<<grid, block, 0, L.stream>>>(g, sp.size, rowptr, colidx);
  L.count++;
  cudaError_t err = cudaStreamSynchronize(L.stream);
  if (err != cudaSuccess) {
    cudaFree(rowptr);
    cudaFree(colidx);
    return fail(GDTB_ERR_CUDA, std::string("k_structured_cg_q1: ") + cudaGetErrorString(err));
  }
  *d_rowptr = rowptr;
  *d_colidx = colidx;
  *nnz_out = nnz;
  return GDTB_OK;
}

} // namespace gdtb

// tools/microbench/fp64_peak.cu -- achievable FP64 throughput of one B200: dependent-chain DFMA (vector pipe) and
// mma.sync.aligned.m8n8k4.f64 (DMMA) loops, 8 independent accumulators per thread, all SMs busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b)
{
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      acc[i] = fma(acc[i], a, b);
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b)
{
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i)
    c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < NACC; ++i)
    s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k4 / m16n8k8 / m16n8k16 f64 shapes (sm_90+)
__device__ __forceinline__ void dmma16x8x4(double (&c)[4], double a0, double a1, double b)
{
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma16x8x8(double (&c)[4], const double (&a)[4], const double (&b)[2])
{
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

__global__ void __launch_bounds__(256) k_dmma16x8x4(double* out, int iters, double a, double b)
{
  double c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      c[i][j] = threadIdx.x + i + j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      dmma16x8x4(c[i], a, b, a);
  }
  double s = 0.;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma16x8x8(double* out, int iters, double a, double b)
{
  double c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      c[i][j] = threadIdx.x + i + j;
  const double av[4] = {a, b, a, b};
  const double bv[2] = {b, a};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      dmma16x8x8(c[i], av, bv);
  }
  double s = 0.;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA and DMMA interleaved in every warp: do the two share one datapath or run side by side?
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b)
{
  double acc[8], c[4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    acc[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dmma(c[i][0], c[i][1], a, b);
      acc[2 * i] = fma(acc[2 * i], a, b);
      acc[2 * i + 1] = fma(acc[2 * i + 1], a, b);
    }
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    s += acc[i];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F f)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  return best;
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  const double n_thr = double(blocks) * threads;
  double ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  const double dfma = 2. * 8 * iters * n_thr / (ms * 1e-3) / 1e12;
  ms = time_ms([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  const double dmma8 = 2. * 256 * 8 * iters * (n_thr / 32) / (ms * 1e-3) / 1e12;
  ms = time_ms([&] { k_dmma<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  const double dmma4 = 2. * 256 * 4 * iters * (n_thr / 32) / (ms * 1e-3) / 1e12;
  ms = time_ms([&] { k_dmma16x8x4<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  const double d1684 = 2. * 512 * 4 * iters * (n_thr / 32) / (ms * 1e-3) / 1e12;
  ms = time_ms([&] { k_dmma16x8x8<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  const double d1688 = 2. * 1024 * 4 * iters * (n_thr / 32) / (ms * 1e-3) / 1e12;
  ms = time_ms([&] { k_mixed<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  // per iteration and warp: 4 DMMA (256 FMA each) + 8 DFMA warp instructions (32 FMA each)
  const double mixed = 2. * (4 * 256 + 8 * 32) * iters * (n_thr / 32) / (ms * 1e-3) / 1e12;
  printf("{\"mixed_dfma_dmma_tflops\": %.2f}\n", mixed);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops_8acc\": %.2f, "
         "\"dmma_m8n8k4_tflops_4acc\": %.2f, \"dmma_m16n8k4_tflops\": %.2f, \"dmma_m16n8k8_tflops\": %.2f}\n",
         p.name, p.multiProcessorCount, dfma, dmma8, dmma4, d1684, d1688);
  return 0;
}

// dune-gdt_b200/csrc/common.cuh -- shared host/device definitions of libgdtb (sm_100a only).
//
// Device-side mirrors of the descriptors in include/gdtb.h plus the small geometric / index helpers
// every kernel needs.  The conventions restated here ([EXT] = dune-grid / dune-geometry /
// dune-localfunctions / dune-xt behaviour):
//   * YaspGrid<d, EquidistantOffsetCoordinates>: element index e = ex + Nx (ey + Ny ez); cell
//     [lower, upper] = origin + {i, i+1} * h; J^{-T} = diag(1/ext), integrationElement = prod ext.
//   * ContinuousMapper (dune/gdt/spaces/mapper/continuous.hh:117-150) on top of MCMGMapper [EXT].
//   * DiscontinuousMapper / FiniteVolumeMapper: e * n_loc + i (spaces/mapper/discontinuous.hh:112-132,
//     finite-volume.hh:92-108).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/gdtb.h"

namespace gdtb {

#ifdef __CUDACC__
// 256-bit read-only global load (LDG.E.256.CONSTANT, sm_100): p must be 32-byte aligned
__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d)
{
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
#endif

constexpr int MAX_Q1D = 8;    // max Gauss points per direction
constexpr int MAX_K = 3;      // max Lagrange order
constexpr int MAX_NLOC = 64;  // (MAX_K+1)^3

// ---------------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define GDTB_CUDA(call)                                                                                                \
  do {                                                                                                                 \
    cudaError_t err__ = (call);                                                                                        \
    if (err__ != cudaSuccess)                                                                                          \
      return ::gdtb::fail(GDTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));                       \
  } while (0)

#define GDTB_TRY(call)                                                                                                 \
  do {                                                                                                                 \
    int st__ = (call);                                                                                                 \
    if (st__ != GDTB_OK)                                                                                               \
      return st__;                                                                                                     \
  } while (0)

// ---------------------------------------------------------------------------------------------
// device descriptors
// ---------------------------------------------------------------------------------------------
struct GridDev
{
  int d;
  int periodic;
  double lo[3];
  double h[3];
  long long n[3];
  long long ne;
  // element layers [layer_lo, layer_hi) along the last direction handled by this process
  long long layer_lo, layer_hi;
};

struct CGMapDev
{
  long long codim_offset[4];
  long long group_offset[8];
  long long block[4];
};

struct SpaceDev
{
  int kind;
  int K; // polynomial order (FV: 0)
  int d;
  int nloc;
  long long size;
  CGMapDev cg;
};

struct FnDev
{
  int kind;
  int order;
  int builtin;
  int nq; // GDTB_FN_QP_*: quadrature points per element of the sampled array
  double c[9];
  double p[8];
  const double* data;    // device pointer
  const SpaceDev* space; // GDTB_FN_DOF_VECTOR: the discrete function's space (device-resident copy)
};

struct IntegrandDev
{
  int kind;
  int hI_kind;
  double prefactor;
  FnDev diffusion;
  FnDev weight;
};

// one local form lowered for the generic kernels: integrand terms + the Gauss rule it is integrated with
// + the 1D Lagrange tables at the rule's points and at the interval ends
struct FormDev
{
  int n_terms;
  int m; // Gauss points per direction
  double scaling;
  IntegrandDev terms[GDTB_MAX_TERMS];
  double qx[MAX_Q1D];
  double qw[MAX_Q1D];
  double phi[MAX_Q1D][MAX_K + 1];  // phi[q][a]  = phi_a(qx[q])
  double dphi[MAX_Q1D][MAX_K + 1]; // dphi[q][a] = phi_a'(qx[q])
  double phi_end[2][MAX_K + 1];    // phi_a(0), phi_a(1)
  double dphi_end[2][MAX_K + 1];
};

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline void elem_coords(const GridDev& g, long long e, long long* idx)
{
  idx[0] = e % g.n[0];
  idx[1] = (e / g.n[0]) % g.n[1];
  idx[2] = e / (g.n[0] * g.n[1]);
}

__host__ __device__ inline long long elem_index(const GridDev& g, const long long* idx)
{
  return idx[0] + g.n[0] * (idx[1] + g.n[1] * idx[2]);
}

// origin + i * h without FMA contraction: the grid coordinates (and with them the ulp noise of the cell extents
// upper - lower) are the ones the CPU path produces
__host__ __device__ inline double grid_coord(double lo, double h, long long i)
{
#ifdef __CUDA_ARCH__
  return __dadd_rn(lo, __dmul_rn(double(i), h));
#else
  volatile double t = double(i) * h;
  return lo + t;
#endif
}

// AxisAlignedCubeGeometry of element idx: lower corner and extents, computed per cell from the
// grid coordinates like YaspGrid does (upper - lower, so extents carry the same ulp noise)
__host__ __device__ inline void cell_geometry(const GridDev& g, const long long* idx, double* lower, double* ext)
{
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < g.d) {
      lower[k] = grid_coord(g.lo[k], g.h[k], idx[k]);
      const double upper = grid_coord(g.lo[k], g.h[k], idx[k] + 1);
      ext[k] = upper - lower[k];
    } else {
      lower[k] = 0.;
      ext[k] = 1.;
    }
  }
}

// neighbour across face (k, s); returns false when there is none; *boundary as Intersection::boundary()
__host__ __device__ inline bool face_neighbor(const GridDev& g, const long long* idx, int k, int s, long long* nb,
                                              bool* boundary)
{
  nb[0] = idx[0];
  nb[1] = idx[1];
  nb[2] = idx[2];
  const long long t = idx[k] + (s ? 1 : -1);
  if (t >= 0 && t < g.n[k]) {
    nb[k] = t;
    *boundary = false;
    return true;
  }
  *boundary = true;
  if (g.periodic & (1 << k)) {
    nb[k] = (t + g.n[k]) % g.n[k];
    return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// mappers
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline long long cg_global_index(const GridDev& g, const SpaceDev& sp, const long long* e,
                                                     const int* a)
{
  const int K = sp.K, d = sp.d;
  int s = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (k < d && a[k] > 0 && a[k] < K)
      s |= 1 << k;
  int pc = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    pc += (s >> k) & 1;
  const int c = d - pc;
  long long lex = 0, stride = 1, key = 0, kstride = 1;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < d) {
      const bool ext = (s >> k) & 1;
      const long long pos = ext ? e[k] : e[k] + (a[k] == K ? 1 : 0);
      lex += pos * stride;
      stride *= ext ? g.n[k] : g.n[k] + 1;
      if (ext) {
        key += (a[k] - 1) * kstride;
        kstride *= (K - 1);
      }
    }
  }
  return sp.cg.codim_offset[c] + (sp.cg.group_offset[s] + lex) * sp.cg.block[c] + key;
}

// global index of local DoF i (lexicographic tensor index, a_0 fastest) of element idx
__host__ __device__ inline long long global_index(const GridDev& g, const SpaceDev& sp, const long long* idx, int i)
{
  if (sp.kind != GDTB_SPACE_CG)
    return elem_index(g, idx) * sp.nloc + i;
  const int n1 = sp.K + 1;
  int a[3];
  a[0] = i % n1;
  a[1] = sp.d > 1 ? (i / n1) % n1 : 0;
  a[2] = sp.d > 2 ? i / (n1 * n1) : 0;
  return cg_global_index(g, sp, idx, a);
}

// ---------------------------------------------------------------------------------------------
// grid functions
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline double builtin_eval(const FnDev& f, int d, const double* x)
{
  switch (f.builtin) {
    case GDTB_BUILTIN_COS_PRODUCT: {
      double v = f.p[0];
      for (int k = 0; k < d; ++k)
        v *= cos(f.p[1] * x[k]);
      return v;
    }
    case GDTB_BUILTIN_AFFINE: {
      double v = f.p[0];
      for (int k = 0; k < d; ++k)
        v += f.p[1 + k] * x[k];
      return v;
    }
    case GDTB_BUILTIN_GAUSSIAN: {
      const double t = x[0] - f.p[0];
      return exp(-(t * t) / (2. * (f.p[1] * f.p[1])));
    }
    case GDTB_BUILTIN_INDICATOR:
      return (f.p[0] <= x[0] && x[0] <= f.p[1]) ? 1. : 0.;
    case GDTB_BUILTIN_QUADRATIC: {
      double s = 0.;
      for (int k = 0; k < d; ++k)
        s += x[k] * x[k];
      return f.p[0] + f.p[1] * s;
    }
    default:
      return 0.;
  }
}

// gradient of the analytic built-ins (norm evaluation against an exact solution, examples/stationary-heat-equation.cc:
// 71-85 passes the jacobian lambda explicitly)
__host__ __device__ inline void builtin_grad(const FnDev& f, int d, const double* x, double* grad)
{
  grad[0] = grad[1] = grad[2] = 0.;
  switch (f.builtin) {
    case GDTB_BUILTIN_COS_PRODUCT:
      for (int k = 0; k < d; ++k) {
        double v = -f.p[0] * f.p[1] * sin(f.p[1] * x[k]);
        for (int j = 0; j < d; ++j)
          if (j != k)
            v *= cos(f.p[1] * x[j]);
        grad[k] = v;
      }
      break;
    case GDTB_BUILTIN_AFFINE:
      for (int k = 0; k < d; ++k)
        grad[k] = f.p[1 + k];
      break;
    case GDTB_BUILTIN_GAUSSIAN: {
      const double t = x[0] - f.p[0];
      grad[0] = -(t / (f.p[1] * f.p[1])) * exp(-(t * t) / (2. * (f.p[1] * f.p[1])));
      break;
    }
    case GDTB_BUILTIN_QUADRATIC:
      for (int k = 0; k < d; ++k)
        grad[k] = 2. * f.p[1] * x[k];
      break;
    default:
      break;
  }
}

// 1D Lagrange basis of order K on the equidistant nodes a / K at x in [0, 1] (values only)
__host__ __device__ inline void lagrange_values_1d(int K, double x, double* v)
{
  if (K == 0) {
    v[0] = 1.;
    return;
  }
  for (int a = 0; a <= K; ++a) {
    const double ta = double(a) / K;
    double val = 1.;
    for (int b = 0; b <= K; ++b)
      if (b != a)
        val *= (x - double(b) / K) / (ta - double(b) / K);
    v[a] = val;
  }
}

// Where a grid function is evaluated: index of the point inside the element's volume rule (GDTB_FN_QP_*: caller-sampled
// arrays are indexed [element][q]; q < 0: no rule context) and the element / reference point (GDTB_FN_DOF_VECTOR)
struct EvalPt
{
  int q;
  const long long* idx;
  const double* xh;
};

// u_h(x) = sum_i dofs[global_index(e, i)] phi_i(xhat) (LocalDiscreteFunction::evaluate, discretefunction/default.hh)
__device__ inline double dof_vector_eval(const FnDev& f, const GridDev& g, const long long* idx, const double* xh)
{
  const SpaceDev sp = *f.space;
  double v[3][MAX_K + 1];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < g.d)
      lagrange_values_1d(sp.K, xh[k], v[k]);
    else
      v[k][0] = 1.;
  }
  const int n1 = sp.K + 1;
  double u = 0.;
  for (int i = 0; i < sp.nloc; ++i) {
    const int a0 = i % n1, a1 = g.d > 1 ? (i / n1) % n1 : 0, a2 = g.d > 2 ? i / (n1 * n1) : 0;
    u += __ldg(f.data + global_index(g, sp, idx, i)) * (v[0][a0] * v[1][a1] * v[2][a2]);
  }
  return u;
}

__device__ inline double fn_scalar(const FnDev& f, const GridDev& g, long long e, const double* x, const EvalPt& pt)
{
  switch (f.kind) {
    case GDTB_FN_ELEM_SCALAR:
      return __ldg(f.data + e);
    case GDTB_FN_BUILTIN:
      return builtin_eval(f, g.d, x);
    case GDTB_FN_QP_SCALAR:
      return __ldg(f.data + e * f.nq + pt.q);
    case GDTB_FN_DOF_VECTOR:
      return dof_vector_eval(f, g, pt.idx, pt.xh);
    default:
      return f.c[0];
  }
}

// d x d tensor, row-major with leading dimension 3; scalar kinds mean c * I (laplace.hh:41)
__device__ inline void fn_tensor(const FnDev& f, const GridDev& g, long long e, const double* x, const EvalPt& pt,
                                 double* T)
{
  const int d = g.d;
#pragma unroll
  for (int i = 0; i < 9; ++i)
    T[i] = 0.;
  if (f.kind == GDTB_FN_CONST_TENSOR) {
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c)
        T[r * 3 + c] = f.c[r * d + c];
  } else if (f.kind == GDTB_FN_ELEM_TENSOR) {
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c)
        T[r * 3 + c] = __ldg(f.data + e * d * d + r * d + c);
  } else if (f.kind == GDTB_FN_QP_TENSOR) {
    const double* src = f.data + (e * f.nq + pt.q) * (d * d);
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c)
        T[r * 3 + c] = __ldg(src + r * d + c);
  } else {
    const double s = fn_scalar(f, g, e, x, pt);
    for (int r = 0; r < d; ++r)
      T[r * 3 + r] = s;
  }
}

// variants without a point context (constants, per-element data, analytic built-ins)
__device__ inline double fn_scalar(const FnDev& f, int d, long long e, const double* x)
{
  switch (f.kind) {
    case GDTB_FN_ELEM_SCALAR:
      return __ldg(f.data + e);
    case GDTB_FN_BUILTIN:
      return builtin_eval(f, d, x);
    default:
      return f.c[0];
  }
}

// d x d tensor, row-major with leading dimension 3; scalar kinds mean c * I (laplace.hh:41)
__device__ inline void fn_tensor(const FnDev& f, int d, long long e, const double* x, double* T)
{
#pragma unroll
  for (int i = 0; i < 9; ++i)
    T[i] = 0.;
  if (f.kind == GDTB_FN_CONST_TENSOR) {
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c)
        T[r * 3 + c] = f.c[r * d + c];
  } else if (f.kind == GDTB_FN_ELEM_TENSOR) {
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c)
        T[r * 3 + c] = __ldg(f.data + e * d * d + r * d + c);
  } else {
    const double s = fn_scalar(f, d, e, x);
    for (int r = 0; r < d; ++r)
      T[r * 3 + r] = s;
  }
}

__host__ __device__ inline int ipow(int b, int e)
{
  int r = 1;
  for (int i = 0; i < e; ++i)
    r *= b;
  return r;
}

// position of column `col` in the sorted CSR row [b, e); -1 if absent
__device__ inline long long csr_find(const int* __restrict__ colidx, long long b, long long e, int col)
{
  long long lo = b, hi = e;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(colidx + mid) < col)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < e && __ldg(colidx + lo) == col) ? lo : -1;
}

} // namespace gdtb

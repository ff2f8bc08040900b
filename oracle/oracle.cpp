// oracle/oracle.cpp -- CPU restatement of dune-gdt's assembly / FV-apply hot path.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Written to follow the reference's loop structure and
// operation order, not to be fast.  Every function cites the reference
// file:line it restates; [EXT] marks conventions of dune-grid / dune-geometry / dune-localfunctions /
// dune-xt, whose sources are not part of the reference checkout ("parity unpinned", see oracle.h).
#include "oracle.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_error;

constexpr int MAXN = 64; // max local DoFs: Q3 in 3D
constexpr int MAXQ1D = 8;

// ------------------------------------------------------------------------------------------------
// grid [EXT: Dune::YaspGrid<d, EquidistantOffsetCoordinates<double, d>>], SURVEY Appendix B
// ------------------------------------------------------------------------------------------------
struct Grid
{
  int d;
  int periodic;
  double lo[3], up[3], h[3];
  int64_t n[3];
  int64_t ne;

  explicit Grid(const orc_grid* g)
  {
    d = g->dim;
    periodic = g->periodic;
    ne = 1;
    for (int k = 0; k < 3; ++k) {
      lo[k] = k < d ? g->lower[k] : 0.;
      up[k] = k < d ? g->upper[k] : 1.;
      n[k] = k < d ? g->n[k] : 1;
      h[k] = (up[k] - lo[k]) / double(n[k]); // EquidistantOffsetCoordinates::_h
      ne *= n[k];
    }
  }

  // element index e = ex + Nx (ey + Ny ez)
  void coords(int64_t e, int64_t* idx) const
  {
    idx[0] = e % n[0];
    idx[1] = (e / n[0]) % n[1];
    idx[2] = e / (n[0] * n[1]);
  }
  int64_t index(const int64_t* idx) const
  {
    return idx[0] + n[0] * (idx[1] + n[1] * idx[2]);
  }
  // AxisAlignedCubeGeometry(lower, upper): corners from coordinate(k, i) = origin + i*h
  void cell(const int64_t* idx, double* lower, double* ext) const
  {
    for (int k = 0; k < 3; ++k) {
      if (k < d) {
        lower[k] = lo[k] + double(idx[k]) * h[k];
        const double upper = lo[k] + double(idx[k] + 1) * h[k];
        ext[k] = upper - lower[k];
      } else {
        lower[k] = 0.;
        ext[k] = 1.;
      }
    }
  }
  double volume(const double* ext) const
  {
    double v = 1.;
    for (int k = 0; k < d; ++k)
      v *= ext[k];
    return v;
  }
  // intersection (direction k, side s) of element idx: neighbour / boundary flags + neighbour coords
  // [EXT] intersection order = indexInInside: x-, x+, y-, y+, z-, z+
  bool neighbor(const int64_t* idx, int k, int s, int64_t* nb, bool* boundary) const
  {
    for (int j = 0; j < 3; ++j)
      nb[j] = idx[j];
    const int64_t t = idx[k] + (s ? 1 : -1);
    if (t >= 0 && t < n[k]) {
      nb[k] = t;
      *boundary = false;
      return true;
    }
    *boundary = true;
    if (periodic & (1 << k)) { // XT::Grid::PeriodicGridView: neighbor() && boundary()
      nb[k] = (t + n[k]) % n[k];
      return true;
    }
    return false;
  }
};

// ------------------------------------------------------------------------------------------------
// Gauss-Legendre rules [EXT dune-geometry QuadratureRules<D,d>::rule(cube, order)]: tensor rule with
// m = floor(order/2)+1 points per direction; points on [0,1].
// ------------------------------------------------------------------------------------------------
int gauss_m(int order)
{
  return std::max(order, 0) / 2 + 1;
}

void gauss01(int m, double* x, double* w)
{
  for (int i = 0; i < m; ++i) {
    long double z = std::cos(M_PIl * (i + 0.75L) / (m + 0.5L));
    long double pp = 1.;
    for (int it = 0; it < 100; ++it) {
      long double p1 = 1., p2 = 0.;
      for (int j = 0; j < m; ++j) {
        const long double p3 = p2;
        p2 = p1;
        p1 = ((2.0L * j + 1.0L) * z * p2 - j * p3) / (j + 1.0L);
      }
      pp = m * (z * p1 - p2) / (z * z - 1.0L);
      const long double z1 = z;
      z = z1 - p1 / pp;
      if (std::fabs((double)(z - z1)) < 1e-19)
        break;
    }
    // ascending order
    x[m - 1 - i] = (double)((1.0L + z) / 2.0L);
    w[m - 1 - i] = (double)(1.0L / ((1.0L - z * z) * pp * pp));
  }
}

struct Rule
{
  int m;
  double x[MAXQ1D], w[MAXQ1D];
  explicit Rule(int order)
  {
    m = gauss_m(order);
    assert(m <= MAXQ1D);
    gauss01(m, x, w);
  }
};

// ------------------------------------------------------------------------------------------------
// Lagrange Q_k shape functions [EXT dune-localfunctions LagrangeLocalFiniteElement<EquidistantPointSet>],
// wrapped by local/finite-elements/lagrange.hh:140-142.  Nodal basis at the equidistant tensor points;
// local order used here: lexicographic (a_0 fastest).  K = 0: the constant 1 (FV basis,
// spaces/basis/finite-volume.hh:134-141).
// ------------------------------------------------------------------------------------------------
void lagrange1d(int K, double x, double* v, double* dv)
{
  if (K == 0) {
    v[0] = 1.;
    dv[0] = 0.;
    return;
  }
  for (int a = 0; a <= K; ++a) {
    const double ta = double(a) / K;
    double val = 1.;
    for (int b = 0; b <= K; ++b)
      if (b != a)
        val *= (x - double(b) / K) / (ta - double(b) / K);
    double der = 0.;
    for (int c = 0; c <= K; ++c) {
      if (c == a)
        continue;
      double t = 1. / (ta - double(c) / K);
      for (int b = 0; b <= K; ++b)
        if (b != a && b != c)
          t *= (x - double(b) / K) / (ta - double(b) / K);
      der += t;
    }
    v[a] = val;
    dv[a] = der;
  }
}

int local_size(int d, int K)
{
  int n = 1;
  for (int k = 0; k < d; ++k)
    n *= (K + 1);
  return n;
}

// values[i], grads[i*3 + r] (reference gradients)
void shape(int d, int K, const double* xh, double* values, double* grads)
{
  double v[3][8], dv[3][8];
  for (int k = 0; k < 3; ++k) {
    if (k < d)
      lagrange1d(K, xh[k], v[k], dv[k]);
    else {
      v[k][0] = 1.;
      dv[k][0] = 0.;
    }
  }
  const int n1 = K + 1;
  const int ny = d > 1 ? n1 : 1, nz = d > 2 ? n1 : 1;
  int i = 0;
  for (int az = 0; az < nz; ++az)
    for (int ay = 0; ay < ny; ++ay)
      for (int ax = 0; ax < n1; ++ax, ++i) {
        if (values)
          values[i] = v[0][ax] * v[1][ay] * v[2][az];
        if (grads) {
          grads[i * 3 + 0] = dv[0][ax] * v[1][ay] * v[2][az];
          grads[i * 3 + 1] = v[0][ax] * dv[1][ay] * v[2][az];
          grads[i * 3 + 2] = v[0][ax] * v[1][ay] * dv[2][az];
        }
      }
}

// ------------------------------------------------------------------------------------------------
// mappers
// ------------------------------------------------------------------------------------------------
// ContinuousMapper (spaces/mapper/continuous.hh:117-150): global = mcmg.subIndex(e, key.subEntity, key.codim)
// + key.index.  [EXT] MCMGMapper offsets loop codim = 0..d, block = number of local
// keys on that sub-entity (continuous.hh:167-176); [EXT] YaspGrid sub-entity index: entities grouped by
// their shift bitset (bit k set <=> entity extends in direction k), groups ordered by increasing bitset
// value, lexicographic (x fastest) inside a group.  For Q1 this reduces to global = vertex index.
struct CGMap
{
  int d, K;
  int64_t n[3];
  int64_t codim_offset[4];
  int64_t group_offset[8]; // offset (in entities) of the shift group inside its codim
  int64_t block[4];
  int64_t size;

  CGMap(const Grid& g, int K_)
  {
    d = g.d;
    K = K_;
    for (int k = 0; k < 3; ++k)
      n[k] = g.n[k];
    int64_t running = 0;
    for (int c = 0; c <= d; ++c) {
      int64_t b = 1;
      for (int j = 0; j < d - c; ++j)
        b *= (K - 1);
      block[c] = b;
      codim_offset[c] = running;
      int64_t entities = 0;
      for (int s = 0; s < (1 << d); ++s) {
        if (__builtin_popcount(s) != d - c)
          continue;
        group_offset[s] = entities;
        int64_t cnt = 1;
        for (int k = 0; k < d; ++k)
          cnt *= (s >> k & 1) ? n[k] : n[k] + 1;
        entities += cnt;
      }
      running += entities * b; // block == 0 <=> geometry type not in the layout
    }
    size = running;
  }

  int64_t index(const int64_t* e, const int* a) const
  {
    int s = 0;
    for (int k = 0; k < d; ++k)
      if (a[k] > 0 && a[k] < K)
        s |= 1 << k;
    const int c = d - __builtin_popcount(s);
    int64_t lex = 0, stride = 1, key = 0, kstride = 1;
    for (int k = 0; k < d; ++k) {
      const bool ext = s >> k & 1;
      const int64_t pos = ext ? e[k] : e[k] + (a[k] == K ? 1 : 0);
      lex += pos * stride;
      stride *= ext ? n[k] : n[k] + 1;
      if (ext) {
        key += (a[k] - 1) * kstride;
        kstride *= (K - 1);
      }
    }
    return codim_offset[c] + (group_offset[s] + lex) * block[c] + key;
  }
};

struct Space
{
  int kind, K, d, nloc;
  int64_t size;
  Grid g;
  CGMap cg;

  Space(const Grid& g_, int kind_, int order)
    : kind(kind_)
    , K(kind_ == ORC_SPACE_FV ? 0 : order)
    , d(g_.d)
    , g(g_)
    , cg(g_, std::max(1, kind_ == ORC_SPACE_FV ? 1 : order))
  {
    nloc = local_size(d, K);
    // DiscontinuousMapper: offset[e] = running sum of local sizes (spaces/mapper/discontinuous.hh:112-132)
    // FiniteVolumeMapper: e*r + i (spaces/mapper/finite-volume.hh:92-108)
    size = kind == ORC_SPACE_CG ? cg.size : g.ne * nloc;
  }

  void global_indices(const int64_t* idx, int64_t* out) const
  {
    if (kind != ORC_SPACE_CG) {
      const int64_t e = g.index(idx);
      for (int i = 0; i < nloc; ++i)
        out[i] = e * nloc + i;
      return;
    }
    const int n1 = K + 1;
    const int ny = d > 1 ? n1 : 1, nz = d > 2 ? n1 : 1;
    int i = 0;
    for (int az = 0; az < nz; ++az)
      for (int ay = 0; ay < ny; ++ay)
        for (int ax = 0; ax < n1; ++ax, ++i) {
          const int a[3] = {ax, ay, az};
          out[i] = cg.index(idx, a);
        }
  }
};

// ------------------------------------------------------------------------------------------------
// grid functions [EXT XT::Functions::GridFunction]: bound to an element, evaluated at a reference point;
// analytic functions see the global coordinate geometry.global(xhat) = lower + xhat * ext.
// ------------------------------------------------------------------------------------------------
double builtin_eval(const orc_function* f, int d, const double* x)
{
  switch (f->builtin) {
    case ORC_BUILTIN_COS_PRODUCT: {
      double v = f->p[0];
      for (int k = 0; k < d; ++k)
        v *= std::cos(f->p[1] * x[k]);
      return v;
    }
    case ORC_BUILTIN_AFFINE: {
      double v = f->p[0];
      for (int k = 0; k < d; ++k)
        v += f->p[1 + k] * x[k];
      return v;
    }
    case ORC_BUILTIN_GAUSSIAN: {
      const double t = x[0] - f->p[0];
      return std::exp(-(t * t) / (2. * (f->p[1] * f->p[1])));
    }
    case ORC_BUILTIN_INDICATOR:
      return (f->p[0] <= x[0] && x[0] <= f->p[1]) ? 1. : 0.;
    case ORC_BUILTIN_QUADRATIC: {
      double s = 0.;
      for (int k = 0; k < d; ++k)
        s += x[k] * x[k];
      return f->p[0] + f->p[1] * s;
    }
    default:
      return 0.;
  }
}

// Where a grid function is evaluated: the quadrature point index inside the element's rule (ORC_FN_QP_*: the
// caller-sampled array is indexed [element][q]) and the element / reference point (ORC_FN_DOF_VECTOR: a discrete
// function is evaluated through its local basis, discretefunction/default.hh).  q < 0: no rule context.
struct Pt
{
  int q = -1;
  int nq = 0;
  const Grid* g = nullptr;
  const int64_t* idx = nullptr;
  const double* xh = nullptr;
};

double dof_vector_eval(const orc_function* f, const Pt& pt)
{
  // u_h(x) = sum_i dofs[global_index(e, i)] * phi_i(xhat)  (LocalDiscreteFunction::evaluate)
  const Space sp(*pt.g, f->space_kind, f->space_order);
  int64_t gi[MAXN];
  double val[MAXN];
  sp.global_indices(pt.idx, gi);
  shape(sp.d, sp.K, pt.xh, val, nullptr);
  double u = 0.;
  for (int i = 0; i < sp.nloc; ++i)
    u += f->data[gi[i]] * val[i];
  return u;
}

double eval_scalar(const orc_function* f, int d, int64_t e, const double* x, const Pt& pt = Pt())
{
  switch (f->kind) {
    case ORC_FN_CONST_SCALAR:
      return f->c[0];
    case ORC_FN_ELEM_SCALAR:
      return f->data[e];
    case ORC_FN_BUILTIN:
      return builtin_eval(f, d, x);
    case ORC_FN_QP_SCALAR:
      assert(pt.q >= 0 && pt.nq == f->qp_per_element);
      return f->data[e * pt.nq + pt.q];
    case ORC_FN_DOF_VECTOR:
      assert(pt.g && pt.idx && pt.xh);
      return dof_vector_eval(f, pt);
    default:
      return f->c[0];
  }
}

// d x d view (row-major, leading dimension 3); scalar functions mean c * I (laplace.hh:41, ipdg.hh:61)
void eval_tensor(const orc_function* f, int d, int64_t e, const double* x, double* T, const Pt& pt = Pt())
{
  for (int i = 0; i < 9; ++i)
    T[i] = 0.;
  switch (f->kind) {
    case ORC_FN_CONST_TENSOR:
      for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c)
          T[r * 3 + c] = f->c[r * d + c];
      break;
    case ORC_FN_ELEM_TENSOR:
      for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c)
          T[r * 3 + c] = f->data[e * d * d + r * d + c];
      break;
    case ORC_FN_QP_TENSOR:
      assert(pt.q >= 0 && pt.nq == f->qp_per_element);
      for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c)
          T[r * 3 + c] = f->data[(e * pt.nq + pt.q) * d * d + r * d + c];
      break;
    default: {
      const double s = eval_scalar(f, d, e, x, pt);
      for (int r = 0; r < d; ++r)
        T[r * 3 + r] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bound local basis: values + physical gradients at a reference point
// DefaultGlobalBasis::Localized::{evaluate,jacobians} (spaces/basis/default.hh:142-175):
//   g_i = J^{-T} ghat_i by a d x d mat-vec; YaspGrid: J^{-T} = diag(1/ext) so the zero products vanish.
// ------------------------------------------------------------------------------------------------
struct Basis
{
  int d, K, n;
  double val[MAXN];
  double grad[MAXN * 3];

  void evaluate(const double* xh, const double* ext)
  {
    shape(d, K, xh, val, grad);
    for (int i = 0; i < n; ++i)
      for (int r = 0; r < 3; ++r)
        grad[i * 3 + r] = r < d ? (1. / ext[r]) * grad[i * 3 + r] : 0.;
  }
};

inline void matvec(int d, const double* T, const double* g, double* y)
{
  for (int r = 0; r < d; ++r) {
    double s = 0.;
    for (int c = 0; c < d; ++c)
      s += T[r * 3 + c] * g[c];
    y[r] = s;
  }
}
inline double dot(int d, const double* a, const double* b)
{
  double s = 0.;
  for (int r = 0; r < d; ++r)
    s += a[r] * b[r];
  return s;
}

// ------------------------------------------------------------------------------------------------
// element integrands + LocalElementIntegralBilinearForm::apply2 (local/bilinear-forms/integrals.hh:97-134)
// ------------------------------------------------------------------------------------------------
int element_integrand_order(const orc_integrand& t, int p)
{
  // laplace.hh:74-79, product.hh:89-100: weight.order + test.order + ansatz.order
  return t.diffusion.order + p + p;
}

int form_order(const orc_form& f, int p, int (*term_order)(const orc_integrand&, int))
{
  int o = 0;
  for (int t = 0; t < f.n_terms; ++t) // combined.hh:293-299: max of the summands
    o = std::max(o, term_order(f.terms[t], p));
  return o + f.over_integrate;
}

// test and ansatz basis are separate arguments like in the reference (laplace.hh:81-102: test_basis.jacobians /
// ansatz_basis.jacobians; the assembler passes the same space twice) -- the pointwise known-answer tests of
// dune/gdt/test/integrands/integrands_laplace.cc:71-91 and integrands_product.cc:58-76 use different ones.
void element_integrand_evaluate2(const orc_integrand& t, const Basis& test, const Basis& ansatz, int d, int64_t e,
                                 const double* x, double* result /* n_test * n_ansatz, overwritten */,
                                 const Pt& pt = Pt())
{
  const int nt = test.n, na = ansatz.n;
  if (t.kind == ORC_INT_LAPLACE) {
    // laplace.hh:81-102: result[ii][jj] += (weight * ansatz_grads[jj][rr]) * test_grads[ii][rr]
    double kappa[9];
    eval_tensor(&t.diffusion, d, e, x, kappa, pt);
    for (int ii = 0; ii < nt; ++ii)
      for (int jj = 0; jj < na; ++jj) {
        double kg[3];
        matvec(d, kappa, &ansatz.grad[jj * 3], kg);
        result[ii * na + jj] = 0. + dot(d, kg, &test.grad[ii * 3]);
      }
  } else { // ORC_INT_PRODUCT, product.hh:104-130: result[ii][jj] = (weight * test[ii]) * ansatz[jj]
    const double w = eval_scalar(&t.diffusion, d, e, x, pt);
    for (int ii = 0; ii < nt; ++ii)
      for (int jj = 0; jj < na; ++jj)
        result[ii * na + jj] = (w * test.val[ii]) * ansatz.val[jj];
  }
}

void element_integrand_evaluate(const orc_integrand& t, const Basis& b, int d, int64_t e, const double* x,
                                double* result /* n*n, overwritten */, const Pt& pt = Pt())
{
  element_integrand_evaluate2(t, b, b, d, e, x, result, pt);
}

void local_element_matrix(const Grid& g, const Space& sp, const orc_form& form, const int64_t* idx, double* L)
{
  const int d = g.d, n = sp.nloc;
  const int64_t e = g.index(idx);
  double lower[3], ext[3];
  g.cell(idx, lower, ext);
  const double ie = g.volume(ext); // integrationElement of an axis-aligned cube
  for (int i = 0; i < n * n; ++i)
    L[i] = 0.;
  const Rule rule(form_order(form, sp.K, element_integrand_order));
  const int m = rule.m;
  const int my = d > 1 ? m : 1, mz = d > 2 ? m : 1;
  Basis b;
  b.d = d;
  b.K = sp.K;
  b.n = n;
  double values[MAXN * MAXN], scratch[MAXN * MAXN];
  for (int qz = 0; qz < mz; ++qz)
    for (int qy = 0; qy < my; ++qy)
      for (int qx = 0; qx < m; ++qx) {
        const double xh[3] = {rule.x[qx], d > 1 ? rule.x[qy] : 0., d > 2 ? rule.x[qz] : 0.};
        const double w = rule.w[qx] * (d > 1 ? rule.w[qy] : 1.) * (d > 2 ? rule.w[qz] : 1.);
        double x[3];
        for (int k = 0; k < 3; ++k)
          x[k] = lower[k] + xh[k] * ext[k];
        const double factor = ie * w; // integrals.hh:119
        b.evaluate(xh, ext);
        Pt pt;
        pt.q = qx + m * (qy + m * qz);
        pt.nq = m * my * mz;
        pt.g = &g;
        pt.idx = idx;
        pt.xh = xh;
        element_integrand_evaluate(form.terms[0], b, d, e, x, values, pt);
        for (int t = 1; t < form.n_terms; ++t) { // combined.hh:212-227
          element_integrand_evaluate(form.terms[t], b, d, e, x, scratch, pt);
          for (int i = 0; i < n * n; ++i)
            values[i] += scratch[i];
        }
        for (int i = 0; i < n * n; ++i) // integrals.hh:129-131
          L[i] += values[i] * factor;
      }
}

// LocalElementIntegralFunctional::apply (local/functionals/integrals.hh:72-98) with
// LocalBinaryToUnaryElementIntegrand (conversion.hh:90-117) around LocalElementProductIntegrand:
//   v_i = (w * psi_i) * f(x);  l_i += v_i * ie * w_q
void local_element_vector(const Grid& g, const Space& sp, const orc_form& form, const int64_t* idx, double* l)
{
  const int d = g.d, n = sp.nloc;
  const int64_t e = g.index(idx);
  double lower[3], ext[3];
  g.cell(idx, lower, ext);
  const double ie = g.volume(ext);
  for (int i = 0; i < n; ++i)
    l[i] = 0.;
  const orc_integrand& t = form.terms[0];
  // conversion.hh:92 -> product.hh:99: weight.order + test.order + f.order
  const int order = t.diffusion.order + sp.K + t.weight.order + form.over_integrate;
  const Rule rule(order);
  const int m = rule.m;
  const int my = d > 1 ? m : 1, mz = d > 2 ? m : 1;
  Basis b;
  b.d = d;
  b.K = sp.K;
  b.n = n;
  for (int qz = 0; qz < mz; ++qz)
    for (int qy = 0; qy < my; ++qy)
      for (int qx = 0; qx < m; ++qx) {
        const double xh[3] = {rule.x[qx], d > 1 ? rule.x[qy] : 0., d > 2 ? rule.x[qz] : 0.};
        const double wq = rule.w[qx] * (d > 1 ? rule.w[qy] : 1.) * (d > 2 ? rule.w[qz] : 1.);
        double x[3];
        for (int k = 0; k < 3; ++k)
          x[k] = lower[k] + xh[k] * ext[k];
        b.evaluate(xh, ext);
        Pt pt;
        pt.q = qx + m * (qy + m * qz);
        pt.nq = m * my * mz;
        pt.g = &g;
        pt.idx = idx;
        pt.xh = xh;
        const double w = eval_scalar(&t.diffusion, d, e, x, pt);
        const double f = eval_scalar(&t.weight, d, e, x, pt);
        for (int i = 0; i < n; ++i) {
          const double v = (w * b.val[i]) * f;
          l[i] += v * ie * wq; // local/functionals/integrals.hh:96
        }
      }
}

// ------------------------------------------------------------------------------------------------
// intersections
// ------------------------------------------------------------------------------------------------
struct Face
{
  int k, s; // direction, side (0: lower, 1: upper) seen from the inside element
  double normal[3];
  double ie;       // intersection.geometry().integrationElement = volume of the face
  double diameter; // XT::Grid::diameter(intersection): max corner distance [EXT]
};

Face make_face(const Grid& g, const double* ext_in, int k, int s)
{
  Face f;
  f.k = k;
  f.s = s;
  for (int j = 0; j < 3; ++j)
    f.normal[j] = 0.;
  f.normal[k] = s ? 1. : -1.;
  f.ie = 1.;
  double d2 = 0.;
  for (int j = 0; j < g.d; ++j)
    if (j != k) {
      f.ie *= ext_in[j];
      d2 += ext_in[j] * ext_in[j];
    }
  f.diameter = std::sqrt(d2);
  return f;
}

// geometryInInside / geometryInOutside .global(xi): insert the fixed coordinate at position k
void face_to_element(int d, int k, int side_coord, const double* xi, double* xh)
{
  int j = 0;
  for (int r = 0; r < 3; ++r) {
    if (r == k)
      xh[r] = side_coord;
    else if (r < d)
      xh[r] = xi[j++];
    else
      xh[r] = 0.;
  }
}

// default_intersection_diameter (ipdg.hh:27-38)
double intersection_h(const Grid& g, const orc_integrand& t, const Face& f, const double* ext_in, const double* ext_out,
                      bool neighbor)
{
  if (t.hI_kind == ORC_HI_VOLUME)
    return f.ie;
  if (g.d == 1) {
    // XT::Grid::diameter(element) of a 1d element = its length
    if (neighbor)
      return 0.5 * (ext_in[0] + ext_out[0]);
    return ext_in[0];
  }
  return f.diameter;
}

int coupling_integrand_order(const orc_integrand& t, int p)
{
  if (t.kind == ORC_INT_IPDG_INNER_COUPLING) // laplace-ipdg.hh:95-105
    return t.diffusion.order + t.weight.order + p + p;
  return t.weight.order + p + p; // ipdg.hh:100-111
}

int boundary_integrand_order(const orc_integrand& t, int p)
{
  if (t.kind == ORC_INT_IPDG_DIRICHLET_COUPLING) // laplace-ipdg.hh:331-338
    return t.diffusion.order + p + p;
  return t.weight.order + p + p; // ipdg.hh:245-252
}

// LocalCouplingIntersectionIntegralBilinearForm::apply2 (integrals.hh:197-267) with the quaternary sum
// (combined.hh:381-431) of InnerCoupling (laplace-ipdg.hh:107-186) and InnerPenalty (ipdg.hh:113-171)
void local_coupling_matrices(const Grid& g, const Space& sp, const orc_form& form, const int64_t* idx_in,
                             const int64_t* idx_out, int k, int s, double* R /* 4 * n*n: in_in,in_out,out_in,out_out */)
{
  const int d = g.d, n = sp.nloc, nn = n * n;
  const int64_t e_in = g.index(idx_in), e_out = g.index(idx_out);
  double lo_in[3], ext_in[3], lo_out[3], ext_out[3];
  g.cell(idx_in, lo_in, ext_in);
  g.cell(idx_out, lo_out, ext_out);
  const Face f = make_face(g, ext_in, k, s);
  for (int i = 0; i < 4 * nn; ++i)
    R[i] = 0.;
  const Rule rule(form_order(form, sp.K, coupling_integrand_order));
  const int m = d > 1 ? rule.m : 1;
  const int m2 = d > 2 ? rule.m : 1;
  Basis bi, bo;
  bi.d = bo.d = d;
  bi.K = bo.K = sp.K;
  bi.n = bo.n = n;
  std::vector<double> V(4 * nn), S(4 * nn);
  for (int q2 = 0; q2 < m2; ++q2)
    for (int q1 = 0; q1 < m; ++q1) {
      const double xi[2] = {d > 1 ? rule.x[q1] : 0., d > 2 ? rule.x[q2] : 0.};
      const double wq = (d > 1 ? rule.w[q1] : 1.) * (d > 2 ? rule.w[q2] : 1.);
      double xh_in[3], xh_out[3], x_in[3], x_out[3];
      face_to_element(d, k, s, xi, xh_in);
      face_to_element(d, k, 1 - s, xi, xh_out);
      for (int r = 0; r < 3; ++r) {
        x_in[r] = lo_in[r] + xh_in[r] * ext_in[r];
        x_out[r] = lo_out[r] + xh_out[r] * ext_out[r];
      }
      bi.evaluate(xh_in, ext_in);
      bo.evaluate(xh_out, ext_out);
      for (int t = 0; t < form.n_terms; ++t) {
        double* T = t == 0 ? V.data() : S.data();
        for (int i = 0; i < 4 * nn; ++i)
          T[i] = 0.;
        const orc_integrand& in = form.terms[t];
        double w_in[9], w_out[9], wn[3];
        Pt p_in, p_out; // no volume-rule index on a face: ORC_FN_QP_* are element-form functions
        p_in.g = p_out.g = &g;
        p_in.idx = idx_in;
        p_in.xh = xh_in;
        p_out.idx = idx_out;
        p_out.xh = xh_out;
        eval_tensor(&in.weight, d, e_in, x_in, w_in, p_in);
        eval_tensor(&in.weight, d, e_out, x_out, w_out, p_out);
        matvec(d, w_out, f.normal, wn);
        const double delta_plus = dot(d, f.normal, wn);
        matvec(d, w_in, f.normal, wn);
        const double delta_minus = dot(d, f.normal, wn);
        if (in.kind == ORC_INT_IPDG_INNER_COUPLING) {
          double k_in[9], k_out[9];
          eval_tensor(&in.diffusion, d, e_in, x_in, k_in, p_in);
          eval_tensor(&in.diffusion, d, e_out, x_out, k_out, p_out);
          const double weight_minus = delta_plus / (delta_plus + delta_minus);
          const double weight_plus = delta_minus / (delta_plus + delta_minus);
          const double sp_ = in.prefactor;
          double fin[MAXN], fout[MAXN]; // (kappa grad phi) . n
          for (int j = 0; j < n; ++j) {
            double kg[3];
            matvec(d, k_in, &bi.grad[j * 3], kg);
            fin[j] = dot(d, kg, f.normal);
            matvec(d, k_out, &bo.grad[j * 3], kg);
            fout[j] = dot(d, kg, f.normal);
          }
          for (int ii = 0; ii < n; ++ii) {
            for (int jj = 0; jj < n; ++jj) {
              T[0 * nn + ii * n + jj] += -1.0 * weight_minus * fin[jj] * bi.val[ii];
              T[0 * nn + ii * n + jj] += -1.0 * sp_ * weight_minus * bi.val[jj] * fin[ii];
            }
            for (int jj = 0; jj < n; ++jj) {
              T[1 * nn + ii * n + jj] += -1.0 * weight_plus * fout[jj] * bi.val[ii];
              T[1 * nn + ii * n + jj] += sp_ * weight_minus * bo.val[jj] * fin[ii];
            }
          }
          for (int ii = 0; ii < n; ++ii) {
            for (int jj = 0; jj < n; ++jj) {
              T[2 * nn + ii * n + jj] += weight_minus * fin[jj] * bo.val[ii];
              T[2 * nn + ii * n + jj] += -1.0 * sp_ * weight_plus * bi.val[jj] * fout[ii];
            }
            for (int jj = 0; jj < n; ++jj) {
              T[3 * nn + ii * n + jj] += weight_plus * fout[jj] * bo.val[ii];
              T[3 * nn + ii * n + jj] += sp_ * weight_plus * bo.val[jj] * fout[ii];
            }
          }
        } else { // ORC_INT_IPDG_INNER_PENALTY
          const double weight = (delta_plus * delta_minus) / (delta_plus + delta_minus);
          const double h = intersection_h(g, in, f, ext_in, ext_out, true);
          const double penalty = (in.prefactor * weight) / h;
          for (int ii = 0; ii < n; ++ii) {
            for (int jj = 0; jj < n; ++jj)
              T[0 * nn + ii * n + jj] += penalty * bi.val[jj] * bi.val[ii];
            for (int jj = 0; jj < n; ++jj)
              T[1 * nn + ii * n + jj] += -1.0 * penalty * bo.val[jj] * bi.val[ii];
          }
          for (int ii = 0; ii < n; ++ii) {
            for (int jj = 0; jj < n; ++jj)
              T[2 * nn + ii * n + jj] += -1.0 * penalty * bi.val[jj] * bo.val[ii];
            for (int jj = 0; jj < n; ++jj)
              T[3 * nn + ii * n + jj] += penalty * bo.val[jj] * bo.val[ii];
          }
        }
        if (t > 0)
          for (int i = 0; i < 4 * nn; ++i)
            V[i] += S[i];
      }
      for (int i = 0; i < 4 * nn; ++i) // integrals.hh:254-265
        R[i] += V[i] * f.ie * wq;
    }
}

// LocalIntersectionIntegralBilinearForm::apply2 (integrals.hh:338-369) with the binary sum (combined.hh:309-318)
// of DirichletCoupling (laplace-ipdg.hh:340-368) and BoundaryPenalty (ipdg.hh:254-282); inside() == true.
void local_boundary_matrix(const Grid& g, const Space& sp, const orc_form& form, const int64_t* idx, int k, int s,
                           double* R /* n*n */)
{
  const int d = g.d, n = sp.nloc, nn = n * n;
  const int64_t e = g.index(idx);
  double lo[3], ext[3];
  g.cell(idx, lo, ext);
  const Face f = make_face(g, ext, k, s);
  for (int i = 0; i < nn; ++i)
    R[i] = 0.;
  const Rule rule(form_order(form, sp.K, boundary_integrand_order));
  const int m = d > 1 ? rule.m : 1;
  const int m2 = d > 2 ? rule.m : 1;
  Basis b;
  b.d = d;
  b.K = sp.K;
  b.n = n;
  std::vector<double> V(nn), S(nn);
  for (int q2 = 0; q2 < m2; ++q2)
    for (int q1 = 0; q1 < m; ++q1) {
      const double xi[2] = {d > 1 ? rule.x[q1] : 0., d > 2 ? rule.x[q2] : 0.};
      const double wq = (d > 1 ? rule.w[q1] : 1.) * (d > 2 ? rule.w[q2] : 1.);
      double xh[3], x[3];
      face_to_element(d, k, s, xi, xh);
      for (int r = 0; r < 3; ++r)
        x[r] = lo[r] + xh[r] * ext[r];
      b.evaluate(xh, ext);
      for (int t = 0; t < form.n_terms; ++t) {
        double* T = t == 0 ? V.data() : S.data();
        for (int i = 0; i < nn; ++i)
          T[i] = 0.;
        const orc_integrand& in = form.terms[t];
        Pt pt;
        pt.g = &g;
        pt.idx = idx;
        pt.xh = xh;
        if (in.kind == ORC_INT_IPDG_DIRICHLET_COUPLING) {
          double kap[9];
          eval_tensor(&in.diffusion, d, e, x, kap, pt);
          double fl[MAXN];
          for (int j = 0; j < n; ++j) {
            double kg[3];
            matvec(d, kap, &b.grad[j * 3], kg);
            fl[j] = dot(d, kg, f.normal);
          }
          for (int ii = 0; ii < n; ++ii)
            for (int jj = 0; jj < n; ++jj) {
              T[ii * n + jj] += -1.0 * fl[jj] * b.val[ii];
              T[ii * n + jj] += -1.0 * in.prefactor * b.val[jj] * fl[ii];
            }
        } else { // ORC_INT_IPDG_BOUNDARY_PENALTY
          double w[9], wn[3];
          eval_tensor(&in.weight, d, e, x, w, pt);
          matvec(d, w, f.normal, wn);
          const double h = intersection_h(g, in, f, ext, ext, false);
          const double penalty = (in.prefactor * dot(d, f.normal, wn)) / h;
          for (int ii = 0; ii < n; ++ii)
            for (int jj = 0; jj < n; ++jj)
              T[ii * n + jj] += penalty * b.val[jj] * b.val[ii];
        }
        if (t > 0)
          for (int i = 0; i < nn; ++i)
            V[i] += S[i];
      }
      for (int i = 0; i < nn; ++i)
        R[i] += V[i] * f.ie * wq;
    }
}

// ------------------------------------------------------------------------------------------------
// global containers [EXT XT::LA]: CSR with add_to_entry = search in the sorted row, guarded by
// row-striped locks when walking thread-parallel.
// ------------------------------------------------------------------------------------------------
struct Csr
{
  const int64_t* rowptr;
  const int32_t* colidx;
  double* values;
  std::mutex* locks;
  int nlocks;
  bool failed = false;

  inline void add_to_entry(int64_t r, int64_t c, double v)
  {
    const int32_t* b = colidx + rowptr[r];
    const int32_t* e = colidx + rowptr[r + 1];
    const int32_t* it = std::lower_bound(b, e, (int32_t)c);
    if (it == e || *it != (int32_t)c) {
      failed = true;
      return;
    }
    if (locks) {
      std::lock_guard<std::mutex> guard(locks[r % nlocks]);
      values[it - colidx] += v;
    } else
      values[it - colidx] += v;
  }
};

struct Vec
{
  double* v;
  std::mutex* locks;
  int nlocks;
  inline void add_to_entry(int64_t i, double x)
  {
    if (locks) {
      std::lock_guard<std::mutex> guard(locks[i % nlocks]);
      v[i] += x;
    } else
      v[i] += x;
  }
};

struct Walk
{
  const Grid* g;
  const Space* sp;
  Csr* A;
  Vec* b;
  int n_ef, n_cf, n_bf, n_rf;
  const orc_form *ef, *cf, *bf, *rf;
};

// XT::Grid::Walker::walk [EXT] over the element range [e0, e1): element functors, then for every
// intersection (order x-,x+,y-,y+,z-,z+) the intersection functors whose filter matches.
void walk_range(const Walk& w, int64_t e0, int64_t e1)
{
  const Grid& g = *w.g;
  const Space& sp = *w.sp;
  const int n = sp.nloc;
  std::vector<double> L(4 * n * n);
  std::vector<double> l(n);
  int64_t gi[MAXN], go[MAXN];
  for (int64_t e = e0; e < e1; ++e) {
    int64_t idx[3];
    g.coords(e, idx);
    if (w.n_ef || w.n_rf || w.n_bf || w.n_cf)
      sp.global_indices(idx, gi);
    // LocalElementBilinearFormAssembler::apply_local (bilinear-form-assemblers.hh:110-128)
    for (int f = 0; f < w.n_ef; ++f) {
      local_element_matrix(g, sp, w.ef[f], idx, L.data());
      const double scaling = w.ef[f].scaling;
      for (int ii = 0; ii < n; ++ii)
        for (int jj = 0; jj < n; ++jj)
          w.A->add_to_entry(gi[ii], gi[jj], scaling * L[ii * n + jj]);
    }
    // LocalElementFunctionalAssembler::apply_local (functional-assemblers.hh:77-86)
    for (int f = 0; f < w.n_rf; ++f) {
      local_element_vector(g, sp, w.rf[f], idx, l.data());
      for (int jj = 0; jj < n; ++jj)
        w.b->add_to_entry(gi[jj], l[jj]);
    }
    if (!w.n_cf && !w.n_bf)
      continue;
    for (int k = 0; k < g.d; ++k)
      for (int s = 0; s < 2; ++s) {
        int64_t nb[3];
        bool boundary;
        const bool neighbor = g.neighbor(idx, k, s, nb, &boundary);
        if (neighbor) {
          // ApplyOn::InnerIntersectionsOnce / PeriodicBoundaryIntersectionsOnce [EXT]: index(inside) < index(outside)
          const int64_t eo = g.index(nb);
          if (!(e < eo))
            continue;
          if (!w.n_cf)
            continue;
          sp.global_indices(nb, go);
          // LocalCouplingIntersectionBilinearFormAssembler::apply_local (bilinear-form-assemblers.hh:238-278)
          for (int f = 0; f < w.n_cf; ++f) {
            local_coupling_matrices(g, sp, w.cf[f], idx, nb, k, s, L.data());
            const double scaling = w.cf[f].scaling;
            const int nn = n * n;
            for (int ii = 0; ii < n; ++ii) {
              for (int jj = 0; jj < n; ++jj)
                w.A->add_to_entry(gi[ii], gi[jj], scaling * L[0 * nn + ii * n + jj]);
              for (int jj = 0; jj < n; ++jj)
                w.A->add_to_entry(gi[ii], go[jj], scaling * L[1 * nn + ii * n + jj]);
            }
            for (int ii = 0; ii < n; ++ii) {
              for (int jj = 0; jj < n; ++jj)
                w.A->add_to_entry(go[ii], gi[jj], scaling * L[2 * nn + ii * n + jj]);
              for (int jj = 0; jj < n; ++jj)
                w.A->add_to_entry(go[ii], go[jj], scaling * L[3 * nn + ii * n + jj]);
            }
          }
        } else if (boundary) {
          // LocalIntersectionBilinearFormAssembler::apply_local (bilinear-form-assemblers.hh:380-396),
          // filter CustomBoundaryIntersections(AllDirichletBoundaryInfo, DirichletBoundary) [EXT]
          for (int f = 0; f < w.n_bf; ++f) {
            local_boundary_matrix(g, sp, w.bf[f], idx, k, s, L.data());
            const double scaling = w.bf[f].scaling;
            for (int ii = 0; ii < n; ++ii)
              for (int jj = 0; jj < n; ++jj)
                w.A->add_to_entry(gi[ii], gi[jj], scaling * L[ii * n + jj]);
          }
        }
      }
  }
}

template <class F>
void run_threads(int64_t total, int num_threads, F&& fn)
{
  if (num_threads <= 1) {
    fn(0, total);
    return;
  }
  std::vector<std::thread> th;
  const int64_t chunk = (total + num_threads - 1) / num_threads;
  for (int t = 0; t < num_threads; ++t) {
    const int64_t a = std::min<int64_t>(total, t * chunk), b = std::min<int64_t>(total, (t + 1) * chunk);
    if (a < b)
      th.emplace_back([=, &fn] { fn(a, b); });
  }
  for (auto& t : th)
    t.join();
}

// ------------------------------------------------------------------------------------------------
// FV: numerical fluxes and the coupling operator
// ------------------------------------------------------------------------------------------------
inline void flux_eval(const orc_flux& fl, int d, double u, double* f)
{
  for (int k = 0; k < d; ++k)
    f[k] = fl.kind == ORC_FLUX_LINEAR ? fl.p[k] * u : 0.5 * u * u;
}
inline void flux_jac(const orc_flux& fl, int d, double u, double* df)
{
  for (int k = 0; k < d; ++k)
    df[k] = fl.kind == ORC_FLUX_LINEAR ? fl.p[k] : u;
}

inline double numerical_flux(const orc_flux& fl, int d, double u, double v, const double* n)
{
  double a[3], b[3];
  if (fl.numflux == ORC_NUMFLUX_UPWIND) {
    // NumericalUpwindFlux<I,d,1>::apply (local/numerical-fluxes/upwind.hh:61-73)
    flux_jac(fl, d, (u + v) / 2., a);
    if (dot(d, n, a) > 0) {
      flux_eval(fl, d, u, b);
      return dot(d, b, n);
    }
    flux_eval(fl, d, v, b);
    return dot(d, b, n);
  }
  // NumericalLaxFriedrichsFlux::apply (local/numerical-fluxes/lax-friedrichs.hh:60-88), lambda_ = 0
  double lambda = 0.;
  flux_jac(fl, d, u, a);
  flux_jac(fl, d, v, b);
  for (int k = 0; k < d; ++k) {
    lambda = std::max(lambda, std::fabs(a[k]));
    lambda = std::max(lambda, std::fabs(b[k]));
  }
  lambda = 1. / lambda;
  flux_eval(fl, d, u, a);
  flux_eval(fl, d, v, b);
  double ret = 0.;
  for (int k = 0; k < d; ++k)
    ret += (a[k] + b[k]) * (n[k] * 0.5);
  ret += (u - v) * (0.5 / lambda);
  return ret;
}

// LocalizableOperator::apply (operators/localizable-operator.hh:352-379) with the coupling operators
// registered by AdvectionFvOperator (operators/advection-fv.hh:77-82) and
// LocalAdvectionFvCouplingOperator::apply (local/operators/advection-fv.hh:127-153); boundary treatments appended by
// AdvectionFvOperator::append (operators/advection-fv.hh:96-123) run on the boundary intersections selected by their
// filter: LocalAdvectionFvBoundaryTreatmentByCustomNumericalFluxOperator::apply (local/operators/advection-fv.hh:281-296)
// and ...ByCustomExtrapolationOperator::apply (:418-443).  The reference takes arbitrary lambdas; the descriptor
// families restated here are  extrapolation v = a u + b  and  numerical boundary flux g = a (f(u) . n) + b.
void fv_walk_range(const Grid& g, const orc_flux& fl, int n_bnd, const orc_fv_boundary* bnd, const double* u, Vec& out,
                   int64_t e0, int64_t e1)
{
  for (int64_t e = e0; e < e1; ++e) {
    int64_t idx[3];
    g.coords(e, idx);
    double lo_in[3], ext_in[3];
    g.cell(idx, lo_in, ext_in);
    for (int k = 0; k < g.d; ++k)
      for (int s = 0; s < 2; ++s) {
        int64_t nb[3];
        bool boundary;
        if (!g.neighbor(idx, k, s, nb, &boundary)) {
          // a domain boundary intersection (not periodic): boundary treatments in append order
          const Face f = make_face(g, ext_in, k, s);
          for (int t = 0; t < n_bnd; ++t) {
            if (!(bnd[t].side_mask >> (2 * k + s) & 1))
              continue;
            const double uu = u[e];
            double gflux;
            if (bnd[t].kind == ORC_FVBND_EXTRAPOLATION) {
              const double vv = bnd[t].a * uu + bnd[t].b;
              gflux = numerical_flux(fl, g.d, uu, vv, f.normal);
            } else {
              double fu[3];
              flux_eval(fl, g.d, uu, fu);
              gflux = bnd[t].a * dot(g.d, fu, f.normal) + bnd[t].b;
            }
            const double factor = f.ie / g.volume(ext_in); // local/operators/advection-fv.hh:293, 440
            out.add_to_entry(e, gflux * factor);
          }
          continue;
        }
        const int64_t eo = g.index(nb);
        if (!(e < eo))
          continue;
        double lo_out[3], ext_out[3];
        g.cell(nb, lo_out, ext_out);
        const Face f = make_face(g, ext_in, k, s);
        const double uu = u[e], vv = u[eo];
        const double gflux = numerical_flux(fl, g.d, uu, vv, f.normal);
        const double h_intersection = f.ie;
        const double hinv_inside = 1. / g.volume(ext_in);
        const double hinv_outside = 1. / g.volume(ext_out);
        const double g_ii = gflux * h_intersection;
        out.add_to_entry(e, g_ii * hinv_inside);
        out.add_to_entry(eo, -g_ii * hinv_outside);
      }
  }
}

void fv_apply(const Grid& g, const orc_flux& fl, int n_bnd, const orc_fv_boundary* bnd, const double* u, double* out,
              int num_threads)
{
  for (int64_t i = 0; i < g.ne; ++i) // range.set_all(0), localizable-operator.hh:359
    out[i] = 0.;
  std::vector<std::mutex> locks(num_threads > 1 ? 4096 : 0);
  Vec v{out, num_threads > 1 ? locks.data() : nullptr, 4096};
  run_threads(g.ne, num_threads, [&](int64_t a, int64_t b) { fv_walk_range(g, fl, n_bnd, bnd, u, v, a, b); });
}

// XT::Common::FloatCmp::{eq,lt,gt}<Style::numpy> with the default epsilons [EXT dune-xt common/float_cmp*.hh]:
// eq(a, b) := |a - b| <= atol + rtol |b|, rtol = atol = Dune::FloatCmp::DefaultEpsilon<double> = 8 * 2^-52;
// lt(a, b) := !eq(a, b) && a < b.  Parity unpinned (third-party), only decides whether a last sliver step is taken.
inline bool floatcmp_eq(double a, double b)
{
  const double eps = 8. * std::numeric_limits<double>::epsilon();
  return std::fabs(a - b) <= eps + eps * std::fabs(b);
}
inline bool floatcmp_lt(double a, double b)
{
  return !floatcmp_eq(a, b) && a < b;
}
inline bool floatcmp_gt(double a, double b)
{
  return !floatcmp_eq(a, b) && a > b;
}

// ExplicitRungeKuttaTimeStepper::step (tools/timestepper/explicit-rungekutta.hh:237-270)
struct RkStepper
{
  const Grid& g;
  const orc_flux& fl;
  int n_bnd;
  const orc_fv_boundary* bnd;
  int s;
  const double *A, *b, *c;
  double r;
  int num_threads;
  std::vector<double> u_i;
  std::vector<std::vector<double>> k;
  RkStepper(const Grid& g_, const orc_flux& fl_, int n_bnd_, const orc_fv_boundary* bnd_, int s_, const double* A_,
            const double* b_, const double* c_, double r_, int nt)
    : g(g_), fl(fl_), n_bnd(n_bnd_), bnd(bnd_), s(s_), A(A_), b(b_), c(c_), r(r_), num_threads(nt), u_i(g_.ne),
      k(s_, std::vector<double>(g_.ne))
  {}
  double step(double* u_n, double& t, double dt, double max_dt)
  {
    const double actual_dt = std::min(dt, max_dt);
    const int64_t n = g.ne;
    for (int ii = 0; ii < s; ++ii) {
      for (int64_t i = 0; i < n; ++i)
        u_i[i] = u_n[i];
      for (int jj = 0; jj < ii; ++jj) {
        const double coef = actual_dt * r * A[ii * s + jj];
        for (int64_t i = 0; i < n; ++i)
          u_i[i] += k[jj][i] * coef;
      }
      // the operator here does not depend on t / dt (the flux families are autonomous)
      fv_apply(g, fl, n_bnd, bnd, u_i.data(), k[ii].data(), num_threads);
    }
    for (int ii = 0; ii < s; ++ii) {
      const double coef = r * actual_dt * b[ii];
      for (int64_t i = 0; i < n; ++i)
        u_n[i] += k[ii][i] * coef;
    }
    t += actual_dt;
    return dt;
  }
};


// ------------------------------------------------------------------------------------------------
// Systems of conservation laws (m > 1): the Euler equations.  EulerTools<d> (tools/euler.hh), d = 1, 2, m = d + 2,
// conservative variables w = (rho, rho v, E); operation order follows the reference.
// ------------------------------------------------------------------------------------------------
struct Euler
{
  int d;
  double gamma;
  int m() const
  {
    return d + 2;
  }
  // primitives(w) (tools/euler.hh:136-147)
  void primitives(const double* w, double& rho, double* v, double& p) const
  {
    rho = w[0];
    double v2 = 0.;
    for (int i = 0; i < d; ++i) {
      v[i] = w[1 + i] / w[0];
      v2 += v[i] * v[i];
    }
    p = (gamma - 1.) * (w[d + 1] - 0.5 * rho * v2);
  }
  // flux(w)[s][i] (:212-236)
  void flux(const double* w, double* f) const
  {
    const int M = m();
    double rho, v[3], p;
    primitives(w, rho, v, p);
    const double E = w[M - 1];
    for (int ss = 0; ss < d; ++ss) {
      double* f_s = f + ss * M;
      f_s[0] = rho * v[ss];
      for (int ii = 0; ii < d; ++ii)
        f_s[1 + ii] = rho * v[ii] * v[ss] + (ss == ii ? 1 : 0) * p;
      f_s[M - 1] = (E + p) * v[ss];
    }
  }
  // flux_jacobian(w)[s][r][c] (:262-316)
  void jacobian(const double* w, double* J) const
  {
    const int M = m();
    for (int i = 0; i < d * M * M; ++i)
      J[i] = 0.;
    const double rho = w[0], E = w[M - 1];
    double v[3] = {0., 0., 0.};
    for (int i = 0; i < d; ++i)
      v[i] = w[1 + i] / w[0];
    const double gamma_1 = gamma - 1.;
    double vnorm2 = 0.;
    for (int i = 0; i < d; ++i)
      vnorm2 += v[i] * v[i];
    const double ek = 0.5 * vnorm2;
    auto at = [&](int s_, int r, int c) -> double& { return J[(s_ * M + r) * M + c]; };
    if (d == 1) {
      at(0, 0, 1) = 1.;
      at(0, 1, 0) = gamma_1 * ek - v[0] * v[0];
      at(0, 1, 1) = (3. - gamma) * v[0];
      at(0, 1, 2) = gamma_1;
      at(0, 2, 0) = v[0] * (gamma_1 * vnorm2 - (gamma * E) / rho);
      at(0, 2, 1) = ((gamma * E) / rho) - gamma_1 * v[0] * v[0] - gamma_1 * ek;
      at(0, 2, 2) = gamma * v[0];
    } else {
      at(0, 0, 1) = 1.;
      at(0, 1, 0) = gamma_1 * ek - v[0] * v[0];
      at(0, 1, 1) = (3. - gamma) * v[0];
      at(0, 1, 2) = -1. * gamma_1 * v[1];
      at(0, 1, 3) = gamma_1;
      at(0, 2, 0) = -1. * v[0] * v[1];
      at(0, 2, 1) = v[1];
      at(0, 2, 2) = v[0];
      at(0, 3, 0) = v[0] * (gamma_1 * vnorm2 - (gamma * E) / rho);
      at(0, 3, 1) = ((gamma * E) / rho) - gamma_1 * v[0] * v[0] - gamma_1 * ek;
      at(0, 3, 2) = -1. * gamma_1 * v[0] * v[1];
      at(0, 3, 3) = gamma * v[0];
      at(1, 0, 2) = 1.;
      at(1, 1, 0) = -1. * v[0] * v[1];
      at(1, 1, 1) = v[1];
      at(1, 1, 2) = v[0];
      at(1, 2, 0) = 0.5 * gamma_1 * vnorm2 - v[1] * v[1];
      at(1, 2, 1) = -1. * gamma_1 * v[0];
      at(1, 2, 2) = (3. - gamma) * v[1];
      at(1, 2, 3) = gamma_1;
      at(1, 3, 0) = v[1] * (gamma_1 * vnorm2 - ((gamma * E) / rho));
      at(1, 3, 1) = -1. * gamma_1 * v[0] * v[1];
      at(1, 3, 2) = ((gamma * E) / rho) - gamma_1 * v[1] * v[1] - gamma_1 * ek;
      at(1, 3, 3) = gamma * v[1];
    }
  }
  // eigenvalues_ / eigenvectors_ / eigenvectors_inv_flux_jacobian(w, n) (:325-462); T, Ti row-major m x m
  void eigen(const double* w, const double* n, double* ev, double* T, double* Ti) const
  {
    const int M = m();
    double rho, v[3] = {0., 0., 0.}, p;
    primitives(w, rho, v, p);
    const double E = w[M - 1];
    const double a = std::sqrt(gamma * p / rho);
    const double H = (E + p) / rho;
    const double rho_over_2a = rho / (2 * a);
    double v2 = 0., vn = 0.;
    for (int i = 0; i < d; ++i) {
      v2 += v[i] * v[i];
      vn += v[i] * n[i];
    }
    const double ek = 0.5 * v2;
    const double Mach = std::sqrt(v2) / a;
    const double gamma_1 = gamma - 1.;
    auto t = [&](int r, int c) -> double& { return T[r * M + c]; };
    auto ti = [&](int r, int c) -> double& { return Ti[r * M + c]; };
    if (d == 1) {
      ev[0] = vn, ev[1] = vn + a, ev[2] = vn - a;
      t(0, 0) = 1., t(0, 1) = rho_over_2a, t(0, 2) = rho_over_2a;
      t(1, 0) = v[0], t(1, 1) = rho_over_2a * (v[0] + a * n[0]), t(1, 2) = rho_over_2a * (v[0] - a * n[0]);
      t(2, 0) = ek, t(2, 1) = rho_over_2a * (H + a * vn), t(2, 2) = rho_over_2a * (H - a * vn);
      ti(0, 0) = 1. - (gamma_1 / 2.) * Mach * Mach;
      ti(0, 1) = gamma_1 * v[0] / (a * a);
      ti(0, 2) = -gamma_1 / (a * a);
      ti(1, 0) = (a / rho) * ((gamma_1 / 2.) * Mach * Mach - vn / a);
      ti(1, 1) = (1. / rho) * (n[0] - gamma_1 * (v[0] / a));
      ti(1, 2) = gamma_1 / (rho * a);
      ti(2, 0) = (a / rho) * ((gamma_1 / 2.) * Mach * Mach + vn / a);
      ti(2, 1) = (-1. / rho) * (n[0] + gamma_1 * (v[0] / a));
      ti(2, 2) = gamma_1 / (rho * a);
    } else {
      ev[0] = vn, ev[1] = vn, ev[2] = vn + a, ev[3] = vn - a;
      t(0, 0) = 1., t(0, 1) = 0., t(0, 2) = rho_over_2a, t(0, 3) = rho_over_2a;
      t(1, 0) = v[0], t(1, 1) = rho * n[1], t(1, 2) = rho_over_2a * (v[0] + a * n[0]), t(1, 3) = rho_over_2a * (v[0] - a * n[0]);
      t(2, 0) = v[1], t(2, 1) = -rho * n[0], t(2, 2) = rho_over_2a * (v[1] + a * n[1]), t(2, 3) = rho_over_2a * (v[1] - a * n[1]);
      t(3, 0) = ek, t(3, 1) = rho * (v[0] * n[1] - v[1] * n[0]), t(3, 2) = rho_over_2a * (H + a * vn), t(3, 3) = rho_over_2a * (H - a * vn);
      ti(0, 0) = 1. - (gamma_1 / 2.) * Mach * Mach;
      ti(0, 1) = gamma_1 * v[0] / (a * a);
      ti(0, 2) = gamma_1 * v[1] / (a * a);
      ti(0, 3) = -gamma_1 / (a * a);
      ti(1, 0) = (1. / rho) * (v[1] * n[0] - v[0] * n[1]);
      ti(1, 1) = n[1] / rho;
      ti(1, 2) = -n[0] / rho;
      ti(1, 3) = 0.;
      ti(2, 0) = (a / rho) * ((gamma_1 / 2.) * Mach * Mach - vn / a);
      ti(2, 1) = (1. / rho) * (n[0] - gamma_1 * (v[0] / a));
      ti(2, 2) = (1. / rho) * (n[1] - gamma_1 * (v[1] / a));
      ti(2, 3) = gamma_1 / (rho * a);
      ti(3, 0) = (a / rho) * ((gamma_1 / 2.) * Mach * Mach + vn / a);
      ti(3, 1) = (-1. / rho) * (n[0] + gamma_1 * (v[0] / a));
      ti(3, 2) = (-1. / rho) * (n[1] + gamma_1 * (v[1] / a));
      ti(3, 3) = gamma_1 / (rho * a);
    }
  }
};

// numerical flux g(u, v, n) for systems: NumericalVijayasundaramFlux::apply (local/numerical-fluxes/vijayasundaram.hh:111-133)
// with the EulerTools eigendecomposition (the lambda of examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:402-409), or
// NumericalLaxFriedrichsFlux::apply (lax-friedrichs.hh:66-88) with the given lambda = p[1]
void system_numerical_flux(const orc_flux& fl, const Euler& eu, const double* u, const double* v, const double* n, double* g)
{
  const int M = eu.m();
  if (fl.numflux == ORC_NUMFLUX_VIJAYASUNDARAM) {
    double w[4], ev[4], T[16], Ti[16];
    for (int i = 0; i < M; ++i)
      w[i] = 0.5 * (u[i] + v[i]);
    eu.eigen(w, n, ev, T, Ti);
    // P_plus = T * lambda_plus * T_inv, P_minus = T * lambda_minus * T_inv (matrix products, left to right)
    double TLp[16], TLm[16], Pp[16], Pm[16];
    for (int r = 0; r < M; ++r)
      for (int c = 0; c < M; ++c) {
        TLp[r * M + c] = T[r * M + c] * std::max(ev[c], 0.);
        TLm[r * M + c] = T[r * M + c] * std::min(ev[c], 0.);
      }
    for (int r = 0; r < M; ++r)
      for (int c = 0; c < M; ++c) {
        double sp = 0., sm = 0.;
        for (int k = 0; k < M; ++k) {
          sp += TLp[r * M + k] * Ti[k * M + c];
          sm += TLm[r * M + k] * Ti[k * M + c];
        }
        Pp[r * M + c] = sp;
        Pm[r * M + c] = sm;
      }
    for (int r = 0; r < M; ++r) {
      double a = 0., b = 0.;
      for (int c = 0; c < M; ++c) {
        a += Pp[r * M + c] * u[c];
        b += Pm[r * M + c] * v[c];
      }
      g[r] = a + b;
    }
    return;
  }
  const double lambda = fl.p[1];
  double fu[8], fv[8];
  eu.flux(u, fu);
  eu.flux(v, fv);
  for (int i = 0; i < M; ++i)
    g[i] = 0.;
  for (int dd = 0; dd < eu.d; ++dd)
    for (int i = 0; i < M; ++i)
      g[i] += (fu[dd * M + i] + fv[dd * M + i]) * (n[dd] * 0.5);
  for (int i = 0; i < M; ++i)
    g[i] += (u[i] - v[i]) * (0.5 / lambda);
}

// LocalizableOperator::apply + LocalAdvectionFvCouplingOperator::apply (local/operators/advection-fv.hh:127-153) for m
// components per cell (DoF m * element + i, spaces/mapper/finite-volume.hh:92-97); no boundary treatments
void fvsys_apply(const Grid& g, const orc_flux& fl, const double* u, double* out, unsigned wall_mask = 0,
                 unsigned mirror_mask = 0)
{
  const Euler eu{g.d, fl.p[0]};
  const int M = eu.m();
  for (int64_t i = 0; i < g.ne * M; ++i)
    out[i] = 0.;
  for (int64_t e = 0; e < g.ne; ++e) {
    int64_t idx[3];
    g.coords(e, idx);
    double lo_in[3], ext_in[3];
    g.cell(idx, lo_in, ext_in);
    for (int k = 0; k < g.d; ++k)
      for (int s = 0; s < 2; ++s) {
        int64_t nb[3];
        bool boundary;
        if (!g.neighbor(idx, k, s, nb, &boundary)) {
          // impermeable walls (test/inviscid-compressible-flow/base.hh:187-241) on the sides selected by the masks:
          // LocalAdvectionFvBoundaryTreatmentByCustomNumericalFluxOperator with g = flux_at_impermeable_walls(u, n)
          // (tools/euler.hh:238-251) / ...ByCustomExtrapolationOperator with the mirrored state ([DF2015, (8.66-8.67)])
          const unsigned bit = 1u << (2 * k + s);
          if (!((wall_mask | mirror_mask) & bit))
            continue;
          const Face f = make_face(g, ext_in, k, s);
          const double factor = f.ie / g.volume(ext_in); // local/operators/advection-fv.hh:293, 440
          const double* uu = u + e * M;
          if (wall_mask & bit) {
            double rho, v[3], p;
            eu.primitives(uu, rho, v, p);
            for (int ii = 0; ii < g.d; ++ii)
              out[e * M + 1 + ii] += (f.normal[ii] * p) * factor;
          }
          if (mirror_mask & bit) {
            double rho, v[3] = {0., 0., 0.}, p;
            eu.primitives(uu, rho, v, p);
            double vn = 0.;
            for (int ii = 0; ii < g.d; ++ii)
              vn += v[ii] * f.normal[ii];
            double vv[4], v2 = 0.;
            for (int ii = 0; ii < g.d; ++ii) {
              v[ii] -= f.normal[ii] * 2. * vn;
              v2 += v[ii] * v[ii];
            }
            vv[0] = rho;
            for (int ii = 0; ii < g.d; ++ii)
              vv[1 + ii] = v[ii] * rho;
            vv[M - 1] = p / (eu.gamma - 1.) + 0.5 * rho * v2; // EulerTools::conservative / energy (:127-131, 153-157)
            double gf[4];
            system_numerical_flux(fl, eu, uu, vv, f.normal, gf);
            for (int ii = 0; ii < M; ++ii)
              out[e * M + ii] += gf[ii] * factor;
          }
          continue;
        }
        const int64_t eo = g.index(nb);
        if (!(e < eo))
          continue;
        double lo_out[3], ext_out[3];
        g.cell(nb, lo_out, ext_out);
        const Face f = make_face(g, ext_in, k, s);
        double gf[4];
        system_numerical_flux(fl, eu, u + e * M, u + eo * M, f.normal, gf);
        const double h_intersection = f.ie;
        const double hinv_inside = 1. / g.volume(ext_in);
        const double hinv_outside = 1. / g.volume(ext_out);
        for (int ii = 0; ii < M; ++ii) {
          const double g_ii = gf[ii] * h_intersection;
          out[e * M + ii] += g_ii * hinv_inside;
          out[eo * M + ii] += -g_ii * hinv_outside;
        }
      }
  }
}

} // namespace

// ==================================================================================================
// C interface
// ==================================================================================================
extern "C" {

const char* orc_last_error(void)
{
  return g_error.c_str();
}

int64_t orc_num_elements(const orc_grid* g)
{
  return Grid(g).ne;
}

int64_t orc_space_size(const orc_grid* g, int kind, int order)
{
  Grid gr(g);
  return Space(gr, kind, order).size;
}

int32_t orc_space_local_size(const orc_grid* g, int kind, int order)
{
  Grid gr(g);
  return Space(gr, kind, order).nloc;
}

void orc_space_global_indices(const orc_grid* g, int kind, int order, int64_t element, int64_t* out)
{
  Grid gr(g);
  Space sp(gr, kind, order);
  int64_t idx[3];
  gr.coords(element, idx);
  sp.global_indices(idx, out);
}

int32_t orc_gauss_rule(int order, double* points01, double* weights)
{
  Rule r(order);
  for (int i = 0; i < r.m; ++i) {
    points01[i] = r.x[i];
    weights[i] = r.w[i];
  }
  return r.m;
}

void orc_shape_values(int dim, int order, const double* xhat, double* values)
{
  double xh[3] = {xhat[0], dim > 1 ? xhat[1] : 0., dim > 2 ? xhat[2] : 0.};
  shape(dim, order, xh, values, nullptr);
}

void orc_shape_gradients(int dim, int order, const double* xhat, double* grads)
{
  double xh[3] = {xhat[0], dim > 1 ? xhat[1] : 0., dim > 2 ? xhat[2] : 0.};
  double g3[MAXN * 3];
  shape(dim, order, xh, nullptr, g3);
  const int n = local_size(dim, order);
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < dim; ++r)
      grads[i * dim + r] = g3[i * 3 + r];
}

struct orc_pattern
{
  int64_t rows;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> colidx;
};

// make_{element,intersection,element_and_intersection}_sparsity_pattern (tools/sparsity-pattern.hh:34-144):
// XT::LA::SparsityPatternDefault::insert appends if absent, sort() sorts each row [EXT] -- restated as
// append-all, then sort + unique per row (same final container).
orc_pattern* orc_pattern_create(const orc_grid* g, int test_kind, int test_order, int ansatz_kind, int ansatz_order,
                                int stencil)
{
  Grid gr(g);
  Space test(gr, test_kind, test_order), ansatz(gr, ansatz_kind, ansatz_order);
  std::vector<std::vector<int32_t>> rows(test.size);
  int64_t ri[MAXN], ci[MAXN];
  for (int64_t e = 0; e < gr.ne; ++e) {
    int64_t idx[3];
    gr.coords(e, idx);
    test.global_indices(idx, ri);
    if (stencil != ORC_STENCIL_INTERSECTION) {
      ansatz.global_indices(idx, ci);
      for (int ii = 0; ii < test.nloc; ++ii)
        for (int jj = 0; jj < ansatz.nloc; ++jj)
          rows[ri[ii]].push_back((int32_t)ci[jj]);
    }
    if (stencil != ORC_STENCIL_ELEMENT) {
      for (int k = 0; k < gr.d; ++k)
        for (int s = 0; s < 2; ++s) {
          int64_t nb[3];
          bool boundary;
          if (!gr.neighbor(idx, k, s, nb, &boundary))
            continue;
          ansatz.global_indices(nb, ci);
          for (int ii = 0; ii < test.nloc; ++ii)
            for (int jj = 0; jj < ansatz.nloc; ++jj)
              rows[ri[ii]].push_back((int32_t)ci[jj]);
        }
    }
  }
  auto* p = new orc_pattern;
  p->rows = test.size;
  p->rowptr.assign(test.size + 1, 0);
  for (int64_t r = 0; r < test.size; ++r) {
    auto& row = rows[r];
    std::sort(row.begin(), row.end());
    row.erase(std::unique(row.begin(), row.end()), row.end());
    p->rowptr[r + 1] = p->rowptr[r] + (int64_t)row.size();
  }
  p->colidx.resize(p->rowptr[test.size]);
  for (int64_t r = 0; r < test.size; ++r)
    std::copy(rows[r].begin(), rows[r].end(), p->colidx.begin() + p->rowptr[r]);
  return p;
}

int64_t orc_pattern_rows(const orc_pattern* p)
{
  return p->rows;
}
int64_t orc_pattern_nnz(const orc_pattern* p)
{
  return (int64_t)p->colidx.size();
}
void orc_pattern_copy(const orc_pattern* p, int64_t* rowptr, int32_t* colidx)
{
  std::copy(p->rowptr.begin(), p->rowptr.end(), rowptr);
  std::copy(p->colidx.begin(), p->colidx.end(), colidx);
}
void orc_pattern_free(orc_pattern* p)
{
  delete p;
}

int orc_assemble(const orc_grid* g, int kind, int order, const int64_t* rowptr, const int32_t* colidx, double* values,
                 int n_element_forms, const orc_form* element_forms, int n_coupling_forms,
                 const orc_form* coupling_forms, int n_boundary_forms, const orc_form* boundary_forms,
                 int n_rhs_forms, const orc_form* rhs_forms, double* rhs, int num_threads)
{
  Grid gr(g);
  Space sp(gr, kind, order);
  if (sp.nloc > MAXN) {
    g_error = "local size exceeds MAXN";
    return 1;
  }
  const int nlocks = 8192;
  std::vector<std::mutex> locks(num_threads > 1 ? nlocks : 0), vlocks(num_threads > 1 ? nlocks : 0);
  Csr A{rowptr, colidx, values, num_threads > 1 ? locks.data() : nullptr, nlocks};
  Vec b{rhs, num_threads > 1 ? vlocks.data() : nullptr, nlocks};
  Walk w{&gr,           &sp,          &A,           &b,           n_element_forms, n_coupling_forms, n_boundary_forms,
         n_rhs_forms,   element_forms, coupling_forms, boundary_forms, rhs_forms};
  run_threads(gr.ne, num_threads, [&](int64_t a, int64_t bb) { walk_range(w, a, bb); });
  if (A.failed) {
    g_error = "add_to_entry: entry not in the sparsity pattern";
    return 2;
  }
  return 0;
}

void orc_local_element_matrix(const orc_grid* g, int kind, int order, const orc_form* form, int64_t element,
                              double* out)
{
  Grid gr(g);
  Space sp(gr, kind, order);
  int64_t idx[3];
  gr.coords(element, idx);
  local_element_matrix(gr, sp, *form, idx, out);
}

int orc_fv_apply(const orc_grid* g, const orc_flux* flux, const double* u, double* out, int num_threads)
{
  Grid gr(g);
  fv_apply(gr, *flux, 0, nullptr, u, out, num_threads);
  return 0;
}

int orc_fv_apply_bnd(const orc_grid* g, const orc_flux* flux, int n_bnd, const orc_fv_boundary* bnd, const double* u,
                     double* out, int num_threads)
{
  Grid gr(g);
  fv_apply(gr, *flux, n_bnd, bnd, u, out, num_threads);
  return 0;
}

int orc_rk_step(const orc_grid* g, const orc_flux* flux, int n_bnd, const orc_fv_boundary* bnd, int num_stages,
                const double* A, const double* b, const double* c, double r, double* u, double* t, double dt,
                double max_dt, int num_threads)
{
  Grid gr(g);
  RkStepper st(gr, *flux, n_bnd, bnd, num_stages, A, b, c, r, num_threads);
  st.step(u, *t, dt, max_dt);
  return 0;
}

// TimeStepperInterface::solve (tools/timestepper/interface.hh:191-263) with num_save_steps = size_t(-1), no output:
// while lt(t, t_end): max_dt = gt(t + dt, t_end) ? t_end - t : dt; dt = step(dt, max_dt)
int orc_rk_solve(const orc_grid* g, const orc_flux* flux, int n_bnd, const orc_fv_boundary* bnd, int num_stages,
                 const double* A, const double* b, const double* c, double r, double* u, double t0, double t_end,
                 double initial_dt, int64_t* n_steps, double* t_final, int num_threads)
{
  Grid gr(g);
  RkStepper st(gr, *flux, n_bnd, bnd, num_stages, A, b, c, r, num_threads);
  double dt = initial_dt, t = t0;
  int64_t steps = 0;
  while (floatcmp_lt(t, t_end)) {
    double max_dt = dt;
    if (floatcmp_gt(t + dt, t_end))
      max_dt = t_end - t;
    dt = st.step(u, t, dt, max_dt);
    ++steps;
  }
  if (n_steps)
    *n_steps = steps;
  if (t_final)
    *t_final = t;
  return 0;
}

// estimate_dt_for_hyperbolic_system (tools/hyperbolic.hh:38-86) for m = 1 and a finite-volume state (order 0: one
// quadrature point per element, the cell value).  boundary_data_range == NULL: the reference's defaults
// {numeric_limits<R>::max(), numeric_limits<R>::min()} (the latter is the smallest positive normal, as in the reference).
double orc_fv_estimate_dt(const orc_grid* g, const orc_flux* flux, const double* u, const double* boundary_data_range)
{
  Grid gr(g);
  double data_minimum = boundary_data_range ? boundary_data_range[0] : std::numeric_limits<double>::max();
  double data_maximum = boundary_data_range ? boundary_data_range[1] : std::numeric_limits<double>::min();
  for (int64_t e = 0; e < gr.ne; ++e) {
    data_minimum = std::min(data_minimum, u[e]);
    data_maximum = std::max(data_maximum, u[e]);
  }
  if (!(data_minimum < data_maximum))
    data_maximum = data_minimum + 1e-6 * data_minimum;
  double max_flux_derivative = std::numeric_limits<double>::min();
  // one-cell YaspGrid [min, max]; Gauss rule of order flux.order() (linear: 1, Burgers: 2; test/linear-transport/
  // base.hh:43, test/burgers/base.hh:41); geometry.global(x) = lower + x (upper - lower) [EXT]
  const Rule rule(flux->kind == ORC_FLUX_LINEAR ? 1 : 2);
  for (int q = 0; q < rule.m; ++q) {
    const double uq = data_minimum + rule.x[q] * (data_maximum - data_minimum);
    double df[3];
    flux_jac(*flux, gr.d, uq, df);
    for (int ss = 0; ss < gr.d; ++ss)
      max_flux_derivative = std::max(max_flux_derivative, std::fabs(df[ss]));
  }
  double perimeter_over_volume = std::numeric_limits<double>::min();
  for (int64_t e = 0; e < gr.ne; ++e) {
    int64_t idx[3];
    gr.coords(e, idx);
    double lo[3], ext[3];
    gr.cell(idx, lo, ext);
    double perimeter = 0.;
    for (int k = 0; k < gr.d; ++k)
      for (int s = 0; s < 2; ++s)
        perimeter += make_face(gr, ext, k, s).ie;
    perimeter_over_volume = std::max(perimeter_over_volume, perimeter / gr.volume(ext));
  }
  return 1. / (perimeter_over_volume * max_flux_derivative);
}

int orc_fv_euler(const orc_grid* g, const orc_flux* flux, double* u, double dt, int64_t n_steps, int num_threads)
{
  Grid gr(g);
  std::vector<double> L(gr.ne);
  for (int64_t s = 0; s < n_steps; ++s) {
    fv_apply(gr, *flux, 0, nullptr, u, L.data(), num_threads);
    for (int64_t i = 0; i < gr.ne; ++i) // u_n - L(u_n) * dt
      u[i] = u[i] - L[i] * dt;
  }
  return 0;
}

void orc_fv_interpolate(const orc_grid* g, const orc_function* f, double* u)
{
  Grid gr(g);
  const int d = gr.d;
  const Rule rule(f->order);
  const int m = rule.m, my = d > 1 ? m : 1, mz = d > 2 ? m : 1;
  for (int64_t e = 0; e < gr.ne; ++e) {
    int64_t idx[3];
    gr.coords(e, idx);
    double lower[3], ext[3];
    gr.cell(idx, lower, ext);
    const double vol = gr.volume(ext);
    double integral = 0.;
    for (int qz = 0; qz < mz; ++qz)
      for (int qy = 0; qy < my; ++qy)
        for (int qx = 0; qx < m; ++qx) {
          const double xh[3] = {rule.x[qx], d > 1 ? rule.x[qy] : 0., d > 2 ? rule.x[qz] : 0.};
          const double w = rule.w[qx] * (d > 1 ? rule.w[qy] : 1.) * (d > 2 ? rule.w[qz] : 1.);
          double x[3];
          for (int k = 0; k < 3; ++k)
            x[k] = lower[k] + xh[k] * ext[k];
          Pt pt;
          pt.q = qx + m * (qy + m * qz);
          pt.nq = m * my * mz;
          pt.g = &gr;
          pt.idx = idx;
          pt.xh = xh;
          integral += eval_scalar(f, d, e, x, pt) * vol * w; // XT::Grid::element_integral [EXT]
        }
    u[e] = integral / vol;
  }
}

double orc_function_eval(const orc_function* f, int dim, const double* x, int64_t element)
{
  double xx[3] = {x[0], dim > 1 ? x[1] : 0., dim > 2 ? x[2] : 0.};
  return eval_scalar(f, dim, element, xx);
}


/* ---- "next" rows of SURVEY.md section 8f: constraints, mat-vec, norms, interpolation ---------------------------- */

// DirichletConstraints::apply_local (tools/dirichlet-constraints.hh:85-110): for every boundary intersection of
// Dirichlet type: the local DoFs attached (by their LocalKey) to the intersection itself (codim 1) and to its
// sub-entities of codim 2..d; collected in a std::set per element, then mapped to global indices and inserted into the
// global std::set.  Reference cube: a sub-entity is described by the set of pinned coordinates and their values in
// {0, 1}; it is a sub-entity of face (k, s) iff direction k is pinned to s [EXT dune-geometry ReferenceElement].
int64_t orc_dirichlet_dofs(const orc_grid* g, int kind, int order, uint32_t boundary_mask, int64_t* out)
{
  const Grid gr(g);
  const Space sp(gr, kind, order);
  std::set<int64_t> dirichlet_dofs;
  if (kind != ORC_SPACE_FV && sp.K > 0) {
    const int d = gr.d, K = sp.K, n1 = K + 1;
    std::vector<int64_t> gi(sp.nloc);
    for (int64_t e = 0; e < gr.ne; ++e) {
      int64_t idx[3];
      gr.coords(e, idx);
      std::set<int> local_dofs;
      for (int k = 0; k < d; ++k)
        for (int s = 0; s < 2; ++s) { // intersections in the order x-, x+, y-, y+, z-, z+
          int64_t nb[3];
          bool boundary = false;
          const bool neighbor = gr.neighbor(idx, k, s, nb, &boundary);
          // a periodic wrap face is boundary() && neighbor(): AllDirichletBoundaryInfo does not see it through the
          // periodic view [EXT]; plain boundary faces are Dirichlet iff their bit is set in the mask
          if (!boundary || neighbor || !(boundary_mask >> (2 * k + s) & 1))
            continue;
          // all sub-entities of the face: pinned sets that contain (k -> s)
          for (int pinned = 0; pinned < (1 << d); ++pinned) {
            if (!(pinned >> k & 1))
              continue;
            for (int vals = 0; vals < (1 << d); ++vals) {
              if ((vals & ~pinned) || ((vals >> k & 1) != s))
                continue;
              // local keys on exactly this sub-entity: a_j in {0, K} as pinned, strictly inside otherwise
              for (int i = 0; i < sp.nloc; ++i) {
                const int a[3] = {i % n1, d > 1 ? (i / n1) % n1 : 0, d > 2 ? i / (n1 * n1) : 0};
                bool on = true;
                for (int j = 0; j < d; ++j) {
                  if (pinned >> j & 1)
                    on = on && a[j] == ((vals >> j & 1) ? K : 0);
                  else
                    on = on && a[j] > 0 && a[j] < K;
                }
                if (on)
                  local_dofs.insert(i);
              }
            }
          }
        }
      if (local_dofs.empty())
        continue;
      sp.global_indices(idx, gi.data());
      for (int i : local_dofs)
        dirichlet_dofs.insert(gi[i]);
    }
  }
  if (out) {
    int64_t t = 0;
    for (int64_t dof : dirichlet_dofs)
      out[t++] = dof;
  }
  return (int64_t)dirichlet_dofs.size();
}

// DirichletConstraints::apply(matrix, vector, only_clear, ensure_symmetry) (dirichlet-constraints.hh:122-184) on a CSR
// matrix: unit_col / unit_row (clear_col / clear_row) [EXT XT::LA::MatrixInterface], vector[DoF] = 0
int orc_dirichlet_apply(int64_t rows, const int64_t* rowptr, const int32_t* colidx, double* values, double* vector,
                        int64_t n_dofs, const int64_t* dofs, int only_clear, int ensure_symmetry)
{
  for (int64_t t = 0; t < n_dofs; ++t) {
    const int64_t dof = dofs[t];
    if (values) {
      if (ensure_symmetry) // clear_col / unit_col
        for (int64_t r = 0; r < rows; ++r)
          for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p)
            if (colidx[p] == dof)
              values[p] = (!only_clear && r == dof) ? 1. : 0.;
      bool diag = false;
      for (int64_t p = rowptr[dof]; p < rowptr[dof + 1]; ++p) { // clear_row / unit_row
        diag = diag || colidx[p] == dof;
        values[p] = (!only_clear && colidx[p] == dof) ? 1. : 0.;
      }
      if (!only_clear && !diag) {
        g_error = "unit_row: the diagonal entry is not part of the pattern";
        return 1;
      }
    }
    if (vector)
      vector[dof] = 0.;
  }
  return 0;
}

// matrix_.mv(source, range) (operators/matrix-based.hh:121-129)
void orc_csr_mv(int64_t rows, const int64_t* rowptr, const int32_t* colidx, const double* values, const double* x,
                double* y)
{
  for (int64_t r = 0; r < rows; ++r) {
    double s = 0.;
    for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p)
      s += values[p] * x[colidx[p]];
    y[r] = s;
  }
}

static void builtin_gradient(const orc_function* f, int d, const double* x, double* grad)
{
  grad[0] = grad[1] = grad[2] = 0.;
  if (f->kind != ORC_FN_BUILTIN)
    return;
  switch (f->builtin) {
    case ORC_BUILTIN_COS_PRODUCT:
      for (int k = 0; k < d; ++k) {
        double v = -f->p[0] * f->p[1] * std::sin(f->p[1] * x[k]);
        for (int j = 0; j < d; ++j)
          if (j != k)
            v *= std::cos(f->p[1] * x[j]);
        grad[k] = v;
      }
      break;
    case ORC_BUILTIN_AFFINE:
      for (int k = 0; k < d; ++k)
        grad[k] = f->p[1 + k];
      break;
    case ORC_BUILTIN_GAUSSIAN: {
      const double t = x[0] - f->p[0];
      grad[0] = -(t / (f->p[1] * f->p[1])) * std::exp(-(t * t) / (2. * (f->p[1] * f->p[1])));
      break;
    }
    case ORC_BUILTIN_QUADRATIC:
      for (int k = 0; k < d; ++k)
        grad[k] = 2. * f->p[1] * x[k];
      break;
    default:
      break;
  }
}

// BilinearForm::apply2 (operators/bilinear-form.hh:340-352, 446-452) with source = range = e = u_h - f: the local
// function is a one-element "basis" for LocalElementIntegralBilinearForm::apply2 (integrals.hh:97-134), so per element
// result += sum_q integrand(e, e)(x_q) * (ie * w_q); the element results are summed in walk order.
double orc_bilinear_form_apply2(const orc_grid* g, int kind, int order, const double* dofs, const orc_function* f,
                                const orc_form* form)
{
  const Grid gr(g);
  const Space sp(gr, kind, order);
  const int d = gr.d, n = sp.nloc;
  const int e_order = std::max(dofs ? sp.K : 0, f ? f->order : 0);
  const Rule rule(form_order(*form, e_order, element_integrand_order));
  const int m = rule.m, my = d > 1 ? m : 1, mz = d > 2 ? m : 1;
  Basis b;
  b.d = d;
  b.K = sp.K;
  b.n = n;
  std::vector<int64_t> gi(n);
  double result = 0.;
  for (int64_t e = 0; e < gr.ne; ++e) {
    int64_t idx[3];
    gr.coords(e, idx);
    double lower[3], ext[3];
    gr.cell(idx, lower, ext);
    const double ie = gr.volume(ext);
    sp.global_indices(idx, gi.data());
    double local = 0.;
    for (int qz = 0; qz < mz; ++qz)
      for (int qy = 0; qy < my; ++qy)
        for (int qx = 0; qx < m; ++qx) {
          const double xh[3] = {rule.x[qx], d > 1 ? rule.x[qy] : 0., d > 2 ? rule.x[qz] : 0.};
          const double w = rule.w[qx] * (d > 1 ? rule.w[qy] : 1.) * (d > 2 ? rule.w[qz] : 1.);
          double x[3];
          for (int k = 0; k < 3; ++k)
            x[k] = lower[k] + xh[k] * ext[k];
          double val = 0., grad[3] = {0., 0., 0.};
          if (dofs) {
            b.evaluate(xh, ext);
            for (int i = 0; i < n; ++i) {
              val += dofs[gi[i]] * b.val[i];
              for (int r = 0; r < d; ++r)
                grad[r] += dofs[gi[i]] * b.grad[i * 3 + r];
            }
          }
          if (f) {
            double fg[3];
            val -= eval_scalar(f, d, e, x);
            builtin_gradient(f, d, x, fg);
            for (int r = 0; r < d; ++r)
              grad[r] -= fg[r];
          }
          double v = 0.;
          Pt pt;
          pt.q = qx + m * (qy + m * qz);
          pt.nq = m * my * mz;
          pt.g = &gr;
          pt.idx = idx;
          pt.xh = xh;
          for (int t = 0; t < form->n_terms; ++t) {
            const orc_integrand& in = form->terms[t];
            if (in.kind == ORC_INT_LAPLACE) {
              double kappa[9], kg[3];
              eval_tensor(&in.diffusion, d, e, x, kappa, pt);
              matvec(d, kappa, grad, kg);
              v += dot(d, kg, grad);
            } else
              v += (eval_scalar(&in.diffusion, d, e, x, pt) * val) * val;
          }
          local += v * (ie * w);
        }
    result += form->scaling * local;
  }
  return result;
}

// default_interpolation into a Lagrange space (interpolations/default.hh:40-83): DoFs = f at the Lagrange points,
// elements in walk order, later elements overwrite shared DoFs
void orc_lagrange_interpolate(const orc_grid* g, int kind, int order, const orc_function* f, double* dofs)
{
  const Grid gr(g);
  const Space sp(gr, kind, order);
  const int d = gr.d, K = sp.K, n1 = K + 1;
  std::vector<int64_t> gi(sp.nloc);
  for (int64_t e = 0; e < gr.ne; ++e) {
    int64_t idx[3];
    gr.coords(e, idx);
    double lower[3], ext[3];
    gr.cell(idx, lower, ext);
    sp.global_indices(idx, gi.data());
    for (int i = 0; i < sp.nloc; ++i) {
      const int a[3] = {i % n1, d > 1 ? (i / n1) % n1 : 0, d > 2 ? i / (n1 * n1) : 0};
      double x[3] = {0., 0., 0.};
      for (int k = 0; k < d; ++k)
        x[k] = lower[k] + (K > 0 ? double(a[k]) / double(K) : 0.5) * ext[k];
      dofs[gi[i]] = eval_scalar(f, d, e, x);
    }
  }
}

// LocalLaplaceIntegrand / LocalElementProductIntegrand::evaluate (laplace.hh:81-102, product.hh:104-130) on
// caller-supplied test / ansatz basis values and (physical) gradients at one point: result[ii * n_ansatz + jj].  The
// same code path orc_assemble uses (element_integrand_evaluate2); lets the tests reproduce the reference's pointwise
// known-answer tests with polynomial bases (dune/gdt/test/integrands/integrands_laplace.cc:71-91,
// integrands_product.cc:58-76), which pin the index convention of a non-symmetric diffusion tensor.
int orc_element_integrand_evaluate(const orc_integrand* integrand, int dim, int n_test, const double* test_values,
                                   const double* test_grads /* n_test * dim */, int n_ansatz,
                                   const double* ansatz_values, const double* ansatz_grads /* n_ansatz * dim */,
                                   const double* x, double* result)
{
  if (n_test > MAXN || n_ansatz > MAXN) {
    g_error = "local size exceeds MAXN";
    return 1;
  }
  Basis test, ansatz;
  test.d = ansatz.d = dim;
  test.K = ansatz.K = 0;
  test.n = n_test;
  ansatz.n = n_ansatz;
  for (int i = 0; i < n_test; ++i) {
    test.val[i] = test_values ? test_values[i] : 0.;
    for (int r = 0; r < 3; ++r)
      test.grad[i * 3 + r] = (test_grads && r < dim) ? test_grads[i * dim + r] : 0.;
  }
  for (int i = 0; i < n_ansatz; ++i) {
    ansatz.val[i] = ansatz_values ? ansatz_values[i] : 0.;
    for (int r = 0; r < 3; ++r)
      ansatz.grad[i * 3 + r] = (ansatz_grads && r < dim) ? ansatz_grads[i * dim + r] : 0.;
  }
  double xx[3] = {x[0], dim > 1 ? x[1] : 0., dim > 2 ? x[2] : 0.};
  element_integrand_evaluate2(*integrand, test, ansatz, dim, 0, xx, result);
  return 0;
}


// ---- systems (m = d + 2): Euler equations ---------------------------------------------------------------------------
int orc_fvsys_apply(const orc_grid* g, const orc_flux* flux, const double* u, double* out)
{
  Grid gr(g);
  if (flux->kind != ORC_FLUX_EULER || gr.d > 2)
    return 1;
  fvsys_apply(gr, *flux, u, out);
  return 0;
}

// the same with impermeable walls on the (non-periodic) domain sides of the two masks (bit 2k + s)
int orc_fvsys_apply_walls(const orc_grid* g, const orc_flux* flux, uint32_t wall_mask, uint32_t mirror_mask, const double* u,
                          double* out)
{
  Grid gr(g);
  if (flux->kind != ORC_FLUX_EULER || gr.d > 2)
    return 1;
  fvsys_apply(gr, *flux, u, out, wall_mask, mirror_mask);
  return 0;
}

// explicit_euler of examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:141-159: u <- u - L(u) dt, n_steps times
int orc_fvsys_euler(const orc_grid* g, const orc_flux* flux, double* u, double dt, int64_t n_steps)
{
  Grid gr(g);
  if (flux->kind != ORC_FLUX_EULER || gr.d > 2)
    return 1;
  const int64_t n = gr.ne * (gr.d + 2);
  std::vector<double> L(n);
  for (int64_t s = 0; s < n_steps; ++s) {
    fvsys_apply(gr, *flux, u, L.data());
    for (int64_t i = 0; i < n; ++i)
      u[i] = u[i] - L[i] * dt;
  }
  return 0;
}

// estimate_dt_for_hyperbolic_system (tools/hyperbolic.hh:38-86) for the Euler flux (order 4) and a finite volume state
double orc_fvsys_estimate_dt(const orc_grid* g, const orc_flux* flux, const double* u)
{
  Grid gr(g);
  const Euler eu{gr.d, flux->p[0]};
  const int M = eu.m();
  double lo[4], hi[4];
  for (int c = 0; c < M; ++c) {
    lo[c] = std::numeric_limits<double>::max();
    hi[c] = std::numeric_limits<double>::min();
  }
  for (int64_t e = 0; e < gr.ne; ++e)
    for (int c = 0; c < M; ++c) {
      lo[c] = std::min(lo[c], u[e * M + c]);
      hi[c] = std::max(hi[c], u[e * M + c]);
    }
  for (int c = 0; c < M; ++c)
    if (!(lo[c] < hi[c]))
      hi[c] = lo[c] + 1e-6 * lo[c];
  double max_flux_derivative = std::numeric_limits<double>::min();
  const int nq = gauss_m(4);
  double qx[8], qw[8];
  gauss01(nq, qx, qw);
  int64_t total = 1;
  for (int c = 0; c < M; ++c)
    total *= nq;
  std::vector<double> J(gr.d * M * M);
  for (int64_t t = 0; t < total; ++t) {
    double w[4];
    int64_t r = t;
    for (int c = 0; c < M; ++c) {
      w[c] = lo[c] + qx[r % nq] * (hi[c] - lo[c]);
      r /= nq;
    }
    eu.jacobian(w, J.data());
    for (int ss = 0; ss < gr.d; ++ss)
      for (int rr = 0; rr < M; ++rr) {
        double sum = 0.;
        for (int cc = 0; cc < M; ++cc)
          sum += std::fabs(J[(ss * M + rr) * M + cc]);
        max_flux_derivative = std::max(max_flux_derivative, sum); // FieldMatrix::infinity_norm
      }
  }
  double perimeter_over_volume = std::numeric_limits<double>::min();
  for (int64_t e = 0; e < gr.ne; ++e) {
    int64_t idx[3];
    gr.coords(e, idx);
    double lower[3], ext[3];
    gr.cell(idx, lower, ext);
    double perimeter = 0;
    for (int k = 0; k < gr.d; ++k)
      for (int s = 0; s < 2; ++s)
        perimeter += make_face(gr, ext, k, s).ie;
    perimeter_over_volume = std::max(perimeter_over_volume, perimeter / gr.volume(ext));
  }
  return 1. / (perimeter_over_volume * max_flux_derivative);
}

// EulerTools pieces for the property tests: flux [d][m], jacobian [d][m][m], eigendecomposition of jacobian . n
void orc_euler_flux(int d, double gamma, const double* w, double* f)
{
  Euler{d, gamma}.flux(w, f);
}
void orc_euler_jacobian(int d, double gamma, const double* w, double* J)
{
  Euler{d, gamma}.jacobian(w, J);
}
void orc_euler_eigen(int d, double gamma, const double* w, const double* n, double* ev, double* T, double* Ti)
{
  Euler{d, gamma}.eigen(w, n, ev, T, Ti);
}

} // extern "C"

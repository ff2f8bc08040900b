// dune-gdt_b200/csrc/capi.cu -- the C ABI of libgdtb (include/gdtb.h): handles, lowering of the appended local
// forms to kernel descriptors, kernel selection, and the host<->device plumbing.  No CPU compute path exists:
// every gdtb_assemble / gdtb_fvop_* call ends in the CUDA kernels of this directory or fails.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include <cstdio>

#include "fv_system.hpp"
#include "kernels.hpp"

namespace gdtb {

static thread_local std::string g_last_error;

void set_error(const std::string& msg)
{
  g_last_error = msg;
}

int fail(int code, const std::string& msg)
{
  g_last_error = msg;
  return code;
}

// ------------------------------------------------------------------------------------------------
// host-side numerics needed for lowering: Gauss-Legendre rules on [0,1] and 1D Lagrange tables
// ([EXT] dune-geometry QuadratureRules / dune-localfunctions Lagrange basis)
// ------------------------------------------------------------------------------------------------
int gauss_points_for_order(int order)
{
  return std::max(order, 0) / 2 + 1;
}

void gauss_legendre_01(int m, double* x, double* w)
{
  // Newton iteration on the Legendre polynomial P_m, extended precision, roots ascending
  for (int i = 0; i < m; ++i) {
    long double z = cosl(3.14159265358979323846264338327950288L * (i + 0.75L) / (m + 0.5L));
    long double dp = 1.0L;
    for (int it = 0; it < 100; ++it) {
      long double p0 = 1.0L, p1 = z; // P_0, P_1
      for (int j = 2; j <= m; ++j) {
        const long double pj = ((2.0L * j - 1.0L) * z * p1 - (j - 1.0L) * p0) / j;
        p0 = p1;
        p1 = pj;
      }
      if (m == 0)
        p1 = 1.0L;
      dp = m * (z * p1 - p0) / (z * z - 1.0L);
      const long double dz = p1 / dp;
      z -= dz;
      if (fabsl(dz) < 1e-19L)
        break;
    }
    x[m - 1 - i] = (double)((1.0L + z) / 2.0L);
    w[m - 1 - i] = (double)(1.0L / ((1.0L - z * z) * dp * dp));
  }
}

void lagrange_1d(int K, double x, double* v, double* dv)
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

} // namespace gdtb

using namespace gdtb;

#include "handles.hpp"

namespace {

int check_ctx(gdtb_ctx* ctx)
{
  if (!ctx)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "context is NULL");
  GDTB_CUDA(cudaSetDevice(ctx->device));
  return GDTB_OK;
}

// SpaceDev of a space on a grid (host arithmetic only): sizes and the MCMG offsets of the continuous mapper
int make_space_dev(const GridDev& g, int kind, int order, SpaceDev& sp)
{
  if (kind < GDTB_SPACE_CG || kind > GDTB_SPACE_FV)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "unknown space kind");
  const int d = g.d;
  int K = order;
  if (kind == GDTB_SPACE_FV)
    K = 0;
  else if (kind == GDTB_SPACE_CG && order < 1)
    return fail(GDTB_ERR_SPACE, "continuous Lagrange spaces need order >= 1");
  else if (order < 0)
    return fail(GDTB_ERR_SPACE, "negative polynomial order");
  if (K > MAX_K)
    return fail(GDTB_ERR_FINITE_ELEMENT, "Lagrange order not supported (max 3)");
  if (kind == GDTB_SPACE_CG && g.periodic)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "continuous Lagrange spaces on periodic grid views are not supported");
  std::memset(&sp, 0, sizeof(sp));
  sp.kind = kind;
  sp.K = K;
  sp.d = d;
  sp.nloc = ipow(K + 1, d);
  if (kind == GDTB_SPACE_CG) {
    // MCMGMapper offsets: codim 0..d, YaspGrid sub-entity groups by shift bitset (common.cuh)
    long long running = 0;
    for (int c = 0; c <= d; ++c) {
      long long b = 1;
      for (int j = 0; j < d - c; ++j)
        b *= (K - 1);
      sp.cg.block[c] = b;
      sp.cg.codim_offset[c] = running;
      long long entities = 0;
      for (int sh = 0; sh < (1 << d); ++sh) {
        int pc = 0;
        for (int k = 0; k < d; ++k)
          pc += (sh >> k) & 1;
        if (pc != d - c)
          continue;
        sp.cg.group_offset[sh] = entities;
        long long cnt = 1;
        for (int k = 0; k < d; ++k)
          cnt *= ((sh >> k) & 1) ? g.n[k] : g.n[k] + 1;
        entities += cnt;
      }
      running += entities * b;
    }
    sp.size = running;
  } else
    sp.size = g.ne * sp.nloc;
  return GDTB_OK;
}


long long function_data_size(const gdtb_function& f, const GridDev& g)
{
  switch (f.kind) {
    case GDTB_FN_ELEM_SCALAR: return g.ne;
    case GDTB_FN_ELEM_TENSOR: return g.ne * g.d * g.d;
    case GDTB_FN_QP_SCALAR: return g.ne * f.qp_per_element;
    case GDTB_FN_QP_TENSOR: return g.ne * f.qp_per_element * g.d * g.d;
    case GDTB_FN_QP_VALUE_GRAD: return g.ne * f.qp_per_element * (1 + g.d);
    case GDTB_FN_DOF_VECTOR: {
      SpaceDev sp;
      return make_space_dev(g, f.space_kind, f.space_order, sp) == GDTB_OK ? sp.size : 0;
    }
    default: return 0;
  }
}

bool fn_has_data(const gdtb_function& f)
{
  return f.kind == GDTB_FN_ELEM_SCALAR || f.kind == GDTB_FN_ELEM_TENSOR || f.kind == GDTB_FN_QP_SCALAR
         || f.kind == GDTB_FN_QP_TENSOR || f.kind == GDTB_FN_DOF_VECTOR || f.kind == GDTB_FN_QP_VALUE_GRAD;
}

int validate_function(const gdtb_function& f, const char* what)
{
  if (f.kind < GDTB_FN_CONST_SCALAR || f.kind > GDTB_FN_QP_VALUE_GRAD)
    return fail(GDTB_ERR_INVALID_ARGUMENT, std::string(what) + ": unknown function kind");
  if (f.kind == GDTB_FN_BUILTIN && (f.builtin < GDTB_BUILTIN_COS_PRODUCT || f.builtin > GDTB_BUILTIN_QUADRATIC))
    return fail(GDTB_ERR_INVALID_ARGUMENT, std::string(what) + ": unknown built-in function id");
  if (fn_has_data(f) && !f.data)
    return fail(GDTB_ERR_INVALID_ARGUMENT, std::string(what) + ": array-backed function without data");
  if ((f.kind == GDTB_FN_QP_SCALAR || f.kind == GDTB_FN_QP_TENSOR || f.kind == GDTB_FN_QP_VALUE_GRAD) && f.qp_per_element < 1)
    return fail(GDTB_ERR_INVALID_ARGUMENT, std::string(what) + ": per-quadrature-point function needs qp_per_element >= 1");
  if (f.order < 0)
    return fail(GDTB_ERR_INVALID_ARGUMENT, std::string(what) + ": negative polynomial order");
  return GDTB_OK;
}

// clone-on-append of a grid function: host arrays are copied to the device; a discrete function also gets a
// device-resident copy of its space description (the mapper the kernels evaluate it through)
int lower_function(gdtb_ctx* ctx, const GridDev& g, gdtb_function& f, LoweredForm& owner)
{
  if (f.kind == GDTB_FN_DOF_VECTOR) {
    SpaceDev sp;
    GDTB_TRY(make_space_dev(g, f.space_kind, f.space_order, sp));
  }
  const long long n = function_data_size(f, g);
  if (n > 0 && !f.data_on_device) {
    double* d = nullptr;
    if (cudaMalloc(&d, sizeof(double) * (size_t)n) != cudaSuccess)
      return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory while cloning an array-backed grid function");
    owner.owned.push_back(d);
    GDTB_CUDA(cudaMemcpyAsync(d, f.data, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->launch.stream));
    GDTB_CUDA(cudaStreamSynchronize(ctx->launch.stream));
    f.data = d;
    f.data_on_device = 1;
  }
  if (f.kind == GDTB_FN_DOF_VECTOR) {
    SpaceDev sp;
    GDTB_TRY(make_space_dev(g, f.space_kind, f.space_order, sp));
    void* d = nullptr;
    if (cudaMalloc(&d, sizeof(SpaceDev)) != cudaSuccess)
      return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory while cloning a discrete function");
    owner.owned.push_back(static_cast<double*>(d));
    GDTB_CUDA(cudaMemcpyAsync(d, &sp, sizeof(SpaceDev), cudaMemcpyHostToDevice, ctx->launch.stream));
    GDTB_CUDA(cudaStreamSynchronize(ctx->launch.stream));
    owner.dof_spaces.emplace_back(f.data, static_cast<const SpaceDev*>(d));
  }
  return GDTB_OK;
}

int lower_form(gdtb_ctx* ctx, const GridDev& g, const gdtb_form* form, int filter, LoweredForm& out)
{
  if (!form)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "form is NULL");
  if (form->n_terms < 1 || form->n_terms > GDTB_MAX_TERMS)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "form: n_terms must be in [1, GDTB_MAX_TERMS]");
  out.form = *form;
  out.filter = filter;
  for (int t = 0; t < form->n_terms; ++t) {
    GDTB_TRY(validate_function(out.form.terms[t].diffusion, "integrand.diffusion"));
    GDTB_TRY(validate_function(out.form.terms[t].weight, "integrand.weight"));
    if (out.form.terms[t].diffusion.kind == GDTB_FN_QP_VALUE_GRAD || out.form.terms[t].weight.kind == GDTB_FN_QP_VALUE_GRAD)
      return fail(GDTB_ERR_INVALID_ARGUMENT, "GDTB_FN_QP_VALUE_GRAD is the `f` of gdtb_bilinear_form_apply2, not a coefficient");
    int st = lower_function(ctx, g, out.form.terms[t].diffusion, out);
    if (st == GDTB_OK)
      st = lower_function(ctx, g, out.form.terms[t].weight, out);
    if (st != GDTB_OK) {
      free_form(out);
      return st;
    }
  }
  return GDTB_OK;
}

FnDev to_dev(const gdtb_function& f, const LoweredForm* owner = nullptr)
{
  FnDev d;
  d.kind = f.kind;
  d.order = f.order;
  d.builtin = f.builtin;
  d.nq = f.qp_per_element;
  std::memcpy(d.c, f.c, sizeof(d.c));
  std::memcpy(d.p, f.p, sizeof(d.p));
  d.data = f.data;
  d.space = nullptr;
  if (owner)
    for (const auto& ds : owner->dof_spaces)
      if (ds.first == f.data)
        d.space = ds.second;
  return d;
}

enum FormRole
{
  ROLE_ELEMENT,
  ROLE_RHS,
  ROLE_COUPLING,
  ROLE_BOUNDARY
};

// quadrature order of a form exactly as the reference derives it
int form_quadrature_order(const gdtb_form& f, int K, FormRole role)
{
  int order = 0;
  for (int t = 0; t < f.n_terms; ++t) {
    const gdtb_integrand& in = f.terms[t];
    int o = 0;
    switch (role) {
      case ROLE_ELEMENT: // laplace.hh:74-79, product.hh:89-100
        o = in.diffusion.order + K + K;
        break;
      case ROLE_RHS: // conversion.hh:92 -> product.hh:99
        o = in.diffusion.order + K + in.weight.order;
        break;
      case ROLE_COUPLING: // laplace-ipdg.hh:95-105, ipdg.hh:100-111
        o = (in.kind == GDTB_INT_IPDG_INNER_COUPLING ? in.diffusion.order : 0) + in.weight.order + K + K;
        break;
      case ROLE_BOUNDARY: // laplace-ipdg.hh:331-338, ipdg.hh:245-252
        o = (in.kind == GDTB_INT_IPDG_DIRICHLET_COUPLING ? in.diffusion.order : in.weight.order) + K + K;
        break;
    }
    order = std::max(order, o); // combined.hh:293-299
  }
  return order + f.over_integrate;
}

int make_form_dev(const gdtb_form& f, int K, FormRole role, FormDev& out, const LoweredForm* owner = nullptr)
{
  std::memset(&out, 0, sizeof(out));
  out.n_terms = f.n_terms;
  out.scaling = f.scaling;
  out.m = gauss_points_for_order(form_quadrature_order(f, K, role));
  if (out.m > MAX_Q1D)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "quadrature order too high (more than 8 Gauss points per direction)");
  gauss_legendre_01(out.m, out.qx, out.qw);
  for (int q = 0; q < out.m; ++q)
    lagrange_1d(K, out.qx[q], out.phi[q], out.dphi[q]);
  lagrange_1d(K, 0., out.phi_end[0], out.dphi_end[0]);
  lagrange_1d(K, 1., out.phi_end[1], out.dphi_end[1]);
  for (int t = 0; t < f.n_terms; ++t) {
    IntegrandDev& d = out.terms[t];
    d.kind = f.terms[t].kind;
    d.hI_kind = f.terms[t].hI_kind;
    d.prefactor = f.terms[t].prefactor;
    d.diffusion = to_dev(f.terms[t].diffusion, owner);
    d.weight = to_dev(f.terms[t].weight, owner);
  }
  return GDTB_OK;
}

// caller-sampled coefficient arrays must have been sampled for the rule the form is integrated with
int check_qp_functions(const gdtb_form& f, int K, FormRole role, int d)
{
  bool any = false;
  for (int t = 0; t < f.n_terms; ++t)
    for (const gdtb_function* fn : {&f.terms[t].diffusion, &f.terms[t].weight})
      any = any || fn->kind == GDTB_FN_QP_SCALAR || fn->kind == GDTB_FN_QP_TENSOR;
  if (!any)
    return GDTB_OK;
  if (role == ROLE_COUPLING || role == ROLE_BOUNDARY)
    return fail(GDTB_ERR_NOT_IMPLEMENTED,
                "per-quadrature-point functions (GDTB_FN_QP_*) are volume-rule data: element forms and functionals only");
  const int m = gauss_points_for_order(form_quadrature_order(f, K, role));
  const int nq = ipow(m, d);
  for (int t = 0; t < f.n_terms; ++t)
    for (const gdtb_function* fn : {&f.terms[t].diffusion, &f.terms[t].weight})
      if ((fn->kind == GDTB_FN_QP_SCALAR || fn->kind == GDTB_FN_QP_TENSOR) && fn->qp_per_element != nq)
        return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH,
                    "per-quadrature-point function: qp_per_element = " + std::to_string(fn->qp_per_element)
                        + " but the form is integrated with " + std::to_string(nq) + " points per element ("
                        + std::to_string(m) + " per direction; see gdtb_form_quadrature_order / gdtb_gauss_rule)");
  return GDTB_OK;
}

// ---- CG-Q1 gather eligibility and lowering ------------------------------------------------------
bool fn_is_const(const gdtb_function& f)
{
  return f.kind == GDTB_FN_CONST_SCALAR || f.kind == GDTB_FN_CONST_TENSOR;
}

bool q1_space(const SpaceDev& sp)
{
  return sp.kind == GDTB_SPACE_CG && sp.K == 1;
}

bool matop_q1_eligible(const gdtb_matop* op)
{
  if (!q1_space(op->test) || !q1_space(op->ansatz) || op->grid.periodic)
    return false;
  if (op->pattern
      && (op->pattern->stencil != GDTB_STENCIL_ELEMENT || !q1_space(op->pattern->test)
          || !q1_space(op->pattern->ansatz)))
    return false;
  if (!op->coupling_forms.empty() || !op->boundary_forms.empty() || op->element_forms.empty())
    return false;
  // element-wise constant coefficients only: then the quadrature sum factorises exactly (kernels.hpp, Q1Group)
  int n_groups = 0;
  for (const auto& lf : op->element_forms)
    for (int t = 0; t < lf.form.n_terms; ++t) {
      const gdtb_integrand& in = lf.form.terms[t];
      if (in.kind == GDTB_INT_LAPLACE) {
        if (in.diffusion.kind >= GDTB_FN_BUILTIN) // varies inside a cell: analytic, per quadrature point, discrete
          return false;
      } else if (in.kind == GDTB_INT_PRODUCT) {
        if (in.diffusion.kind != GDTB_FN_CONST_SCALAR && in.diffusion.kind != GDTB_FN_ELEM_SCALAR)
          return false;
      } else
        return false;
      ++n_groups;
    }
  return n_groups <= Q1G_MAX_GROUPS;
}

// CG-Q2 row-gather path (assemble_q2_gather.cu): element forms, Laplace with kappa = c I or mass, coefficients
// constant or one scalar per element; needs the element pattern (its rowptr is read, colidx is not)
bool matop_q2_eligible(const gdtb_matop* op)
{
  const auto q2 = [](const SpaceDev& sp) { return sp.kind == GDTB_SPACE_CG && sp.K == 2; };
  if (!q2(op->test) || !q2(op->ansatz) || op->grid.periodic || (op->grid.d != 2 && op->grid.d != 3))
    return false;
  // the kernel places every value by closed forms; a caller-provided pattern must be the element stencil they describe
  if (op->pattern
      && (op->pattern->stencil != GDTB_STENCIL_ELEMENT || !q2(op->pattern->test) || !q2(op->pattern->ansatz)))
    return false;
  if (!op->coupling_forms.empty() || !op->boundary_forms.empty() || op->element_forms.empty())
    return false;
  int n_groups = 0;
  for (const auto& lf : op->element_forms)
    for (int t = 0; t < lf.form.n_terms; ++t) {
      const gdtb_integrand& in = lf.form.terms[t];
      if (in.kind != GDTB_INT_LAPLACE && in.kind != GDTB_INT_PRODUCT)
        return false;
      if (in.diffusion.kind != GDTB_FN_CONST_SCALAR && in.diffusion.kind != GDTB_FN_ELEM_SCALAR)
        return false;
      ++n_groups;
    }
  return n_groups <= Q2G_MAX_GROUPS;
}

// DG row-gather path (assemble_dg_gather.cu): any mix of element / inner-coupling / boundary forms on a non-periodic
// grid with the element_and_intersection pattern
// every coefficient of the DG forms is a constant or one scalar per element: the factorised kernels apply
bool matop_dg_scalar_coefficients(const gdtb_matop* op)
{
  const auto scalar = [](const gdtb_function& f) { return f.kind == GDTB_FN_CONST_SCALAR || f.kind == GDTB_FN_ELEM_SCALAR; };
  for (const auto* list : {&op->element_forms, &op->coupling_forms, &op->boundary_forms})
    for (const auto& lf : *list)
      for (int t = 0; t < lf.form.n_terms; ++t)
        if (!scalar(lf.form.terms[t].diffusion) || !scalar(lf.form.terms[t].weight))
          return false;
  return true;
}

bool grid_dg_closed_form(const GridDev& g)
{
  for (int k = 0; k < g.d; ++k)
    if (((g.periodic >> k) & 1) && g.n[k] < 3)
      return false;
  return true;
}

bool matop_dg_eligible(const gdtb_matop* op)
{
  if (op->test.kind != GDTB_SPACE_DG || op->ansatz.kind != GDTB_SPACE_DG)
    return false;
  if (!dg_gather_supported(op->grid.d, op->test.K))
    return false;
  // periodic grid views: the factorised kernels know the wrap neighbours (periodic directions with >= 3 cells); the
  // quadrature-faithful gather kernel does not
  if (op->grid.periodic
      && !(grid_dg_closed_form(op->grid) && dg_gather_fast_supported(op->grid, op->test.K)
           && matop_dg_scalar_coefficients(op) && !std::getenv("GDTB_DG_NO_FAST")))
    return false;
  // a pattern-free operator follows the closed-form element_and_intersection stencil; a given pattern must be that one
  if (op->pattern
      && (op->pattern->stencil != GDTB_STENCIL_ELEMENT_AND_INTERSECTION || op->pattern->test.kind != GDTB_SPACE_DG
          || op->pattern->test.K != op->test.K))
    return false;
  const size_t n = op->element_forms.size() + op->coupling_forms.size() + op->boundary_forms.size();
  if (n == 0 || n > (size_t)DGG_MAX_FORMS)
    return false;
  for (const auto& lf : op->boundary_forms)
    if (lf.filter != GDTB_FILTER_ALL_BOUNDARY)
      return false;
  return true;
}

// CG Q1 / Q2 row gather with coefficients that vary inside a cell (assemble_q1_gather.cu::k_q1_gather_qp,
// assemble_q2_gather.cu SF = 3): Laplace (scalar or full-tensor kappa) / product element forms whose coefficients are
// arbitrary grid functions -- sampled by the caller (GDTB_FN_QP_*), analytic, discrete (GDTB_FN_DOF_VECTOR) or, mixed in
// with those, constants / per-element values.  Every coefficient is turned into one value per quadrature point.
bool fn_is_tensor(const gdtb_function& f)
{
  return f.kind == GDTB_FN_CONST_TENSOR || f.kind == GDTB_FN_ELEM_TENSOR || f.kind == GDTB_FN_QP_TENSOR;
}

bool matop_cg_qp_eligible(const gdtb_matop* op)
{
  const SpaceDev& sp = op->test;
  if (sp.kind != GDTB_SPACE_CG || std::memcmp(&op->test, &op->ansatz, sizeof(SpaceDev)) != 0 || op->grid.periodic)
    return false;
  const int d = op->grid.d;
  if (!(sp.K == 1 || (sp.K == 2 && (d == 2 || d == 3))))
    return false;
  if (op->pattern
      && (op->pattern->stencil != GDTB_STENCIL_ELEMENT || std::memcmp(&op->pattern->test, &sp, sizeof(SpaceDev)) != 0
          || std::memcmp(&op->pattern->ansatz, &sp, sizeof(SpaceDev)) != 0))
    return false;
  if (!op->coupling_forms.empty() || !op->boundary_forms.empty() || op->element_forms.empty())
    return false;
  if (std::getenv("GDTB_NO_QP_GATHER"))
    return false;
  for (const auto& lf : op->element_forms) {
    const int m = gauss_points_for_order(form_quadrature_order(lf.form, sp.K, ROLE_ELEMENT));
    for (int t = 0; t < lf.form.n_terms; ++t) {
      const gdtb_integrand& in = lf.form.terms[t];
      if (in.kind != GDTB_INT_LAPLACE && in.kind != GDTB_INT_PRODUCT)
        return false;
      if (in.kind == GDTB_INT_PRODUCT && fn_is_tensor(in.diffusion))
        return false;
      const int kind = in.kind == GDTB_INT_PRODUCT ? Q1G_MASS
                                                    : (fn_is_tensor(in.diffusion) ? Q1G_LAPLACE_TENSOR : Q1G_LAPLACE_SCALAR);
      if (!(sp.K == 1 ? q1_qp_supported(d, m, kind) : q2_qp_supported(d, m, kind)))
        return false;
    }
  }
  return true;
}

// 1D point tables of a rule: pt[t][q][a][b] = w_q D^ta phi_a(x_q) D^tb phi_b(x_q) (kernels.hpp, QPT_*)
void qp_point_tables(int K, int m, double qx[MAX_Q1D], double pt[4][MAX_Q1D][3][3])
{
  double qw[MAX_Q1D];
  gauss_legendre_01(m, qx, qw);
  std::memset(pt, 0, sizeof(double) * 4 * MAX_Q1D * 9);
  for (int q = 0; q < m; ++q) {
    double v[MAX_K + 1], dv[MAX_K + 1];
    lagrange_1d(K, qx[q], v, dv);
    for (int a = 0; a <= K; ++a)
      for (int b = 0; b <= K; ++b) {
        pt[QPT_MM][q][a][b] = qw[q] * v[a] * v[b];
        pt[QPT_KK][q][a][b] = qw[q] * dv[a] * dv[b];
        pt[QPT_KM][q][a][b] = qw[q] * dv[a] * v[b];
        pt[QPT_MK][q][a][b] = qw[q] * v[a] * dv[b];
      }
  }
}

int build_q2_params(const gdtb_matop* op, Q2GatherParams& p)
{
  std::memset(&p, 0, sizeof(p));
  p.g = op->grid;
  p.rowptr = op->pattern ? op->pattern->d_rowptr : nullptr;
  for (const auto& lf : op->element_forms) {
    // 1D reference tables with the form's own Gauss rule (order logic of laplace.hh:74-79 / product.hh:89-100)
    const int m = gauss_points_for_order(form_quadrature_order(lf.form, 2, ROLE_ELEMENT));
    if (m > MAX_Q1D)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "quadrature order too high (more than 8 Gauss points per direction)");
    double qx[MAX_Q1D], qw[MAX_Q1D];
    gauss_legendre_01(m, qx, qw);
    double TM[3][3] = {{0}}, TK[3][3] = {{0}};
    for (int q = 0; q < m; ++q) {
      double v[3], dv[3];
      lagrange_1d(2, qx[q], v, dv);
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          TM[a][b] += qw[q] * v[a] * v[b];
          TK[a][b] += qw[q] * dv[a] * dv[b];
        }
    }
    for (int t = 0; t < lf.form.n_terms; ++t) {
      const gdtb_integrand& in = lf.form.terms[t];
      Q2Group& G = p.group[p.n_groups++];
      G.kind = in.kind == GDTB_INT_PRODUCT ? Q1G_MASS : Q1G_LAPLACE_SCALAR;
      G.coef_elem = in.diffusion.kind == GDTB_FN_ELEM_SCALAR ? 1 : 0;
      G.scale = lf.form.scaling * (G.coef_elem ? 1. : in.diffusion.c[0]);
      G.coef = in.diffusion.data;
      std::memcpy(G.TM, TM, sizeof(TM));
      std::memcpy(G.TK, TK, sizeof(TK));
    }
  }
  return GDTB_OK;
}

bool builtin_is_separable(int id)
{
  return id == GDTB_BUILTIN_COS_PRODUCT || id == GDTB_BUILTIN_GAUSSIAN || id == GDTB_BUILTIN_INDICATOR;
}

bool vecfun_q1_eligible(const gdtb_vecfun* fun)
{
  if (!q1_space(fun->space) || fun->grid.periodic || fun->forms.empty())
    return false;
  int n_elem = 0, n_sep = 0;
  for (const auto& lf : fun->forms) {
    if (lf.form.n_terms != 1)
      return false;
    const gdtb_integrand& in = lf.form.terms[0];
    if (in.kind != GDTB_INT_PRODUCT || in.diffusion.kind != GDTB_FN_CONST_SCALAR)
      return false;
    if (in.weight.kind == GDTB_FN_ELEM_SCALAR)
      ++n_elem;
    else if (in.weight.kind == GDTB_FN_BUILTIN && builtin_is_separable(in.weight.builtin))
      ++n_sep;
    else if (in.weight.kind != GDTB_FN_CONST_SCALAR)
      return false;
  }
  return n_elem <= 1 && n_sep <= 1;
}

// 1D integrals of the Q1 shape functions with an m-point Gauss rule:
// G[ta][tb][a][b] = sum_q w_q D^ta phi_a(x_q) D^tb phi_b(x_q), s1[a] = sum_q w_q phi_a(x_q)
struct Q1Tables
{
  double G[2][2][2][2];
  double s1[2];
  int m;
  double qx[MAX_Q1D], qw[MAX_Q1D], phi[MAX_Q1D][2];
};

void q1_tables(int order, Q1Tables& t)
{
  std::memset(&t, 0, sizeof(t));
  t.m = gauss_points_for_order(order);
  gauss_legendre_01(t.m, t.qx, t.qw);
  for (int q = 0; q < t.m; ++q) {
    double v[2], dv[2];
    lagrange_1d(1, t.qx[q], v, dv);
    t.phi[q][0] = v[0];
    t.phi[q][1] = v[1];
    const double* D[2] = {v, dv};
    for (int ta = 0; ta < 2; ++ta)
      for (int tb = 0; tb < 2; ++tb)
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b)
            t.G[ta][tb][a][b] += t.qw[q] * D[ta][a] * D[tb][b];
    t.s1[0] += t.qw[q] * v[0];
    t.s1[1] += t.qw[q] * v[1];
  }
}

// reference-element tensors of one integrand for the gather kernel (kernels.hpp, Q1Group):
// M[r*3+c][o][s] = sum_q w_q d_r phihat_i(x_q) d_c phihat_j(x_q) with i = (2^d - 1) ^ o (the vertex seen from the
// element at offset o), j = s; tensor-product rule => product of the 1D tables.  Mass: M[0][o][s] = sum_q w_q phihat_i phihat_j.
void q1_group_tensors(int kind, int d, const Q1Tables& tab, double M[9][8][8])
{
  const int n = 1 << d;
  std::memset(M, 0, sizeof(double) * 9 * 64);
  for (int o = 0; o < n; ++o) {
    const int i = (n - 1) ^ o; // a_k = 1 - o_k
    for (int j = 0; j < n; ++j) {
      if (kind == Q1G_MASS) {
        double prod = 1.;
        for (int k = 0; k < d; ++k)
          prod *= tab.G[0][0][(i >> k) & 1][(j >> k) & 1];
        M[0][o][j] = prod;
        continue;
      }
      // v_ij = sum_{r,c} kappa_rc d_c phi_j d_r phi_i (laplace.hh:98-101)
      for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c) {
          if (kind == Q1G_LAPLACE_SCALAR && r != c)
            continue;
          double prod = 1.;
          for (int k = 0; k < d; ++k)
            prod *= tab.G[k == r ? 1 : 0][k == c ? 1 : 0][(i >> k) & 1][(j >> k) & 1];
          M[r * 3 + c][o][j] = prod;
        }
    }
  }
}

// closed-form CSR row pointer of the CG-Q1 element stencil (same formula as the kernels)
long long q1_S(long long i, long long N)
{
  return i == 0 ? 0 : (i > N ? 3 * N + 1 : 3 * i - 1);
}

// global CSR position of the first entry of vertex layer `layer` along the last direction
long long q1_layer_rowptr(const GridDev& g, long long layer)
{
  long long w = 1;
  for (int k = 0; k < g.d - 1; ++k)
    w *= 3 * g.n[k] + 1;
  return q1_S(layer, g.n[g.d - 1]) * w;
}

long long q1_layer_rows(const GridDev& g)
{
  long long v = 1;
  for (int k = 0; k < g.d - 1; ++k)
    v *= g.n[k] + 1;
  return v;
}

// owner-computes-rows ranges of the slab [begin, end) of element layers: vertex layers [begin, end) plus the top
// layer on the last slab, produced from the element layers [begin - 1, end)
void q1_slab_ranges(const GridDev& g, long long begin, long long end, long long& row_lo, long long& row_hi,
                    long long& elem_lo, long long& elem_hi)
{
  const long long n_last = g.n[g.d - 1];
  row_lo = begin;
  row_hi = end == n_last ? n_last + 1 : end;
  elem_lo = std::max<long long>(begin - 1, 0);
  elem_hi = end;
}

// per-axis geometry tables of a grid (ignoring the slab), built on the device at first use
int q1_axis_tables(gdtb_ctx* ctx, const GridDev& grid, const double* (&axis_tab)[3], long long& axis_tab_inv);

int q1_axis_tables(gdtb_ctx* ctx, const GridDev& grid, Q1GatherParams& p)
{
  return q1_axis_tables(ctx, grid, p.axis_tab, p.axis_tab_inv);
}

int q1_axis_tables(gdtb_ctx* ctx, const GridDev& grid, const double* (&axis_tab)[3], long long& axis_tab_inv)
{
  GridDev key = grid;
  key.layer_lo = 0;
  key.layer_hi = 0;
  const gdtb_ctx::AxisTables* found = nullptr;
  for (const auto& t : ctx->axis_tables)
    if (std::memcmp(&t.grid, &key, sizeof(GridDev)) == 0)
      found = &t;
  if (!found) {
    gdtb_ctx::AxisTables t;
    t.grid = key;
    long long total = 0;
    for (int k = 0; k < 3; ++k) {
      t.offset[k] = total;
      total += (k < key.d ? key.n[k] : 0) + 2;
    }
    t.inv = total;
    if (cudaMalloc(&t.d_tab, sizeof(double) * 2 * (size_t)total) != cudaSuccess)
      return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory for the grid geometry tables");
    double* tabs[3] = {t.d_tab + t.offset[0], t.d_tab + t.offset[1], t.d_tab + t.offset[2]};
    GDTB_TRY(launch_q1_axis_tables(ctx->launch, key, tabs, t.inv));
    ctx->axis_tables.push_back(t);
    found = &ctx->axis_tables.back();
  }
  for (int k = 0; k < 3; ++k)
    axis_tab[k] = found->d_tab + found->offset[k];
  axis_tab_inv = found->inv;
  return GDTB_OK;
}

int build_q1_params(gdtb_matop* op, gdtb_vecfun* fun, Q1GatherParams& p)
{
  std::memset(&p, 0, sizeof(p));
  const GridDev& g = op ? op->grid : fun->grid;
  p.g = g;
  const int d = g.d;
  p.div_vx = make_fast_div((unsigned)(g.n[0] + 1));
  p.div_vy = make_fast_div((unsigned)(d > 1 ? g.n[1] + 1 : 1));
  GDTB_TRY(q1_axis_tables(op ? op->ctx : fun->ctx, g, p));
  p.row_lo = op ? op->row_lo : fun->row_lo;
  p.row_hi = op ? op->row_hi : fun->row_hi;
  p.elem_lo = op ? op->elem_lo : fun->elem_lo;
  p.elem_hi = op ? op->elem_hi : fun->elem_hi;
  p.value_offset = q1_layer_rowptr(g, p.row_lo);
  p.row_offset = p.row_lo * q1_layer_rows(g);
  if (op) {
    for (const auto& lf : op->element_forms) {
      Q1Tables tab;
      q1_tables(form_quadrature_order(lf.form, 1, ROLE_ELEMENT), tab);
      for (int t = 0; t < lf.form.n_terms; ++t) {
        const gdtb_integrand& in = lf.form.terms[t];
        Q1Group& G = p.group[p.n_groups++];
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) {
            G.K1[a][b] = tab.G[1][1][a][b];
            G.M1[a][b] = tab.G[0][0][a][b];
          }
        const gdtb_function& f = in.diffusion;
        G.scale = lf.form.scaling;
        G.coef = f.data;
        if (in.kind == GDTB_INT_PRODUCT)
          G.kind = Q1G_MASS;
        else
          G.kind = (f.kind == GDTB_FN_CONST_TENSOR || f.kind == GDTB_FN_ELEM_TENSOR) ? Q1G_LAPLACE_TENSOR
                                                                                   : Q1G_LAPLACE_SCALAR;
        G.coef_elem = (f.kind == GDTB_FN_ELEM_SCALAR || f.kind == GDTB_FN_ELEM_TENSOR) ? 1 : 0;
        if (f.kind == GDTB_FN_CONST_SCALAR)
          G.scale *= f.c[0];
        if (f.kind == GDTB_FN_CONST_TENSOR)
          for (int r = 0; r < d; ++r)
            for (int c = 0; c < d; ++c)
              G.kappa[r * 3 + c] = f.c[r * d + c];
        q1_group_tensors(G.kind, d, tab, G.M);
      }
    }
  }
  return GDTB_OK;
}

} // namespace

// ==================================================================================================
// C ABI
// ==================================================================================================
namespace gdtb {
// sets *flag when x holds an inf or a nan (the reference's VectorType::valid())
__global__ void k_flag_non_finite(const double* __restrict__ x, long long n, int* __restrict__ flag)
{
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    bad = bad || !isfinite(x[i]);
  if (bad)
    *flag = 1;
}
} // namespace gdtb
using gdtb::k_flag_non_finite;

extern "C" {

const char* gdtb_last_error(void)
{
  return g_last_error.c_str();
}

int gdtb_version(void)
{
  return GDTB_VERSION;
}

int gdtb_ctx_create(int device, gdtb_ctx** out)
{
  if (!out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "ctx out pointer is NULL");
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0)
    return fail(GDTB_ERR_CUDA,
                std::string("no usable CUDA device (libgdtb has no CPU fallback): ") + cudaGetErrorString(err));
  if (device < 0 || device >= count)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "device index out of range");
  GDTB_CUDA(cudaSetDevice(device));
  auto ctx = new gdtb_ctx();
  ctx->device = device;
  ctx->d_error_flag = nullptr;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return fail(GDTB_ERR_CUDA, "cudaStreamCreate failed");
  }
  ctx->launch.stream = ctx->own_stream;
  ctx->launch.count = 0;
  ctx->launch.timing = &ctx->timing;
  ctx->error_flag_pending = false;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return fail(GDTB_ERR_CUDA, "cudaGetDeviceProperties failed");
  }
  ctx->launch.sm_count = prop.multiProcessorCount;
  if (cudaMalloc(&ctx->d_error_flag, sizeof(int)) != cudaSuccess
      || cudaMemset(ctx->d_error_flag, 0, sizeof(int)) != cudaSuccess) {
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return fail(GDTB_ERR_CUDA, "cudaMalloc failed");
  }
  *out = ctx;
  return GDTB_OK;
}

int gdtb_ctx_destroy(gdtb_ctx* ctx)
{
  if (!ctx)
    return GDTB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->launch.stream);
  cudaFree(ctx->d_error_flag);
  for (auto& t : ctx->axis_tables)
    cudaFree(t.d_tab);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return GDTB_OK;
}

int gdtb_ctx_set_stream(gdtb_ctx* ctx, void* cuda_stream)
{
  GDTB_TRY(check_ctx(ctx));
  ctx->launch.stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return GDTB_OK;
}

int gdtb_ctx_synchronize(gdtb_ctx* ctx)
{
  GDTB_TRY(check_ctx(ctx));
  if (ctx->error_flag_pending) {
    // scatter kernels flag local entries that are missing from the pattern (the reference's add_to_entry throws)
    int flag = 0;
    GDTB_CUDA(cudaMemcpyAsync(&flag, ctx->d_error_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->launch.stream));
    GDTB_CUDA(cudaStreamSynchronize(ctx->launch.stream));
    ctx->error_flag_pending = false;
    if (flag) {
      GDTB_CUDA(cudaMemset(ctx->d_error_flag, 0, sizeof(int)));
      return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH,
                  "add_to_entry: a local entry is not part of the sparsity pattern (wrong stencil for the appended "
                  "forms?)");
    }
    return GDTB_OK;
  }
  GDTB_CUDA(cudaStreamSynchronize(ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_ctx_enable_timing(gdtb_ctx* ctx, int enabled)
{
  GDTB_TRY(check_ctx(ctx));
  ctx->timing.enabled = enabled != 0;
  return GDTB_OK;
}

int gdtb_ctx_kernel_time(gdtb_ctx* ctx, const char* family, double* total_ms, int64_t* launches)
{
  GDTB_TRY(check_ctx(ctx));
  static const char* names[KF_COUNT] = {"q1_gather",      "q2_gather",      "dg_gather",       "fv_apply",
                                        "element_matrix", "element_vector", "coupling_matrix", "boundary_matrix"};
  int fam = -1;
  for (int i = 0; i < KF_COUNT; ++i)
    if (family && std::strcmp(family, names[i]) == 0)
      fam = i;
  if (fam < 0)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "unknown kernel family");
  GDTB_CUDA(cudaStreamSynchronize(ctx->launch.stream));
  auto& st = ctx->timing.start[fam];
  auto& en = ctx->timing.stop[fam];
  double total = 0.;
  const size_t n = std::min(st.size(), en.size());
  for (size_t i = 0; i < n; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, st[i], en[i]);
    total += ms;
  }
  for (auto e : st)
    cudaEventDestroy(e);
  for (auto e : en)
    cudaEventDestroy(e);
  st.clear();
  en.clear();
  if (total_ms)
    *total_ms = total;
  if (launches)
    *launches = (int64_t)n;
  return GDTB_OK;
}

static const char* const kFamilyNames[KF_COUNT] = {"q1_gather",      "q2_gather",      "dg_gather",       "fv_apply",
                                                   "element_matrix", "element_vector", "coupling_matrix", "boundary_matrix"};

const char* gdtb_ctx_kernel_name(const gdtb_ctx* ctx, const char* family)
{
  if (!ctx || !family)
    return "";
  for (int i = 0; i < KF_COUNT; ++i)
    if (std::strcmp(family, kFamilyNames[i]) == 0)
      return ctx->timing.last_kernel[i].c_str();
  return "";
}

int64_t gdtb_ctx_launch_count(const gdtb_ctx* ctx)
{
  return ctx ? ctx->launch.count : 0;
}

// ---- grid / spaces ---------------------------------------------------------------------------
// GridDev of a cube grid description (host arithmetic only)
static int make_grid_dev(const gdtb_grid_desc* desc, GridDev& d)
{
  if (desc->dim < 1 || desc->dim > 3)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "grid dimension must be 1, 2 or 3");
  std::memset(&d, 0, sizeof(d));
  d.d = desc->dim;
  d.periodic = desc->periodic & ((1 << desc->dim) - 1);
  d.ne = 1;
  for (int k = 0; k < 3; ++k) {
    if (k < desc->dim) {
      if (desc->n[k] < 1 || !(desc->upper[k] > desc->lower[k]))
        return fail(GDTB_ERR_INVALID_ARGUMENT, "grid needs n >= 1 and upper > lower in every direction");
      d.lo[k] = desc->lower[k];
      d.n[k] = desc->n[k];
      d.h[k] = (desc->upper[k] - desc->lower[k]) / double(desc->n[k]);
    } else {
      d.lo[k] = 0.;
      d.n[k] = 1;
      d.h[k] = 1.;
    }
    d.ne *= d.n[k];
  }
  d.layer_lo = 0;
  d.layer_hi = d.n[d.d - 1];
  return GDTB_OK;
}

int gdtb_grid_create_cube(gdtb_ctx* ctx, const gdtb_grid_desc* desc, gdtb_grid** out)
{
  if (!ctx || !desc || !out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_grid_create_cube: NULL argument");
  GridDev d;
  GDTB_TRY(make_grid_dev(desc, d));
  auto g = new gdtb_grid();
  g->ctx = ctx;
  g->desc = *desc;
  g->dev = d;
  *out = g;
  return GDTB_OK;
}

int gdtb_grid_destroy(gdtb_grid* grid)
{
  delete grid;
  return GDTB_OK;
}

int64_t gdtb_grid_num_elements(const gdtb_grid* grid)
{
  return grid ? grid->dev.ne : 0;
}

int gdtb_space_create(gdtb_ctx* ctx, const gdtb_grid* grid, int kind, int order, gdtb_space** out)
{
  if (!ctx || !grid || !out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_space_create: NULL argument");
  SpaceDev sp;
  GDTB_TRY(make_space_dev(grid->dev, kind, order, sp));
  auto s = new gdtb_space();
  s->ctx = ctx;
  s->grid = grid->dev;
  s->dev = sp;
  *out = s;
  return GDTB_OK;
}

int gdtb_fv_space_create(gdtb_ctx* ctx, const gdtb_grid* grid, int range_dim, gdtb_space** out)
{
  if (!ctx || !grid || !out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fv_space_create: NULL argument");
  if (range_dim < 1 || range_dim > 4)
    return fail(GDTB_ERR_SPACE, "finite volume spaces: 1 <= range_dim <= 4");
  SpaceDev sp;
  GDTB_TRY(make_space_dev(grid->dev, GDTB_SPACE_FV, 0, sp));
  // FiniteVolumeMapper<GV, m>: m consecutive DoFs per element (spaces/mapper/finite-volume.hh:92-108)
  sp.nloc = range_dim;
  sp.size = grid->dev.ne * range_dim;
  auto s = new gdtb_space();
  s->ctx = ctx;
  s->grid = grid->dev;
  s->dev = sp;
  *out = s;
  return GDTB_OK;
}

// Host-side view of the closed-form CSR geometry the gather kernels and the structured pattern generators use (no
// device needed): row pointers of the CG Q1 / CG Q2 element stencil or the DG element_and_intersection stencil on a
// non-periodic grid, and the row ranges a slab of element layers owns.  Lets a binder size its containers before any
// device work -- and lets the CPU test-suite check the layout arithmetic against the oracle's patterns.
int gdtb_host_closed_form_rowptr(const gdtb_grid_desc* grid, int kind, int order, int64_t* rowptr)
{
  if (!grid || !rowptr)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_host_closed_form_rowptr: NULL argument");
  GridDev g;
  SpaceDev sp;
  GDTB_TRY(make_grid_dev(grid, g));
  GDTB_TRY(make_space_dev(g, kind, order, sp));
  if (g.periodic && !(kind == GDTB_SPACE_DG && grid_dg_closed_form(g)))
    return fail(GDTB_ERR_NOT_IMPLEMENTED,
                "closed-form row pointers exist for non-periodic grids (DG: also periodic directions with >= 3 cells)");
  if (kind == GDTB_SPACE_CG && sp.K == 1)
    return q1_host_rowptr(g, sp, (long long*)rowptr);
  if (kind == GDTB_SPACE_CG && sp.K == 2 && (g.d == 2 || g.d == 3))
    return q2_host_rowptr(g, sp, (long long*)rowptr);
  if (kind == GDTB_SPACE_DG)
    return dg_host_rowptr(g, sp, (long long*)rowptr);
  return fail(GDTB_ERR_NOT_IMPLEMENTED, "closed-form row pointers: CG Q1, CG Q2 (2D / 3D) and DG spaces");
}

int gdtb_host_slab_row_ranges(const gdtb_grid_desc* grid, int kind, int order, int64_t layer_begin, int64_t layer_end,
                              int32_t max_ranges, int64_t* row_begin, int64_t* row_end, int64_t* value_offset,
                              int64_t* value_count, int32_t* n_ranges)
{
  if (!grid || !n_ranges)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_host_slab_row_ranges: NULL argument");
  GridDev g;
  SpaceDev sp;
  GDTB_TRY(make_grid_dev(grid, g));
  GDTB_TRY(make_space_dev(g, kind, order, sp));
  const long long n_last = g.n[g.d - 1];
  if (layer_begin < 0 || layer_end > n_last || layer_begin >= layer_end)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "slab must satisfy 0 <= begin < end <= n[last]");
  const bool dg = kind == GDTB_SPACE_DG;
  if (g.periodic || (!dg && (kind != GDTB_SPACE_CG || (sp.K != 1 && !(sp.K == 2 && (g.d == 2 || g.d == 3))))))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "slab row ranges: CG Q1, CG Q2 (2D / 3D) and DG on non-periodic grids");
  g.layer_lo = layer_begin;
  g.layer_hi = layer_end;
  Q2SlabRange ranges[8];
  int n = 1;
  if (dg) {
    const long long plane = g.ne / n_last;
    ranges[0].row_begin = layer_begin * plane * sp.nloc;
    ranges[0].row_end = layer_end * plane * sp.nloc;
    ranges[0].value_offset = dg_value_offset(g, sp, layer_begin * plane);
    ranges[0].count = dg_value_offset(g, sp, layer_end * plane) - ranges[0].value_offset;
  } else if (sp.K == 2)
    n = q2_slab_ranges(g, sp, ranges);
  else {
    long long row_lo, row_hi, elem_lo, elem_hi;
    q1_slab_ranges(g, layer_begin, layer_end, row_lo, row_hi, elem_lo, elem_hi);
    ranges[0].row_begin = row_lo * q1_layer_rows(g);
    ranges[0].row_end = row_hi * q1_layer_rows(g);
    ranges[0].value_offset = q1_layer_rowptr(g, row_lo);
    ranges[0].count = q1_layer_rowptr(g, row_hi) - ranges[0].value_offset;
  }
  *n_ranges = n;
  if (max_ranges < n)
    return (row_begin || row_end || value_offset || value_count) ? fail(GDTB_ERR_INVALID_ARGUMENT, "arrays too short") : GDTB_OK;
  for (int r = 0; r < n; ++r) {
    if (row_begin)
      row_begin[r] = ranges[r].row_begin;
    if (row_end)
      row_end[r] = ranges[r].row_end;
    if (value_offset)
      value_offset[r] = ranges[r].value_offset;
    if (value_count)
      value_count[r] = ranges[r].count;
  }
  return GDTB_OK;
}

int gdtb_space_destroy(gdtb_space* space)
{
  delete space;
  return GDTB_OK;
}

int64_t gdtb_space_size(const gdtb_space* space)
{
  return space ? space->dev.size : 0;
}

int32_t gdtb_space_max_local_size(const gdtb_space* space)
{
  return space ? space->dev.nloc : 0;
}

int gdtb_space_global_indices(const gdtb_space* space, int64_t element, int64_t* out)
{
  if (!space || !out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_space_global_indices: NULL argument");
  if (element < 0 || element >= space->grid.ne)
    return fail(GDTB_ERR_SPACE, "element index out of range"); // Exceptions::mapper_error
  long long idx[3];
  elem_coords(space->grid, element, idx);
  for (int i = 0; i < space->dev.nloc; ++i)
    out[i] = global_index(space->grid, space->dev, idx, i);
  return GDTB_OK;
}

// ---- quadrature ----------------------------------------------------------------------------------
int gdtb_form_quadrature_order(const gdtb_space* space, const gdtb_form* form, int role, int32_t* order)
{
  if (!space || !form || !order)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_form_quadrature_order: NULL argument");
  if (form->n_terms < 1 || form->n_terms > GDTB_MAX_TERMS)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "form: n_terms must be in [1, GDTB_MAX_TERMS]");
  FormRole r;
  switch (role) {
    case GDTB_ROLE_ELEMENT: r = ROLE_ELEMENT; break;
    case GDTB_ROLE_FUNCTIONAL: r = ROLE_RHS; break;
    case GDTB_ROLE_COUPLING: r = ROLE_COUPLING; break;
    case GDTB_ROLE_BOUNDARY: r = ROLE_BOUNDARY; break;
    default: return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_form_quadrature_order: unknown role");
  }
  *order = form_quadrature_order(*form, space->dev.K, r);
  return GDTB_OK;
}

int gdtb_gauss_rule(int order, int32_t* m, double* points01, double* weights)
{
  if (!m)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_gauss_rule: NULL argument");
  const int mm = gauss_points_for_order(order);
  *m = mm;
  if (mm > MAX_Q1D)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "quadrature order too high (more than 8 Gauss points per direction)");
  if (points01 && weights)
    gauss_legendre_01(mm, points01, weights);
  return GDTB_OK;
}

// ---- sparsity pattern --------------------------------------------------------------------------
int gdtb_pattern_create(gdtb_ctx* ctx, const gdtb_space* test, const gdtb_space* ansatz, int stencil, int method,
                        gdtb_pattern** out)
{
  if (!test || !ansatz || !out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_pattern_create: NULL argument");
  GDTB_TRY(check_ctx(ctx));
  if (stencil < GDTB_STENCIL_ELEMENT || stencil > GDTB_STENCIL_ELEMENT_AND_INTERSECTION)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "Unknown Stencil encountered"); // sparsity-pattern.hh:174-177
  if (std::memcmp(&test->grid, &ansatz->grid, sizeof(GridDev)) != 0)
    return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "test and ansatz space live on different grids");
  const bool same_space = std::memcmp(&test->dev, &ansatz->dev, sizeof(SpaceDev)) == 0;
  const bool structured_q1 =
      stencil == GDTB_STENCIL_ELEMENT && q1_space(test->dev) && q1_space(ansatz->dev) && !test->grid.periodic;
  const bool structured_q2 = stencil == GDTB_STENCIL_ELEMENT && same_space && test->dev.kind == GDTB_SPACE_CG
                             && test->dev.K == 2 && (test->grid.d == 2 || test->grid.d == 3) && !test->grid.periodic
                             && test->dev.size < (1LL << 31);
  const bool structured_dg = stencil == GDTB_STENCIL_ELEMENT_AND_INTERSECTION && same_space
                             && test->dev.kind == GDTB_SPACE_DG && grid_dg_closed_form(test->grid)
                             && test->dev.size < (1LL << 31);
  const bool structured_ok = structured_q1 || structured_q2 || structured_dg;
  if (method == GDTB_PATTERN_STRUCTURED && !structured_ok)
    return fail(GDTB_ERR_NOT_IMPLEMENTED,
                "structured pattern generators cover the CG Q1 / Q2 element stencils and the DG element_and_intersection "
                "stencil on non-periodic grids");
  auto p = std::make_unique<gdtb_pattern>();
  p->ctx = ctx;
  p->grid = test->grid;
  p->test = test->dev;
  p->ansatz = ansatz->dev;
  p->stencil = stencil;
  p->rows = test->dev.size;
  p->cols = ansatz->dev.size;
  p->d_rowptr = nullptr;
  p->d_colidx = nullptr;
  const bool use_structured = method == GDTB_PATTERN_STRUCTURED || (method == GDTB_PATTERN_AUTO && structured_ok);
  if (use_structured && structured_q1)
    GDTB_TRY(pattern_structured_cg_q1(ctx->launch, p->grid, p->test, &p->d_rowptr, &p->d_colidx, &p->nnz));
  else if (use_structured && structured_q2)
    GDTB_TRY(pattern_structured_cg_q2(ctx->launch, p->grid, p->test, &p->d_rowptr, &p->d_colidx, &p->nnz));
  else if (use_structured)
    GDTB_TRY(pattern_structured_dg(ctx->launch, p->grid, p->test, &p->d_rowptr, &p->d_colidx, &p->nnz));
  else
    GDTB_TRY(
        pattern_sort_unique(ctx->launch, p->grid, p->test, p->ansatz, stencil, &p->d_rowptr, &p->d_colidx, &p->nnz));
  *out = p.release();
  return GDTB_OK;
}

int gdtb_pattern_destroy(gdtb_pattern* p)
{
  if (!p)
    return GDTB_OK;
  cudaSetDevice(p->ctx->device);
  cudaFree(p->d_rowptr);
  cudaFree(p->d_colidx);
  delete p;
  return GDTB_OK;
}

int64_t gdtb_pattern_rows(const gdtb_pattern* p)
{
  return p ? p->rows : 0;
}
int64_t gdtb_pattern_cols(const gdtb_pattern* p)
{
  return p ? p->cols : 0;
}
int64_t gdtb_pattern_nnz(const gdtb_pattern* p)
{
  return p ? p->nnz : 0;
}

int gdtb_pattern_download(const gdtb_pattern* p, int64_t* rowptr, int32_t* colidx)
{
  if (!p)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "pattern is NULL");
  GDTB_TRY(check_ctx(p->ctx));
  if (rowptr)
    GDTB_CUDA(cudaMemcpy(rowptr, p->d_rowptr, sizeof(int64_t) * (size_t)(p->rows + 1), cudaMemcpyDeviceToHost));
  if (colidx)
    GDTB_CUDA(cudaMemcpy(colidx, p->d_colidx, sizeof(int32_t) * (size_t)p->nnz, cudaMemcpyDeviceToHost));
  return GDTB_OK;
}

int gdtb_pattern_device(const gdtb_pattern* p, const int64_t** d_rowptr, const int32_t** d_colidx)
{
  if (!p)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "pattern is NULL");
  if (d_rowptr)
    *d_rowptr = (const int64_t*)p->d_rowptr;
  if (d_colidx)
    *d_colidx = p->d_colidx;
  return GDTB_OK;
}

// ---- MatrixOperator ----------------------------------------------------------------------------
int gdtb_matop_create(gdtb_ctx* ctx, const gdtb_space* test, const gdtb_space* ansatz, const gdtb_pattern* pattern,
                      gdtb_matop** out)
{
  if (!test || !ansatz || !out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_create: NULL argument");
  GDTB_TRY(check_ctx(ctx));
  const auto q2_space = [](const SpaceDev& sp) { return sp.kind == GDTB_SPACE_CG && sp.K == 2 && (sp.d == 2 || sp.d == 3); };
  const bool closed_q1 = q1_space(test->dev) && q1_space(ansatz->dev) && !test->grid.periodic;
  const bool closed_q2 = q2_space(test->dev) && q2_space(ansatz->dev) && !test->grid.periodic;
  // discontinuous spaces: the element_and_intersection stencil (what an IPDG operator needs) has a closed form too
  const bool closed_dg = test->dev.kind == GDTB_SPACE_DG && ansatz->dev.kind == GDTB_SPACE_DG
                         && grid_dg_closed_form(test->grid) && test->dev.size < (1LL << 31);
  if (!pattern && !closed_q1 && !closed_q2 && !closed_dg)
    return fail(GDTB_ERR_INVALID_ARGUMENT,
                "gdtb_matop_create: a pattern is required (only the CG Q1 / Q2 element stencils and the DG "
                "element_and_intersection stencil on non-periodic grids have closed forms)");
  // matrix-based.hh:73-80: matrix.rows() == range_space.mapper().size(), cols == source_space.mapper().size()
  if (pattern && (pattern->rows != test->dev.size || pattern->cols != ansatz->dev.size))
    return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "pattern shape does not match the spaces (rows = test, cols = ansatz)");
  if (std::memcmp(&test->grid, &ansatz->grid, sizeof(GridDev)) != 0)
    return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "test and ansatz space live on different grids");
  if (std::memcmp(&test->dev, &ansatz->dev, sizeof(SpaceDev)) != 0)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "rectangular operators (test space != ansatz space) are not supported yet");
  auto op = std::make_unique<gdtb_matop>();
  op->ctx = ctx;
  op->grid = test->grid;
  op->test = test->dev;
  op->ansatz = ansatz->dev;
  op->pattern = pattern;
  op->d_values = nullptr;
  op->owns_values = true;
  op->slab = false;
  op->row_begin = 0;
  op->row_end = test->dev.size;
  op->value_offset = 0;
  if (pattern)
    op->nnz_local = pattern->nnz;
  else if (closed_q1) {
    op->nnz_local = 1;
    for (int k = 0; k < op->grid.d; ++k)
      op->nnz_local *= 3 * op->grid.n[k] + 1;
  } else if (closed_dg)
    op->nnz_local = dg_value_offset(op->grid, op->test, op->grid.ne);
  else { // CG Q2 element stencil: prod_k (8 n_k + 1) lattice couplings
    Q2SlabRange ranges[8];
    const int nr = q2_slab_ranges(op->grid, op->test, ranges);
    op->nnz_local = 0;
    for (int r = 0; r < nr; ++r)
      op->nnz_local += ranges[r].count;
  }
  op->n_ranges = 0;
  op->row_lo = 0;
  op->row_hi = op->grid.n[op->grid.d - 1] + 1;
  op->elem_lo = 0;
  op->elem_hi = op->grid.n[op->grid.d - 1];
  if (pattern || true) {
    if (cudaMalloc(&op->d_values, sizeof(double) * (size_t)std::max<long long>(op->nnz_local, 1)) != cudaSuccess)
      return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory for the matrix values");
    GDTB_CUDA(cudaMemsetAsync(op->d_values, 0, sizeof(double) * (size_t)op->nnz_local, ctx->launch.stream));
  }
  *out = op.release();
  return GDTB_OK;
}

static void halo_p2p_release(gdtb_matop* op);

int gdtb_matop_clear_forms(gdtb_matop* op)
{
  if (!op)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "operator is NULL");
  cudaSetDevice(op->ctx->device);
  cudaStreamSynchronize(op->ctx->launch.stream);
  for (auto* v : {&op->element_forms, &op->coupling_forms, &op->boundary_forms}) {
    for (auto& f : *v)
      free_form(f);
    v->clear();
  }
  return GDTB_OK;
}

int gdtb_matop_destroy(gdtb_matop* op)
{
  if (!op)
    return GDTB_OK;
  gdtb_matop_clear_forms(op);
  halo_p2p_release(op);
  if (op->owns_values)
    cudaFree(op->d_values);
  cudaFree(op->d_forms);
  cudaFree(op->d_q2_tab);
  cudaFree(op->d_q2_items);
  cudaFree(op->d_q1_items);
  cudaFree(op->d_qp_scratch);
  cudaFree(op->d_own_rowptr);
  cudaFree(op->d_own_colidx);
  delete op;
  return GDTB_OK;
}

int gdtb_matop_append_element(gdtb_matop* op, const gdtb_form* form)
{
  if (!op)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "operator is NULL");
  GDTB_TRY(check_ctx(op->ctx));
  if (form)
    for (int t = 0; t < form->n_terms && t < GDTB_MAX_TERMS; ++t)
      if (form->terms[t].kind != GDTB_INT_LAPLACE && form->terms[t].kind != GDTB_INT_PRODUCT)
        return fail(GDTB_ERR_INTEGRAND, "element bilinear forms take Laplace / product integrands");
  if (form && form->n_terms >= 1 && form->n_terms <= GDTB_MAX_TERMS)
    GDTB_TRY(check_qp_functions(*form, op->test.K, ROLE_ELEMENT, op->grid.d));
  LoweredForm lf;
  GDTB_TRY(lower_form(op->ctx, op->grid, form, 0, lf));
  op->element_forms.push_back(std::move(lf));
  return GDTB_OK;
}

int gdtb_matop_append_coupling(gdtb_matop* op, const gdtb_form* form, int filter)
{
  if (!op)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "operator is NULL");
  GDTB_TRY(check_ctx(op->ctx));
  if (filter != GDTB_FILTER_INNER_ONCE && filter != GDTB_FILTER_INNER_AND_PERIODIC_ONCE)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "coupling forms need an inner-intersections-once filter");
  if (op->test.kind == GDTB_SPACE_CG)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "intersection forms are implemented for DG / FV spaces only");
  if (form)
    for (int t = 0; t < form->n_terms && t < GDTB_MAX_TERMS; ++t)
      if (form->terms[t].kind != GDTB_INT_IPDG_INNER_COUPLING && form->terms[t].kind != GDTB_INT_IPDG_INNER_PENALTY)
        return fail(GDTB_ERR_INTEGRAND, "This integrand cannot be used on an inner intersection!");
  if (form && form->n_terms >= 1 && form->n_terms <= GDTB_MAX_TERMS)
    GDTB_TRY(check_qp_functions(*form, op->test.K, ROLE_COUPLING, op->grid.d));
  LoweredForm lf;
  GDTB_TRY(lower_form(op->ctx, op->grid, form, filter, lf));
  op->coupling_forms.push_back(std::move(lf));
  return GDTB_OK;
}

int gdtb_matop_append_boundary(gdtb_matop* op, const gdtb_form* form, int filter)
{
  if (!op)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "operator is NULL");
  GDTB_TRY(check_ctx(op->ctx));
  if (filter != GDTB_FILTER_ALL_BOUNDARY)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "boundary forms need a boundary filter");
  if (op->test.kind == GDTB_SPACE_CG)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "intersection forms are implemented for DG / FV spaces only");
  if (form)
    for (int t = 0; t < form->n_terms && t < GDTB_MAX_TERMS; ++t)
      if (form->terms[t].kind != GDTB_INT_IPDG_DIRICHLET_COUPLING
          && form->terms[t].kind != GDTB_INT_IPDG_BOUNDARY_PENALTY)
        return fail(GDTB_ERR_INTEGRAND, "This integrand cannot be used on a boundary intersection!"); // ipdg.hh:93-94
  if (form && form->n_terms >= 1 && form->n_terms <= GDTB_MAX_TERMS)
    GDTB_TRY(check_qp_functions(*form, op->test.K, ROLE_BOUNDARY, op->grid.d));
  LoweredForm lf;
  GDTB_TRY(lower_form(op->ctx, op->grid, form, filter, lf));
  op->boundary_forms.push_back(std::move(lf));
  return GDTB_OK;
}

int gdtb_matop_num_forms(const gdtb_matop* op)
{
  return op ? int(op->element_forms.size() + op->coupling_forms.size() + op->boundary_forms.size()) : 0;
}

const char* gdtb_matop_plan(gdtb_matop* op)
{
  if (!op)
    return "";
  if (matop_q1_eligible(op))
    op->plan = "q1_gather";
  else if (matop_q2_eligible(op))
    op->plan = "q2_gather";
  else if (matop_cg_qp_eligible(op))
    op->plan = op->test.K == 1 ? "q1_gather_qp" : "q2_gather_qp";
  else if (matop_dg_eligible(op))
    op->plan = "dg_gather";
  else
    op->plan = "generic_coloured";
  return op->plan.c_str();
}

const char* gdtb_matop_plan_reason(gdtb_matop* op)
{
  if (!op)
    return "";
  if (std::string(gdtb_matop_plan(op)) != "generic_coloured")
    return "";
  const SpaceDev& sp = op->test;
  const GridDev& g = op->grid;
  const bool cg = sp.kind == GDTB_SPACE_CG, dg = sp.kind == GDTB_SPACE_DG;
  const size_t n_forms = op->element_forms.size() + op->coupling_forms.size() + op->boundary_forms.size();
  if (n_forms == 0)
    return "no forms appended";
  if (std::memcmp(&op->test, &op->ansatz, sizeof(SpaceDev)) != 0)
    return "test and ansatz space differ";
  if (sp.kind == GDTB_SPACE_FV)
    return "finite volume spaces have no row-gather assembly kernel";
  if (cg && sp.K >= 3)
    return "continuous Lagrange order >= 3: only the quadrature-faithful kernels cover it";
  if (cg && sp.K == 2 && g.d == 1)
    return "the CG Q2 row-gather kernels are 2D / 3D";
  if (cg && (!op->coupling_forms.empty() || !op->boundary_forms.empty()))
    return "intersection forms on a continuous space";
  if (cg && op->pattern && op->pattern->stencil != GDTB_STENCIL_ELEMENT)
    return "the CG row-gather kernels follow the element stencil";
  if (dg && !dg_gather_supported(g.d, sp.K))
    return "DG row gather: order 1 (1D - 3D) or order 2 (1D / 2D)";
  if (dg && op->pattern && op->pattern->stencil != GDTB_STENCIL_ELEMENT_AND_INTERSECTION)
    return "the DG row-gather kernels follow the element_and_intersection stencil";
  if (dg && g.periodic && !grid_dg_closed_form(g))
    return "periodic direction with fewer than 3 cells";
  if (dg && g.periodic && (sp.K != 1 || !matop_dg_scalar_coefficients(op)))
    return "periodic grid view: only the factorised DG kernels (order 1, constant / element-wise scalar coefficients) know "
           "the wrap neighbours";
  if (n_forms > (size_t)DGG_MAX_FORMS)
    return "too many forms for one gather pass";
  return "this combination of integrands / coefficient kinds has no row-gather kernel";
}

int gdtb_matop_set_zero(gdtb_matop* op)
{
  if (!op)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "operator is NULL");
  GDTB_TRY(check_ctx(op->ctx));
  GDTB_CUDA(cudaMemsetAsync(op->d_values, 0, sizeof(double) * (size_t)op->nnz_local, op->ctx->launch.stream));
  GDTB_CUDA(cudaStreamSynchronize(op->ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_matop_values_download(const gdtb_matop* op, double* values)
{
  if (!op || !values)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_values_download: NULL argument");
  GDTB_TRY(check_ctx(op->ctx));
  GDTB_CUDA(cudaMemcpyAsync(values, op->d_values, sizeof(double) * (size_t)op->nnz_local, cudaMemcpyDeviceToHost,
                            op->ctx->launch.stream));
  GDTB_CUDA(cudaStreamSynchronize(op->ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_matop_values_upload(gdtb_matop* op, const double* values)
{
  if (!op || !values)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_values_upload: NULL argument");
  GDTB_TRY(check_ctx(op->ctx));
  GDTB_CUDA(cudaMemcpyAsync(op->d_values, values, sizeof(double) * (size_t)op->nnz_local, cudaMemcpyHostToDevice,
                            op->ctx->launch.stream));
  GDTB_CUDA(cudaStreamSynchronize(op->ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_matop_values_device(const gdtb_matop* op, double** d_values)
{
  if (!op || !d_values)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_values_device: NULL argument");
  *d_values = op->d_values;
  return GDTB_OK;
}

int gdtb_matop_set_values_device(gdtb_matop* op, double* d_values)
{
  if (!op || !d_values)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_set_values_device: NULL argument");
  if (op->owns_values)
    cudaFree(op->d_values);
  op->d_values = d_values;
  op->owns_values = false;
  return GDTB_OK;
}

// interface-row halo partition: rows of the vertex layers [begin, end] (the top one is the interface layer owned by
// the slab above) from the slab's OWN element layers [begin, end) only
static void q1_halo_ranges(const GridDev& g, long long begin, long long end, long long& row_lo, long long& row_hi,
                           long long& elem_lo, long long& elem_hi)
{
  row_lo = begin;
  row_hi = end + 1;
  elem_lo = begin;
  elem_hi = end;
  (void)g;
}

static int matop_set_slab_impl(gdtb_matop* op, int64_t layer_begin, int64_t layer_end, bool halo);

int gdtb_matop_set_slab(gdtb_matop* op, int64_t layer_begin, int64_t layer_end)
{
  return matop_set_slab_impl(op, layer_begin, layer_end, false);
}

int gdtb_matop_set_slab_halo(gdtb_matop* op, int64_t layer_begin, int64_t layer_end)
{
  if (op && !(q1_space(op->test) && q1_space(op->ansatz)))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "the interface-row halo partition is implemented for CG Q1 spaces");
  return matop_set_slab_impl(op, layer_begin, layer_end, true);
}

int gdtb_matop_halo_layout(const gdtb_matop* op, int64_t* recv_offset, int64_t* send_offset, int64_t* count)
{
  if (!op || !op->slab || !op->halo)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_halo_layout: the operator is not in interface-row halo mode");
  // one vertex layer of rows: the first local layer receives the partial sums of the slab below, the last local layer
  // (the interface owned by the slab above) is sent up; interface layers are interior along the last direction, so
  // both have the same CSR shape
  const long long layer = q1_layer_rowptr(op->grid, op->row_lo + 1) - q1_layer_rowptr(op->grid, op->row_lo);
  const long long n_last = op->grid.n[op->grid.d - 1];
  const bool has_lower = op->grid.layer_lo > 0, has_upper = op->grid.layer_hi < n_last;
  const long long interior = has_lower ? layer : (has_upper ? q1_layer_rowptr(op->grid, op->row_hi) - q1_layer_rowptr(op->grid, op->row_hi - 1) : 0);
  if (recv_offset)
    *recv_offset = has_lower ? 0 : -1;
  if (send_offset)
    *send_offset = has_upper ? op->nnz_local - (q1_layer_rowptr(op->grid, op->row_hi) - q1_layer_rowptr(op->grid, op->row_hi - 1)) : -1;
  if (count)
    *count = interior;
  return GDTB_OK;
}

int gdtb_vecfun_set_slab_halo(gdtb_vecfun* fun, int64_t layer_begin, int64_t layer_end);

// ---- peer-memory interface-row halo ---------------------------------------------------------------------------
static void halo_p2p_release(gdtb_matop* op)
{
  if (op->halo_opened_upper) {
    cudaIpcCloseMemHandle(op->halo_peer_recv);
    cudaIpcCloseMemHandle(op->halo_peer_flags);
  }
  if (op->halo_opened_lower)
    cudaIpcCloseMemHandle(op->halo_lower_flags);
  cudaFree(op->halo_recv);
  cudaFree(op->halo_flags);
  op->halo_recv = op->halo_peer_recv = nullptr;
  op->halo_flags = op->halo_peer_flags = op->halo_lower_flags = nullptr;
  op->halo_opened_lower = op->halo_opened_upper = op->halo_connected = false;
  op->halo_step = 0;
}

static void halo_layer_sizes(const gdtb_matop* op, long long& layer_rows, long long& layer_values)
{
  layer_rows = q1_layer_rows(op->grid);
  // interface layers are interior along the last direction: every one has the same CSR shape
  layer_values = q1_layer_rowptr(op->grid, 2) - q1_layer_rowptr(op->grid, 1);
}

int gdtb_halo_p2p_alloc(gdtb_matop* op, void* handles)
{
  if (!op || !handles)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_halo_p2p_alloc: NULL argument");
  GDTB_TRY(check_ctx(op->ctx));
  if (!op->slab || !op->halo)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_halo_p2p_alloc: call gdtb_matop_set_slab_halo first");
  if (op->grid.n[op->grid.d - 1] < 2)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_halo_p2p_alloc: the grid has no interior vertex layer");
  halo_p2p_release(op);
  long long layer_rows, layer_values;
  halo_layer_sizes(op, layer_rows, layer_values);
  const size_t bytes = sizeof(double) * 2 * (size_t)(layer_values + layer_rows);
  if (cudaMalloc(&op->halo_recv, bytes) != cudaSuccess || cudaMalloc(&op->halo_flags, 64 * sizeof(int)) != cudaSuccess) {
    halo_p2p_release(op);
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (peer-memory halo buffers)");
  }
  GDTB_CUDA(cudaMemset(op->halo_recv, 0, bytes));
  GDTB_CUDA(cudaMemset(op->halo_flags, 0, 64 * sizeof(int)));
  cudaIpcMemHandle_t h[2];
  if (cudaIpcGetMemHandle(&h[0], op->halo_recv) != cudaSuccess || cudaIpcGetMemHandle(&h[1], op->halo_flags) != cudaSuccess) {
    const std::string why = cudaGetErrorString(cudaGetLastError());
    halo_p2p_release(op);
    return fail(GDTB_ERR_CUDA, "cudaIpcGetMemHandle failed: " + why);
  }
  std::memcpy(handles, h, sizeof(h));
  return GDTB_OK;
}

int gdtb_halo_p2p_connect(gdtb_matop* op, const void* lower_handles, int64_t lower_layers, const void* upper_handles)
{
  if (!op || !op->halo_recv)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_halo_p2p_connect: call gdtb_halo_p2p_alloc first");
  GDTB_TRY(check_ctx(op->ctx));
  const long long n_last = op->grid.n[op->grid.d - 1];
  const bool has_lower = op->grid.layer_lo > 0, has_upper = op->grid.layer_hi < n_last;
  if ((has_lower && (!lower_handles || lower_layers < 1)) || (has_upper && !upper_handles))
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_halo_p2p_connect: a neighbour's handles are missing");
  if (has_upper) {
    cudaIpcMemHandle_t h[2];
    std::memcpy(h, upper_handles, sizeof(h));
    void* ptr[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i)
      if (cudaIpcOpenMemHandle(&ptr[i], h[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        return fail(GDTB_ERR_CUDA, std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError()));
    op->halo_peer_recv = static_cast<double*>(ptr[0]);
    op->halo_peer_flags = static_cast<int*>(ptr[1]);
    op->halo_opened_upper = true;
  }
  if (has_lower) {
    cudaIpcMemHandle_t h[2];
    std::memcpy(h, lower_handles, sizeof(h));
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, h[1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
      return fail(GDTB_ERR_CUDA, std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError()));
    op->halo_lower_flags = static_cast<int*>(ptr);
    op->halo_opened_lower = true;
    op->halo_lower_layers = lower_layers;
  }
  op->halo_connected = true;
  op->halo_step = 0;
  return GDTB_OK;
}

int gdtb_halo_p2p_check(gdtb_matop* op)
{
  if (!op || !op->halo_flags)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_halo_p2p_check: no peer-memory halo");
  GDTB_TRY(check_ctx(op->ctx));
  GDTB_CUDA(cudaStreamSynchronize(op->ctx->launch.stream));
  int flags[4] = {0, 0, 0, 0};
  GDTB_CUDA(cudaMemcpy(flags, op->halo_flags, sizeof(flags), cudaMemcpyDeviceToHost));
  if (flags[2])
    return fail(GDTB_ERR_OPERATOR, "peer-memory halo: a wait for the neighbour's counter timed out");
  return GDTB_OK;
}

int gdtb_vector_add(gdtb_ctx* ctx, double* d_y, const double* d_x, int64_t n)
{
  if (!d_y || !d_x || n < 0)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_vector_add: invalid argument");
  GDTB_TRY(check_ctx(ctx));
  RkAxpyParams q;
  q.n = n;
  q.nv = 1;
  q.v[0] = d_x;
  q.c[0] = 1.;
  return launch_rk_axpy(ctx->launch, q, d_y, d_y);
}

int gdtb_vector_upload(gdtb_ctx* ctx, double* d_dst, const double* src, int64_t n)
{
  if (!d_dst || !src || n < 0)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_vector_upload: invalid argument");
  GDTB_TRY(check_ctx(ctx));
  GDTB_CUDA(cudaMemcpyAsync(d_dst, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->launch.stream));
  GDTB_CUDA(cudaStreamSynchronize(ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_vector_download(gdtb_ctx* ctx, double* dst, const double* d_src, int64_t n)
{
  if (!dst || !d_src || n < 0)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_vector_download: invalid argument");
  GDTB_TRY(check_ctx(ctx));
  GDTB_CUDA(cudaMemcpyAsync(dst, d_src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->launch.stream));
  GDTB_CUDA(cudaStreamSynchronize(ctx->launch.stream));
  return GDTB_OK;
}

static int matop_set_slab_impl(gdtb_matop* op, int64_t layer_begin, int64_t layer_end, bool halo)
{
  if (!op)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "operator is NULL");
  GDTB_TRY(check_ctx(op->ctx));
  const long long n_last = op->grid.n[op->grid.d - 1];
  if (layer_begin < 0 || layer_end > n_last || layer_begin >= layer_end)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "slab must satisfy 0 <= begin < end <= n[last]");
  const bool q2 = op->test.kind == GDTB_SPACE_CG && op->test.K == 2 && (op->grid.d == 2 || op->grid.d == 3)
                  && !op->grid.periodic && std::memcmp(&op->test, &op->ansatz, sizeof(SpaceDev)) == 0;
  const bool dg = op->test.kind == GDTB_SPACE_DG && !op->grid.periodic
                  && std::memcmp(&op->test, &op->ansatz, sizeof(SpaceDev)) == 0 && !halo;
  if (!(q1_space(op->test) && q1_space(op->ansatz)) && !q2 && !dg)
    return fail(GDTB_ERR_NOT_IMPLEMENTED,
                "slab-partitioned assembly is implemented for CG Q1 / Q2 spaces and DG spaces on non-periodic grids");
  if (dg && op->pattern
      && (op->pattern->stencil != GDTB_STENCIL_ELEMENT_AND_INTERSECTION || op->pattern->test.kind != GDTB_SPACE_DG))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "DG slabs follow the element_and_intersection stencil");
  if (!op->owns_values)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_set_slab must be called before lending a value buffer");
  op->grid.layer_lo = layer_begin;
  op->grid.layer_hi = layer_end;
  op->slab = true;
  if (dg) {
    // DG rows are element-owned: the slab's rows are those of its own elements, contiguous in the global numbering;
    // the neighbours across the slab faces enter through their index, geometry and coefficients only (no exchange)
    const long long plane = op->grid.ne / n_last;
    op->elem_lo = layer_begin;
    op->elem_hi = layer_end;
    op->row_begin = layer_begin * plane * op->test.nloc;
    op->row_end = layer_end * plane * op->test.nloc;
    op->value_offset = dg_value_offset(op->grid, op->test, layer_begin * plane);
    op->nnz_local = dg_value_offset(op->grid, op->test, layer_end * plane) - op->value_offset;
    op->n_ranges = 0;
  } else if (q2) {
    // the MCMG numbering groups the rows by sub-entity kind: a slab owns one contiguous row range per group
    Q2SlabRange ranges[8];
    op->n_ranges = q2_slab_ranges(op->grid, op->test, ranges);
    op->nnz_local = 0;
    for (int r = 0; r < op->n_ranges; ++r) {
      op->range_row_begin[r] = ranges[r].row_begin;
      op->range_row_end[r] = ranges[r].row_end;
      op->range_value_offset[r] = ranges[r].value_offset;
      op->range_count[r] = ranges[r].count;
      op->nnz_local += ranges[r].count;
    }
    op->row_begin = ranges[0].row_begin;
    op->row_end = ranges[0].row_end;
    op->value_offset = ranges[0].value_offset;
  } else {
    if (halo)
      q1_halo_ranges(op->grid, layer_begin, layer_end, op->row_lo, op->row_hi, op->elem_lo, op->elem_hi);
    else
      q1_slab_ranges(op->grid, layer_begin, layer_end, op->row_lo, op->row_hi, op->elem_lo, op->elem_hi);
    op->halo = halo;
    op->row_begin = op->row_lo * q1_layer_rows(op->grid);
    op->row_end = op->row_hi * q1_layer_rows(op->grid);
    op->value_offset = q1_layer_rowptr(op->grid, op->row_lo);
    op->nnz_local = q1_layer_rowptr(op->grid, op->row_hi) - op->value_offset;
    op->n_ranges = 0;
  }
  cudaFree(op->d_values);
  op->d_values = nullptr;
  if (cudaMalloc(&op->d_values, sizeof(double) * (size_t)op->nnz_local) != cudaSuccess)
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory for the matrix values of the slab");
  GDTB_CUDA(cudaMemsetAsync(op->d_values, 0, sizeof(double) * (size_t)op->nnz_local, op->ctx->launch.stream));
  return GDTB_OK;
}

int64_t gdtb_matop_local_nnz(const gdtb_matop* op)
{
  return op ? op->nnz_local : 0;
}

int gdtb_matop_local_rows(const gdtb_matop* op, int64_t* row_begin, int64_t* row_end, int64_t* value_offset)
{
  if (!op)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "operator is NULL");
  if (row_begin)
    *row_begin = op->row_begin;
  if (row_end)
    *row_end = op->row_end;
  if (value_offset)
    *value_offset = op->value_offset;
  return GDTB_OK;
}

int gdtb_matop_local_row_ranges(const gdtb_matop* op, int32_t max_ranges, int64_t* row_begin, int64_t* row_end,
                                int64_t* value_offset, int64_t* value_count, int32_t* n_ranges)
{
  if (!op || !n_ranges)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_local_row_ranges: NULL argument");
  const int n = op->n_ranges > 0 ? op->n_ranges : 1;
  *n_ranges = n;
  if (max_ranges < n)
    return (row_begin || row_end || value_offset || value_count) ? fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_matop_local_row_ranges: arrays too short") : GDTB_OK;
  for (int r = 0; r < n; ++r) {
    if (row_begin)
      row_begin[r] = op->n_ranges > 0 ? op->range_row_begin[r] : op->row_begin;
    if (row_end)
      row_end[r] = op->n_ranges > 0 ? op->range_row_end[r] : op->row_end;
    if (value_offset)
      value_offset[r] = op->n_ranges > 0 ? op->range_value_offset[r] : op->value_offset;
    if (value_count)
      value_count[r] = op->n_ranges > 0 ? op->range_count[r] : op->nnz_local;
  }
  return GDTB_OK;
}

// ---- VectorBasedFunctional ---------------------------------------------------------------------
int gdtb_vecfun_create(gdtb_ctx* ctx, const gdtb_space* space, gdtb_vecfun** out)
{
  if (!space || !out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_vecfun_create: NULL argument");
  GDTB_TRY(check_ctx(ctx));
  auto f = std::make_unique<gdtb_vecfun>();
  f->ctx = ctx;
  f->grid = space->grid;
  f->space = space->dev;
  f->d_vec = nullptr;
  f->owns_vec = true;
  f->d_sep_tab = nullptr;
  f->d_rule = nullptr;
  f->slab = false;
  f->rule_uploaded = false;
  f->row_begin = 0;
  f->row_end = space->dev.size;
  f->local_size = space->dev.size;
  f->row_lo = 0;
  f->row_hi = space->grid.n[space->grid.d - 1] + 1;
  f->elem_lo = 0;
  f->elem_hi = space->grid.n[space->grid.d - 1];
  if (cudaMalloc(&f->d_vec, sizeof(double) * (size_t)std::max<long long>(space->dev.size, 1)) != cudaSuccess)
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory for the vector");
  GDTB_CUDA(cudaMemsetAsync(f->d_vec, 0, sizeof(double) * (size_t)space->dev.size, ctx->launch.stream));
  *out = f.release();
  return GDTB_OK;
}

int gdtb_vecfun_clear_forms(gdtb_vecfun* fun)
{
  if (!fun)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "functional is NULL");
  cudaSetDevice(fun->ctx->device);
  cudaStreamSynchronize(fun->ctx->launch.stream);
  for (auto& f : fun->forms)
    free_form(f);
  fun->forms.clear();
  return GDTB_OK;
}

int gdtb_vecfun_destroy(gdtb_vecfun* fun)
{
  if (!fun)
    return GDTB_OK;
  gdtb_vecfun_clear_forms(fun);
  if (fun->owns_vec)
    cudaFree(fun->d_vec);
  cudaFree(fun->d_sep_tab);
  cudaFree(fun->d_rule);
  delete fun;
  return GDTB_OK;
}

int gdtb_vecfun_append_element(gdtb_vecfun* fun, const gdtb_form* form)
{
  if (!fun)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "functional is NULL");
  GDTB_TRY(check_ctx(fun->ctx));
  if (form && (form->n_terms != 1 || form->terms[0].kind != GDTB_INT_PRODUCT))
    return fail(GDTB_ERR_INTEGRAND, "element functionals take LocalProductIntegrand(w).with_ansatz(f) (one term)");
  if (form)
    GDTB_TRY(check_qp_functions(*form, fun->space.K, ROLE_RHS, fun->grid.d));
  LoweredForm lf;
  GDTB_TRY(lower_form(fun->ctx, fun->grid, form, 0, lf));
  fun->forms.push_back(std::move(lf));
  return GDTB_OK;
}

int gdtb_vecfun_set_zero(gdtb_vecfun* fun)
{
  if (!fun)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "functional is NULL");
  GDTB_TRY(check_ctx(fun->ctx));
  GDTB_CUDA(cudaMemsetAsync(fun->d_vec, 0, sizeof(double) * (size_t)fun->local_size, fun->ctx->launch.stream));
  GDTB_CUDA(cudaStreamSynchronize(fun->ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_vecfun_download(const gdtb_vecfun* fun, double* vector)
{
  if (!fun || !vector)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_vecfun_download: NULL argument");
  GDTB_TRY(check_ctx(fun->ctx));
  GDTB_CUDA(cudaMemcpyAsync(vector, fun->d_vec, sizeof(double) * (size_t)fun->local_size, cudaMemcpyDeviceToHost,
                            fun->ctx->launch.stream));
  GDTB_CUDA(cudaStreamSynchronize(fun->ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_vecfun_device(const gdtb_vecfun* fun, double** d_vector)
{
  if (!fun || !d_vector)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_vecfun_device: NULL argument");
  *d_vector = fun->d_vec;
  return GDTB_OK;
}

int gdtb_vecfun_set_device(gdtb_vecfun* fun, double* d_vector)
{
  if (!fun || !d_vector)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_vecfun_set_device: NULL argument");
  if (fun->owns_vec)
    cudaFree(fun->d_vec);
  fun->d_vec = d_vector;
  fun->owns_vec = false;
  return GDTB_OK;
}

static int vecfun_set_slab_impl(gdtb_vecfun* fun, int64_t layer_begin, int64_t layer_end, bool halo);

int gdtb_vecfun_set_slab(gdtb_vecfun* fun, int64_t layer_begin, int64_t layer_end)
{
  return vecfun_set_slab_impl(fun, layer_begin, layer_end, false);
}

int gdtb_vecfun_set_slab_halo(gdtb_vecfun* fun, int64_t layer_begin, int64_t layer_end)
{
  return vecfun_set_slab_impl(fun, layer_begin, layer_end, true);
}

int gdtb_vecfun_halo_layout(const gdtb_vecfun* fun, int64_t* recv_offset, int64_t* send_offset, int64_t* count)
{
  if (!fun || !fun->slab || !fun->halo)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_vecfun_halo_layout: the functional is not in interface-row halo mode");
  const long long layer = q1_layer_rows(fun->grid);
  const long long n_last = fun->grid.n[fun->grid.d - 1];
  const bool has_lower = fun->grid.layer_lo > 0, has_upper = fun->grid.layer_hi < n_last;
  if (recv_offset)
    *recv_offset = has_lower ? 0 : -1;
  if (send_offset)
    *send_offset = has_upper ? (fun->row_end - fun->row_begin) - layer : -1;
  if (count)
    *count = (has_lower || has_upper) ? layer : 0;
  return GDTB_OK;
}

static int vecfun_set_slab_impl(gdtb_vecfun* fun, int64_t layer_begin, int64_t layer_end, bool halo)
{
  if (!fun)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "functional is NULL");
  GDTB_TRY(check_ctx(fun->ctx));
  const long long n_last = fun->grid.n[fun->grid.d - 1];
  if (layer_begin < 0 || layer_end > n_last || layer_begin >= layer_end)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "slab must satisfy 0 <= begin < end <= n[last]");
  const SpaceDev& sp = fun->space;
  const bool q1 = q1_space(sp);
  const bool q2 = sp.kind == GDTB_SPACE_CG && sp.K == 2 && (fun->grid.d == 2 || fun->grid.d == 3);
  const bool element_owned = sp.kind != GDTB_SPACE_CG;
  if (halo && !q1)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "the interface-row halo partition is implemented for CG Q1 spaces");
  if (!q1 && !q2 && !element_owned)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "slab-partitioned functionals: CG Q1 / Q2, DG and FV spaces");
  if (!fun->owns_vec)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_vecfun_set_slab must be called before lending a vector buffer");
  fun->grid.layer_lo = layer_begin;
  fun->grid.layer_hi = layer_end;
  fun->slab = true;
  fun->halo = halo;
  fun->n_ranges = 0;
  if (q1) {
    if (halo)
      q1_halo_ranges(fun->grid, layer_begin, layer_end, fun->row_lo, fun->row_hi, fun->elem_lo, fun->elem_hi);
    else
      q1_slab_ranges(fun->grid, layer_begin, layer_end, fun->row_lo, fun->row_hi, fun->elem_lo, fun->elem_hi);
    fun->row_begin = fun->row_lo * q1_layer_rows(fun->grid);
    fun->row_end = fun->row_hi * q1_layer_rows(fun->grid);
    fun->local_size = fun->row_end - fun->row_begin;
  } else if (q2) {
    // one owned row range per sub-entity group of the MCMG numbering, back to back in the local vector
    Q2SlabRange ranges[8];
    fun->n_ranges = q2_slab_ranges(fun->grid, sp, ranges);
    fun->local_size = 0;
    for (int r = 0; r < fun->n_ranges; ++r) {
      fun->range_row_begin[r] = ranges[r].row_begin;
      fun->range_row_end[r] = ranges[r].row_end;
      fun->range_local[r] = fun->local_size;
      fun->local_size += ranges[r].row_end - ranges[r].row_begin;
    }
    fun->row_begin = ranges[0].row_begin;
    fun->row_end = ranges[0].row_end;
    fun->elem_lo = std::max<long long>(layer_begin - 1, 0);
    fun->elem_hi = layer_end;
  } else {
    // DG / FV: the rows of the slab's own elements
    const long long plane = fun->grid.ne / n_last;
    fun->elem_lo = layer_begin;
    fun->elem_hi = layer_end;
    fun->row_begin = layer_begin * plane * sp.nloc;
    fun->row_end = layer_end * plane * sp.nloc;
    fun->local_size = fun->row_end - fun->row_begin;
  }
  cudaFree(fun->d_vec);
  fun->d_vec = nullptr;
  const size_t bytes = sizeof(double) * (size_t)fun->local_size;
  if (cudaMalloc(&fun->d_vec, std::max<size_t>(bytes, 8)) != cudaSuccess)
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory for the vector of the slab");
  GDTB_CUDA(cudaMemsetAsync(fun->d_vec, 0, bytes, fun->ctx->launch.stream));
  return GDTB_OK;
}

// ---- the grid walk -----------------------------------------------------------------------------
static int q1_rhs_params(gdtb_vecfun* fun, Q1GatherParams& p)
{
  const GridDev& g = fun->grid;
  const int d = g.d;
  const int n = 1 << d;
  p.has_rhs = 1;
  for (const auto& lf : fun->forms) {
    const gdtb_integrand& in = lf.form.terms[0];
    Q1Tables tab;
    q1_tables(form_quadrature_order(lf.form, 1, ROLE_RHS), tab);
    const double w = in.diffusion.c[0];
    // S[o] = prod_k sum_q w_q phihat_{a_k}(x_q), a_k = 1 - o_k: the reference-element part of
    // l_i = sum_q (w psi_i f) |det J| w_q (conversion.hh:109-116, local/functionals/integrals.hh:95-96) for constant f
    double S[8];
    for (int o = 0; o < n; ++o) {
      S[o] = 1.;
      for (int k = 0; k < d; ++k)
        S[o] *= tab.s1[1 - ((o >> k) & 1)];
    }
    if (in.weight.kind == GDTB_FN_CONST_SCALAR) {
      p.rhs_has_const = 1;
      for (int o = 0; o < n; ++o)
        p.rhs_S_const[o] += w * in.weight.c[0] * S[o];
    } else if (in.weight.kind == GDTB_FN_ELEM_SCALAR) {
      p.rhs_has_elem = 1;
      for (int o = 0; o < n; ++o)
        p.rhs_S_elem[o] = w * S[o];
      p.rhs_elem = in.weight.data;
    } else {
      // separable built-in: 1D tables on the device (they carry the cells' extents, i.e. the integration element)
      const long long stride = std::max(std::max(g.n[0], g.n[1]), g.n[2]) + 1;
      gdtb_ctx* ctx = fun->ctx;
      if (!fun->d_sep_tab)
        if (cudaMalloc(&fun->d_sep_tab, sizeof(double) * 3 * (size_t)stride) != cudaSuccess)
          return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory for the right-hand-side tables");
      if (!fun->d_rule)
        if (cudaMalloc(&fun->d_rule, sizeof(double) * 4 * MAX_Q1D) != cudaSuccess)
          return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory for the quadrature rule");
      double host[4 * MAX_Q1D];
      for (int q = 0; q < MAX_Q1D; ++q) {
        host[q] = tab.qx[q];
        host[MAX_Q1D + q] = tab.qw[q];
        host[2 * MAX_Q1D + 2 * q] = tab.phi[q][0];
        host[2 * MAX_Q1D + 2 * q + 1] = tab.phi[q][1];
      }
      if (!fun->rule_uploaded || std::memcmp(host, fun->h_rule, sizeof(host)) != 0) {
        std::memcpy(fun->h_rule, host, sizeof(host));
        GDTB_CUDA(cudaMemcpyAsync(fun->d_rule, fun->h_rule, sizeof(host), cudaMemcpyHostToDevice, ctx->launch.stream));
        GDTB_CUDA(cudaStreamSynchronize(ctx->launch.stream));
        fun->rule_uploaded = true;
        fun->sep_valid = false;
      }
      // the tables depend on (source, rule, slab) only: rebuilt when one of them changed since the last walk
      const FnDev src = to_dev(in.weight);
      if (!fun->sep_valid || std::memcmp(&src, &fun->sep_fn, sizeof(FnDev)) != 0 || fun->sep_m != tab.m
          || fun->sep_lo != fun->elem_lo || fun->sep_hi != fun->elem_hi) {
        GDTB_TRY(launch_q1_rhs_tables(ctx->launch, g, fun->elem_lo, fun->elem_hi, src, tab.m, fun->d_rule,
                                      fun->d_rule + MAX_Q1D, fun->d_rule + 2 * MAX_Q1D, fun->d_sep_tab, stride));
        fun->sep_fn = src;
        fun->sep_m = tab.m;
        fun->sep_lo = fun->elem_lo;
        fun->sep_hi = fun->elem_hi;
        fun->sep_valid = true;
      }
      p.rhs_has_sep = 1;
      p.rhs_sep_scale = w * (in.weight.builtin == GDTB_BUILTIN_COS_PRODUCT ? in.weight.p[0] : 1.);
      p.rhs_sep_tab = fun->d_sep_tab;
      p.rhs_sep_stride = stride;
    }
  }
  return GDTB_OK;
}

// CG Q1 / Q2 element forms with coefficients that vary inside a cell: one gather launch per integrand (the first one
// overwrites, the others accumulate), every coefficient given / sampled per quadrature point of its form's rule
static int assemble_cg_qp(gdtb_matop* op, bool accumulate)
{
  gdtb_ctx* ctx = op->ctx;
  Launch& L = ctx->launch;
  const GridDev& g = op->grid;
  const int d = g.d, K = op->test.K;
  const long long n_last = g.n[d - 1];
  const long long plane = g.ne / n_last; // elements per layer of the last direction
  // element layers the owned rows need (the slab plus the ghost layer below it)
  const long long lay_lo = K == 1 ? op->elem_lo : std::max<long long>(g.layer_lo - 1, 0);
  const long long lay_hi = K == 1 ? op->elem_hi : g.layer_hi;
  const long long e_begin = lay_lo * plane, e_end = lay_hi * plane;
  bool first = !accumulate;
  for (const auto& lf : op->element_forms) {
    const int m = gauss_points_for_order(form_quadrature_order(lf.form, K, ROLE_ELEMENT));
    const int nq = ipow(m, d);
    CgQpGroup G;
    std::memset(&G, 0, sizeof(G));
    double qx[MAX_Q1D];
    qp_point_tables(K, m, qx, G.pt);
    G.m = m;
    for (int t = 0; t < lf.form.n_terms; ++t) {
      const gdtb_integrand& in = lf.form.terms[t];
      const gdtb_function& f = in.diffusion;
      const bool tensor = fn_is_tensor(f);
      G.kind = in.kind == GDTB_INT_PRODUCT ? Q1G_MASS : (tensor ? Q1G_LAPLACE_TENSOR : Q1G_LAPLACE_SCALAR);
      G.scale = lf.form.scaling;
      const int comps = tensor ? d * d : 1;
      if (f.kind == GDTB_FN_QP_SCALAR || f.kind == GDTB_FN_QP_TENSOR)
        G.coef = f.data; // already one value per quadrature point, indexed by the global element index
      else {
        const size_t bytes = sizeof(double) * (size_t)(e_end - e_begin) * nq * comps;
        if (op->d_qp_scratch_bytes < bytes) {
          cudaFree(op->d_qp_scratch);
          op->d_qp_scratch = nullptr;
          op->d_qp_scratch_bytes = 0;
          if (cudaMalloc(&op->d_qp_scratch, bytes) != cudaSuccess)
            return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (coefficient samples)");
          op->d_qp_scratch_bytes = bytes;
        }
        GDTB_TRY(launch_sample_function(L, g, to_dev(f, &lf), m, qx, tensor ? 1 : 0, e_begin, e_end, op->d_qp_scratch));
        G.coef = op->d_qp_scratch - e_begin * (long long)(nq * comps);
      }
      if (K == 1) {
        Q1QpParams p;
        std::memset(&p, 0, sizeof(p));
        p.g = g;
        p.div_vx = make_fast_div((unsigned)(g.n[0] + 1));
        p.div_vy = make_fast_div((unsigned)(d > 1 ? g.n[1] + 1 : 1));
        GDTB_TRY(q1_axis_tables(ctx, g, p.axis_tab, p.axis_tab_inv));
        p.group = G;
        p.row_lo = op->row_lo;
        p.row_hi = op->row_hi;
        p.elem_lo = op->elem_lo;
        p.elem_hi = op->elem_hi;
        p.value_offset = q1_layer_rowptr(g, p.row_lo);
        p.row_offset = p.row_lo * q1_layer_rows(g);
        GDTB_TRY(launch_q1_gather_qp(L, p, op->d_values, !first));
      } else if (q2_qp_xfused_supported(d, m, G.kind)) {
        const bool sampled = !(f.kind == GDTB_FN_QP_SCALAR || f.kind == GDTB_FN_QP_TENSOR);
        GDTB_TRY(launch_q2_qp_xfused(L, g, G, op->test, sampled ? e_begin : 0, sampled ? e_end : g.ne, op->d_values, !first));
      } else {
        Q2GatherParams p;
        std::memset(&p, 0, sizeof(p));
        p.g = g;
        GDTB_TRY(launch_q2_gather_qp(L, p, G, op->test, op->d_values, !first));
      }
      first = false;
    }
  }
  return GDTB_OK;
}

static int assemble_impl(gdtb_matop* op, gdtb_vecfun* fun, int mode, bool synchronize)
{
  if (!op && !fun)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_assemble: nothing to assemble");
  if (mode != GDTB_ASSEMBLE_OVERWRITE && mode != GDTB_ASSEMBLE_ACCUMULATE)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_assemble: unknown mode");
  gdtb_ctx* ctx = op ? op->ctx : fun->ctx;
  GDTB_TRY(check_ctx(ctx));
  if (op && fun && std::memcmp(&op->grid, &fun->grid, sizeof(GridDev)) != 0)
    return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "operator and functional live on different grids / slabs");
  Launch& L = ctx->launch;
  const bool accumulate = mode == GDTB_ASSEMBLE_ACCUMULATE;
  const bool op_fast = op && matop_q1_eligible(op);
  const bool fun_fast = fun && vecfun_q1_eligible(fun);
  const bool op_q2 = op && !op_fast && matop_q2_eligible(op);
  const bool op_qp = op && !op_fast && !op_q2 && matop_cg_qp_eligible(op);
  const bool op_dg = op && !op_fast && !op_q2 && !op_qp && matop_dg_eligible(op);
  if (op && !op_fast && !op_q2 && !op_qp && !op_dg && (op->slab || !op->pattern))
    return fail(GDTB_ERR_NOT_IMPLEMENTED,
                "slab-partitioned / pattern-free operators only support forms the row-gather kernels cover (CG Q1 / Q2 "
                "element forms, DG forms with the element_and_intersection stencil)");
  if (fun && !fun_fast && fun->slab && fun->halo)
    return fail(GDTB_ERR_NOT_IMPLEMENTED,
                "the interface-row halo partition only supports sources the CG Q1 row-gather kernel covers");

  // --- CG-Q1 row-gather path: matrix and right-hand side in ONE pass over the vertices ---------
  if (op_fast || fun_fast) {
    Q1GatherParams p;
    GDTB_TRY(build_q1_params(op_fast ? op : nullptr, fun_fast ? fun : nullptr, p));
    if (fun_fast)
      GDTB_TRY(q1_rhs_params(fun, p));
    const bool halo_p2p = op_fast && op->halo && op->halo_connected;
    if (halo_p2p) {
      if (fun && !fun_fast)
        return fail(GDTB_ERR_NOT_IMPLEMENTED, "peer-memory halo: the functional must take the CG Q1 gather path");
      const long long n_last = op->grid.n[op->grid.d - 1];
      long long layer_rows, layer_values;
      halo_layer_sizes(op, layer_rows, layer_values);
      Q1HaloP2p& H = p.halo;
      p.halo_p2p = 1;
      p.halo_top_value_start = op->nnz_local - layer_values;
      H.has_lower = op->grid.layer_lo > 0;
      H.has_upper = op->grid.layer_hi < n_last;
      H.layer_rows = layer_rows;
      H.layer_values = layer_values;
      const long long s = op->halo_step, par = s & 1, stride = layer_values + layer_rows;
      H.recv_values = op->halo_recv + par * stride;
      H.recv_rhs = H.recv_values + layer_values;
      H.my_flags = op->halo_flags;
      if (H.has_upper) {
        H.peer_values = op->halo_peer_recv + par * stride;
        H.peer_rhs = H.peer_values + layer_values;
        H.peer_flags = op->halo_peer_flags;
        // the rank above acknowledges every bottom item of every step; the buffer of this parity was last used at
        // step s - 2, i.e. steps 0 .. s - 2 must have been consumed
        const long long n_bottom = q1_halo_items(layer_rows, 1, false);
        H.expect_ack = (int)(std::max<long long>(s - 1, 0) * n_bottom);
      }
      if (H.has_lower) {
        H.lower_flags = op->halo_lower_flags;
        H.expect_data = (int)((s + 1) * q1_halo_items(layer_rows, op->halo_lower_layers, true));
      }
      op->halo_step++;
    }
    // one kappa per element (3D, one Laplace integrand): the kernel variant with prefetched coefficients takes its
    // per-item bookkeeping from records computed once per grid / slab and kept with the operator
    bool q1_items = false;
    if (op_fast && !accumulate && p.n_groups == 1 && p.group[0].coef_elem && p.group[0].kind == Q1G_LAPLACE_SCALAR
        && !p.halo_p2p) {
      const long long cap = q1_pref_item_capacity(p.g, p.row_lo, p.row_hi);
      if (cap > 0) {
        const size_t bytes = (size_t)cap * 32;
        if (op->d_q1_items_bytes < bytes) {
          cudaFree(op->d_q1_items);
          op->d_q1_items = nullptr;
          op->d_q1_items_bytes = 0;
          op->q1_items_key[0] = -1;
          if (cudaMalloc(&op->d_q1_items, bytes) != cudaSuccess)
            return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (Q1 work-item records)");
          op->d_q1_items_bytes = bytes;
        }
        p.items = op->d_q1_items;
        p.items_ready = op->q1_items_key[0] == p.row_lo && op->q1_items_key[1] == p.row_hi
                        && op->q1_items_key[2] == p.value_offset;
        q1_items = true;
      }
    }
    GDTB_TRY(launch_q1_gather(L, p, op_fast ? op->d_values : nullptr, fun_fast ? fun->d_vec : nullptr, accumulate));
    if (q1_items && p.items_ready) {
      op->q1_items_key[0] = p.row_lo;
      op->q1_items_key[1] = p.row_hi;
      op->q1_items_key[2] = p.value_offset;
    } else if (q1_items)
      op->q1_items_key[0] = -1;
  }

  // --- CG-Q2 row-gather path ---------------------------------------------------------------
  if (op_q2) {
    Q2GatherParams p;
    GDTB_TRY(build_q2_params(op, p));
    bool all_const = true;
    for (int gi = 0; gi < p.n_groups; ++gi)
      all_const = all_const && !p.group[gi].coef_elem;
    if (all_const && !std::getenv("GDTB_Q2_NO_SF")) {
      const size_t bytes = sizeof(double) * (size_t)q2_sf_table_doubles(op->grid) * (size_t)p.n_groups;
      if (op->d_q2_tab_bytes < bytes) {
        cudaFree(op->d_q2_tab);
  cudaFree(op->d_q2_items);
  cudaFree(op->d_q1_items);
        op->d_q2_tab = nullptr;
        op->d_q2_tab_bytes = 0;
        if (cudaMalloc(&op->d_q2_tab, bytes) != cudaSuccess)
          return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (Q2 sum-factorisation tables)");
        op->d_q2_tab_bytes = bytes;
      }
      p.sf = 1;
      p.sf_tab = op->d_q2_tab;
    } else if (p.n_groups == 1 && p.group[0].coef_elem
               && (p.group[0].kind == Q1G_MASS || p.group[0].kind == Q1G_LAPLACE_SCALAR) && !std::getenv("GDTB_Q2_NO_SF")
               && !std::getenv("GDTB_Q2_NO_PE")) {
      // one integrand with one coefficient value per element: per-element 1D factor tables (q2_row_plane_pe)
      const size_t bytes = sizeof(double) * (size_t)q2_pe_table_doubles(op->grid);
      if (op->d_q2_tab_bytes < bytes) {
        cudaFree(op->d_q2_tab);
        op->d_q2_tab = nullptr;
        op->d_q2_tab_bytes = 0;
        if (cudaMalloc(&op->d_q2_tab, bytes) != cudaSuccess)
          return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (Q2 per-element factor tables)");
        op->d_q2_tab_bytes = bytes;
      }
      p.sf = 2;
      p.sf_tab = op->d_q2_tab;
    }
    // work-item records of the gather kernel: computed on the device once per grid / slab, kept with the operator
    const long long n_items = accumulate ? 0 : q2_item_count(op->grid, op->test);
    if (n_items > 0) {
      const size_t bytes = (size_t)n_items * 32;
      if (op->d_q2_items_bytes < bytes) {
        cudaFree(op->d_q2_items);
  cudaFree(op->d_q1_items);
        op->d_q2_items = nullptr;
        op->d_q2_items_bytes = 0;
        op->q2_items_key[2] = -1;
        if (cudaMalloc(&op->d_q2_items, bytes) != cudaSuccess)
          return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (Q2 work-item records)");
        op->d_q2_items_bytes = bytes;
      }
      p.items = op->d_q2_items;
      p.items_ready = op->q2_items_key[0] == op->grid.layer_lo && op->q2_items_key[1] == op->grid.layer_hi
                      && op->q2_items_key[2] == n_items;
    }
    GDTB_TRY(launch_q2_gather(L, p, op->test, op->d_values, accumulate));
    if (n_items > 0 && p.items_ready) {
      op->q2_items_key[0] = op->grid.layer_lo;
      op->q2_items_key[1] = op->grid.layer_hi;
      op->q2_items_key[2] = n_items;
    }
  }

  // --- CG row gather with coefficients per quadrature point ----------------------------------
  if (op_qp)
    GDTB_TRY(assemble_cg_qp(op, accumulate));

  // --- DG row-gather path ------------------------------------------------------------------
  if (op_dg) {
    std::vector<FormDev> forms;
    for (const auto& lf : op->element_forms) {
      forms.emplace_back();
      GDTB_TRY(make_form_dev(lf.form, op->test.K, ROLE_ELEMENT, forms.back(), &lf));
    }
    for (const auto& lf : op->coupling_forms) {
      forms.emplace_back();
      GDTB_TRY(make_form_dev(lf.form, op->test.K, ROLE_COUPLING, forms.back(), &lf));
    }
    for (const auto& lf : op->boundary_forms) {
      forms.emplace_back();
      GDTB_TRY(make_form_dev(lf.form, op->test.K, ROLE_BOUNDARY, forms.back(), &lf));
    }
    const size_t bytes = sizeof(FormDev) * forms.size();
    if (op->d_forms_bytes < bytes) {
      cudaFree(op->d_forms);
      op->d_forms = nullptr;
      if (cudaMalloc(&op->d_forms, bytes) != cudaSuccess)
        return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory for the lowered forms");
      op->d_forms_bytes = bytes;
    }
    // upload only when the lowered forms changed since the last assembly (the host vector dies at the end of this
    // scope: synchronous copy)
    if (op->h_forms_cache.size() != bytes || std::memcmp(op->h_forms_cache.data(), forms.data(), bytes) != 0) {
      GDTB_CUDA(cudaMemcpyAsync(op->d_forms, forms.data(), bytes, cudaMemcpyHostToDevice, L.stream));
      GDTB_CUDA(cudaStreamSynchronize(L.stream));
      op->h_forms_cache.assign(reinterpret_cast<const char*>(forms.data()), reinterpret_cast<const char*>(forms.data()) + bytes);
    }
    DgGatherParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = op->grid;
    p.sp = op->test;
    p.forms = static_cast<const FormDev*>(op->d_forms);
    p.n_elem = (int)op->element_forms.size();
    p.n_coup = (int)op->coupling_forms.size();
    p.n_bnd = (int)op->boundary_forms.size();
    for (size_t f = 0; f < op->coupling_forms.size(); ++f)
      if (op->coupling_forms[f].filter == GDTB_FILTER_INNER_AND_PERIODIC_ONCE)
        p.coup_on_periodic |= 1u << f;
    const long long plane = op->grid.ne / op->grid.n[op->grid.d - 1];
    p.e_begin = op->slab ? op->elem_lo * plane : 0;
    p.e_end = op->slab ? op->elem_hi * plane : op->grid.ne;
    p.value_offset = op->slab ? op->value_offset : 0;
    // factorised kernel: order 1 and every coefficient a constant or element-wise scalar
    const auto scalar = [](const FnDev& f) { return f.kind == GDTB_FN_CONST_SCALAR || f.kind == GDTB_FN_ELEM_SCALAR; };
    bool fast = dg_gather_fast_supported(op->grid, op->test.K) && !std::getenv("GDTB_DG_NO_FAST");
    for (const FormDev& f : forms)
      for (int t = 0; t < f.n_terms; ++t)
        fast = fast && scalar(f.terms[t].diffusion) && scalar(f.terms[t].weight);
    // ... and the cheaper variant when every coefficient is a constant (face matrices tabulated once per launch)
    bool all_const = true;
    for (const FormDev& f : forms)
      for (int t = 0; t < f.n_terms; ++t)
        all_const = all_const && f.terms[t].diffusion.kind == GDTB_FN_CONST_SCALAR && f.terms[t].weight.kind == GDTB_FN_CONST_SCALAR;
    p.fast = fast ? ((all_const && !std::getenv("GDTB_DG_NO_CC")) ? 2 : 1) : 0;
    p.swip = forms.size() == 3 && p.n_elem == 1 && p.n_coup == 1 && p.n_bnd == 1 && forms[0].n_terms == 1
             && forms[0].terms[0].kind == GDTB_INT_LAPLACE && forms[1].n_terms == 2
             && forms[1].terms[0].kind == GDTB_INT_IPDG_INNER_COUPLING && forms[1].terms[1].kind == GDTB_INT_IPDG_INNER_PENALTY
             && forms[2].n_terms == 2 && forms[2].terms[0].kind == GDTB_INT_IPDG_DIRICHLET_COUPLING
             && forms[2].terms[1].kind == GDTB_INT_IPDG_BOUNDARY_PENALTY;
    if (p.swip && p.fast == 1) {
      const auto sw_fn = [](const FnDev& f) {
        DgGatherParams::SwFn r;
        r.data = f.kind == GDTB_FN_ELEM_SCALAR ? f.data : nullptr;
        r.c = f.c[0];
        return r;
      };
      p.sw.elem_kappa = sw_fn(forms[0].terms[0].diffusion);
      p.sw.coup_kappa = sw_fn(forms[1].terms[0].diffusion);
      p.sw.coup_weight = sw_fn(forms[1].terms[0].weight);
      p.sw.pen_weight = sw_fn(forms[1].terms[1].weight);
      p.sw.bnd_kappa = sw_fn(forms[2].terms[0].diffusion);
      p.sw.bnd_weight = sw_fn(forms[2].terms[1].weight);
      p.sw.coup_prefactor = forms[1].terms[0].prefactor;
      p.sw.pen_prefactor = forms[1].terms[1].prefactor;
      p.sw.bnd_prefactor = forms[2].terms[0].prefactor;
      p.sw.bndpen_prefactor = forms[2].terms[1].prefactor;
      p.sw.s_elem = forms[0].scaling;
      p.sw.s_coup = forms[1].scaling;
      p.sw.s_bnd = forms[2].scaling;
      p.sw.pen_hI = forms[1].terms[1].hI_kind;
      p.sw.bndpen_hI = forms[2].terms[1].hI_kind;
      const auto same = [](const DgGatherParams::SwFn& a, const DgGatherParams::SwFn& b) {
        return a.data == b.data && (a.data || a.c == b.c);
      };
      p.sw.coup_same = same(p.sw.coup_kappa, p.sw.coup_weight) && same(p.sw.coup_kappa, p.sw.pen_weight) ? 1 : 0;
    }
    if (p.fast)
      GDTB_TRY(q1_axis_tables(op->ctx, op->grid, p.axis_tab, p.axis_tab_inv));
    if (!p.fast) { // the quadrature-faithful kernel reads the row pointer (materialised for a pattern-free operator)
      const long long* rp = nullptr;
      const int* ci = nullptr;
      GDTB_TRY(internal_matop_pattern(op, &rp, &ci));
      p.rowptr = rp;
    }
    GDTB_TRY(launch_dg_gather(L, p, op->d_values, accumulate));
  }

  // --- generic path --------------------------------------------------------------------------
  if (op && !op_fast && !op_q2 && !op_qp && !op_dg) {
    static const bool warn_generic = std::getenv("GDTB_WARN_GENERIC") != nullptr;
    if (warn_generic && !op->warned_generic) {
      op->warned_generic = true;
      std::fprintf(stderr, "gdtb: operator %p assembles through the generic coloured-scatter kernels: %s\n", (void*)op,
                   gdtb_matop_plan_reason(op));
    }
    const gdtb_pattern* pat = op->pattern;
    if (!accumulate)
      GDTB_CUDA(cudaMemsetAsync(op->d_values, 0, sizeof(double) * (size_t)pat->nnz, L.stream));
    for (const auto& lf : op->element_forms) {
      FormDev fd;
      GDTB_TRY(make_form_dev(lf.form, op->test.K, ROLE_ELEMENT, fd, &lf));
      GDTB_TRY(launch_element_matrix(L, op->grid, op->test, fd, pat->d_rowptr, pat->d_colidx, op->d_values,
                                     ctx->d_error_flag));
    }
    for (const auto& lf : op->coupling_forms) {
      FormDev fd;
      GDTB_TRY(make_form_dev(lf.form, op->test.K, ROLE_COUPLING, fd, &lf));
      GDTB_TRY(launch_coupling_matrix(L, op->grid, op->test, fd, lf.filter, pat->d_rowptr, pat->d_colidx,
                                      op->d_values, ctx->d_error_flag));
    }
    for (const auto& lf : op->boundary_forms) {
      FormDev fd;
      GDTB_TRY(make_form_dev(lf.form, op->test.K, ROLE_BOUNDARY, fd, &lf));
      GDTB_TRY(launch_boundary_matrix(L, op->grid, op->test, fd, pat->d_rowptr, pat->d_colidx, op->d_values,
                                      ctx->d_error_flag));
    }
    ctx->error_flag_pending = true;
  }
  if (fun && !fun_fast) {
    if (!accumulate)
      GDTB_CUDA(cudaMemsetAsync(fun->d_vec, 0, sizeof(double) * (size_t)fun->local_size, L.stream));
    // slabs: walk the element layers the owned rows need (continuous spaces: plus the ghost layer below) and keep the
    // contributions to the owned rows only -- complete rows without communication, like the matrix
    RowMap rows;
    std::memset(&rows, 0, sizeof(rows));
    GridDev walk = fun->grid;
    if (fun->slab) {
      walk.layer_lo = fun->elem_lo;
      walk.layer_hi = fun->elem_hi;
      rows.n = fun->n_ranges > 0 ? fun->n_ranges : 1;
      for (int r = 0; r < rows.n; ++r) {
        rows.begin[r] = fun->n_ranges > 0 ? fun->range_row_begin[r] : fun->row_begin;
        rows.end[r] = fun->n_ranges > 0 ? fun->range_row_end[r] : fun->row_end;
        rows.local[r] = fun->n_ranges > 0 ? fun->range_local[r] : 0;
      }
    }
    for (const auto& lf : fun->forms) {
      FormDev fd;
      GDTB_TRY(make_form_dev(lf.form, fun->space.K, ROLE_RHS, fd, &lf));
      GDTB_TRY(launch_element_vector(L, walk, fun->space, fd, fun->d_vec, rows));
    }
  }
  if (synchronize)
    return gdtb_ctx_synchronize(ctx);
  return GDTB_OK;
}

int gdtb_assemble(gdtb_matop* op, gdtb_vecfun* fun, int mode)
{
  return assemble_impl(op, fun, mode, true);
}

int gdtb_assemble_async(gdtb_matop* op, gdtb_vecfun* fun, int mode)
{
  return assemble_impl(op, fun, mode, false);
}

int gdtb_assemble_host(gdtb_matop* op, gdtb_vecfun* fun, double* values, double* vector)
{
  GDTB_TRY(gdtb_assemble(op, fun, GDTB_ASSEMBLE_OVERWRITE));
  if (op && values)
    GDTB_TRY(gdtb_matop_values_download(op, values));
  if (fun && vector)
    GDTB_TRY(gdtb_vecfun_download(fun, vector));
  return GDTB_OK;
}

// ---- AdvectionFvOperator -----------------------------------------------------------------------
// operators on a finite volume space with m > 1 components (fv_system.cu)
static bool fv_is_system(const gdtb_fvop* L)
{
  return L->space.nloc > 1;
}

static void fvsys_fill_params(const gdtb_fvop* L, FvSysParams& p)
{
  p.g = L->grid;
  p.m = L->space.nloc;
  p.numflux = L->flux.numflux;
  p.gamma = L->flux.p[0];
  p.half_over_lambda = L->flux.numflux == GDTB_NUMFLUX_LAX_FRIEDRICHS ? 0.5 / L->flux.p[1] : 0.;
  for (int k = 0; k < 3; ++k)
    p.inv_ext[k] = L->d_ext + L->inv_ext_shift + L->ext_offset[k];
  p.euler = 0;
  p.dt = 0.;
  p.wall_mask = L->bnd_nf_mask;
  p.mirror_mask = L->bnd_ext_mask;
}

#define GDTB_NO_SYSTEMS(L, what)                                                                                            \
  if (fv_is_system(L))                                                                                                      \
  return fail(GDTB_ERR_NOT_IMPLEMENTED, what ": not available for systems (m > 1) yet")

int gdtb_fvop_create(gdtb_ctx* ctx, const gdtb_space* space, const gdtb_flux* flux, gdtb_fvop** out)
{
  if (!space || !flux || !out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_create: NULL argument");
  GDTB_TRY(check_ctx(ctx));
  if (space->dev.kind != GDTB_SPACE_FV)
    return fail(GDTB_ERR_OPERATOR, "Use LocalAdvectionDgCouplingOperator instead!"); // local/operators/advection-fv.hh:131-134
  if (flux->kind == GDTB_FLUX_EULER) {
    // systems (m = d + 2): tools/euler.hh covers d = 1, 2
    if (space->grid.d > 2)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "Euler equations: d = 1, 2 (tools/euler.hh:310, 348)");
    if (space->dev.nloc != space->grid.d + 2)
      return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "the Euler flux needs make_finite_volume_space<d + 2>");
    if (flux->numflux == GDTB_NUMFLUX_UPWIND)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "NumericalUpwindFlux is only available for m = 1 (upwind.hh:44)");
    if (flux->numflux != GDTB_NUMFLUX_VIJAYASUNDARAM && flux->numflux != GDTB_NUMFLUX_LAX_FRIEDRICHS)
      return fail(GDTB_ERR_INVALID_ARGUMENT, "unknown numerical flux");
    if (flux->numflux == GDTB_NUMFLUX_LAX_FRIEDRICHS && !(flux->p[1] > 0.))
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "Not yet implemented for m > 1 if lambda is not provided!"); // lax-friedrichs.hh:40-41
    if (!(flux->p[0] > 1.))
      return fail(GDTB_ERR_INVALID_ARGUMENT, "Euler equations: gamma > 1 expected in p[0]");
  } else {
    if (flux->kind != GDTB_FLUX_LINEAR && flux->kind != GDTB_FLUX_BURGERS)
      return fail(GDTB_ERR_INVALID_ARGUMENT, "unknown flux function");
    if (flux->numflux != GDTB_NUMFLUX_UPWIND && flux->numflux != GDTB_NUMFLUX_LAX_FRIEDRICHS)
      return fail(GDTB_ERR_INVALID_ARGUMENT, "unknown numerical flux");
    if (space->dev.nloc != 1)
      return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH, "scalar flux functions need a finite volume space with one component");
  }
  auto L = new gdtb_fvop();
  L->ctx = ctx;
  L->grid = space->grid;
  L->space = space->dev;
  L->flux = *flux;
  L->ghosted = false;
  L->d_tmp = L->d_src = L->d_dst = L->d_ext = nullptr;
  L->rows_per_block = 0;
  if (const char* env = std::getenv("GDTB_FV_ROWS"))
    L->rows_per_block = std::atoi(env);
  // cell extents per axis, exactly as YaspGrid's EquidistantOffsetCoordinates produce them: upper - lower with
  // lower = origin + i * h, upper = origin + (i + 1) * h [EXT]
  std::vector<double> ext;
  for (int k = 0; k < 3; ++k) {
    L->ext_offset[k] = (long long)ext.size();
    for (long long i = 0; i < L->grid.n[k]; ++i) {
      volatile double lower = L->grid.lo[k] + double(i) * L->grid.h[k];
      volatile double upper = L->grid.lo[k] + double(i + 1) * L->grid.h[k];
      ext.push_back(k < L->grid.d ? upper - lower : 1.);
    }
    if (ext.size() & 1) // every axis table starts 16-byte aligned (the marching kernel reads pairs)
      ext.push_back(1.);
  }
  // ... followed by the reciprocals 1 / ext, same layout
  L->inv_ext_shift = (long long)ext.size();
  for (long long i = 0; i < L->inv_ext_shift; ++i)
    ext.push_back(1. / ext[(size_t)i]);
  if (cudaMalloc(&L->d_ext, sizeof(double) * ext.size()) != cudaSuccess
      || cudaMemcpyAsync(L->d_ext, ext.data(), sizeof(double) * ext.size(), cudaMemcpyHostToDevice, ctx->launch.stream)
             != cudaSuccess
      || cudaStreamSynchronize(ctx->launch.stream) != cudaSuccess) {
    cudaFree(L->d_ext);
    delete L;
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (fv geometry tables)");
  }
  *out = L;
  return GDTB_OK;
}

static void fv_fill_params(const gdtb_fvop* L, FvParams& p)
{
  p.g = L->grid;
  p.flux = L->flux;
  p.ghosted = L->ghosted ? 1 : 0;
  p.euler = 0;
  p.dt = 0.;
  p.lf_lambda_linear = 0.;
  for (int k = 0; k < L->grid.d; ++k)
    p.lf_lambda_linear = std::max(p.lf_lambda_linear, std::fabs(L->flux.p[k]));
  for (int k = 0; k < 3; ++k) {
    p.ext[k] = L->d_ext + L->ext_offset[k];
    p.inv_ext[k] = L->d_ext + L->inv_ext_shift + L->ext_offset[k];
  }
  p.rows_per_block = L->rows_per_block;
  p.apply_lo = L->grid.layer_lo;
  p.apply_hi = L->grid.layer_hi;
  p.n_stage = -1; // no Runge-Kutta stage fused into the apply
  p.out_mode = p.n_out = 0;
  p.out_cL = 0.;
  for (int j = 0; j < 3; ++j) {
    p.stage_v[j] = p.out_v[j] = nullptr;
    p.stage_c[j] = p.out_c[j] = 0.;
  }
  p.p2p = 0;
  p.wait_lo = p.wait_hi = 0;
  p.peer_lo_ghost = p.peer_hi_ghost = nullptr;
  p.peer_lo_flag = p.peer_hi_flag = nullptr;
  p.my_flags = nullptr;
  p.expect = 0;
  p.timeout_flag = nullptr;
  p.edge_count = nullptr;
  p.bnd_ext_mask = L->bnd_ext_mask;
  p.bnd_nf_mask = L->bnd_nf_mask;
  for (int i = 0; i < 6; ++i) {
    p.bnd_ext_a[i] = L->bnd_ext_a[i];
    p.bnd_ext_b[i] = L->bnd_ext_b[i];
    p.bnd_nf_a[i] = L->bnd_nf_a[i];
    p.bnd_nf_b[i] = L->bnd_nf_b[i];
  }
}

static void fv_p2p_release(gdtb_fvop* L);
static long long fv_local_size(const gdtb_fvop* L);

int gdtb_fvop_destroy(gdtb_fvop* L)
{
  if (!L)
    return GDTB_OK;
  cudaSetDevice(L->ctx->device);
  cudaFree(L->d_tmp);
  cudaFree(L->d_src);
  cudaFree(L->d_dst);
  cudaFree(L->d_ext);
  cudaFree(L->d_partial);
  fv_p2p_release(L);
  delete L;
  return GDTB_OK;
}

int gdtb_fvop_append_boundary(gdtb_fvop* L, const gdtb_fv_boundary* t)
{
  if (!L || !t)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_append_boundary: NULL argument");
  const unsigned all = (1u << (2 * L->grid.d)) - 1u;
  if (t->side_mask & ~all)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "boundary treatment: side_mask names a side the grid does not have");
  if (fv_is_system(L)) {
    // systems: the impermeable-wall treatments of the Euler equations (the masks of the scalar families are reused:
    // numerical boundary flux -> wall flux, extrapolation -> mirrored state)
    if (t->kind != GDTB_FVBND_EULER_IMPERMEABLE_WALL && t->kind != GDTB_FVBND_EULER_INVISCID_MIRROR)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "boundary treatments of systems: GDTB_FVBND_EULER_IMPERMEABLE_WALL / _INVISCID_MIRROR");
    if (t->kind == GDTB_FVBND_EULER_INVISCID_MIRROR && (t->side_mask & L->bnd_ext_mask))
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "two extrapolation treatments on the same boundary side");
    (t->kind == GDTB_FVBND_EULER_INVISCID_MIRROR ? L->bnd_ext_mask : L->bnd_nf_mask) |= t->side_mask;
    return GDTB_OK;
  }
  if (t->kind != GDTB_FVBND_EXTRAPOLATION && t->kind != GDTB_FVBND_NUMERICAL_FLUX)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "unknown boundary treatment");
  if (t->kind == GDTB_FVBND_EXTRAPOLATION && (t->side_mask & L->bnd_ext_mask))
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "two extrapolation treatments on the same boundary side");
  for (int side = 0; side < 2 * L->grid.d; ++side) {
    if (!(t->side_mask >> side & 1))
      continue;
    if (t->kind == GDTB_FVBND_EXTRAPOLATION) {
      L->bnd_ext_a[side] = t->a;
      L->bnd_ext_b[side] = t->b;
    } else { // numerical boundary fluxes of several treatments add up: g is affine in f(u) . n
      L->bnd_nf_a[side] += t->a;
      L->bnd_nf_b[side] += t->b;
    }
  }
  (t->kind == GDTB_FVBND_EXTRAPOLATION ? L->bnd_ext_mask : L->bnd_nf_mask) |= t->side_mask;
  return GDTB_OK;
}

// ---- peer-memory ghost exchange (multi-GPU FV time loop without host-launched collectives) -----------------------
// The slab vectors of the time loop live in library-owned cudaMalloc memory that the neighbour processes open through
// CUDA IPC.  Step s reads u[s % 2] and writes u[(s + 1) % 2]; k_fv_march<..., P2P> stores the first / last owned layer of
// the result into the neighbours' ghost layers (NVLink peer stores) and raises their step counters; the blocks that
// read a ghost layer wait for the counter of the previous step.  Replaces the DataHandle communicate() of
// tools/timestepper/explicit-rungekutta.hh:252-257 for the fused Euler loop.
static void fv_p2p_release(gdtb_fvop* L)
{
  for (int side = 0; side < 2; ++side) {
    if (L->p2p_opened[side]) {
      for (void* ptr : {(void*)L->p2p_peer_u[side][0], (void*)L->p2p_peer_u[side][1], (void*)L->p2p_peer_flags[side]})
        if (ptr)
          cudaIpcCloseMemHandle(ptr);
    }
    L->p2p_opened[side] = false;
    L->p2p_peer_u[side][0] = L->p2p_peer_u[side][1] = nullptr;
    L->p2p_peer_flags[side] = nullptr;
  }
  cudaFree(L->p2p_u[0]);
  cudaFree(L->p2p_u[1]);
  cudaFree(L->p2p_flags);
  L->p2p_u[0] = L->p2p_u[1] = nullptr;
  L->p2p_flags = nullptr;
}

int gdtb_fvop_p2p_alloc(gdtb_fvop* L, double** d_u0, double** d_u1, void* handles)
{
  if (!L || !handles)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_p2p_alloc: NULL argument");
  GDTB_TRY(check_ctx(L->ctx));
  if (!L->ghosted)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_p2p_alloc: call gdtb_fvop_set_slab first");
  if (L->grid.d < 2)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "gdtb_fvop_p2p_alloc: 2D and 3D grids only");
  fv_p2p_release(L);
  const size_t bytes = sizeof(double) * (size_t)fv_local_size(L);
  // flags: [0] lower ghost filled, [1] upper ghost filled, [2] wait timed out, [4..5] edge block counters
  if (cudaMalloc(&L->p2p_u[0], bytes) != cudaSuccess || cudaMalloc(&L->p2p_u[1], bytes) != cudaSuccess
      || cudaMalloc(&L->p2p_flags, 64 * sizeof(int)) != cudaSuccess) {
    fv_p2p_release(L);
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (peer-memory slab vectors)");
  }
  GDTB_CUDA(cudaMemset(L->p2p_u[0], 0, bytes));
  GDTB_CUDA(cudaMemset(L->p2p_u[1], 0, bytes));
  GDTB_CUDA(cudaMemset(L->p2p_flags, 0, 64 * sizeof(int)));
  cudaIpcMemHandle_t h[3];
  if (cudaIpcGetMemHandle(&h[0], L->p2p_u[0]) != cudaSuccess || cudaIpcGetMemHandle(&h[1], L->p2p_u[1]) != cudaSuccess
      || cudaIpcGetMemHandle(&h[2], L->p2p_flags) != cudaSuccess) {
    const std::string why = cudaGetErrorString(cudaGetLastError());
    fv_p2p_release(L);
    return fail(GDTB_ERR_CUDA, "cudaIpcGetMemHandle failed: " + why);
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == GDTB_IPC_HANDLE_BYTES, "IPC handle size");
  std::memcpy(handles, h, sizeof(h));
  if (d_u0)
    *d_u0 = L->p2p_u[0];
  if (d_u1)
    *d_u1 = L->p2p_u[1];
  L->p2p_step = 0;
  return GDTB_OK;
}

int gdtb_fvop_p2p_connect(gdtb_fvop* L, const void* lower_handles, int64_t lower_layers, int lower_is_self,
                          const void* upper_handles, int64_t upper_layers, int upper_is_self)
{
  if (!L || !L->p2p_u[0])
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_p2p_connect: call gdtb_fvop_p2p_alloc first");
  GDTB_TRY(check_ctx(L->ctx));
  const void* hs[2] = {lower_handles, upper_handles};
  const int self[2] = {lower_is_self, upper_is_self};
  const int64_t layers[2] = {lower_layers, upper_layers};
  for (int side = 0; side < 2; ++side) {
    L->p2p_peer_layers[side] = layers[side];
    if (self[side]) { // the periodic neighbour is this rank itself (one slab)
      L->p2p_peer_u[side][0] = L->p2p_u[0];
      L->p2p_peer_u[side][1] = L->p2p_u[1];
      L->p2p_peer_flags[side] = L->p2p_flags;
      L->p2p_peer_layers[side] = L->grid.layer_hi - L->grid.layer_lo;
      continue;
    }
    if (!hs[side])
      continue; // domain boundary without periodicity
    if (side == 1 && hs[0] && !self[0] && std::memcmp(hs[0], hs[1], 3 * GDTB_IPC_HANDLE_BYTES) == 0) {
      // two slabs, periodic: both neighbours are the same process; a handle can be opened once per process
      L->p2p_peer_u[1][0] = L->p2p_peer_u[0][0];
      L->p2p_peer_u[1][1] = L->p2p_peer_u[0][1];
      L->p2p_peer_flags[1] = L->p2p_peer_flags[0];
      continue;
    }
    cudaIpcMemHandle_t h[3];
    std::memcpy(h, hs[side], sizeof(h));
    void* ptr[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < 3; ++i)
      if (cudaIpcOpenMemHandle(&ptr[i], h[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        return fail(GDTB_ERR_CUDA, std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError()));
    L->p2p_peer_u[side][0] = static_cast<double*>(ptr[0]);
    L->p2p_peer_u[side][1] = static_cast<double*>(ptr[1]);
    L->p2p_peer_flags[side] = static_cast<int*>(ptr[2]);
    L->p2p_opened[side] = true;
  }
  return GDTB_OK;
}

int gdtb_fvop_p2p_step(gdtb_fvop* L, int euler, double dt)
{
  if (!L || !L->p2p_u[0])
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_p2p_step: call gdtb_fvop_p2p_alloc / connect first");
  GDTB_TRY(check_ctx(L->ctx));
  FvParams p;
  fv_fill_params(L, p);
  p.euler = euler ? 1 : 0;
  p.dt = dt;
  const long long plane = gdtb_fvop_ghost_layer_size(L);
  const int src = (int)(L->p2p_step & 1), dst = src ^ 1;
  p.p2p = 1;
  if (L->p2p_peer_u[0][dst]) { // my first layer -> the lower neighbour's upper ghost layer
    p.peer_lo_ghost = L->p2p_peer_u[0][dst] + (L->p2p_peer_layers[0] + 1) * plane;
    p.peer_lo_flag = L->p2p_peer_flags[0] + 1;
  }
  if (L->p2p_peer_u[1][dst]) { // my last layer -> the upper neighbour's lower ghost layer
    p.peer_hi_ghost = L->p2p_peer_u[1][dst];
    p.peer_hi_flag = L->p2p_peer_flags[1] + 0;
  }
  p.wait_lo = p.peer_lo_ghost != nullptr;
  p.wait_hi = p.peer_hi_ghost != nullptr;
  p.my_flags = L->p2p_flags;
  p.timeout_flag = L->p2p_flags + 2;
  p.edge_count = L->p2p_flags + 4;
  p.expect = (int)L->p2p_step;
  GDTB_TRY(launch_fv_apply(L->ctx->launch, p, L->p2p_u[src], L->p2p_u[dst]));
  L->p2p_step++;
  return GDTB_OK;
}

int gdtb_fvop_p2p_current(gdtb_fvop* L, double** d_u, int64_t* step)
{
  if (!L || !L->p2p_u[0])
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_p2p_current: no peer-memory vectors");
  if (d_u)
    *d_u = L->p2p_u[L->p2p_step & 1];
  if (step)
    *step = L->p2p_step;
  return GDTB_OK;
}

int gdtb_fvop_p2p_check(gdtb_fvop* L)
{
  if (!L || !L->p2p_flags)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_p2p_check: no peer-memory vectors");
  GDTB_TRY(check_ctx(L->ctx));
  int flags[3] = {0, 0, 0};
  GDTB_CUDA(cudaStreamSynchronize(L->ctx->launch.stream));
  GDTB_CUDA(cudaMemcpy(flags, L->p2p_flags, sizeof(flags), cudaMemcpyDeviceToHost));
  if (flags[2])
    return fail(GDTB_ERR_OPERATOR, "peer-memory ghost exchange: a wait for the neighbour's step counter timed out");
  return GDTB_OK;
}

int64_t gdtb_fvop_ghost_layer_size(const gdtb_fvop* L)
{
  if (!L)
    return 0;
  long long plane = 1;
  for (int k = 0; k < L->grid.d - 1; ++k)
    plane *= L->grid.n[k];
  return plane;
}

static long long fv_local_size(const gdtb_fvop* L)
{
  const long long plane = gdtb_fvop_ghost_layer_size(L);
  return plane * (L->grid.layer_hi - L->grid.layer_lo + (L->ghosted ? 2 : 0)) * L->space.nloc;
}

int gdtb_fvop_set_slab(gdtb_fvop* L, int64_t layer_begin, int64_t layer_end)
{
  if (!L)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "operator is NULL");
  GDTB_NO_SYSTEMS(L, "gdtb_fvop_set_slab");
  const long long n_last = L->grid.n[L->grid.d - 1];
  if (layer_begin < 0 || layer_end > n_last || layer_begin >= layer_end)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "slab must satisfy 0 <= begin < end <= n[last]");
  L->grid.layer_lo = layer_begin;
  L->grid.layer_hi = layer_end;
  L->ghosted = true;
  cudaFree(L->d_tmp);
  L->d_tmp = nullptr;
  return GDTB_OK;
}

int gdtb_fvop_apply(gdtb_fvop* L, const double* d_source, double* d_range)
{
  if (!L || !d_source || !d_range)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_apply: NULL argument");
  GDTB_TRY(check_ctx(L->ctx));
  if (fv_is_system(L)) {
    FvSysParams q;
    fvsys_fill_params(L, q);
    GDTB_TRY(launch_fvsys_apply(L->ctx->launch, q, d_source, d_range));
    GDTB_CUDA(cudaStreamSynchronize(L->ctx->launch.stream));
    return GDTB_OK;
  }
  FvParams p;
  fv_fill_params(L, p);
  GDTB_TRY(launch_fv_apply(L->ctx->launch, p, d_source, d_range));
  GDTB_CUDA(cudaStreamSynchronize(L->ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_fvop_step_async(gdtb_fvop* L, const double* d_source, double* d_range, int euler, double dt,
                         int64_t layer_begin, int64_t layer_end)
{
  if (!L || !d_source || !d_range)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_step_async: NULL argument");
  GDTB_TRY(check_ctx(L->ctx));
  GDTB_NO_SYSTEMS(L, "gdtb_fvop_step_async");
  FvParams p;
  fv_fill_params(L, p);
  if (layer_begin != 0 || layer_end != 0) {
    if (layer_begin < L->grid.layer_lo || layer_end > L->grid.layer_hi || layer_begin > layer_end)
      return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_step_async: layers must lie inside the operator's slab");
    p.apply_lo = layer_begin;
    p.apply_hi = layer_end;
  }
  p.euler = euler ? 1 : 0;
  p.dt = dt;
  return launch_fv_apply(L->ctx->launch, p, d_source, d_range);
}

static int fv_stage(gdtb_fvop* L)
{
  const size_t bytes = sizeof(double) * (size_t)fv_local_size(L);
  if (!L->d_src && cudaMalloc(&L->d_src, bytes) != cudaSuccess)
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (fv staging)");
  if (!L->d_dst && cudaMalloc(&L->d_dst, bytes) != cudaSuccess)
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (fv staging)");
  return GDTB_OK;
}

int gdtb_fvop_apply_host(gdtb_fvop* L, const double* source, double* range)
{
  if (!L || !source || !range)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_apply_host: NULL argument");
  GDTB_TRY(check_ctx(L->ctx));
  GDTB_TRY(fv_stage(L));
  const size_t bytes = sizeof(double) * (size_t)fv_local_size(L);
  cudaStream_t s = L->ctx->launch.stream;
  GDTB_CUDA(cudaMemcpyAsync(L->d_src, source, bytes, cudaMemcpyHostToDevice, s));
  // apply(VectorType source, ...) of the reference refuses a source with inf / nan
  // (operators/localizable-operator.hh:383-385); checked on the device copy.  The device-pointer entry points are the
  // discrete-function overload, which has no such check.
  {
    gdtb_ctx* ctx = L->ctx;
    const long long n = (long long)fv_local_size(L);
    if (ctx->error_flag_pending) // an unchecked asynchronous assembly uses the same flag: report it first
      GDTB_TRY(gdtb_ctx_synchronize(ctx));
    GDTB_CUDA(cudaMemsetAsync(ctx->d_error_flag, 0, sizeof(int), s));
    if (n > 0) {
      k_flag_non_finite<<<(unsigned)std::min<long long>((n + 255) / 256, 4LL * ctx->launch.sm_count), 256, 0, s>>>(
          L->d_src, n, ctx->d_error_flag);
      ctx->launch.count++;
    }
    int flag = 0;
    GDTB_CUDA(cudaMemcpyAsync(&flag, ctx->d_error_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    GDTB_CUDA(cudaStreamSynchronize(s));
    if (flag)
      return fail(GDTB_ERR_OPERATOR, "source contains inf or nan! (Exceptions::operator_error, localizable-operator.hh:385)");
  }
  GDTB_TRY(gdtb_fvop_apply(L, L->d_src, L->d_dst));
  GDTB_CUDA(cudaMemcpyAsync(range, L->d_dst, bytes, cudaMemcpyDeviceToHost, s));
  GDTB_CUDA(cudaStreamSynchronize(s));
  return GDTB_OK;
}

int gdtb_fvop_euler(gdtb_fvop* L, double* d_u, double dt, int64_t n_steps)
{
  if (!L || !d_u)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_euler: NULL argument");
  if (n_steps < 0)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_euler: negative step count");
  GDTB_TRY(check_ctx(L->ctx));
  if (L->ghosted && n_steps > 1)
    return fail(GDTB_ERR_INVALID_ARGUMENT,
                "gdtb_fvop_euler: on a slab the ghost layers must be exchanged between steps (n_steps <= 1)");
  const size_t bytes = sizeof(double) * (size_t)fv_local_size(L);
  if (!L->d_tmp && cudaMalloc(&L->d_tmp, bytes) != cudaSuccess)
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (fv euler buffer)");
  FvParams p;
  fv_fill_params(L, p);
  p.euler = 1;
  p.dt = dt;
  FvSysParams q;
  if (fv_is_system(L)) {
    fvsys_fill_params(L, q);
    q.euler = 1;
    q.dt = dt;
  }
  double* a = d_u;
  double* b = L->d_tmp;
  for (int64_t s = 0; s < n_steps; ++s) {
    if (fv_is_system(L))
      GDTB_TRY(launch_fvsys_apply(L->ctx->launch, q, a, b));
    else
      GDTB_TRY(launch_fv_apply(L->ctx->launch, p, a, b));
    std::swap(a, b);
  }
  if (a != d_u) // odd number of steps: result sits in the internal buffer
    GDTB_CUDA(cudaMemcpyAsync(d_u, a, bytes, cudaMemcpyDeviceToDevice, L->ctx->launch.stream));
  GDTB_CUDA(cudaStreamSynchronize(L->ctx->launch.stream));
  return GDTB_OK;
}

int gdtb_fvop_euler_host(gdtb_fvop* L, double* u, double dt, int64_t n_steps)
{
  if (!L || !u)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fvop_euler_host: NULL argument");
  GDTB_TRY(check_ctx(L->ctx));
  GDTB_TRY(fv_stage(L));
  const size_t bytes = sizeof(double) * (size_t)fv_local_size(L);
  cudaStream_t s = L->ctx->launch.stream;
  GDTB_CUDA(cudaMemcpyAsync(L->d_src, u, bytes, cudaMemcpyHostToDevice, s));
  GDTB_TRY(gdtb_fvop_euler(L, L->d_src, dt, n_steps));
  GDTB_CUDA(cudaMemcpyAsync(u, L->d_src, bytes, cudaMemcpyDeviceToHost, s));
  GDTB_CUDA(cudaStreamSynchronize(s));
  return GDTB_OK;
}

// ---- estimate_dt_for_hyperbolic_system (tools/hyperbolic.hh:38-86) ------------------------------------------------
int gdtb_fv_estimate_dt(gdtb_fvop* L, const double* d_u, const double* boundary_data_range, double* dt)
{
  if (!L || !d_u || !dt)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fv_estimate_dt: NULL argument");
  GDTB_TRY(check_ctx(L->ctx));
  if (L->ghosted)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "gdtb_fv_estimate_dt: not available on a slab (reduce over the ranks yourself)");
  const int blocks = (int)std::max<long long>(1, std::min<long long>((L->grid.ne + 255) / 256, (long long)L->ctx->launch.sm_count * 8));
  const int max_blocks = L->ctx->launch.sm_count * 8;
  if (!L->d_partial && cudaMalloc(&L->d_partial, sizeof(double) * 8 * (size_t)max_blocks) != cudaSuccess)
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (dt estimate)");
  FvParams p;
  fv_fill_params(L, p);
  // (for a system the scalar reduction is only used for max perimeter / volume: it reads the first ne entries of d_u)
  GDTB_TRY(launch_fv_dt_reduce(L->ctx->launch, p, d_u, L->d_partial, blocks));
  std::vector<double> h(3 * (size_t)blocks);
  GDTB_CUDA(cudaMemcpyAsync(h.data(), L->d_partial, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, L->ctx->launch.stream));
  GDTB_CUDA(cudaStreamSynchronize(L->ctx->launch.stream));
  if (fv_is_system(L)) {
    // hyperbolic.hh:50-86 for m > 1: data range per component, Gauss rule of order flux.order() (EulerTools::flux_order()
    // = 4) on the one-cell grid [min, max] in R^m, largest infinity norm of the flux jacobians there
    if (boundary_data_range)
      return fail(GDTB_ERR_NOT_IMPLEMENTED, "gdtb_fv_estimate_dt: boundary data ranges for systems");
    const int m = L->space.nloc;
    double perimeter_over_volume = std::numeric_limits<double>::min();
    for (int i = 0; i < blocks; ++i)
      perimeter_over_volume = std::max(perimeter_over_volume, h[3 * i + 2]);
    FvSysParams q;
    fvsys_fill_params(L, q);
    GDTB_TRY(launch_fvsys_minmax(L->ctx->launch, q, d_u, L->d_partial, blocks));
    std::vector<double> mm(2 * (size_t)m * blocks);
    GDTB_CUDA(cudaMemcpyAsync(mm.data(), L->d_partial, sizeof(double) * mm.size(), cudaMemcpyDeviceToHost, L->ctx->launch.stream));
    GDTB_CUDA(cudaStreamSynchronize(L->ctx->launch.stream));
    double lo[4], hi[4];
    for (int c = 0; c < m; ++c) {
      lo[c] = std::numeric_limits<double>::max();
      hi[c] = std::numeric_limits<double>::min();
      for (int i = 0; i < blocks; ++i) {
        lo[c] = std::min(lo[c], mm[(size_t)i * 2 * m + c]);
        hi[c] = std::max(hi[c], mm[(size_t)i * 2 * m + m + c]);
      }
      if (!(lo[c] < hi[c]))
        hi[c] = lo[c] + 1e-6 * lo[c];
    }
    const int nq = gauss_points_for_order(4);
    double qx[MAX_Q1D], qw[MAX_Q1D];
    gauss_legendre_01(nq, qx, qw);
    double max_flux_derivative = std::numeric_limits<double>::min();
    long long total = 1;
    for (int c = 0; c < m; ++c)
      total *= nq;
    for (long long t = 0; t < total; ++t) {
      double w[4];
      long long r = t;
      for (int c = 0; c < m; ++c) {
        w[c] = lo[c] + qx[r % nq] * (hi[c] - lo[c]);
        r /= nq;
      }
      max_flux_derivative = std::max(max_flux_derivative, euler_jacobian_inf_norm(L->grid.d, L->flux.p[0], w));
    }
    *dt = 1. / (perimeter_over_volume * max_flux_derivative);
    return GDTB_OK;
  }
  // hyperbolic.hh:47-48: {numeric_limits<R>::max(), numeric_limits<R>::min()} -- min() is the smallest positive normal
  double data_minimum = boundary_data_range ? boundary_data_range[0] : std::numeric_limits<double>::max();
  double data_maximum = boundary_data_range ? boundary_data_range[1] : std::numeric_limits<double>::min();
  double perimeter_over_volume = std::numeric_limits<double>::min();
  for (int i = 0; i < blocks; ++i) {
    data_minimum = std::min(data_minimum, h[3 * i]);
    data_maximum = std::max(data_maximum, h[3 * i + 1]);
    perimeter_over_volume = std::max(perimeter_over_volume, h[3 * i + 2]);
  }
  if (!(data_minimum < data_maximum)) // :62-64
    data_maximum = data_minimum + 1e-6 * data_minimum;
  // :66-74: Gauss rule of order flux.order() on the one-cell grid [min, max]
  double max_flux_derivative = std::numeric_limits<double>::min();
  const int m = gauss_points_for_order(L->flux.kind == GDTB_FLUX_LINEAR ? 1 : 2);
  double qx[MAX_Q1D], qw[MAX_Q1D];
  gauss_legendre_01(m, qx, qw);
  for (int q = 0; q < m; ++q) {
    const double uq = data_minimum + qx[q] * (data_maximum - data_minimum);
    for (int ss = 0; ss < L->grid.d; ++ss) {
      const double df = L->flux.kind == GDTB_FLUX_LINEAR ? L->flux.p[ss] : uq;
      max_flux_derivative = std::max(max_flux_derivative, std::fabs(df));
    }
  }
  *dt = 1. / (perimeter_over_volume * max_flux_derivative);
  return GDTB_OK;
}

int gdtb_fv_estimate_dt_host(gdtb_fvop* L, const double* u, const double* boundary_data_range, double* dt)
{
  if (!L || !u || !dt)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fv_estimate_dt_host: NULL argument");
  GDTB_TRY(check_ctx(L->ctx));
  GDTB_TRY(fv_stage(L));
  GDTB_CUDA(cudaMemcpyAsync(L->d_src, u, sizeof(double) * (size_t)fv_local_size(L), cudaMemcpyHostToDevice, L->ctx->launch.stream));
  return gdtb_fv_estimate_dt(L, L->d_src, boundary_data_range, dt);
}

// ---- ExplicitRungeKuttaTimeStepper (tools/timestepper/explicit-rungekutta.hh:158-270) ------------------------------
int gdtb_rk_create(gdtb_fvop* L, int method, int num_stages, const double* A, const double* b, const double* c, double r,
                   double t0, gdtb_rk** out)
{
  if (!L || !out)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_rk_create: NULL argument");
  GDTB_TRY(check_ctx(L->ctx));
  GDTB_NO_SYSTEMS(L, "gdtb_rk_create");
  auto ts = new gdtb_rk();
  ts->op = L;
  ts->slab = L->ghosted;
  ts->r = r;
  ts->t = t0;
  // internal::ButcherArrayProvider (:63-141)
  static const double A2[] = {0., 0., 1., 0.}, b2[] = {0.5, 0.5}, c2[] = {0., 1.};
  static const double A3[] = {0., 0., 0., 1., 0., 0., 0.25, 0.25, 0.}, b3[] = {1. / 6., 1. / 6., 2. / 3.}, c3[] = {0., 1., 0.5};
  static const double A4[] = {0., 0., 0., 0., 0.5, 0., 0., 0., 0., 0.5, 0., 0., 0., 0., 1., 0.};
  static const double b4[] = {1. / 6., 1. / 3., 1. / 3., 1. / 6.}, c4[] = {0., 0.5, 0.5, 1.};
  static const double A1[] = {0.}, b1[] = {1.}, c1[] = {0.};
  switch (method) {
    case GDTB_RK_EULER: num_stages = 1, A = A1, b = b1, c = c1; break;
    case GDTB_RK_SSP2: num_stages = 2, A = A2, b = b2, c = c2; break;
    case GDTB_RK_SSP3: num_stages = 3, A = A3, b = b3, c = c3; break;
    case GDTB_RK_CLASSIC4: num_stages = 4, A = A4, b = b4, c = c4; break;
    case GDTB_RK_OTHER:
      if (!A || !b || !c) {
        delete ts;
        return fail(GDTB_ERR_NOT_IMPLEMENTED, "You have to provide a Butcher array in ExplicitRungeKuttaTimeStepper's constructor for this method!"); // :40-58
      }
      break;
    default: delete ts; return fail(GDTB_ERR_INVALID_ARGUMENT, "unknown TimeStepperMethods value");
  }
  if (num_stages < 1 || num_stages > GDTB_RK_MAX_STAGES) {
    delete ts;
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_rk_create: 1 <= num_stages <= GDTB_RK_MAX_STAGES");
  }
  ts->s = num_stages;
  for (int i = 0; i < num_stages; ++i) {
    ts->b[i] = b[i];
    ts->c[i] = c[i];
    for (int j = 0; j < num_stages; ++j) {
      ts->A[i * num_stages + j] = A[i * num_stages + j];
      // FloatCmp::ne(A[ii][jj], 0.) for jj >= ii (:216-222)
      if (j >= i && std::fabs(A[i * num_stages + j]) > 8. * std::numeric_limits<double>::epsilon()) {
        delete ts;
        return fail(GDTB_ERR_INVALID_ARGUMENT, "A has to be a lower triangular matrix with 0 on the main diagonal");
      }
    }
  }
  const size_t bytes = sizeof(double) * (size_t)fv_local_size(L);
  if (ts->slab && (num_stages < 2 || L->grid.d < 2)) {
    delete ts;
    return fail(GDTB_ERR_NOT_IMPLEMENTED,
                "Runge-Kutta stepping on a slab needs >= 2 stages and a 2D / 3D grid (explicit Euler: gdtb_fvop_p2p_step)");
  }
  bool ok = cudaMalloc(&ts->d_ui, bytes) == cudaSuccess;
  for (int i = 0; ok && i < (num_stages > 1 ? num_stages : 0); ++i)
    ok = cudaMalloc(&ts->d_k[i], bytes) == cudaSuccess && cudaMemset(ts->d_k[i], 0, bytes) == cudaSuccess;
  if (ok && ts->slab) {
    ok = cudaMalloc(&ts->p2p_un, bytes) == cudaSuccess && cudaMalloc(&ts->p2p_ui[0], bytes) == cudaSuccess
         && cudaMalloc(&ts->p2p_ui[1], bytes) == cudaSuccess && cudaMalloc(&ts->p2p_flags, 64 * sizeof(int)) == cudaSuccess
         && cudaMemset(ts->p2p_un, 0, bytes) == cudaSuccess && cudaMemset(ts->p2p_ui[0], 0, bytes) == cudaSuccess
         && cudaMemset(ts->p2p_ui[1], 0, bytes) == cudaSuccess && cudaMemset(ts->p2p_flags, 0, 64 * sizeof(int)) == cudaSuccess;
  }
  if (!ok) {
    gdtb_rk_destroy(ts);
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory (Runge-Kutta stages)");
  }
  *out = ts;
  return GDTB_OK;
}

int gdtb_rk_destroy(gdtb_rk* ts)
{
  if (!ts)
    return GDTB_OK;
  cudaSetDevice(ts->op->ctx->device);
  cudaFree(ts->d_ui);
  for (double* k : ts->d_k)
    cudaFree(k);
  for (int side = 0; side < 2; ++side)
    if (ts->peer_opened[side])
      for (void* ptr : {(void*)ts->peer_un[side], (void*)ts->peer_ui[side][0], (void*)ts->peer_ui[side][1], (void*)ts->peer_flags[side]})
        if (ptr)
          cudaIpcCloseMemHandle(ptr);
  cudaFree(ts->p2p_un);
  cudaFree(ts->p2p_ui[0]);
  cudaFree(ts->p2p_ui[1]);
  cudaFree(ts->p2p_flags);
  delete ts;
  return GDTB_OK;
}

// ---- Runge-Kutta on slabs: stage vectors handed over by peer stores (explicit-rungekutta.hh:252-257) ----------------
int gdtb_rk_p2p_handles(gdtb_rk* ts, double** d_un, void* handles)
{
  if (!ts || !ts->slab || !handles)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_rk_p2p_handles: the stepper was not created on a slab operator");
  GDTB_TRY(check_ctx(ts->op->ctx));
  cudaIpcMemHandle_t h[4];
  void* ptr[4] = {ts->p2p_un, ts->p2p_ui[0], ts->p2p_ui[1], ts->p2p_flags};
  for (int i = 0; i < 4; ++i)
    if (cudaIpcGetMemHandle(&h[i], ptr[i]) != cudaSuccess)
      return fail(GDTB_ERR_CUDA, std::string("cudaIpcGetMemHandle failed: ") + cudaGetErrorString(cudaGetLastError()));
  std::memcpy(handles, h, sizeof(h));
  if (d_un)
    *d_un = ts->p2p_un;
  return GDTB_OK;
}

int gdtb_rk_p2p_connect(gdtb_rk* ts, const void* lower_handles, int64_t lower_layers, int lower_is_self,
                        const void* upper_handles, int64_t upper_layers, int upper_is_self)
{
  if (!ts || !ts->slab)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_rk_p2p_connect: the stepper was not created on a slab operator");
  gdtb_fvop* L = ts->op;
  GDTB_TRY(check_ctx(L->ctx));
  const void* hs[2] = {lower_handles, upper_handles};
  const int self[2] = {lower_is_self, upper_is_self};
  const int64_t layers[2] = {lower_layers, upper_layers};
  for (int side = 0; side < 2; ++side) {
    ts->peer_layers[side] = layers[side];
    ts->has_peer[side] = false;
    if (self[side]) {
      ts->peer_un[side] = ts->p2p_un;
      ts->peer_ui[side][0] = ts->p2p_ui[0];
      ts->peer_ui[side][1] = ts->p2p_ui[1];
      ts->peer_flags[side] = ts->p2p_flags;
      ts->peer_layers[side] = L->grid.layer_hi - L->grid.layer_lo;
      ts->has_peer[side] = true;
      continue;
    }
    if (!hs[side])
      continue;
    if (side == 1 && hs[0] && !self[0] && std::memcmp(hs[0], hs[1], 4 * GDTB_IPC_HANDLE_BYTES) == 0) {
      ts->peer_un[1] = ts->peer_un[0];
      ts->peer_ui[1][0] = ts->peer_ui[0][0];
      ts->peer_ui[1][1] = ts->peer_ui[0][1];
      ts->peer_flags[1] = ts->peer_flags[0];
      ts->has_peer[1] = true;
      continue;
    }
    cudaIpcMemHandle_t h[4];
    std::memcpy(h, hs[side], sizeof(h));
    void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < 4; ++i)
      if (cudaIpcOpenMemHandle(&ptr[i], h[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        return fail(GDTB_ERR_CUDA, std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError()));
    ts->peer_un[side] = static_cast<double*>(ptr[0]);
    ts->peer_ui[side][0] = static_cast<double*>(ptr[1]);
    ts->peer_ui[side][1] = static_cast<double*>(ptr[2]);
    ts->peer_flags[side] = static_cast<int*>(ptr[3]);
    ts->peer_opened[side] = true;
    ts->has_peer[side] = true;
  }
  return GDTB_OK;
}

int gdtb_rk_p2p_check(gdtb_rk* ts)
{
  if (!ts || !ts->slab)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_rk_p2p_check: the stepper was not created on a slab operator");
  GDTB_TRY(check_ctx(ts->op->ctx));
  int flags[3] = {0, 0, 0};
  GDTB_CUDA(cudaStreamSynchronize(ts->op->ctx->launch.stream));
  GDTB_CUDA(cudaMemcpy(flags, ts->p2p_flags, sizeof(flags), cudaMemcpyDeviceToHost));
  if (flags[2])
    return fail(GDTB_ERR_OPERATOR, "Runge-Kutta on slabs: a wait for the neighbour's counter timed out");
  return GDTB_OK;
}

// which = 0: the solution vector u_n, 1 / 2: the two stage buffers
static int rk_p2p_send(gdtb_rk* ts, int which)
{
  gdtb_fvop* L = ts->op;
  const long long plane = gdtb_fvop_ghost_layer_size(L);
  P2pSendParams q;
  q.src = which == 0 ? ts->p2p_un : ts->p2p_ui[which - 1];
  q.plane = plane;
  q.layers = L->grid.layer_hi - L->grid.layer_lo;
  double* lo = !ts->has_peer[0] ? nullptr : (which == 0 ? ts->peer_un[0] : ts->peer_ui[0][which - 1]);
  double* hi = !ts->has_peer[1] ? nullptr : (which == 0 ? ts->peer_un[1] : ts->peer_ui[1][which - 1]);
  q.peer_lo_ghost = lo ? lo + (ts->peer_layers[0] + 1) * plane : nullptr; // the lower neighbour's upper ghost layer
  q.peer_hi_ghost = hi;                                                   // the upper neighbour's lower ghost layer
  q.peer_lo_flag = lo ? ts->peer_flags[0] + 1 : nullptr;
  q.peer_hi_flag = hi ? ts->peer_flags[1] + 0 : nullptr;
  q.edge_count = ts->p2p_flags + 4;
  GDTB_TRY(launch_p2p_send_layers(L->ctx->launch, q));
  ts->sends++;
  return GDTB_OK;
}

// one step on the slab: the stage vectors are formed on the owned layers, their boundary layers handed to the
// neighbours, and every apply waits (inside the kernel) until the neighbours' hand-over of its source has arrived
static int rk_enqueue_step_slab(gdtb_rk* ts, double actual_dt)
{
  gdtb_fvop* L = ts->op;
  Launch& la = L->ctx->launch;
  const int s = ts->s;
  const long long plane = gdtb_fvop_ghost_layer_size(L);
  const long long owned = plane * (L->grid.layer_hi - L->grid.layer_lo);
  double* un = ts->p2p_un;
  for (int ii = 0; ii < s; ++ii) {
    const double* ui = un;
    if (ii > 0) {
      double* dst = ts->p2p_ui[ts->ui_parity];
      RkAxpyParams q;
      q.n = owned;
      q.nv = 0;
      const double* base = un + plane;
      bool wrote = false;
      for (int jj = 0; jj < ii; ++jj) {
        const double coef = actual_dt * ts->r * ts->A[ii * s + jj];
        if (coef == 0.)
          continue;
        q.v[q.nv] = ts->d_k[jj] + plane;
        q.c[q.nv] = coef;
        if (++q.nv == RK_MAX_TERMS) {
          GDTB_TRY(launch_rk_axpy(la, q, base, dst + plane));
          base = dst + plane;
          q.nv = 0;
          wrote = true;
        }
      }
      if (q.nv > 0 || !wrote) // (all coefficients zero: u_i = u_n, still a distinct buffer so that the hand-over order holds)
        GDTB_TRY(launch_rk_axpy(la, q, base, dst + plane));
      GDTB_TRY(rk_p2p_send(ts, 1 + ts->ui_parity));
      ui = dst;
      ts->ui_parity ^= 1;
    }
    FvParams p;
    fv_fill_params(L, p);
    p.p2p = 1;
    p.wait_lo = ts->has_peer[0];
    p.wait_hi = ts->has_peer[1];
    p.my_flags = ts->p2p_flags;
    p.timeout_flag = ts->p2p_flags + 2;
    p.edge_count = ts->p2p_flags + 5;
    p.expect = (int)ts->sends;
    GDTB_TRY(launch_fv_apply(la, p, ui, ts->d_k[ii]));
  }
  RkAxpyParams q;
  q.n = owned;
  q.nv = 0;
  for (int ii = 0; ii < s; ++ii) {
    const double coef = ts->r * actual_dt * ts->b[ii];
    if (coef == 0.)
      continue;
    q.v[q.nv] = ts->d_k[ii] + plane;
    q.c[q.nv] = coef;
    if (++q.nv == RK_MAX_TERMS) {
      GDTB_TRY(launch_rk_axpy(la, q, un + plane, un + plane));
      q.nv = 0;
    }
  }
  if (q.nv > 0)
    GDTB_TRY(launch_rk_axpy(la, q, un + plane, un + plane));
  return rk_p2p_send(ts, 0);
}

double gdtb_rk_current_time(const gdtb_rk* ts)
{
  return ts ? ts->t : 0.;
}

// enqueue one step of length actual_dt: reads `in`, leaves the new solution in `outp` (Euler: outp != in, fused
// apply + update; multi-stage: outp == in, updated in place)
static int rk_enqueue_step(gdtb_rk* ts, double* in, double* outp, double actual_dt)
{
  gdtb_fvop* L = ts->op;
  Launch& la = L->ctx->launch;
  FvParams p;
  fv_fill_params(L, p);
  const int s = ts->s;
  if (s == 1) {
    // u_n + k_0 (r dt b_0) with k_0 = L(u_n): the fused kernel's u - acc * dt' with dt' = -(r dt b_0) (same roundings)
    p.euler = 1;
    p.dt = -(ts->r * actual_dt * ts->b[0]);
    return launch_fv_apply(la, p, in, outp);
  }
  const long long n = fv_local_size(L);
  // Fused stages: k_ii = L(u_n + sum_jj k_jj c_jj) with the stage vector formed inside the apply (never written), and
  // the step's update u_n + sum_ii k_ii (r dt b_ii) produced by the last apply -- 9 vector passes per SSP3 step
  // instead of 18.  Plain kernel only (2D / 3D, no boundary treatments), at most 3 terms per stage.
  bool fuse = L->grid.d >= 2 && !L->ghosted && !(L->bnd_ext_mask | L->bnd_nf_mask) && !std::getenv("GDTB_RK_NO_FUSE");
  int n_final = 0;
  for (int ii = 0; ii < s && fuse; ++ii) {
    int nz = 0;
    for (int jj = 0; jj < ii; ++jj)
      nz += (actual_dt * ts->r * ts->A[ii * s + jj] != 0.) ? 1 : 0;
    fuse = nz <= 3;
    if (ii < s - 1 && ts->r * actual_dt * ts->b[ii] != 0.)
      ++n_final;
  }
  fuse = fuse && n_final <= 3;
  if (fuse) {
    for (int ii = 0; ii < s; ++ii) {
      FvParams q = p;
      q.n_stage = 0;
      for (int jj = 0; jj < ii; ++jj) {
        const double coef = actual_dt * ts->r * ts->A[ii * s + jj];
        if (coef == 0.)
          continue;
        q.stage_v[q.n_stage] = ts->d_k[jj];
        q.stage_c[q.n_stage] = coef;
        ++q.n_stage;
      }
      if (ii < s - 1) {
        GDTB_TRY(launch_fv_apply(la, q, in, ts->d_k[ii])); // k_ii = L(u_i)
        continue;
      }
      // last stage: u_{n+1} = u_n + sum_{ii < s-1} k_ii (r dt b_ii) + L(u_{s-1}) (r dt b_{s-1}), into the scratch vector
      q.out_mode = 1;
      for (int kk = 0; kk < s - 1; ++kk) {
        const double coef = ts->r * actual_dt * ts->b[kk];
        if (coef == 0.)
          continue;
        q.out_v[q.n_out] = ts->d_k[kk];
        q.out_c[q.n_out] = coef;
        ++q.n_out;
      }
      q.out_cL = ts->r * actual_dt * ts->b[s - 1];
      GDTB_TRY(launch_fv_apply(la, q, in, ts->d_ui));
    }
    // the solution lives in the caller's vector (the reference keeps a reference to initial_values)
    GDTB_CUDA(cudaMemcpyAsync(in, ts->d_ui, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, la.stream));
    (void)outp;
    return GDTB_OK;
  }
  for (int ii = 0; ii < s; ++ii) {
    const double* ui = in; // stage 0: u_i = u_n
    if (ii > 0) {
      // u_i = u_n + sum_{jj < ii} k_jj (dt r A[ii][jj]) (:248-250); terms with A[ii][jj] == 0 add exact zeros and are skipped
      RkAxpyParams q;
      q.n = n;
      q.nv = 0;
      const double* base = in;
      bool wrote = false;
      for (int jj = 0; jj < ii; ++jj) {
        const double coef = actual_dt * ts->r * ts->A[ii * s + jj];
        if (coef == 0.)
          continue;
        q.v[q.nv] = ts->d_k[jj];
        q.c[q.nv] = coef;
        if (++q.nv == RK_MAX_TERMS) {
          GDTB_TRY(launch_rk_axpy(la, q, base, ts->d_ui));
          base = ts->d_ui;
          q.nv = 0;
          wrote = true;
        }
      }
      if (q.nv > 0) {
        GDTB_TRY(launch_rk_axpy(la, q, base, ts->d_ui));
        wrote = true;
      }
      ui = wrote ? ts->d_ui : in;
    }
    GDTB_TRY(launch_fv_apply(la, p, ui, ts->d_k[ii])); // k_ii = L(u_i) (:253-255; the operator is autonomous)
  }
  // u_n += sum_ii k_ii (r dt b_ii) (:261-263)
  RkAxpyParams q;
  q.n = n;
  q.nv = 0;
  for (int ii = 0; ii < s; ++ii) {
    const double coef = ts->r * actual_dt * ts->b[ii];
    if (coef == 0.)
      continue;
    q.v[q.nv] = ts->d_k[ii];
    q.c[q.nv] = coef;
    if (++q.nv == RK_MAX_TERMS) {
      GDTB_TRY(launch_rk_axpy(la, q, in, in));
      q.nv = 0;
    }
  }
  if (q.nv > 0)
    GDTB_TRY(launch_rk_axpy(la, q, in, in));
  (void)outp;
  return GDTB_OK;
}

int gdtb_rk_step(gdtb_rk* ts, double* d_u, double dt, double max_dt, double* returned_dt)
{
  if (!ts || !d_u)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_rk_step: NULL argument");
  gdtb_fvop* L = ts->op;
  GDTB_TRY(check_ctx(L->ctx));
  const double actual_dt = std::min(dt, max_dt); // :239
  cudaStream_t st = L->ctx->launch.stream;
  if (ts->slab) {
    if (d_u != ts->p2p_un)
      return fail(GDTB_ERR_INVALID_ARGUMENT, "on a slab the solution lives in the stepper's own vector (gdtb_rk_p2p_handles)");
    GDTB_TRY(rk_enqueue_step_slab(ts, actual_dt));
  } else if (ts->s == 1) {
    GDTB_TRY(rk_enqueue_step(ts, d_u, ts->d_ui, actual_dt));
    GDTB_CUDA(cudaMemcpyAsync(d_u, ts->d_ui, sizeof(double) * (size_t)fv_local_size(L), cudaMemcpyDeviceToDevice, st));
  } else
    GDTB_TRY(rk_enqueue_step(ts, d_u, d_u, actual_dt));
  GDTB_CUDA(cudaStreamSynchronize(st));
  ts->t += actual_dt; // :266
  if (returned_dt)
    *returned_dt = dt; // :268
  return GDTB_OK;
}

namespace {
// XT::Common::FloatCmp::{lt,gt} (numpy style, default epsilons) [EXT dune-xt]: eq(a,b) = |a-b| <= eps + eps |b|,
// eps = 8 * 2^-52 (Dune::FloatCmp::DefaultEpsilon<double>)
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
} // namespace

int gdtb_rk_solve(gdtb_rk* ts, double* d_u, double t_end, double initial_dt, int64_t* n_steps, double* next_dt)
{
  if (!ts || !d_u)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_rk_solve: NULL argument");
  if (!(initial_dt > 0.))
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_rk_solve: initial_dt must be positive");
  gdtb_fvop* L = ts->op;
  GDTB_TRY(check_ctx(L->ctx));
  Launch& la = L->ctx->launch;
  cudaStream_t st = la.stream;
  const size_t bytes = sizeof(double) * (size_t)fv_local_size(L);
  // the step sequence of TimeStepperInterface::solve (interface.hh:216-255) depends on t alone: plan it on the host
  double dt = initial_dt, t = ts->t;
  std::vector<double> plan;
  while (floatcmp_lt(t, t_end)) {
    double max_dt = dt;
    if (floatcmp_gt(t + dt, t_end))
      max_dt = t_end - t;
    plan.push_back(std::min(dt, max_dt));
    t += plan.back();
  }
  // full steps run as replays of one captured graph (Euler: two steps per graph, ping-pong d_u -> buffer -> d_u)
  size_t n_full = 0;
  while (n_full < plan.size() && plan[n_full] == initial_dt)
    ++n_full;
  const int per_graph = ts->s == 1 ? 2 : 1;
  if (ts->slab && d_u != ts->p2p_un)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "on a slab the solution lives in the stepper's own vector (gdtb_rk_p2p_handles)");
  // (slab mode: the counters each apply waits for advance from step to step, so the steps are enqueued one by one)
  const size_t replays = (n_full >= 8 && !ts->slab) ? n_full / per_graph : 0;
  int status = GDTB_OK;
  double* cur = d_u; // where the current solution lives (Euler ping-pong)
  size_t done = 0;
  if (replays > 0) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const bool timing_was = L->ctx->timing.enabled;
    L->ctx->timing.enabled = false; // event records are not capturable
    const long long count_before = la.count;
    {
      const cudaError_t berr = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
      if (berr != cudaSuccess) {
        L->ctx->timing.enabled = timing_was; // (restored on this exit as well)
        return fail(GDTB_ERR_CUDA, std::string("Runge-Kutta graph capture: ") + cudaGetErrorString(berr));
      }
    }
    if (ts->s == 1) {
      status = rk_enqueue_step(ts, d_u, ts->d_ui, initial_dt);
      if (status == GDTB_OK)
        status = rk_enqueue_step(ts, ts->d_ui, d_u, initial_dt);
    } else
      status = rk_enqueue_step(ts, d_u, d_u, initial_dt);
    cudaError_t err = cudaStreamEndCapture(st, &graph);
    const long long launched = la.count - count_before;
    la.count = count_before;
    L->ctx->timing.enabled = timing_was;
    if (status == GDTB_OK && err != cudaSuccess)
      status = fail(GDTB_ERR_CUDA, std::string("Runge-Kutta graph capture: ") + cudaGetErrorString(err));
    if (status == GDTB_OK && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess)
      status = fail(GDTB_ERR_CUDA, "Runge-Kutta graph instantiation failed");
    for (size_t k = 0; status == GDTB_OK && k < replays; ++k) {
      if (cudaGraphLaunch(exec, st) != cudaSuccess)
        status = fail(GDTB_ERR_CUDA, "Runge-Kutta graph launch failed");
      la.count += launched;
    }
    done = replays * per_graph;
    if (exec)
      cudaGraphExecDestroy(exec);
    if (graph)
      cudaGraphDestroy(graph);
  }
  for (size_t k = done; status == GDTB_OK && k < plan.size(); ++k) {
    if (ts->slab)
      status = rk_enqueue_step_slab(ts, plan[k]);
    else if (ts->s == 1) {
      double* nxt = cur == d_u ? ts->d_ui : d_u;
      status = rk_enqueue_step(ts, cur, nxt, plan[k]);
      cur = nxt;
    } else
      status = rk_enqueue_step(ts, d_u, d_u, plan[k]);
  }
  if (status == GDTB_OK && cur != d_u)
    GDTB_CUDA(cudaMemcpyAsync(d_u, cur, bytes, cudaMemcpyDeviceToDevice, st));
  cudaError_t e2 = cudaStreamSynchronize(st);
  if (status == GDTB_OK && e2 != cudaSuccess)
    status = fail(GDTB_ERR_CUDA, std::string("gdtb_rk_solve: ") + cudaGetErrorString(e2));
  if (status != GDTB_OK)
    return status;
  for (double a : plan) // t += actual_dt per step (:266), same roundings as the planning loop
    ts->t += a;
  if (n_steps)
    *n_steps = (int64_t)plan.size();
  if (next_dt)
    *next_dt = dt; // step() returns the dt it was given (interface.hh:226, explicit-rungekutta.hh:268)
  return GDTB_OK;
}

int gdtb_rk_step_host(gdtb_rk* ts, double* u, double dt, double max_dt, double* returned_dt)
{
  if (!ts || !u)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_rk_step_host: NULL argument");
  gdtb_fvop* L = ts->op;
  GDTB_TRY(check_ctx(L->ctx));
  GDTB_TRY(fv_stage(L));
  const size_t bytes = sizeof(double) * (size_t)fv_local_size(L);
  cudaStream_t s = L->ctx->launch.stream;
  GDTB_CUDA(cudaMemcpyAsync(L->d_src, u, bytes, cudaMemcpyHostToDevice, s));
  GDTB_TRY(gdtb_rk_step(ts, L->d_src, dt, max_dt, returned_dt));
  GDTB_CUDA(cudaMemcpyAsync(u, L->d_src, bytes, cudaMemcpyDeviceToHost, s));
  GDTB_CUDA(cudaStreamSynchronize(s));
  return GDTB_OK;
}

int gdtb_rk_solve_host(gdtb_rk* ts, double* u, double t_end, double initial_dt, int64_t* n_steps, double* next_dt)
{
  if (!ts || !u)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_rk_solve_host: NULL argument");
  gdtb_fvop* L = ts->op;
  GDTB_TRY(check_ctx(L->ctx));
  GDTB_TRY(fv_stage(L));
  const size_t bytes = sizeof(double) * (size_t)fv_local_size(L);
  cudaStream_t s = L->ctx->launch.stream;
  GDTB_CUDA(cudaMemcpyAsync(L->d_src, u, bytes, cudaMemcpyHostToDevice, s));
  GDTB_TRY(gdtb_rk_solve(ts, L->d_src, t_end, initial_dt, n_steps, next_dt));
  GDTB_CUDA(cudaMemcpyAsync(u, L->d_src, bytes, cudaMemcpyDeviceToHost, s));
  GDTB_CUDA(cudaStreamSynchronize(s));
  return GDTB_OK;
}

int gdtb_fv_interpolate(gdtb_ctx* ctx, const gdtb_space* space, const gdtb_function* f, double* d_u)
{
  if (!space || !f || !d_u)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fv_interpolate: NULL argument");
  GDTB_TRY(check_ctx(ctx));
  if (space->dev.kind != GDTB_SPACE_FV)
    return fail(GDTB_ERR_SPACE, "gdtb_fv_interpolate needs a finite volume space");
  GDTB_TRY(validate_function(*f, "function"));
  if (fn_has_data(*f) && !f->data_on_device)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fv_interpolate: array-backed function data must live on the device");
  if (f->kind == GDTB_FN_DOF_VECTOR || f->kind == GDTB_FN_QP_TENSOR || f->kind == GDTB_FN_ELEM_TENSOR
      || f->kind == GDTB_FN_CONST_TENSOR)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "gdtb_fv_interpolate: f must be a scalar constant, per-element, "
                                          "per-quadrature-point or analytic function");
  const int m = gauss_points_for_order(f->order);
  if (m > MAX_Q1D)
    return fail(GDTB_ERR_NOT_IMPLEMENTED, "quadrature order too high");
  if (f->kind == GDTB_FN_QP_SCALAR && f->qp_per_element != ipow(m, space->grid.d))
    return fail(GDTB_ERR_SHAPES_DO_NOT_MATCH,
                "gdtb_fv_interpolate: qp_per_element does not match the Gauss rule of the function's declared order");
  double host[2 * MAX_Q1D] = {0};
  gauss_legendre_01(m, host, host + MAX_Q1D);
  double* d_rule = nullptr;
  if (cudaMalloc(&d_rule, sizeof(host)) != cudaSuccess)
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory");
  // on the context's own stream: a blocking copy on the legacy stream is not ordered against a non-blocking stream
  cudaError_t err = cudaMemcpyAsync(d_rule, host, sizeof(host), cudaMemcpyHostToDevice, ctx->launch.stream);
  int st = GDTB_OK;
  if (err == cudaSuccess)
    st = launch_fv_interpolate(ctx->launch, space->grid, to_dev(*f), m, d_rule, d_rule + MAX_Q1D, d_u);
  if (st == GDTB_OK)
    err = cudaStreamSynchronize(ctx->launch.stream);
  cudaFree(d_rule);
  if (st != GDTB_OK)
    return st;
  if (err != cudaSuccess)
    return fail(GDTB_ERR_CUDA, std::string("gdtb_fv_interpolate: ") + cudaGetErrorString(err));
  return GDTB_OK;
}

int gdtb_fv_interpolate_host(gdtb_ctx* ctx, const gdtb_space* space, const gdtb_function* f, double* u)
{
  if (!space || !f || !u)
    return fail(GDTB_ERR_INVALID_ARGUMENT, "gdtb_fv_interpolate_host: NULL argument");
  GDTB_TRY(check_ctx(ctx));
  GDTB_TRY(validate_function(*f, "function"));
  const size_t n = (size_t)space->dev.size;
  double* d_u = nullptr;
  if (cudaMalloc(&d_u, sizeof(double) * std::max<size_t>(n, 1)) != cudaSuccess)
    return fail(GDTB_ERR_OUT_OF_MEMORY, "out of device memory");
  LoweredForm owner; // owns the device clone of per-element host data
  gdtb_function fd = *f;
  int st = lower_function(ctx, space->grid, fd, owner);
  if (st == GDTB_OK)
    st = gdtb_fv_interpolate(ctx, space, &fd, d_u);
  if (st == GDTB_OK && cudaMemcpy(u, d_u, sizeof(double) * n, cudaMemcpyDeviceToHost) != cudaSuccess)
    st = fail(GDTB_ERR_CUDA, "gdtb_fv_interpolate_host: copy to host failed");
  free_form(owner);
  cudaFree(d_u);
  return st;
}

} // extern "C"

// ---- internal helpers shared with solve.cu (declared in handles.hpp) ----------------------------------------------
namespace gdtb {
int internal_check_ctx(gdtb_ctx* ctx)
{
  return check_ctx(ctx);
}
int internal_validate_function(const gdtb_function& f, const char* what)
{
  return validate_function(f, what);
}
int internal_lower_function(gdtb_ctx* ctx, const GridDev& g, gdtb_function& f, LoweredForm& owner)
{
  return lower_function(ctx, g, f, owner);
}
FnDev internal_to_dev(const gdtb_function& f, const LoweredForm* owner)
{
  return to_dev(f, owner);
}
} // namespace gdtb

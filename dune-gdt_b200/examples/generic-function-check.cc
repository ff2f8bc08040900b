// dune-gdt_b200/examples/generic-function-check.cc -- coefficients given as XT::Functions::GenericFunction lambdas
// (local/integrands/laplace.hh:40-48 takes any GridFunction<E, d, d>; examples/stationary-heat-equation.cc:67-70 shows
// the lambda form) against the same coefficients given without a lambda: the facade samples the lambda at the
// quadrature points of the form, the library either consumes those samples (GDTB_FN_QP_*) or samples the analytic
// built-in / constant itself on the device -- both must assemble the same matrix.
//
//   ./generic-function-check [num_elements_per_direction = 24]
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>

#include <dune/gdt/b200.hh>

using namespace Dune;
using namespace Dune::GDT;

template <class M>
double max_diff(const M& a, const M& b, double& scale)
{
  double diff = 0.;
  scale = 0.;
  for (size_t k = 0; k < a.values().size(); ++k) {
    diff = std::max(diff, std::abs(a.values()[k] - b.values()[k]));
    scale = std::max(scale, std::abs(b.values()[k]));
  }
  return diff;
}

template <class G>
bool run(const unsigned int num_elements, const int order)
{
  static const constexpr size_t d = G::dimension;
  using GV = typename G::LeafGridView;
  using E = XT::Grid::extract_entity_t<GV>;
  using M = XT::LA::IstlRowMajorSparseMatrix<double>;

  auto grid = XT::Grid::make_cube_grid<G>(-1., 1., num_elements);
  auto grid_view = grid.leaf_view();
  auto space = make_continuous_lagrange_space(grid_view, order);
  bool ok = true;

  // (i) scalar kappa(x) = 1 + 0.5 |x|^2 of declared order 2: lambda vs the built-in quadratic
  {
    const XT::Functions::GenericFunction<d> kappa(2, [](const auto& x, const auto& /*param*/) {
      double s = 0.;
      for (size_t k = 0; k < d; ++k)
        s += x[k] * x[k];
      return 1. + 0.5 * s;
    });
    gdtb_function builtin{};
    builtin.kind = GDTB_FN_BUILTIN;
    builtin.builtin = GDTB_BUILTIN_QUADRATIC;
    builtin.order = 2;
    builtin.p[0] = 1.;
    builtin.p[1] = 0.5;
    auto lambda_op = make_matrix_operator<M>(space, Stencil::element);
    lambda_op.append(LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(XT::Functions::GridFunction<E, d, d>(kappa))));
    lambda_op.assemble();
    auto builtin_op = make_matrix_operator<M>(space, Stencil::element);
    builtin_op.append(LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(XT::Functions::GridFunction<E, d, d>(builtin))));
    builtin_op.assemble();
    double scale;
    const double diff = max_diff(lambda_op.matrix(), builtin_op.matrix(), scale);
    std::cout << d << "D Q" << order << " scalar kappa lambda vs built-in: max diff " << diff << " (scale " << scale << ")\n";
    ok = ok && diff <= 1e-13 * scale;
  }
  // (ii) a full, non-symmetric constant tensor: matrix-valued lambda vs FieldMatrix
  {
    FieldMatrix<double, int(d), int(d)> K;
    for (size_t r = 0; r < d; ++r)
      for (size_t c = 0; c < d; ++c)
        K[r][c] = (r == c ? 1. + 0.25 * r : 0.) + 0.1 * double(r) - 0.05 * double(c);
    const XT::Functions::GenericFunction<d, d, d> kappa(0, [K](const auto& /*x*/, const auto& /*param*/) { return K; });
    auto lambda_op = make_matrix_operator<M>(space, Stencil::element);
    lambda_op.append(LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(XT::Functions::GridFunction<E, d, d>(kappa))));
    lambda_op.assemble();
    auto const_op = make_matrix_operator<M>(space, Stencil::element);
    const_op.append(LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(XT::Functions::GridFunction<E, d, d>(K))));
    const_op.assemble();
    double scale;
    const double diff = max_diff(lambda_op.matrix(), const_op.matrix(), scale);
    std::cout << d << "D Q" << order << " tensor kappa lambda vs FieldMatrix: max diff " << diff << " (scale " << scale << ")\n";
    ok = ok && diff <= 1e-13 * scale;
  }
  return ok;
}

int main(int argc, char* argv[])
{
  try {
    const unsigned int n = argc > 1 ? std::atoi(argv[1]) : 24;
    bool ok = run<YASP_2D_EQUIDISTANT_OFFSET>(n, 1) && run<YASP_2D_EQUIDISTANT_OFFSET>(n, 2)
              && run<YASP_3D_EQUIDISTANT_OFFSET>(std::max(n / 2, 2u), 1) && run<YASP_3D_EQUIDISTANT_OFFSET>(std::max(n / 3, 2u), 2);
    std::cout << (ok ? "OK" : "FAILED") << std::endl;
    return ok ? EXIT_SUCCESS : EXIT_FAILURE;
  } catch (Exception& e) {
    std::cerr << "\nDUNE reported error: " << e.what() << std::endl;
    return EXIT_FAILURE;
  } catch (std::exception& e) {
    std::cerr << "\nstl reported error: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
}

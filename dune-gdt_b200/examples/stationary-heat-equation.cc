// dune-gdt_b200/examples/stationary-heat-equation.cc -- dune-gdt's examples/stationary-heat-equation.cc (lines 60-127:
// assembly, Dirichlet constraints, solve, error norms) written against the B200 facade: same types, same calls.
// What differs from the reference driver is only (i) the include and (ii) the VTK output, which is out of scope.  In 2D
// the GenericFunction lambda block is the reference's own (examples/stationary-heat-equation.cc:67-85), character for
// character: the facade samples the lambdas at the quadrature points the reference would evaluate them at and hands
// the samples to the library; the 3D run uses the same formulas written for any dimension.  Every assembly / solve /
// norm step runs on the device (there is no CPU path behind the facade).
//
//   ./stationary-heat-equation [num_elements_per_direction = 128] [dim = 2]
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <memory>

#include <dune/gdt/b200.hh>

using namespace Dune;
using namespace Dune::GDT;

template <class G>
int run(const unsigned int num_elements)
{
  static const constexpr size_t d = G::dimension;
  using GV = typename G::LeafGridView;
  using E = XT::Grid::extract_entity_t<GV>;
  using M = XT::LA::IstlRowMajorSparseMatrix<double>;
  using V = XT::LA::IstlDenseVector<double>;

  const double diffusion = 1;
  std::unique_ptr<XT::Functions::GridFunction<E>> source_ptr, exact_ptr;
  if constexpr (d == 2) {
    // ---- examples/stationary-heat-equation.cc:67-85, unchanged --------------------------------------------------
    const XT::Functions::GenericFunction<d> source(3, [](const auto& x, const auto& /*param*/) {
      return M_PI_2 * M_PI * std::cos(M_PI_2 * x[0]) * std::cos(M_PI_2 * x[1]);
    });
    const XT::Functions::GridFunction<E> exact_solution(XT::Functions::GenericFunction<d>(
        3,
        /*evaluate=*/
        [](const auto& x, const auto& /*param*/) { return std::cos(M_PI_2 * x[0]) * std::cos(M_PI_2 * x[1]); },
        /*name=*/"exact_solution",
        /*parameter_type=*/{},
        /*jacobian=*/
        [](const auto& x, const auto& /*param*/) {
          const auto pre = -0.5 * M_PI;
          const auto x_arg = M_PI_2 * x[0];
          const auto y_arg = M_PI_2 * x[1];
          FieldMatrix<double, 1, d> result;
          result[0] = {pre * std::sin(x_arg) * std::cos(y_arg), pre * std::cos(x_arg) * std::sin(y_arg)};
          return result;
        }));
    // --------------------------------------------------------------------------------------------------------------
    source_ptr = std::make_unique<XT::Functions::GridFunction<E>>(source);
    exact_ptr = std::make_unique<XT::Functions::GridFunction<E>>(exact_solution);
  } else {
    // the same problem in any dimension: u = prod_i cos(pi/2 x_i), source = (d pi^2 / 4) u, declared order 3
    source_ptr = std::make_unique<XT::Functions::GridFunction<E>>(
        XT::Functions::GenericFunction<d>(3, [](const auto& x, const auto& /*param*/) {
          double v = d * M_PI_2 * M_PI_2;
          for (size_t k = 0; k < d; ++k)
            v *= std::cos(M_PI_2 * x[k]);
          return v;
        }));
    exact_ptr = std::make_unique<XT::Functions::GridFunction<E>>(XT::Functions::GenericFunction<d>(
        3,
        [](const auto& x, const auto& /*param*/) {
          double v = 1.;
          for (size_t k = 0; k < d; ++k)
            v *= std::cos(M_PI_2 * x[k]);
          return v;
        },
        "exact_solution",
        {},
        [](const auto& x, const auto& /*param*/) {
          FieldMatrix<double, 1, d> result;
          for (size_t r = 0; r < d; ++r) {
            double v = -M_PI_2;
            for (size_t k = 0; k < d; ++k)
              v *= k == r ? std::sin(M_PI_2 * x[k]) : std::cos(M_PI_2 * x[k]);
            result[0][r] = v;
          }
          return result;
        }));
  }
  const XT::Functions::GridFunction<E>& source = *source_ptr;
  const XT::Functions::GridFunction<E>& exact_solution = *exact_ptr;

  auto grid = XT::Grid::make_cube_grid<G>(/*lower_left=*/-1., /*upper_right=*/1., /*num_elements=*/num_elements);
  auto grid_view = grid.leaf_view();

  auto space = make_continuous_lagrange_space(grid_view, /*polorder=*/1);

  auto lhs_op = make_matrix_operator<M>(space, Stencil::element);
  lhs_op.append(LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(diffusion)));

  auto rhs_func = make_vector_functional<V>(space);
  rhs_func.append(LocalElementIntegralFunctional<E>(LocalProductIntegrand<E>().with_ansatz(source)));

  using I = XT::Grid::extract_intersection_t<GV>;
  const XT::Grid::AllDirichletBoundaryInfo<I> boundary_info;
  auto dirichlet_constraints = make_dirichlet_constraints(space, boundary_info);

  auto walker = XT::Grid::make_walker(grid_view);
  walker.append(lhs_op);
  walker.append(rhs_func);
  walker.append(dirichlet_constraints);
  walker.walk(/*thread_parallel=*/true);

  const auto& A = lhs_op.matrix();
  const auto& b = rhs_func.vector();
  // checks that hold for any grid size: constants are in the kernel of the stiffness matrix, the right-hand side
  // sums to the integral of the source over [-1,1]^d = (d pi^2/4) (4/pi)^d
  V ones(A.cols(), 1.), Aones(A.rows(), 0.);
  A.mv(ones, Aones);
  double rhs_sum = 0.;
  for (size_t i = 0; i < b.size(); ++i)
    rhs_sum += b[i];
  const double expected = d * M_PI_2 * M_PI_2 * std::pow(4. / M_PI, double(d));
  std::cout << "dofs: " << space.mapper().size() << "  nnz: " << A.non_zeros() << "\n"
            << "|A 1|_inf = " << Aones.sup_norm() << "\n"
            << "sum(b) = " << rhs_sum << " (integral of the source: " << expected << ")" << std::endl;
  bool ok = Aones.sup_norm() < 1e-12 && std::abs(rhs_sum - expected) < 1e-3 * expected;

  // examples/stationary-heat-equation.cc:108-127: constraints, solve, error against the exact solution
  // u = prod_i cos(pi/2 x_i) (homogeneous Dirichlet values on [-1, 1]^d)
  dirichlet_constraints.apply(lhs_op.matrix(), rhs_func.vector());
  auto solution = make_discrete_function<V>(space);
  auto solver = XT::LA::make_solver(lhs_op.matrix());
  solver.apply(rhs_func.vector(), solution.dofs().vector());

  const auto error = solution - exact_solution;

  auto h1_prod = make_bilinear_form(grid_view, error, error);
  h1_prod += LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(diffusion));

  auto l2_prod = make_bilinear_form(grid_view, error, error);
  l2_prod += LocalElementIntegralBilinearForm<E>(LocalProductIntegrand<E>());

  walker.append(h1_prod);
  walker.append(l2_prod);
  walker.walk(/*thread_parallel=*/true);

  const double h = 2. / num_elements;
  const double h1_error = std::sqrt(h1_prod.result()), l2_error = std::sqrt(l2_prod.result());
  std::cout << "Dirichlet DoFs: " << dirichlet_constraints.dirichlet_DoFs().size() << ", solver iterations: "
            << solver.info().iterations << "\n"
            << "error in H^1 semi-norm: " << h1_error << "\n"
            << "error in L^2 norm:      " << l2_error << std::endl;
  // Q1 on a uniform grid: first order in H^1, second order in L^2
  ok = ok && solver.info().converged && h1_error < 3. * h && l2_error < 3. * h * h;
  std::cout << (ok ? "OK" : "FAILED") << std::endl;
  return ok ? EXIT_SUCCESS : EXIT_FAILURE;
}

int main(int argc, char* argv[])
{
  try {
    const unsigned int n = argc > 1 ? std::atoi(argv[1]) : 128;
    const int dim = argc > 2 ? std::atoi(argv[2]) : 2;
    if (dim == 3)
      return run<YASP_3D_EQUIDISTANT_OFFSET>(n);
    return run<YASP_2D_EQUIDISTANT_OFFSET>(n);
  } catch (Exception& e) {
    std::cerr << "\nDUNE reported error: " << e.what() << std::endl;
    return EXIT_FAILURE;
  } catch (std::exception& e) {
    std::cerr << "\nstl reported error: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
}

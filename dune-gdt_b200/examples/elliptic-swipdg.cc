// dune-gdt_b200/examples/elliptic-swipdg.cc -- the assemble-and-solve step of dune-gdt's
// examples/adaptive_elliptic_swipdg.cc (lines 216-255) on the ESV2007 problem of the reference's own test
// (dune/gdt/test/stationary-heat-equation/ESV2007.hh:58-112, YaspGrid variant), written against the B200 facade:
// element Laplace form + SWIP coupling / penalty forms on the inner faces + Dirichlet coupling / penalty on the boundary
// faces + the force functional, all in ONE grid walk with the matrix operator as the walker, then the solve and the
// broken H^1 semi-norm of the error.  The reference's table (stationary_heat_equation__ESV2007__table_1.mini:31-36,
// cubic grid) lists norm.H_1_semi = 2.52e-01, 1.26e-01, 6.30e-02 for 8^2, 16^2, 32^2 elements.
// Not here: the adaptation loop around it (estimators, marking, refinement are out of scope).
//
//   ./elliptic-swipdg [num_elements_per_direction = 0: the reference's three grids]
#include <cmath>
#include <cstdlib>
#include <iostream>

#include <dune/gdt/b200.hh>

using namespace Dune;
using namespace Dune::GDT;

using G = YASP_2D_EQUIDISTANT_OFFSET;
using GV = typename G::LeafGridView;
using E = XT::Grid::extract_entity_t<GV>;
using I = XT::Grid::extract_intersection_t<GV>;
using M = XT::LA::IstlRowMajorSparseMatrix<double>;
using V = XT::LA::IstlDenseVector<double>;

static double solve_on(const unsigned int num_elements)
{
  // ESV2007DiffusionProblem: diffusion 1, force (pi^2 / 2) cos(pi/2 x) cos(pi/2 y) of order 3, homogeneous Dirichlet
  const double diffusion = 1.;
  const double weight_function = 1.;
  const auto force = XT::Functions::make_cosine_product<E>(3, 0.5 * M_PI * M_PI, M_PI_2);
  const auto exact_solution = XT::Functions::make_cosine_product<E>(4, 1., M_PI_2);
  const double symmetry_prefactor = 1; // SIPDG
  const double penalty_inner = 8, penalty_dirichlet = 14; // ESV2007.hh:108-112, h_I = |I|
  const XT::Grid::AllDirichletBoundaryInfo<I> boundary_info;

  auto grid = XT::Grid::make_cube_grid<G>(-1., 1., num_elements);
  auto grid_view = grid.leaf_view();
  auto dg_space = make_discontinuous_lagrange_space(grid_view, 1);
  auto current_solution = make_discrete_function<V>(dg_space);

  auto lhs_op = make_matrix_operator<M>(dg_space, Stencil::element_and_intersection);
  lhs_op.append(LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(diffusion)));
  lhs_op.append(LocalCouplingIntersectionIntegralBilinearForm<I>(
                    LocalLaplaceIPDGIntegrands::InnerCoupling<I>(symmetry_prefactor, diffusion, weight_function)
                    + LocalIPDGIntegrands::InnerPenalty<I>(penalty_inner, weight_function, IntersectionDiameter::volume)),
                {},
                XT::Grid::ApplyOn::InnerIntersectionsOnce<GV>());
  lhs_op.append(LocalIntersectionIntegralBilinearForm<I>(
                    LocalIPDGIntegrands::BoundaryPenalty<I>(penalty_dirichlet, weight_function, IntersectionDiameter::volume)
                    + LocalLaplaceIPDGIntegrands::DirichletCoupling<I>(symmetry_prefactor, diffusion)),
                {},
                XT::Grid::ApplyOn::CustomBoundaryIntersections<GV>(boundary_info, new XT::Grid::DirichletBoundary()));
  auto rhs_func = make_vector_functional<V>(dg_space);
  rhs_func.append(LocalElementIntegralFunctional<E>(LocalProductIntegrand<E>().with_ansatz(force)));
  // assemble everything in one grid walk (uses the lhs_op as grid walker)
  lhs_op.append(rhs_func);
  lhs_op.assemble(true);

  auto solver = XT::LA::make_solver(lhs_op.matrix());
  solver.apply(rhs_func.vector(), current_solution.dofs().vector(),
               gdtb_solver_opts{GDTB_SOLVER_CG, GDTB_PRECOND_JACOBI, 0, 0, 1e-12});

  const auto error = current_solution - exact_solution;
  auto h1_prod = make_bilinear_form(grid_view, error, error);
  h1_prod += LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(diffusion));
  h1_prod.assemble();
  const double h1 = std::sqrt(h1_prod.result());
  std::cout << num_elements << "^2 elements, " << dg_space.mapper().size() << " DoFs, " << lhs_op.matrix().non_zeros()
            << " nnz, " << solver.info().iterations << " CG iterations: broken H^1 semi-norm of the error " << h1
            << std::endl;
  return h1;
}

int main(int argc, char* argv[])
{
  try {
    const unsigned int n = argc > 1 ? std::atoi(argv[1]) : 0;
    bool ok = true;
    if (n > 0)
      ok = solve_on(n) < 4. / n; // first order: |e|_{H^1} ~ 2 h
    else {
      const unsigned int sizes[3] = {8, 16, 32};
      const double reference[3] = {2.52e-01, 1.26e-01, 6.30e-02};
      for (int k = 0; k < 3; ++k)
        ok = ok && std::abs(solve_on(sizes[k]) - reference[k]) < 6e-3 * reference[k];
    }
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

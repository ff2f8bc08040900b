// dune-gdt_b200/examples/stationary-heat-equation.cc -- the assembly part of dune-gdt's
// examples/stationary-heat-equation.cc (lines 60-106) written against the B200 facade: same types, same calls.
// What differs from the reference driver is only (i) the include, (ii) the GenericFunction lambda, which cannot
// cross the C ABI as code and is replaced by the built-in analytic source of the same formula, and (iii) the steps
// after the walk (constraints, solve, norms), which SURVEY.md section 8(f) lists as "next" rows.
//
//   ./stationary-heat-equation [num_elements_per_direction = 128] [dim = 2]
#include <cmath>
#include <cstdlib>
#include <iostream>

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
  // source(x) = (d pi^2 / 4) prod_i cos(pi/2 x_i), declared order 3 (examples/stationary-heat-equation.cc:68-70)
  const auto source = XT::Functions::make_cosine_product<E>(3, d * M_PI_2 * M_PI_2, M_PI_2);

  auto grid = XT::Grid::make_cube_grid<G>(/*lower_left=*/-1., /*upper_right=*/1., /*num_elements=*/num_elements);
  auto grid_view = grid.leaf_view();

  auto space = make_continuous_lagrange_space(grid_view, /*polorder=*/1);

  auto lhs_op = make_matrix_operator<M>(space, Stencil::element);
  lhs_op.append(LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(diffusion)));

  auto rhs_func = make_vector_functional<V>(space);
  rhs_func.append(LocalElementIntegralFunctional<E>(LocalProductIntegrand<E>().with_ansatz(source)));

  auto walker = XT::Grid::make_walker(grid_view);
  walker.append(lhs_op);
  walker.append(rhs_func);
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
  const bool ok = Aones.sup_norm() < 1e-12 && std::abs(rhs_sum - expected) < 1e-3 * expected;
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

// dune-gdt_b200/examples/euler-2d.cc -- the 2d_euler driver of dune-gdt's
// examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:381-434 written against the B200 facade: the Euler equations
// (m = 4) on a periodic 2D YaspGrid, finite volume space with 4 components, Vijayasundaram flux with EulerTools'
// eigendecomposition, initial values 4 / 1.6 inside [-0.5, 0]^2 and 1 / 0.4 outside, dt from
// estimate_dt_for_hyperbolic_system, explicit Euler up to T_end.  The flux / jacobian / eigendecomposition lambdas of the
// reference are EulerTools calls; here EulerTools itself crosses the C ABI as a tag (GDTB_FLUX_EULER).
//
//   ./euler-2d [num_elements = 128] [T_end = 1]
#include <cmath>
#include <cstdlib>
#include <iostream>

#include <dune/gdt/b200.hh>

using namespace Dune;
using namespace Dune::GDT;

using V = XT::LA::IstlDenseVector<double>;
using M = XT::LA::IstlRowMajorSparseMatrix<double>;

int main(int argc, char* argv[])
{
  try {
    const unsigned int N = argc > 1 ? std::atoi(argv[1]) : 128;
    const double T_end = argc > 2 ? std::atof(argv[2]) : 1.;
    using G = YASP_2D_EQUIDISTANT_OFFSET;
    static const size_t d = G::dimension;
    using DomainType = FieldVector<double, d>;
    static const size_t m = EulerTools<d>::m;
    const EulerTools<d> euler_tools(/*gamma=*/1.4);
    const auto f = make_euler_flux(euler_tools);

    auto grid = XT::Grid::make_cube_grid<G>(-1., 1., N);
    auto grid_view = XT::Grid::make_periodic_grid_view(grid.leaf_view());
    using GV = decltype(grid_view);
    using I = XT::Grid::extract_intersection_t<GV>;

    auto V_h_0 = make_finite_volume_space<m>(grid_view);

    const NumericalVijayasundaramFlux<I, d, m> g(f);
    auto L_h = make_advection_fv_operator<M>(grid_view, g, V_h_0, V_h_0);

    auto w_0 = default_interpolation<V>(
        0,
        [&](const DomainType& xx, const auto& /*mu*/) {
          bool inside = true;
          for (size_t k = 0; k < d; ++k)
            inside = inside && xx[k] >= -0.5 && xx[k] <= 0.;
          if (inside)
            return euler_tools.conservative(/*density=*/4., /*velocity=*/0., /*pressure=*/1.6);
          else
            return euler_tools.conservative(/*density=*/1., /*velocity=*/0., /*pressure=*/0.4);
        },
        V_h_0);

    const double dt = estimate_dt_for_hyperbolic_system(L_h, w_0);
    auto w_h = explicit_euler(w_0, L_h, T_end, dt);

    // what can be checked without the reference at hand: conservation of mass / momentum / energy on the periodic grid,
    // positivity of density and pressure, and the x <-> y symmetry of the problem
    double total0[m] = {}, total1[m] = {}, rho_min = 1e300, p_min = 1e300, asym = 0.;
    for (size_t e = 0; e < size_t(N) * N; ++e) {
      FieldVector<double, int(m)> w{};
      for (size_t c = 0; c < m; ++c) {
        total0[c] += w_0[e * m + c];
        total1[c] += w_h[e * m + c];
        w[c] = w_h[e * m + c];
      }
      rho_min = std::min(rho_min, euler_tools.density(w));
      p_min = std::min(p_min, euler_tools.pressure(w));
      const size_t ix = e % N, iy = e / N, et = iy + ix * N; // the transposed cell
      asym = std::max(asym, std::abs(w_h[e * m] - w_h[et * m]));
      asym = std::max(asym, std::abs(w_h[e * m + 1] - w_h[et * m + 2]));
    }
    std::cout << "2d euler, " << N << "^2 elements, dt = " << dt << ", T_end = " << T_end << "\n  mass " << total0[0] << " -> "
              << total1[0] << ", energy " << total0[3] << " -> " << total1[3] << "\n  min density " << rho_min
              << ", min pressure " << p_min << ", x/y asymmetry " << asym << std::endl;
    bool ok = rho_min > 0.9 && p_min > 0.3 && asym < 1e-9;
    for (size_t c = 0; c < m; ++c)
      ok = ok && std::abs(total1[c] - total0[c]) < 1e-9 * double(N) * N;
    std::cout << (ok ? "OK" : "FAILED") << std::endl;
    return ok ? EXIT_SUCCESS : EXIT_FAILURE;
  } catch (Exception& e) {
    std::cerr << "\nDUNE reported error: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
}

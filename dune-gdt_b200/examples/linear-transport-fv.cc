// dune-gdt_b200/examples/linear-transport-fv.cc -- the explicit first-order FV drivers of dune-gdt's
// examples/mpi_2019_02_talk_on_hyperbolic_equations.cc (linear_transport :255-298, burgers :300-337) written against
// the B200 facade.  The flux / initial-value lambdas of the reference cannot cross the C ABI as code; they are the
// built-in LinearFlux / BurgersFlux tags and the analytic indicator / Gaussian functions of the same formulas.
//
//   ./linear-transport-fv [num_elements = 1024]
#include <cmath>
#include <cstdlib>
#include <iostream>

#include <dune/gdt/b200.hh>

using namespace Dune;
using namespace Dune::GDT;

using M = XT::LA::IstlRowMajorSparseMatrix<double>;
using V = XT::LA::IstlDenseVector<double>;

static double mass(const V& u)
{
  double s = 0.;
  for (size_t i = 0; i < u.size(); ++i)
    s += u[i];
  return s;
}

int main(int argc, char* argv[])
{
  try {
    const unsigned int N = argc > 1 ? std::atoi(argv[1]) : 1024;
    using G = YASP_1D_EQUIDISTANT_OFFSET;
    static const size_t d = G::dimension;
    auto grid = XT::Grid::make_cube_grid<G>(0., 1., N);
    auto grid_view = XT::Grid::make_periodic_grid_view(grid.leaf_view());
    using GV = decltype(grid_view);
    using E = XT::Grid::extract_entity_t<GV>;
    using I = XT::Grid::extract_intersection_t<GV>;

    auto V_h_0 = make_finite_volume_space(grid_view);
    bool ok = true;

    { // linear transport to the right, u_0 = indicator of [1/4, 1/2]; dt = h makes the upwind scheme an exact shift
      const NumericalUpwindFlux<I, d, 1> g(LinearFlux{});
      auto L_h = make_advection_fv_operator<M>(grid_view, g, V_h_0, V_h_0);
      auto w_0 = default_interpolation<V>(XT::Functions::make_indicator<E>(0, 0.25, 0.5), V_h_0);
      const double T_end = 1.;
      const double dt = 1. / N;
      auto w_h = explicit_euler(w_0, L_h, T_end, dt);
      // the loop `while (time < T_end + dt)` takes N + 1 or N + 2 steps: the profile comes back shifted by 1-2 cells
      double best = 1e300;
      for (size_t shift = 0; shift < 4; ++shift) {
        double err = 0.;
        for (size_t i = 0; i < w_0.size(); ++i)
          err = std::max(err, std::abs(w_h[(i + shift) % w_0.size()] - w_0[i]));
        best = std::min(best, err);
      }
      std::cout << "linear transport: mass " << mass(w_0) / N << " -> " << mass(w_h) / N << ", shift error " << best
                << std::endl;
      ok = ok && best < 1e-12 && std::abs(mass(w_h) - mass(w_0)) < 1e-9;
    }
    { // Burgers, Gaussian initial values (order 3 => two Gauss points per cell average)
      const NumericalUpwindFlux<I, d, 1> g(BurgersFlux{});
      auto L_h = make_advection_fv_operator<M>(grid_view, g, V_h_0, V_h_0);
      auto w_0 = default_interpolation<V>(XT::Functions::make_gaussian<E>(3, 0.33, 0.075), V_h_0);
      auto w_h = explicit_euler(w_0, L_h, 0.5, 0.5 / N);
      std::cout << "burgers: mass " << mass(w_0) / N << " -> " << mass(w_h) / N << ", max " << w_h.sup_norm()
                << std::endl;
      ok = ok && std::abs(mass(w_h) - mass(w_0)) < 1e-9 * N && w_h.sup_norm() <= 1. + 1e-12;
    }
    { // the same Burgers problem through the reference's time stepper (tools/timestepper/explicit-rungekutta.hh) with
      // dt from estimate_dt_for_hyperbolic_system (tools/hyperbolic.hh), third order SSP Runge-Kutta, u_t = -L(u)
      const NumericalLaxFriedrichsFlux<I, d, 1> g(BurgersFlux{});
      auto L_h = make_advection_fv_operator<M>(grid_view, g, V_h_0, V_h_0);
      using Op = decltype(L_h);
      auto w = default_interpolation<V>(XT::Functions::make_gaussian<E>(3, 0.33, 0.075), V_h_0);
      const double m0 = mass(w), max0 = w.sup_norm();
      const double dt = estimate_dt_for_hyperbolic_system(L_h, w);
      ExplicitRungeKuttaTimeStepper<Op, TimeStepperMethods::explicit_rungekutta_third_order_ssp> stepper(L_h, w, -1.);
      stepper.solve(0.5, dt);
      std::cout << "burgers (SSP3, dt = " << dt << ", " << stepper.num_steps() << " steps): t = " << stepper.current_time()
                << ", mass " << m0 / N << " -> " << mass(w) / N << ", max " << w.sup_norm() << std::endl;
      ok = ok && std::abs(stepper.current_time() - 0.5) < 1e-12 && std::abs(mass(w) - m0) < 1e-9 * N
           && w.sup_norm() <= max0 + 1e-12;
    }
    { // non-periodic: inflow value 0 on the left by extrapolation, outflow of the physical flux on the right
      auto plain_view = grid.leaf_view();
      auto V_np = make_finite_volume_space(plain_view);
      const NumericalUpwindFlux<I, d, 1> g(LinearFlux{});
      auto L_h = make_advection_fv_operator<M>(plain_view, g, V_np, V_np);
      L_h.append(BoundaryTreatmentByCustomExtrapolation{0., 0.}, 0b01).append(BoundaryTreatmentByCustomNumericalFlux{1., 0.}, 0b10);
      using Op = decltype(L_h);
      auto w = default_interpolation<V>(XT::Functions::make_indicator<E>(0, 0.25, 0.5), V_np);
      ExplicitRungeKuttaTimeStepper<Op> stepper(L_h, w, -1.);
      stepper.solve(1., 1. / N); // exact shift: after t = 1 the indicator has left through the right boundary
      std::cout << "linear transport with in/outflow boundaries: remaining mass " << mass(w) / N << std::endl;
      ok = ok && std::abs(mass(w)) < 1e-12 * N;
    }
    std::cout << (ok ? "OK" : "FAILED") << std::endl;
    return ok ? EXIT_SUCCESS : EXIT_FAILURE;
  } catch (Exception& e) {
    std::cerr << "\nDUNE reported error: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
}

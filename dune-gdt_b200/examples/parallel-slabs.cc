// dune-gdt_b200/examples/parallel-slabs.cc -- the multi-GPU entry points of the C++ facade, one process per GPU:
//   tools/mprun.sh 2 dune-gdt_b200/examples/parallel-slabs
// (i)   3D CG Q1 Laplace + right-hand side on z-slabs (Parallel::SlabAssembler, ghost layer recomputed) and with the
//       interface-row halo inside the gather kernel (Parallel::HaloSlabAssembler), both compared with the rows of the
//       unpartitioned assembly of the same rank;
// (ii)  2D linear transport, SSP3 Runge-Kutta on slabs (Parallel::PeerMemoryRungeKuttaTimeStepper, the halo site of
//       tools/timestepper/explicit-rungekutta.hh:252-257) and the explicit Euler loop of
//       examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:141-159 (Parallel::PeerMemoryEulerTimeLoop), both compared
//       with the single-GPU steppers.
// Exit code 0 = every comparison within 1e-12 (relative to the largest entry).
#include <cmath>
#include <cstdio>

#include <dune/gdt/b200-parallel.hh>

using namespace Dune;
using namespace Dune::GDT;

using V = XT::LA::IstlDenseVector<double>;
using M = XT::LA::IstlRowMajorSparseMatrix<double>;

static double rel_diff(const double* a, const double* b, std::size_t n)
{
  double d = 0., m = 0.;
  for (std::size_t i = 0; i < n; ++i) {
    d = std::max(d, std::abs(a[i] - b[i]));
    m = std::max(m, std::abs(b[i]));
  }
  return m > 0. ? d / m : d;
}

int main()
{
  auto comm = Parallel::FileRendezvous::from_environment();
  Parallel::bind_device(comm);
  int failures = 0;
  auto report = [&](const char* what, double err) {
    std::printf("[rank %d/%d] %-58s max rel diff %.3e %s\n", comm.rank(), comm.size(), what, err, err <= 1e-12 ? "ok" : "FAILED");
    if (!(err <= 1e-12))
      ++failures;
  };

  { // ---- (i) assembly --------------------------------------------------------------------------------------------
    using G = YASP_3D_EQUIDISTANT_OFFSET;
    auto grid = XT::Grid::make_cube_grid<G>(-1., 1., 24);
    auto grid_view = grid.leaf_view();
    using GV = decltype(grid_view);
    using E = XT::Grid::extract_entity_t<GV>;
    auto space = make_continuous_lagrange_space(grid_view, 1);
    const auto f = XT::Functions::make_cosine_product<E>(3, 0.75 * M_PI * M_PI, 0.5 * M_PI);

    // the unpartitioned assembly (every rank computes it for the comparison)
    auto op = make_matrix_operator<M>(space, Stencil::element);
    op.append(LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(1.)));
    auto rhs = make_vector_functional<V>(space);
    rhs.append(LocalElementIntegralFunctional<E>(LocalElementProductIntegrand<E>().with_ansatz(f)));
    op.append(rhs);
    op.assemble();
    const auto& A = op.matrix().values();
    const auto& b = rhs.vector();

    Parallel::SlabAssembler<GV> slab(space, comm);
    slab.append(LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(1.)));
    slab.append(LocalElementIntegralFunctional<E>(LocalElementProductIntegrand<E>().with_ansatz(f)));
    slab.assemble();
    {
      double err_a = 0., err_b = 0.;
      std::size_t at_v = 0, at_r = 0;
      for (const auto& r : slab.row_ranges()) {
        err_a = std::max(err_a, rel_diff(slab.values().data() + at_v, A.data() + r.value_offset, std::size_t(r.count)));
        err_b = std::max(err_b, rel_diff(slab.vector().data() + at_r, b.data() + r.row_begin, std::size_t(r.row_end - r.row_begin)));
        at_v += std::size_t(r.count);
        at_r += std::size_t(r.row_end - r.row_begin);
      }
      report("SlabAssembler: CG Q1 Laplace values (owned rows)", err_a);
      report("SlabAssembler: right-hand side (owned rows)", err_b);
    }

    Parallel::HaloSlabAssembler<GV> halo(space, comm);
    halo.append(LocalElementIntegralBilinearForm<E>(LocalLaplaceIntegrand<E>(1.)));
    halo.append(LocalElementIntegralFunctional<E>(LocalElementProductIntegrand<E>().with_ansatz(f)));
    for (int rep = 0; rep < 3; ++rep) // both step parities of the receive buffers
      halo.assemble();
    {
      const auto& r = halo.row_ranges()[0];
      report("HaloSlabAssembler: CG Q1 Laplace values (owned rows)",
             rel_diff(halo.values().data(), A.data() + r.value_offset, std::size_t(r.count)));
      report("HaloSlabAssembler: right-hand side (owned rows)",
             rel_diff(halo.vector().data(), b.data() + r.row_begin, std::size_t(r.row_end - r.row_begin)));
    }
  }

  { // ---- (ii) explicit time stepping on slabs ---------------------------------------------------------------------
    using G = YASP_2D_EQUIDISTANT_OFFSET;
    auto grid = XT::Grid::make_cube_grid<G>(0., 1., 96);
    auto grid_view = XT::Grid::make_periodic_grid_view(grid.leaf_view());
    using GV = decltype(grid_view);
    using E = XT::Grid::extract_entity_t<GV>;
    using I = XT::Grid::extract_intersection_t<GV>;
    auto V_h = make_finite_volume_space(grid_view);
    const NumericalUpwindFlux<I, 2> g(LinearFlux{{1.0, 0.5}});
    const auto w_0 = default_interpolation<V>(XT::Functions::make_gaussian<E>(3, 0.33, 0.075), V_h);
    const double dt = 0.25 / 96, T_end = 40 * dt;

    // single GPU
    auto L_ref = make_advection_fv_operator<M>(grid_view, g, V_h, V_h);
    V u_ref(w_0);
    ExplicitRungeKuttaTimeStepper<decltype(L_ref), TimeStepperMethods::explicit_rungekutta_third_order_ssp> ts_ref(L_ref, u_ref, -1.);
    ts_ref.solve(T_end, dt);
    const V u_euler_ref = explicit_euler(w_0, L_ref, T_end, dt);

    {
      auto L_h = make_advection_fv_operator<M>(grid_view, g, V_h, V_h);
      Parallel::PeerMemoryRungeKuttaTimeStepper<M, GV> ts(L_h, comm, -1.);
      ts.set_initial_values(w_0);
      ts.solve(T_end, dt);
      const V mine = ts.owned_solution();
      report("PeerMemoryRungeKuttaTimeStepper (SSP3): owned cells",
             rel_diff(mine.data(), u_ref.data() + ts.layer_begin() * ts.cells_per_layer(), mine.size()));
      if (ts.num_steps() != ts_ref.num_steps())
        report("PeerMemoryRungeKuttaTimeStepper: number of steps", 1.);
    }
    {
      auto L_h = make_advection_fv_operator<M>(grid_view, g, V_h, V_h);
      Parallel::PeerMemoryEulerTimeLoop<M, GV> loop(L_h, comm);
      loop.set_initial_values(w_0);
      std::int64_t steps = 0;
      for (double time = 0.; time < T_end + dt; time += dt)
        ++steps;
      loop.euler_steps(dt, steps);
      const V mine = loop.owned_solution();
      report("PeerMemoryEulerTimeLoop: owned cells",
             rel_diff(mine.data(), u_euler_ref.data() + loop.layer_begin() * loop.cells_per_layer(), mine.size()));
    }
  }
  comm.barrier();
  if (failures == 0 && comm.rank() == 0)
    std::printf("OK\n");
  return failures == 0 ? 0 : 1;
}

// dune-gdt_b200/include/dune/gdt/b200-parallel.hh -- multi-GPU host side of the C++ facade (one process per GPU).
//
// The reference's parallel sites on this path are C++: the grid view's CollectiveCommunication / DataHandle
// communicate() after every Runge-Kutta stage (tools/timestepper/explicit-rungekutta.hh:252-257) and YaspGrid's
// overlap for the assembly.  Here:
//   * Parallel::SlabAssembler           owner-computes-rows assembly of one rank's element slab, ghost layer recomputed,
//                                       no communication (gdtb_matop_set_slab / gdtb_vecfun_set_slab);
//   * Parallel::HaloSlabAssembler       the interface-row halo partition with the hand-over inside the gather kernel
//                                       (gdtb_matop_set_slab_halo + gdtb_halo_p2p_*; CG Q1);
//   * Parallel::PeerMemoryRungeKuttaTimeStepper  ExplicitRungeKuttaTimeStepper on slabs: stage vectors handed over by
//                                       peer stores, applies wait in-kernel (gdtb_rk_p2p_*);
//   * Parallel::PeerMemoryEulerTimeLoop the explicit Euler loop of examples/mpi_2019_02_talk_on_hyperbolic_equations.cc
//                                       with the ghost exchange fused into the apply kernel (gdtb_fvop_p2p_*).
// The only thing the ranks exchange on the host are CUDA IPC handles (64 bytes each), once, through a
// Parallel::CollectiveCommunication: the interface below is the subset of Dune::CollectiveCommunication the reference
// uses (rank / size / barrier / allgather); a binder on an MPI build forwards it to grid_view.comm(), the
// FileRendezvous implementation needs nothing but a directory all ranks see (no MPI in this image).
#ifndef DUNE_GDT_B200_PARALLEL_HH
#define DUNE_GDT_B200_PARALLEL_HH

#include <dune/gdt/b200.hh>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <thread>
#include <unistd.h>

namespace Dune {
namespace GDT {
namespace Parallel {


class CollectiveCommunication
{
public:
  virtual ~CollectiveCommunication() = default;
  virtual int rank() const = 0;
  virtual int size() const = 0;
  virtual void barrier() = 0;
  // every rank contributes `bytes` bytes; recv holds size() * bytes, rank r's block at r * bytes
  virtual void allgather(const void* send, std::size_t bytes, void* recv) = 0;
};


// Rendezvous through files in a directory every rank sees (one node: /dev/shm/...): rank r publishes the block of
// collective number s as <dir>/<s>.<r> (written under a temporary name, then renamed: readers never see a partial
// file) and reads the others' as they appear.
class FileRendezvous : public CollectiveCommunication
{
public:
  FileRendezvous(const int rank, const int size, std::string directory, const double timeout_seconds = 120.)
    : rank_(rank)
    , size_(size)
    , dir_(std::move(directory))
    , timeout_(timeout_seconds)
  {
    if (rank < 0 || size < 1 || rank >= size)
      throw XT::Common::Exceptions::wrong_input_given("FileRendezvous: 0 <= rank < size required");
  }

  // RANK / WORLD_SIZE as torchrun, mpirun wrappers and tools/mprun.sh set them; GDTB_RENDEZVOUS_DIR names the directory
  static FileRendezvous from_environment()
  {
    const char* r = std::getenv("RANK");
    const char* w = std::getenv("WORLD_SIZE");
    const char* d = std::getenv("GDTB_RENDEZVOUS_DIR");
    const int size = w ? std::atoi(w) : 1;
    if (size > 1 && !d)
      throw XT::Common::Exceptions::wrong_input_given("GDTB_RENDEZVOUS_DIR must name a directory shared by all ranks");
    return FileRendezvous(r ? std::atoi(r) : 0, size, d ? d : "/tmp");
  }

  int rank() const override
  {
    return rank_;
  }
  int size() const override
  {
    return size_;
  }
  void barrier() override
  {
    char token = 1;
    std::vector<char> all(std::size_t(size_), 0);
    allgather(&token, 1, all.data());
  }
  void allgather(const void* send, const std::size_t bytes, void* recv) override
  {
    char* out = static_cast<char*>(recv);
    if (size_ == 1) {
      std::copy(static_cast<const char*>(send), static_cast<const char*>(send) + bytes, out);
      return;
    }
    const long seq = seq_++;
    const std::string mine = path(seq, rank_);
    {
      std::ofstream f(mine + ".tmp", std::ios::binary);
      f.write(static_cast<const char*>(send), std::streamsize(bytes));
    }
    if (std::rename((mine + ".tmp").c_str(), mine.c_str()) != 0)
      throw Dune::InvalidStateException("FileRendezvous: cannot publish " + mine);
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < size_; ++r) {
      for (;;) {
        std::ifstream f(path(seq, r), std::ios::binary);
        if (f && f.read(out + std::size_t(r) * bytes, std::streamsize(bytes)))
          break;
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_)
          throw Dune::InvalidStateException("FileRendezvous: rank " + std::to_string(r) + " did not arrive");
        std::this_thread::sleep_for(std::chrono::microseconds(200));
      }
    }
    // everybody who reaches collective s has read all files of s - 1, hence nobody still needs those of s - 2
    if (seq >= 2)
      std::remove(path(seq - 2, rank_).c_str());
  }

private:
  std::string path(const long seq, const int r) const
  {
    return dir_ + "/gdtb_rv." + std::to_string(seq) + "." + std::to_string(r);
  }
  int rank_, size_;
  std::string dir_;
  double timeout_;
  long seq_ = 0;
};


// Binds this process to its GPU (LOCAL_RANK, else the rank): must run before the first facade object is created
inline void bind_device(const CollectiveCommunication& comm)
{
  const char* l = std::getenv("LOCAL_RANK");
  GDT::internal::context(l ? std::atoi(l) : comm.rank());
}

// element layers [begin, end) of `rank` along the last direction: as even as possible, the first ranks one more
inline std::pair<std::int64_t, std::int64_t> slab_layers(const std::int64_t n_last, const int rank, const int size)
{
  const std::int64_t base = n_last / size, extra = n_last % size;
  const std::int64_t begin = rank * base + std::min<std::int64_t>(rank, extra);
  return {begin, begin + base + (rank < extra ? 1 : 0)};
}

namespace internal {

// lower / upper neighbour of a slab (periodic wrap along the last direction), -1: none
inline std::pair<int, int> neighbours(const int rank, const int size, const bool periodic_last)
{
  const int lower = rank > 0 ? rank - 1 : (periodic_last ? size - 1 : -1);
  const int upper = rank + 1 < size ? rank + 1 : (periodic_last ? 0 : -1);
  return {lower, upper};
}

template <class GV>
std::int64_t n_last(const SpaceInterface<GV>& space)
{
  const gdtb_grid_desc& g = space.grid_view().desc();
  return g.n[g.dim - 1];
}

template <class GV>
bool periodic_last(const SpaceInterface<GV>& space)
{
  const gdtb_grid_desc& g = space.grid_view().desc();
  return ((g.periodic >> (g.dim - 1)) & 1) && g.n[g.dim - 1] > 1;
}

// exchanges one block of IPC handles per rank and hands the neighbours' blocks to `connect`
template <std::size_t BYTES, class Connect>
void connect_neighbours(CollectiveCommunication& comm, const std::int64_t n_last, const bool periodic,
                        const unsigned char (&mine)[BYTES], Connect&& connect)
{
  std::vector<unsigned char> all(std::size_t(comm.size()) * BYTES);
  comm.allgather(mine, BYTES, all.data());
  const auto nb = neighbours(comm.rank(), comm.size(), periodic);
  auto layers = [&](const int r) {
    const auto s = slab_layers(n_last, r, comm.size());
    return s.second - s.first;
  };
  const bool lo_self = nb.first == comm.rank(), hi_self = nb.second == comm.rank();
  connect(nb.first < 0 || lo_self ? nullptr : all.data() + std::size_t(nb.first) * BYTES, nb.first < 0 || lo_self ? 0 : layers(nb.first),
          lo_self ? 1 : 0, nb.second < 0 || hi_self ? nullptr : all.data() + std::size_t(nb.second) * BYTES,
          nb.second < 0 || hi_self ? 0 : layers(nb.second), hi_self ? 1 : 0);
  comm.barrier(); // everybody is connected before anybody computes
}

} // namespace internal


// global row range [row_begin, row_end) of a slab and where its values sit in the global CSR array
struct RowRange
{
  std::int64_t row_begin, row_end, value_offset, count;
};


// MatrixOperator (+ VectorBasedFunctional riding along) of ONE rank on the rank's element slab: owned rows are complete
// because the ghost element layer is walked redundantly (YaspGrid overlap in the reference), so there is no exchange.
// CG Q1 / Q2 and DG spaces; the global CSR matrix is the concatenation of the ranks' row ranges (row_ranges()).
template <class GV>
class SlabAssembler
{
public:
  using E = typename GV::Element;
  using I = typename GV::Intersection;

  SlabAssembler(const SpaceInterface<GV>& space, CollectiveCommunication& comm, const bool with_functional = true)
    : space_(space)
  {
    const auto s = slab_layers(internal::n_last(space), comm.rank(), comm.size());
    begin_ = s.first;
    end_ = s.second;
    gdtb_matop* op = nullptr;
    GDT::internal::check(gdtb_matop_create(GDT::internal::context(), space_.handle(), space_.handle(), nullptr, &op));
    op_ = GDT::internal::Handle<gdtb_matop, gdtb_matop_destroy>(op);
    GDT::internal::check(gdtb_matop_set_slab(op, begin_, end_));
    if (with_functional) {
      gdtb_vecfun* fun = nullptr;
      GDT::internal::check(gdtb_vecfun_create(GDT::internal::context(), space_.handle(), &fun));
      fun_ = GDT::internal::Handle<gdtb_vecfun, gdtb_vecfun_destroy>(fun);
      GDT::internal::check(gdtb_vecfun_set_slab(fun, begin_, end_));
    }
    std::int32_t n = 0;
    GDT::internal::check(gdtb_matop_local_row_ranges(op, 0, nullptr, nullptr, nullptr, nullptr, &n));
    std::vector<std::int64_t> rb(n), re(n), vo(n), cnt(n);
    GDT::internal::check(gdtb_matop_local_row_ranges(op, n, rb.data(), re.data(), vo.data(), cnt.data(), &n));
    std::int64_t rows = 0;
    for (std::int32_t i = 0; i < n; ++i) {
      ranges_.push_back({rb[i], re[i], vo[i], cnt[i]});
      rows += re[i] - rb[i];
    }
    values_.resize(std::size_t(gdtb_matop_local_nnz(op)));
    if (with_functional)
      vector_.resize(std::size_t(rows));
  }

  SlabAssembler& append(const LocalElementBilinearFormInterface<E>& form)
  {
    const auto f = form.lowered(1., space_.handle(), space_.grid_view().desc());
    GDT::internal::check(gdtb_matop_append_element(op_.get(), &f.form));
    return *this;
  }
  SlabAssembler& append(const LocalCouplingIntersectionBilinearFormInterface<I>& form,
                        const XT::Grid::IntersectionFilter<GV>& filter = XT::Grid::ApplyOn::InnerIntersectionsOnce<GV>())
  {
    const auto f = form.lowered(1., space_.handle(), space_.grid_view().desc());
    GDT::internal::check(gdtb_matop_append_coupling(op_.get(), &f.form, filter.gdtb_filter()));
    return *this;
  }
  SlabAssembler& append(const LocalIntersectionBilinearFormInterface<I>& form)
  {
    const auto f = form.lowered(1., space_.handle(), space_.grid_view().desc());
    GDT::internal::check(gdtb_matop_append_boundary(op_.get(), &f.form, GDTB_FILTER_ALL_BOUNDARY));
    return *this;
  }
  SlabAssembler& append(const LocalElementFunctionalInterface<E>& functional)
  {
    if (!fun_.get())
      throw Dune::InvalidStateException("this SlabAssembler was created without a functional");
    const auto f = functional.lowered(space_.handle(), space_.grid_view().desc());
    GDT::internal::check(gdtb_vecfun_append_element(fun_.get(), &f.form));
    return *this;
  }

  // one grid walk over the slab; values() / vector() hold the owned rows afterwards
  void assemble()
  {
    GDT::internal::check(gdtb_assemble_host(op_.get(), fun_.get(), values_.data(), fun_.get() ? vector_.data() : nullptr));
  }
  // the same without the device-to-host copy (values stay on the GPU: device_values())
  void assemble_on_device()
  {
    GDT::internal::check(gdtb_assemble(op_.get(), fun_.get(), GDTB_ASSEMBLE_OVERWRITE));
  }

  const std::vector<double>& values() const
  {
    return values_;
  }
  const std::vector<double>& vector() const
  {
    return vector_;
  }
  const std::vector<RowRange>& row_ranges() const
  {
    return ranges_;
  }
  std::int64_t layer_begin() const
  {
    return begin_;
  }
  std::int64_t layer_end() const
  {
    return end_;
  }
  double* device_values() const
  {
    double* p = nullptr;
    GDT::internal::check(gdtb_matop_values_device(op_.get(), &p));
    return p;
  }
  gdtb_matop* handle() const
  {
    return op_.get();
  }

private:
  SpaceInterface<GV> space_;
  std::int64_t begin_ = 0, end_ = 0;
  GDT::internal::Handle<gdtb_matop, gdtb_matop_destroy> op_;
  GDT::internal::Handle<gdtb_vecfun, gdtb_vecfun_destroy> fun_;
  std::vector<RowRange> ranges_;
  std::vector<double> values_, vector_;
};


// ExplicitRungeKuttaTimeStepper (tools/timestepper/explicit-rungekutta.hh:158-270) on slabs of a finite volume space.
// The stepper owns the solution vector ([ghost | owned layers | ghost]) and two alternating stage vectors in memory the
// neighbour ranks open through CUDA IPC; the hand-over of the boundary layers after every stage (the reference's
// communicate(), :252-257) happens on the device.  >= 2 stages (use PeerMemoryEulerTimeLoop for explicit Euler).
template <class M, class GV, TimeStepperMethods method = TimeStepperMethods::explicit_rungekutta_third_order_ssp>
class PeerMemoryRungeKuttaTimeStepper
{
public:
  using V = XT::LA::IstlDenseVector<double>;

  // `op` is this rank's operator: the constructor restricts it to the rank's slab
  PeerMemoryRungeKuttaTimeStepper(AdvectionFvOperator<M, GV>& op, CollectiveCommunication& comm, const double r = 1.0,
                                  const double t_0 = 0.0)
    : comm_(comm)
  {
    const auto& space = op.source_space();
    const auto s = slab_layers(internal::n_last(space), comm.rank(), comm.size());
    begin_ = s.first;
    end_ = s.second;
    GDT::internal::check(gdtb_fvop_set_slab(op.handle(), begin_, end_));
    plane_ = gdtb_fvop_ghost_layer_size(op.handle());
    owned_ = (end_ - begin_) * plane_;
    gdtb_rk* raw = nullptr;
    GDT::internal::check(gdtb_rk_create(op.handle(), int(method), 0, nullptr, nullptr, nullptr, r, t_0, &raw));
    handle_ = GDT::internal::Handle<gdtb_rk, gdtb_rk_destroy>(raw);
    unsigned char mine[4 * GDTB_IPC_HANDLE_BYTES];
    GDT::internal::check(gdtb_rk_p2p_handles(raw, &d_u_, mine));
    periodic_ = internal::periodic_last(space);
    internal::connect_neighbours(comm, internal::n_last(space), periodic_, mine,
                                 [&](const void* lo, std::int64_t lo_layers, int lo_self, const void* hi,
                                     std::int64_t hi_layers, int hi_self) {
                                   GDT::internal::check(
                                       gdtb_rk_p2p_connect(raw, lo, lo_layers, lo_self, hi, hi_layers, hi_self));
                                 });
  }
  ~PeerMemoryRungeKuttaTimeStepper()
  {
    // collective: nobody may free its buffers while a neighbour can still store into them
    gdtb_ctx_synchronize(GDT::internal::context());
    try {
      comm_.barrier();
    } catch (...) {
    }
  }

  // every rank passes the same global initial values (cell averages of the whole grid); owned part and ghost layers
  // are cut out here, so no exchange is needed for the start
  void set_initial_values(const V& u_global)
  {
    const std::int64_t total = std::int64_t(u_global.size());
    std::vector<double> loc(std::size_t(owned_ + 2 * plane_), 0.);
    const std::int64_t first = begin_ * plane_;
    std::copy(u_global.data() + first, u_global.data() + first + owned_, loc.begin() + plane_);
    const std::int64_t below = begin_ > 0 ? first - plane_ : (periodic_ ? total - plane_ : -1);
    const std::int64_t above = first + owned_ < total ? first + owned_ : (periodic_ ? 0 : -1);
    if (below >= 0)
      std::copy(u_global.data() + below, u_global.data() + below + plane_, loc.begin());
    if (above >= 0)
      std::copy(u_global.data() + above, u_global.data() + above + plane_, loc.begin() + plane_ + owned_);
    GDT::internal::check(gdtb_vector_upload(GDT::internal::context(), d_u_, loc.data(), std::int64_t(loc.size())));
    comm_.barrier();
  }

  double current_time() const
  {
    return gdtb_rk_current_time(handle_.get());
  }
  double step(const double dt, const double max_dt)
  {
    double ret = dt;
    GDT::internal::check(gdtb_rk_step(handle_.get(), d_u_, dt, max_dt, &ret));
    return ret;
  }
  // TimeStepperInterface::solve(t_end, initial_dt) (tools/timestepper/interface.hh:191-263), nothing saved or printed
  double solve(const double t_end, const double initial_dt)
  {
    std::int64_t steps = 0;
    double next = initial_dt;
    GDT::internal::check(gdtb_rk_solve(handle_.get(), d_u_, t_end, initial_dt, &steps, &next));
    GDT::internal::check(gdtb_rk_p2p_check(handle_.get()));
    num_steps_ = steps;
    return next;
  }
  std::int64_t num_steps() const
  {
    return num_steps_;
  }
  // the owned cells of the current solution (global cell indices [layer_begin * plane, layer_end * plane))
  V owned_solution() const
  {
    GDT::internal::check(gdtb_rk_p2p_check(handle_.get()));
    V u(std::size_t(owned_), 0.);
    GDT::internal::check(gdtb_vector_download(GDT::internal::context(), u.data(), d_u_ + plane_, owned_));
    return u;
  }
  std::int64_t layer_begin() const
  {
    return begin_;
  }
  std::int64_t layer_end() const
  {
    return end_;
  }
  std::int64_t cells_per_layer() const
  {
    return plane_;
  }

private:
  CollectiveCommunication& comm_;
  std::int64_t begin_ = 0, end_ = 0, plane_ = 0, owned_ = 0;
  bool periodic_ = false;
  double* d_u_ = nullptr;
  GDT::internal::Handle<gdtb_rk, gdtb_rk_destroy> handle_;
  std::int64_t num_steps_ = 0;
};


// explicit_euler of examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:141-159 on slabs: one fused apply + update +
// ghost hand-over kernel per step and rank (gdtb_fvop_p2p_*), no host-launched collective inside the loop.
template <class M, class GV>
class PeerMemoryEulerTimeLoop
{
public:
  using V = XT::LA::IstlDenseVector<double>;

  PeerMemoryEulerTimeLoop(AdvectionFvOperator<M, GV>& op, CollectiveCommunication& comm)
    : comm_(comm)
    , op_(op.handle())
  {
    const auto& space = op.source_space();
    const auto s = slab_layers(internal::n_last(space), comm.rank(), comm.size());
    begin_ = s.first;
    end_ = s.second;
    GDT::internal::check(gdtb_fvop_set_slab(op_, begin_, end_));
    plane_ = gdtb_fvop_ghost_layer_size(op_);
    owned_ = (end_ - begin_) * plane_;
    unsigned char mine[3 * GDTB_IPC_HANDLE_BYTES];
    GDT::internal::check(gdtb_fvop_p2p_alloc(op_, &d_u0_, &d_u1_, mine));
    periodic_ = internal::periodic_last(space);
    internal::connect_neighbours(comm, internal::n_last(space), periodic_, mine,
                                 [&](const void* lo, std::int64_t lo_layers, int lo_self, const void* hi,
                                     std::int64_t hi_layers, int hi_self) {
                                   GDT::internal::check(
                                       gdtb_fvop_p2p_connect(op_, lo, lo_layers, lo_self, hi, hi_layers, hi_self));
                                 });
  }
  ~PeerMemoryEulerTimeLoop()
  {
    gdtb_ctx_synchronize(GDT::internal::context());
    try {
      comm_.barrier();
    } catch (...) {
    }
  }
  void set_initial_values(const V& u_global)
  {
    const std::int64_t total = std::int64_t(u_global.size());
    std::vector<double> loc(std::size_t(owned_ + 2 * plane_), 0.);
    const std::int64_t first = begin_ * plane_;
    std::copy(u_global.data() + first, u_global.data() + first + owned_, loc.begin() + plane_);
    const std::int64_t below = begin_ > 0 ? first - plane_ : (periodic_ ? total - plane_ : -1);
    const std::int64_t above = first + owned_ < total ? first + owned_ : (periodic_ ? 0 : -1);
    if (below >= 0)
      std::copy(u_global.data() + below, u_global.data() + below + plane_, loc.begin());
    if (above >= 0)
      std::copy(u_global.data() + above, u_global.data() + above + plane_, loc.begin() + plane_ + owned_);
    GDT::internal::check(gdtb_vector_upload(GDT::internal::context(), d_u0_, loc.data(), std::int64_t(loc.size())));
    comm_.barrier();
  }
  // `while (time < T_end + dt)` of the reference's loop: u <- u - dt L(u), n_steps times, enqueued back to back
  void euler_steps(const double dt, const std::int64_t n_steps)
  {
    for (std::int64_t s = 0; s < n_steps; ++s)
      GDT::internal::check(gdtb_fvop_p2p_step(op_, 1, dt));
    GDT::internal::check(gdtb_fvop_p2p_check(op_));
  }
  V owned_solution() const
  {
    double* cur = nullptr;
    std::int64_t step = 0;
    GDT::internal::check(gdtb_fvop_p2p_check(op_));
    GDT::internal::check(gdtb_fvop_p2p_current(op_, &cur, &step));
    V u(std::size_t(owned_), 0.);
    GDT::internal::check(gdtb_vector_download(GDT::internal::context(), u.data(), cur + plane_, owned_));
    return u;
  }
  std::int64_t layer_begin() const
  {
    return begin_;
  }
  std::int64_t layer_end() const
  {
    return end_;
  }
  std::int64_t cells_per_layer() const
  {
    return plane_;
  }

private:
  CollectiveCommunication& comm_;
  gdtb_fvop* op_;
  std::int64_t begin_ = 0, end_ = 0, plane_ = 0, owned_ = 0;
  bool periodic_ = false;
  double *d_u0_ = nullptr, *d_u1_ = nullptr;
};


// Interface-row halo partition of a CG Q1 assembly with the hand-over inside the gather kernel (gdtb_halo_p2p_*): every
// rank walks only its own elements; the partial sums of the interface rows travel to the owner above by peer stores.
// values() / vector() hold the rows of the vertex layers [layer_begin, layer_end]; row_ranges()[0] describes the OWNED
// ones (the top layer of a rank that has an upper neighbour is the interface layer that neighbour owns).
template <class GV>
class HaloSlabAssembler
{
public:
  using E = typename GV::Element;

  HaloSlabAssembler(const SpaceInterface<GV>& space, CollectiveCommunication& comm)
    : space_(space)
    , comm_(comm)
  {
    const auto s = slab_layers(internal::n_last(space), comm.rank(), comm.size());
    begin_ = s.first;
    end_ = s.second;
    gdtb_matop* op = nullptr;
    GDT::internal::check(gdtb_matop_create(GDT::internal::context(), space_.handle(), space_.handle(), nullptr, &op));
    op_ = GDT::internal::Handle<gdtb_matop, gdtb_matop_destroy>(op);
    GDT::internal::check(gdtb_matop_set_slab_halo(op, begin_, end_));
    gdtb_vecfun* fun = nullptr;
    GDT::internal::check(gdtb_vecfun_create(GDT::internal::context(), space_.handle(), &fun));
    fun_ = GDT::internal::Handle<gdtb_vecfun, gdtb_vecfun_destroy>(fun);
    GDT::internal::check(gdtb_vecfun_set_slab_halo(fun, begin_, end_));
    if (comm.size() > 1) {
      unsigned char mine[2 * GDTB_IPC_HANDLE_BYTES];
      GDT::internal::check(gdtb_halo_p2p_alloc(op, mine));
      internal::connect_neighbours(comm, internal::n_last(space), false, mine,
                                   [&](const void* lo, std::int64_t lo_layers, int, const void* hi, std::int64_t, int) {
                                     GDT::internal::check(gdtb_halo_p2p_connect(op, lo, lo_layers, hi));
                                   });
    }
    std::int64_t rb = 0, re = 0, vo = 0;
    GDT::internal::check(gdtb_matop_local_rows(op, &rb, &re, &vo));
    // the interface layer at the top belongs to the rank above: held here, but not an owned row
    std::int64_t recv = 0, send = 0, count_m = 0, count_v = 0;
    GDT::internal::check(gdtb_matop_halo_layout(op, &recv, &send, &count_m));
    GDT::internal::check(gdtb_vecfun_halo_layout(fun, &recv, &send, &count_v));
    const bool top = comm.rank() + 1 < comm.size();
    const std::int64_t nnz = gdtb_matop_local_nnz(op);
    ranges_.push_back({rb, re - (top ? count_v : 0), vo, nnz - (top ? count_m : 0)});
    values_.resize(std::size_t(nnz));
    vector_.resize(std::size_t(re - rb));
  }
  ~HaloSlabAssembler()
  {
    gdtb_ctx_synchronize(GDT::internal::context());
    try {
      comm_.barrier();
    } catch (...) {
    }
  }
  HaloSlabAssembler& append(const LocalElementBilinearFormInterface<E>& form)
  {
    const auto f = form.lowered(1., space_.handle(), space_.grid_view().desc());
    GDT::internal::check(gdtb_matop_append_element(op_.get(), &f.form));
    return *this;
  }
  HaloSlabAssembler& append(const LocalElementFunctionalInterface<E>& functional)
  {
    const auto f = functional.lowered(space_.handle(), space_.grid_view().desc());
    GDT::internal::check(gdtb_vecfun_append_element(fun_.get(), &f.form));
    return *this;
  }
  void assemble()
  {
    GDT::internal::check(gdtb_assemble(op_.get(), fun_.get(), GDTB_ASSEMBLE_OVERWRITE));
    if (comm_.size() > 1)
      GDT::internal::check(gdtb_halo_p2p_check(op_.get()));
    GDT::internal::check(gdtb_matop_values_download(op_.get(), values_.data()));
    GDT::internal::check(gdtb_vecfun_download(fun_.get(), vector_.data()));
  }
  const std::vector<double>& values() const
  {
    return values_;
  }
  const std::vector<double>& vector() const
  {
    return vector_;
  }
  const std::vector<RowRange>& row_ranges() const
  {
    return ranges_;
  }

private:
  SpaceInterface<GV> space_;
  CollectiveCommunication& comm_;
  std::int64_t begin_ = 0, end_ = 0;
  GDT::internal::Handle<gdtb_matop, gdtb_matop_destroy> op_;
  GDT::internal::Handle<gdtb_vecfun, gdtb_vecfun_destroy> fun_;
  std::vector<RowRange> ranges_;
  std::vector<double> values_, vector_;
};


} // namespace Parallel
} // namespace GDT
} // namespace Dune

#endif // DUNE_GDT_B200_PARALLEL_HH

// dune-gdt_b200/include/dune/gdt/b200.hh -- header-only C++ facade: dune-gdt's class and function names for the
// assembly / FV-apply hot path, lowered through the C ABI (include/gdtb.h) to the sm_100a kernels of libgdtb.
//
// Scope: exactly what the three reference drivers call on this path:
//   examples/stationary-heat-equation.cc:87-106, examples/adaptive_elliptic_swipdg.cc:216-251,
//   examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:141-158,255-298.
// Same spelling, argument order and meaning as the reference; value types are cloned on append like
// copy()/copy_as_*_integrand() (operators/matrix-based.hh:346-408); errors are Dune::Exception subclasses with the
// reference's names.  The dune-xt / dune-grid types the drivers touch (grid provider, grid view, GridFunction,
// LA containers, walker, ApplyOn filters) are provided as thin stand-ins in namespace Dune::XT -- they are
// [EXT] to dune-gdt and only cover what the hot path needs.
//
// What cannot cross a C ABI as code: user lambdas.  XT::Functions::GenericFunction keeps the lambda on the host; when a
// local form that uses it is appended to an operator / functional, the facade asks the library for the rule the
// reference would integrate that form with (gdtb_form_quadrature_order, gdtb_gauss_rule), evaluates the lambda at
// every quadrature point of every element -- exactly the calls the reference's integrand makes -- and hands the
// samples over as GDTB_FN_QP_* arrays.  Constants, per-element arrays and the built-in analytic functions of gdtb.h
// need no sampling.
#ifndef DUNE_GDT_B200_HH
#define DUNE_GDT_B200_HH

#include <array>
#include <cstddef>
#include <cstdint>
#include <exception>
#include <functional>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include <gdtb.h>

namespace Dune {

// ---- exceptions (dune-common Exception + dune/gdt/exceptions.hh:24-75) ------------------------------------
class Exception : public std::exception
{
public:
  Exception() = default;
  explicit Exception(std::string msg)
    : msg_(std::move(msg))
  {}
  const char* what() const noexcept override
  {
    return msg_.c_str();
  }
  void message(const std::string& m)
  {
    msg_ = m;
  }

private:
  std::string msg_;
};

class NotImplemented : public Exception
{
  using Exception::Exception;
};
class InvalidStateException : public Exception
{
  using Exception::Exception;
};

namespace XT {
namespace Common {
namespace Exceptions {
class wrong_input_given : public Dune::Exception
{
  using Exception::Exception;
};
class shapes_do_not_match : public Dune::Exception
{
  using Exception::Exception;
};
} // namespace Exceptions

// XT::Common::Parameter: only carried for signature compatibility ({} everywhere on this path)
struct Parameter
{
  Parameter() = default;
  Parameter(std::initializer_list<std::pair<std::string, std::vector<double>>>) {}
};
struct ParameterType
{
  ParameterType() = default;
  ParameterType(std::initializer_list<std::pair<std::string, std::size_t>>) {}
};
} // namespace Common
} // namespace XT

namespace GDT {
namespace Exceptions {
class integrand_error : public Dune::Exception
{
  using Exception::Exception;
};
class finite_element_error : public Dune::Exception
{
  using Exception::Exception;
};
class space_error : public Dune::Exception
{
  using Exception::Exception;
};
class mapper_error : public space_error
{
  using space_error::space_error;
};
class operator_error : public Dune::Exception
{
  using Exception::Exception;
};
class device_error : public Dune::Exception // CUDA failure / no device: there is no CPU fallback
{
  using Exception::Exception;
};
} // namespace Exceptions

namespace internal {
inline void check(int status)
{
  if (status == GDTB_OK)
    return;
  const std::string msg = gdtb_last_error();
  switch (status) {
    case GDTB_ERR_INVALID_ARGUMENT: throw XT::Common::Exceptions::wrong_input_given(msg);
    case GDTB_ERR_SHAPES_DO_NOT_MATCH: throw XT::Common::Exceptions::shapes_do_not_match(msg);
    case GDTB_ERR_INTEGRAND: throw Exceptions::integrand_error(msg);
    case GDTB_ERR_FINITE_ELEMENT: throw Exceptions::finite_element_error(msg);
    case GDTB_ERR_SPACE: throw Exceptions::space_error(msg);
    case GDTB_ERR_OPERATOR: throw Exceptions::operator_error(msg);
    case GDTB_ERR_NOT_IMPLEMENTED: throw Dune::NotImplemented(msg);
    default: throw Exceptions::device_error(msg);
  }
}

// process-wide context (one GPU per process, like one MPI rank per GPU)
inline gdtb_ctx* context(int device = -1)
{
  struct Holder
  {
    gdtb_ctx* ctx = nullptr;
    ~Holder()
    {
      gdtb_ctx_destroy(ctx);
    }
  };
  static Holder holder;
  if (!holder.ctx)
    check(gdtb_ctx_create(device < 0 ? 0 : device, &holder.ctx));
  return holder.ctx;
}

template <class T, int (*Destroy)(T*)>
struct Handle
{
  std::shared_ptr<T> ptr;
  Handle() = default;
  explicit Handle(T* raw)
    : ptr(raw, [](T* p) { Destroy(p); })
  {}
  T* get() const
  {
    return ptr.get();
  }
};
} // namespace internal
} // namespace GDT

// ==========================================================================================================
// [EXT] stand-ins for the dune-grid / dune-xt types the drivers touch
// ==========================================================================================================
template <class K, int n>
using FieldVector = std::array<K, n>;

template <class K, int r, int c>
struct FieldMatrix
{
  std::array<std::array<K, c>, r> rows{};
  std::array<K, c>& operator[](std::size_t i)
  {
    return rows[i];
  }
  const std::array<K, c>& operator[](std::size_t i) const
  {
    return rows[i];
  }
};

namespace XT {
namespace Grid {

template <std::size_t d>
struct CubeEntity
{
  static constexpr std::size_t dimension = d;
};
template <std::size_t d>
struct CubeIntersection
{
  using Entity = CubeEntity<d>;
};

// Dune::YaspGrid<d, EquidistantOffsetCoordinates<double, d>>::LeafGridView stand-in
template <std::size_t d>
class CubeGridView
{
public:
  static constexpr std::size_t dimension = d;
  using ctype = double;
  using Element = CubeEntity<d>;
  using Intersection = CubeIntersection<d>;

  CubeGridView() = default;
  CubeGridView(const gdtb_grid_desc& desc)
    : desc_(desc)
  {
    gdtb_grid* raw = nullptr;
    GDT::internal::check(gdtb_grid_create_cube(GDT::internal::context(), &desc_, &raw));
    handle_ = GDT::internal::Handle<gdtb_grid, gdtb_grid_destroy>(raw);
  }
  gdtb_grid* handle() const
  {
    return handle_.get();
  }
  const gdtb_grid_desc& desc() const
  {
    return desc_;
  }
  std::int64_t size(int codim) const
  {
    if (codim != 0)
      throw Dune::NotImplemented("CubeGridView::size: only codim 0");
    return gdtb_grid_num_elements(handle_.get());
  }

private:
  gdtb_grid_desc desc_{};
  GDT::internal::Handle<gdtb_grid, gdtb_grid_destroy> handle_;
};

template <std::size_t d>
struct YaspEquidistantOffset
{
  static constexpr std::size_t dimension = d;
  using LeafGridView = CubeGridView<d>;
};

template <class GV>
using extract_entity_t = typename GV::Element;
template <class GV>
using extract_intersection_t = typename GV::Intersection;

template <class G>
class GridProvider
{
public:
  explicit GridProvider(const gdtb_grid_desc& desc)
    : desc_(desc)
  {}
  typename G::LeafGridView leaf_view() const
  {
    return typename G::LeafGridView(desc_);
  }
  const gdtb_grid_desc& desc() const
  {
    return desc_;
  }

private:
  gdtb_grid_desc desc_;
};

// XT::Grid::make_cube_grid<G>(lower, upper, num_elements) (examples/stationary-heat-equation.cc:87)
template <class G>
GridProvider<G> make_cube_grid(const std::array<double, G::dimension>& lower,
                               const std::array<double, G::dimension>& upper,
                               const std::array<unsigned int, G::dimension>& num_elements)
{
  gdtb_grid_desc desc{};
  desc.dim = int(G::dimension);
  desc.periodic = 0;
  for (std::size_t k = 0; k < 3; ++k) {
    desc.lower[k] = k < G::dimension ? lower[k] : 0.;
    desc.upper[k] = k < G::dimension ? upper[k] : 1.;
    desc.n[k] = k < G::dimension ? num_elements[k] : 1;
  }
  return GridProvider<G>(desc);
}

template <class G>
GridProvider<G> make_cube_grid(const double lower, const double upper, const unsigned int num_elements)
{
  std::array<double, G::dimension> lo, up;
  std::array<unsigned int, G::dimension> n;
  lo.fill(lower);
  up.fill(upper);
  n.fill(num_elements);
  return make_cube_grid<G>(lo, up, n);
}

// XT::Grid::make_periodic_grid_view(view) (examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:261)
template <std::size_t d>
CubeGridView<d> make_periodic_grid_view(const CubeGridView<d>& view)
{
  gdtb_grid_desc desc = view.desc();
  desc.periodic = (1 << d) - 1;
  return CubeGridView<d>(desc);
}

// intersection / element filters used on this path (XT::Grid::ApplyOn)
template <class GV>
struct ElementFilter
{
  virtual ~ElementFilter() = default;
};
template <class GV>
struct IntersectionFilter
{
  virtual ~IntersectionFilter() = default;
  virtual int gdtb_filter() const = 0;
};
namespace ApplyOn {
template <class GV>
struct AllElements : ElementFilter<GV>
{};
template <class GV>
struct InnerIntersectionsOnce : IntersectionFilter<GV>
{
  int gdtb_filter() const override
  {
    return GDTB_FILTER_INNER_ONCE;
  }
};
template <class GV>
struct PeriodicBoundaryIntersectionsOnce : IntersectionFilter<GV>
{
  int gdtb_filter() const override
  {
    return GDTB_FILTER_INNER_AND_PERIODIC_ONCE;
  }
};
template <class GV>
struct AllIntersections : IntersectionFilter<GV>
{
  int gdtb_filter() const override
  {
    return -1;
  }
};
} // namespace ApplyOn

struct DirichletBoundary
{};
template <class I>
struct AllDirichletBoundaryInfo
{};
namespace ApplyOn {
// CustomBoundaryIntersections(boundary_info, new DirichletBoundary()) with AllDirichletBoundaryInfo
// (examples/adaptive_elliptic_swipdg.cc:243)
template <class GV>
struct CustomBoundaryIntersections : IntersectionFilter<GV>
{
  template <class Info>
  CustomBoundaryIntersections(const Info&, DirichletBoundary* type)
  {
    delete type;
  }
  int gdtb_filter() const override
  {
    return GDTB_FILTER_ALL_BOUNDARY;
  }
};
} // namespace ApplyOn
} // namespace Grid

namespace Functions {

namespace internal {
// a host-side function of the global coordinate: what is left of a GenericFunction lambda after type erasure
struct Generic
{
  int order = 0;
  int comps = 1; // 1: scalar, d * d: matrix-valued (row-major)
  std::function<void(const double* x, double* out)> evaluate;
  std::function<void(const double* x, double* grad)> jacobian; // scalar functions only; may be empty
};
} // namespace internal

// XT::Functions::GenericFunction<d, r, rC> [EXT dune-xt functions/generic/function.hh]: order + evaluate lambda
// (+ name, parameter type, jacobian lambda), examples/stationary-heat-equation.cc:67-85.  Scalar (r = rC = 1) and
// matrix-valued (r = rC = d) functions.
template <std::size_t d, std::size_t r = 1, std::size_t rC = 1, class R = double>
class GenericFunction
{
public:
  using DomainType = FieldVector<double, int(d)>;
  using RangeReturnType = typename std::conditional<r == 1 && rC == 1, double, FieldMatrix<double, int(r), int(rC)>>::type;
  using DerivativeRangeReturnType = FieldMatrix<double, int(r), int(d)>;
  using EvaluateType = std::function<RangeReturnType(const DomainType&, const Common::Parameter&)>;
  using JacobianType = std::function<DerivativeRangeReturnType(const DomainType&, const Common::Parameter&)>;

  GenericFunction(const int ord, EvaluateType evaluate, const std::string& /*name*/ = "GenericFunction",
                  const Common::ParameterType& /*param_type*/ = {}, JacobianType jacobian = nullptr)
    : generic_(std::make_shared<internal::Generic>())
  {
    static_assert((r == 1 && rC == 1) || (r == d && rC == d), "scalar or d x d matrix-valued functions");
    generic_->order = ord;
    generic_->comps = int(r * rC);
    generic_->evaluate = [evaluate](const double* x, double* out) {
      DomainType xx;
      for (std::size_t k = 0; k < d; ++k)
        xx[k] = x[k];
      store(evaluate(xx, Common::Parameter()), out);
    };
    if (jacobian)
      generic_->jacobian = [jacobian](const double* x, double* grad) {
        DomainType xx;
        for (std::size_t k = 0; k < d; ++k)
          xx[k] = x[k];
        const auto J = jacobian(xx, Common::Parameter());
        for (std::size_t k = 0; k < d; ++k)
          grad[k] = J[0][k];
      };
  }
  int order() const
  {
    return generic_->order;
  }
  const std::shared_ptr<internal::Generic>& generic() const
  {
    return generic_;
  }

private:
  static void store(const double v, double* out)
  {
    out[0] = v;
  }
  static void store(const FieldMatrix<double, int(r), int(rC)>& v, double* out)
  {
    for (std::size_t i = 0; i < r; ++i)
      for (std::size_t j = 0; j < rC; ++j)
        out[i * rC + j] = v[i][j];
  }
  std::shared_ptr<internal::Generic> generic_;
};

// XT::Functions::GridFunction<E, r, rC>: a constant (scalar -> c * I for r = rC = d, laplace.hh:41), a constant
// matrix, a per-element array, a built-in analytic function, or a GenericFunction (sampled when the local form it is
// used in is appended).  Cloned by value.
template <class E, std::size_t r = 1, std::size_t rC = 1, class R = double>
class GridFunction
{
public:
  // a GenericFunction lambda: scalar, or matrix-valued; a scalar lambda used as a d x d function means c(x) * I
  template <std::size_t fr, std::size_t frC>
  GridFunction(const GenericFunction<E::dimension, fr, frC>& function)
    : generic_(function.generic())
  {
    static_assert((fr == 1 && frC == 1) || (fr == r && frC == rC), "function shape does not match");
    fn_.kind = GDTB_FN_CONST_SCALAR; // placeholder: replaced by the samples at append time
    fn_.order = function.order();
    fn_.c[0] = 1.;
  }
  const std::shared_ptr<internal::Generic>& generic() const
  {
    return generic_;
  }

  GridFunction(const double value = 1., const int order = 0)
  {
    fn_.kind = GDTB_FN_CONST_SCALAR;
    fn_.order = order;
    fn_.c[0] = value;
  }
  GridFunction(const FieldMatrix<double, int(r), int(rC)>& value)
  {
    fn_.kind = GDTB_FN_CONST_TENSOR;
    for (std::size_t i = 0; i < r; ++i)
      for (std::size_t j = 0; j < rC; ++j)
        fn_.c[i * rC + j] = value[i][j];
  }
  explicit GridFunction(const gdtb_function& fn, std::shared_ptr<std::vector<double>> storage = nullptr)
    : fn_(fn)
    , storage_(std::move(storage))
  {}
  const gdtb_function& descriptor() const
  {
    return fn_;
  }
  int order() const
  {
    return fn_.order;
  }

private:
  gdtb_function fn_{};
  std::shared_ptr<std::vector<double>> storage_;
  std::shared_ptr<internal::Generic> generic_;
};

// element-wise constant data (e.g. a heterogeneous diffusion field), values[e] for element index e
template <class E>
GridFunction<E> make_elementwise_constant(std::vector<double> values, const int order = 0)
{
  auto storage = std::make_shared<std::vector<double>>(std::move(values));
  gdtb_function fn{};
  fn.kind = GDTB_FN_ELEM_SCALAR;
  fn.order = order;
  fn.data = storage->data();
  return GridFunction<E>(fn, storage);
}

// stand-ins for the GenericFunction lambdas of the reference drivers
template <class E>
GridFunction<E> make_cosine_product(const int order, const double factor, const double frequency)
{
  gdtb_function fn{};
  fn.kind = GDTB_FN_BUILTIN;
  fn.builtin = GDTB_BUILTIN_COS_PRODUCT;
  fn.order = order;
  fn.p[0] = factor;
  fn.p[1] = frequency;
  return GridFunction<E>(fn);
}
template <class E>
GridFunction<E> make_indicator(const int order, const double lower, const double upper)
{
  gdtb_function fn{};
  fn.kind = GDTB_FN_BUILTIN;
  fn.builtin = GDTB_BUILTIN_INDICATOR;
  fn.order = order;
  fn.p[0] = lower;
  fn.p[1] = upper;
  return GridFunction<E>(fn);
}
template <class E>
GridFunction<E> make_gaussian(const int order, const double center, const double sigma)
{
  gdtb_function fn{};
  fn.kind = GDTB_FN_BUILTIN;
  fn.builtin = GDTB_BUILTIN_GAUSSIAN;
  fn.order = order;
  fn.p[0] = center;
  fn.p[1] = sigma;
  return GridFunction<E>(fn);
}

} // namespace Functions

namespace LA {

// XT::LA::IstlDenseVector<double> stand-in (host storage)
template <class R = double>
class IstlDenseVector
{
public:
  using ScalarType = R;
  explicit IstlDenseVector(std::size_t size = 0, R value = 0)
    : data_(size, value)
  {}
  std::size_t size() const
  {
    return data_.size();
  }
  R& operator[](std::size_t i)
  {
    return data_[i];
  }
  const R& operator[](std::size_t i) const
  {
    return data_[i];
  }
  R get_entry(std::size_t i) const
  {
    return data_[i];
  }
  R* data()
  {
    return data_.data();
  }
  const R* data() const
  {
    return data_.data();
  }
  IstlDenseVector operator*(R alpha) const
  {
    IstlDenseVector out(*this);
    for (auto& x : out.data_)
      x *= alpha;
    return out;
  }
  IstlDenseVector operator-(const IstlDenseVector& other) const
  {
    IstlDenseVector out(*this);
    for (std::size_t i = 0; i < data_.size(); ++i)
      out.data_[i] -= other.data_[i];
    return out;
  }
  R sup_norm() const
  {
    R m = 0;
    for (auto x : data_)
      m = std::max(m, x < 0 ? -x : x);
    return m;
  }
  // the device-resident functional this host copy mirrors (set by VectorBasedFunctional; non-owning)
  void bind_device(gdtb_vecfun* f)
  {
    device_ = f;
  }
  gdtb_vecfun* device() const
  {
    return device_;
  }

private:
  std::vector<R> data_;
  gdtb_vecfun* device_ = nullptr;
};

// XT::LA::SparsityPatternDefault stand-in: the CSR pattern lives on the device, a host copy is made on demand
class SparsityPatternDefault
{
public:
  SparsityPatternDefault() = default;
  explicit SparsityPatternDefault(gdtb_pattern* raw)
    : handle_(raw)
  {}
  gdtb_pattern* handle() const
  {
    return handle_.get();
  }
  std::size_t size() const
  {
    return std::size_t(gdtb_pattern_rows(handle_.get()));
  }
  std::int64_t nnz() const
  {
    return gdtb_pattern_nnz(handle_.get());
  }
  // inner(row): the sorted column indices of a row (as SparsityPatternDefault::inner)
  std::vector<std::size_t> inner(std::size_t row) const
  {
    fetch();
    std::vector<std::size_t> cols;
    for (std::int64_t k = rowptr_[row]; k < rowptr_[row + 1]; ++k)
      cols.push_back(std::size_t(colidx_[std::size_t(k)]));
    return cols;
  }
  const std::vector<std::int64_t>& rowptr() const
  {
    fetch();
    return rowptr_;
  }
  const std::vector<std::int32_t>& colidx() const
  {
    fetch();
    return colidx_;
  }

private:
  void fetch() const
  {
    if (!rowptr_.empty())
      return;
    rowptr_.resize(size() + 1);
    colidx_.resize(std::size_t(nnz()));
    GDT::internal::check(gdtb_pattern_download(handle_.get(), rowptr_.data(), colidx_.data()));
  }
  GDT::internal::Handle<gdtb_pattern, gdtb_pattern_destroy> handle_;
  mutable std::vector<std::int64_t> rowptr_;
  mutable std::vector<std::int32_t> colidx_;
};

// XT::LA::IstlRowMajorSparseMatrix<double> stand-in: CSR, values on the host after assemble()
template <class R = double>
class IstlRowMajorSparseMatrix
{
public:
  using ScalarType = R;
  IstlRowMajorSparseMatrix() = default;
  IstlRowMajorSparseMatrix(std::size_t rows, std::size_t cols, const SparsityPatternDefault& pattern)
    : rows_(rows)
    , cols_(cols)
    , pattern_(pattern)
    , values_(std::size_t(pattern.nnz()), R(0))
  {}
  std::size_t rows() const
  {
    return rows_;
  }
  std::size_t cols() const
  {
    return cols_;
  }
  std::size_t non_zeros() const
  {
    return values_.size();
  }
  const SparsityPatternDefault& pattern() const
  {
    return pattern_;
  }
  R get_entry(std::size_t i, std::size_t j) const
  {
    const auto& rp = pattern_.rowptr();
    const auto& ci = pattern_.colidx();
    for (std::int64_t k = rp[i]; k < rp[i + 1]; ++k)
      if (std::size_t(ci[std::size_t(k)]) == j)
        return values_[std::size_t(k)];
    return R(0);
  }
  // y = A x (ConstMatrixOperator::apply, operators/matrix-based.hh:121-129): the CSR mat-vec runs on the device with
  // the operator's values (this host container mirrors them); a matrix without a device operator cannot multiply
  void mv(const IstlDenseVector<R>& x, IstlDenseVector<R>& y) const
  {
    if (!device_)
      throw GDT::Exceptions::device_error("matrix is not attached to a device operator (there is no CPU path)");
    if (x.size() != cols_ || y.size() != rows_)
      throw Common::Exceptions::shapes_do_not_match("mv: vector sizes do not match the matrix");
    GDT::internal::check(gdtb_matop_apply_host(device_, x.data(), y.data()));
  }
  // the device-resident operator this host copy mirrors (set by MatrixOperator; non-owning)
  void bind_device(gdtb_matop* op)
  {
    device_ = op;
  }
  gdtb_matop* device() const
  {
    return device_;
  }
  std::vector<R>& values()
  {
    return values_;
  }
  const std::vector<R>& values() const
  {
    return values_;
  }

private:
  std::size_t rows_ = 0, cols_ = 0;
  SparsityPatternDefault pattern_;
  std::vector<R> values_;
  gdtb_matop* device_ = nullptr;
};

// XT::LA::make_solver(matrix).apply(rhs, solution) (examples/stationary-heat-equation.cc:110) [EXT dune-xt]: Krylov
// solve on the device with the operator's values (CG / BiCGStab as CUDA graphs, gdtb_matop_apply_inverse)
template <class M>
class Solver
{
public:
  explicit Solver(const M& matrix)
    : matrix_(matrix)
  {}
  void apply(const IstlDenseVector<double>& rhs, IstlDenseVector<double>& solution) const
  {
    apply(rhs, solution, gdtb_solver_opts{GDTB_SOLVER_CG, GDTB_PRECOND_JACOBI, 0, 0, 0.});
  }
  void apply(const IstlDenseVector<double>& rhs, IstlDenseVector<double>& solution, const gdtb_solver_opts& opts) const
  {
    if (!matrix_.device())
      throw GDT::Exceptions::device_error("matrix is not attached to a device operator (there is no CPU path)");
    if (rhs.size() != matrix_.rows() || solution.size() != matrix_.cols())
      throw Common::Exceptions::shapes_do_not_match("when applying linear solver: shapes do not match");
    gdtb_solver_info info{};
    GDT::internal::check(gdtb_matop_apply_inverse_host(matrix_.device(), rhs.data(), solution.data(), &opts, &info));
    last_ = info;
  }
  const gdtb_solver_info& info() const
  {
    return last_;
  }

private:
  const M& matrix_;
  mutable gdtb_solver_info last_{};
};
template <class M>
Solver<M> make_solver(const M& matrix)
{
  return Solver<M>(matrix);
}

} // namespace LA
} // namespace XT

// ==========================================================================================================
// dune-gdt
// ==========================================================================================================
namespace GDT {

// dune/gdt/type_traits.hh:24-32,55-60
enum class SpaceType
{
  continuous_lagrange,
  discontinuous_lagrange,
  finite_volume
};
enum class Stencil
{
  element,
  intersection,
  element_and_intersection
};

// ---- spaces (spaces/interface.hh:92-98, spaces/mapper/interfaces.hh) ----------------------------------------
template <class GV>
class SpaceInterface;

template <class GV>
class MapperInterface
{
public:
  explicit MapperInterface(const SpaceInterface<GV>& space)
    : space_(space)
  {}
  std::size_t size() const;
  std::size_t max_local_size() const;
  std::size_t local_size(std::int64_t /*element*/) const
  {
    return max_local_size();
  }
  std::vector<std::size_t> global_indices(std::int64_t element) const;

private:
  const SpaceInterface<GV>& space_;
};

template <class GV>
class SpaceInterface
{
public:
  using GridViewType = GV;
  static constexpr std::size_t d = GV::dimension;
  using ElementType = typename GV::Element;

  SpaceInterface(const GV& grid_view, int kind, int order)
    : grid_view_(grid_view)
    , kind_(kind)
    , order_(order)
    , mapper_(*this)
  {
    gdtb_space* raw = nullptr;
    internal::check(gdtb_space_create(internal::context(), grid_view_.handle(), kind, order, &raw));
    handle_ = internal::Handle<gdtb_space, gdtb_space_destroy>(raw);
  }
  // make_finite_volume_space<m>: m components per element (spaces/l2/finite-volume.hh:208-230)
  struct FiniteVolumeSystem
  {
    int range_dim;
  };
  SpaceInterface(const GV& grid_view, const FiniteVolumeSystem fv)
    : grid_view_(grid_view)
    , kind_(GDTB_SPACE_FV)
    , order_(0)
    , range_dim_(fv.range_dim)
    , mapper_(*this)
  {
    gdtb_space* raw = nullptr;
    internal::check(gdtb_fv_space_create(internal::context(), grid_view_.handle(), fv.range_dim, &raw));
    handle_ = internal::Handle<gdtb_space, gdtb_space_destroy>(raw);
  }
  SpaceInterface(const SpaceInterface& other)
    : grid_view_(other.grid_view_)
    , kind_(other.kind_)
    , order_(other.order_)
    , range_dim_(other.range_dim_)
    , mapper_(*this)
    , handle_(other.handle_)
  {}
  int range_dim() const
  {
    return range_dim_;
  }
  const GV& grid_view() const
  {
    return grid_view_;
  }
  const MapperInterface<GV>& mapper() const
  {
    return mapper_;
  }
  SpaceType type() const
  {
    return kind_ == GDTB_SPACE_CG ? SpaceType::continuous_lagrange
                                  : (kind_ == GDTB_SPACE_DG ? SpaceType::discontinuous_lagrange
                                                            : SpaceType::finite_volume);
  }
  int min_polorder() const
  {
    return order_;
  }
  int max_polorder() const
  {
    return order_;
  }
  gdtb_space* handle() const
  {
    return handle_.get();
  }

private:
  GV grid_view_;
  int kind_, order_;
  int range_dim_ = 1;
  MapperInterface<GV> mapper_;
  internal::Handle<gdtb_space, gdtb_space_destroy> handle_;
};

template <class GV>
std::size_t MapperInterface<GV>::size() const
{
  return std::size_t(gdtb_space_size(space_.handle()));
}
template <class GV>
std::size_t MapperInterface<GV>::max_local_size() const
{
  return std::size_t(gdtb_space_max_local_size(space_.handle()));
}
template <class GV>
std::vector<std::size_t> MapperInterface<GV>::global_indices(std::int64_t element) const
{
  std::vector<std::int64_t> tmp(max_local_size());
  internal::check(gdtb_space_global_indices(space_.handle(), element, tmp.data()));
  return std::vector<std::size_t>(tmp.begin(), tmp.end());
}

template <class GV>
using ContinuousLagrangeSpace = SpaceInterface<GV>;
template <class GV>
using DiscontinuousLagrangeSpace = SpaceInterface<GV>;
template <class GV>
using FiniteVolumeSpace = SpaceInterface<GV>;

// spaces/h1/continuous-lagrange.hh:189-203
template <class GV>
SpaceInterface<GV> make_continuous_lagrange_space(const GV& grid_view, const int order)
{
  return SpaceInterface<GV>(grid_view, GDTB_SPACE_CG, order);
}
// spaces/l2/discontinuous-lagrange.hh:191-205
template <class GV>
SpaceInterface<GV> make_discontinuous_lagrange_space(const GV& grid_view, const int order)
{
  return SpaceInterface<GV>(grid_view, GDTB_SPACE_DG, order);
}
// spaces/l2/finite-volume.hh:208-230
template <class GV>
SpaceInterface<GV> make_finite_volume_space(const GV& grid_view)
{
  return SpaceInterface<GV>(grid_view, GDTB_SPACE_FV, 0);
}

// make_finite_volume_space<m>(grid_view) (spaces/l2/finite-volume.hh:208-230, m > 1: systems)
template <std::size_t m, class GV>
SpaceInterface<GV> make_finite_volume_space(const GV& grid_view)
{
  if (m == 1)
    return SpaceInterface<GV>(grid_view, GDTB_SPACE_FV, 0);
  return SpaceInterface<GV>(grid_view, typename SpaceInterface<GV>::FiniteVolumeSystem{int(m)});
}

// ---- sparsity patterns (tools/sparsity-pattern.hh:34-178) ---------------------------------------------------
template <class GV>
XT::LA::SparsityPatternDefault make_sparsity_pattern(const SpaceInterface<GV>& test_space,
                                                     const SpaceInterface<GV>& ansatz_space,
                                                     const GV& /*grid_view*/,
                                                     const Stencil stencil)
{
  gdtb_pattern* raw = nullptr;
  internal::check(gdtb_pattern_create(
      internal::context(), test_space.handle(), ansatz_space.handle(), int(stencil), GDTB_PATTERN_AUTO, &raw));
  return XT::LA::SparsityPatternDefault(raw);
}
template <class GV>
XT::LA::SparsityPatternDefault make_element_sparsity_pattern(const SpaceInterface<GV>& space)
{
  return make_sparsity_pattern(space, space, space.grid_view(), Stencil::element);
}
template <class GV>
XT::LA::SparsityPatternDefault make_element_sparsity_pattern(const SpaceInterface<GV>& test,
                                                             const SpaceInterface<GV>& ansatz,
                                                             const GV& grid_view)
{
  return make_sparsity_pattern(test, ansatz, grid_view, Stencil::element);
}
template <class GV>
XT::LA::SparsityPatternDefault make_intersection_sparsity_pattern(const SpaceInterface<GV>& space)
{
  return make_sparsity_pattern(space, space, space.grid_view(), Stencil::intersection);
}
template <class GV>
XT::LA::SparsityPatternDefault make_element_and_intersection_sparsity_pattern(const SpaceInterface<GV>& space)
{
  return make_sparsity_pattern(space, space, space.grid_view(), Stencil::element_and_intersection);
}

// ---- integrands (local/integrands/*.hh) -----------------------------------------------------------------------
namespace internal {
using GenericPtr = std::shared_ptr<XT::Functions::internal::Generic>;
struct IntegrandTerms
{
  std::vector<gdtb_integrand> terms;
  // GenericFunction lambdas behind terms[i].diffusion / terms[i].weight (null: the descriptor is complete)
  std::vector<std::pair<GenericPtr, GenericPtr>> generic;
  // keep per-element storage of the wrapped grid functions alive until the form has been appended (= cloned)
  std::vector<std::shared_ptr<void>> keep;
};
inline IntegrandTerms concat(const IntegrandTerms& a, const IntegrandTerms& b)
{
  IntegrandTerms out = a;
  out.terms.insert(out.terms.end(), b.terms.begin(), b.terms.end());
  out.generic.insert(out.generic.end(), b.generic.begin(), b.generic.end());
  out.keep.insert(out.keep.end(), b.keep.begin(), b.keep.end());
  return out;
}
template <class F>
GenericPtr generic_of(const F* f)
{
  return f ? f->generic() : GenericPtr();
}
// appends one integrand term together with the lambdas behind its functions
template <class FD, class FW>
void push_term(IntegrandTerms& t, int kind, const FD* diffusion, const FW* weight, double prefactor, int hI_kind);
inline gdtb_form make_form(const IntegrandTerms& t, int over_integrate, double scaling)
{
  if (t.terms.empty() || t.terms.size() > GDTB_MAX_TERMS)
    throw Exceptions::integrand_error("an integrand sum may have 1 to GDTB_MAX_TERMS summands");
  gdtb_form f{};
  f.n_terms = int(t.terms.size());
  f.over_integrate = over_integrate;
  f.scaling = scaling;
  for (std::size_t i = 0; i < t.terms.size(); ++i)
    f.terms[i] = t.terms[i];
  return f;
}
template <class F>
gdtb_integrand make_term(int kind, const F* diffusion, const F* weight, double prefactor, int hI_kind)
{
  gdtb_integrand in{};
  in.kind = kind;
  in.hI_kind = hI_kind;
  in.prefactor = prefactor;
  in.diffusion.kind = GDTB_FN_CONST_SCALAR;
  in.diffusion.c[0] = 1.;
  in.weight.kind = GDTB_FN_CONST_SCALAR;
  in.weight.c[0] = 1.;
  if (diffusion)
    in.diffusion = diffusion->descriptor();
  if (weight)
    in.weight = weight->descriptor();
  return in;
}
template <class FD, class FW>
void push_term(IntegrandTerms& t, int kind, const FD* diffusion, const FW* weight, double prefactor, int hI_kind)
{
  gdtb_integrand in{};
  in.kind = kind;
  in.hI_kind = hI_kind;
  in.prefactor = prefactor;
  in.diffusion.kind = GDTB_FN_CONST_SCALAR;
  in.diffusion.c[0] = 1.;
  in.weight.kind = GDTB_FN_CONST_SCALAR;
  in.weight.c[0] = 1.;
  if (diffusion)
    in.diffusion = diffusion->descriptor();
  if (weight)
    in.weight = weight->descriptor();
  t.terms.push_back(in);
  t.generic.emplace_back(generic_of(diffusion), generic_of(weight));
}

// A local form ready for gdtb_*_append_*: the descriptor plus the sample arrays its GDTB_FN_QP_* functions point to
struct LoweredForm
{
  gdtb_form form{};
  std::vector<std::shared_ptr<std::vector<double>>> samples;
};

// x_q = lower_e + xhat_q * (upper_e - lower_e) with YaspGrid's coordinates lower = origin + i h, upper = origin +
// (i + 1) h [EXT]: the points the reference hands to a bound local function (geometry.global(xhat))
inline void sample_generic(const XT::Functions::internal::Generic& g, const gdtb_grid_desc& grid, int m, const double* xh,
                           bool with_jacobian, std::vector<double>& out)
{
  const int d = grid.dim;
  long long ne = 1;
  double h[3] = {1., 1., 1.};
  for (int k = 0; k < d; ++k) {
    ne *= grid.n[k];
    h[k] = (grid.upper[k] - grid.lower[k]) / double(grid.n[k]);
  }
  int nq = 1;
  for (int k = 0; k < d; ++k)
    nq *= m;
  const int comps = with_jacobian ? 1 + d : g.comps;
  out.assign(std::size_t(ne) * nq * comps, 0.);
  double x[3] = {0., 0., 0.};
  for (long long e = 0; e < ne; ++e) {
    const long long idx[3] = {e % grid.n[0], (e / grid.n[0]) % grid.n[1], e / (grid.n[0] * grid.n[1])};
    double lower[3], ext[3];
    for (int k = 0; k < d; ++k) {
      volatile double lo = grid.lower[k] + double(idx[k]) * h[k];
      volatile double up = grid.lower[k] + double(idx[k] + 1) * h[k];
      lower[k] = lo;
      ext[k] = up - lo;
    }
    for (int q = 0; q < nq; ++q) {
      const int qk[3] = {q % m, (q / m) % m, q / (m * m)};
      for (int k = 0; k < d; ++k)
        x[k] = lower[k] + xh[qk[k]] * ext[k];
      double* dst = out.data() + (std::size_t(e) * nq + q) * comps;
      g.evaluate(x, dst);
      if (with_jacobian)
        g.jacobian(x, dst + 1);
    }
  }
}

// samples every GenericFunction lambda of the form at the points of the rule the reference integrates the form with
inline LoweredForm lower_form(const IntegrandTerms& t, int over_integrate, double scaling, int role, gdtb_space* space,
                              const gdtb_grid_desc& grid)
{
  LoweredForm out;
  out.form = make_form(t, over_integrate, scaling);
  bool any = false;
  for (const auto& g : t.generic)
    any = any || g.first || g.second;
  if (!any)
    return out;
  if (role == GDTB_ROLE_COUPLING || role == GDTB_ROLE_BOUNDARY)
    throw Dune::NotImplemented("GenericFunction lambdas in intersection integrands (use constants, per-element data or "
                               "the built-in functions there)");
  std::int32_t order = 0, m = 0;
  check(gdtb_form_quadrature_order(space, &out.form, role, &order)); // the placeholders carry the declared orders
  double xh[8], w[8];
  check(gdtb_gauss_rule(order, &m, xh, w));
  int nq = 1;
  for (int k = 0; k < grid.dim; ++k)
    nq *= m;
  for (std::size_t i = 0; i < t.terms.size(); ++i)
    for (int which = 0; which < 2; ++which) {
      const GenericPtr& g = which == 0 ? t.generic[i].first : t.generic[i].second;
      if (!g)
        continue;
      gdtb_function& fn = which == 0 ? out.form.terms[i].diffusion : out.form.terms[i].weight;
      auto samples = std::make_shared<std::vector<double>>();
      sample_generic(*g, grid, m, xh, false, *samples);
      fn.kind = g->comps == 1 ? GDTB_FN_QP_SCALAR : GDTB_FN_QP_TENSOR;
      fn.order = g->order;
      fn.qp_per_element = nq;
      fn.data = samples->data();
      fn.data_on_device = 0;
      out.samples.push_back(samples);
    }
  return out;
}
} // namespace internal

// local/integrands/interfaces.hh: binary element integrands, summable with operator+ (:233-236)
template <class E>
class LocalBinaryElementIntegrandInterface
{
public:
  explicit LocalBinaryElementIntegrandInterface(internal::IntegrandTerms terms = {})
    : terms_(std::move(terms))
  {}
  const internal::IntegrandTerms& terms() const
  {
    return terms_;
  }
  LocalBinaryElementIntegrandInterface operator+(const LocalBinaryElementIntegrandInterface& other) const
  {
    return LocalBinaryElementIntegrandInterface(internal::concat(terms_, other.terms_));
  }

protected:
  internal::IntegrandTerms terms_;
};

template <class E>
class LocalUnaryElementIntegrandInterface
{
public:
  explicit LocalUnaryElementIntegrandInterface(internal::IntegrandTerms terms = {})
    : terms_(std::move(terms))
  {}
  const internal::IntegrandTerms& terms() const
  {
    return terms_;
  }

protected:
  internal::IntegrandTerms terms_;
};

// local/integrands/laplace.hh:40-48
template <class E, std::size_t r = 1, class F = double>
class LocalLaplaceIntegrand : public LocalBinaryElementIntegrandInterface<E>
{
  static constexpr std::size_t d = E::dimension;

public:
  LocalLaplaceIntegrand(XT::Functions::GridFunction<E, d, d> diffusion = 1.)
  {
    internal::push_term(this->terms_, GDTB_INT_LAPLACE, &diffusion, (decltype(&diffusion)) nullptr, 0., GDTB_HI_DIAMETER);
    this->terms_.keep.push_back(std::make_shared<XT::Functions::GridFunction<E, d, d>>(diffusion));
  }
  // element-wise constant scalar diffusion (kappa_e * I)
  LocalLaplaceIntegrand(const XT::Functions::GridFunction<E>& diffusion, int /*scalar tag*/)
  {
    internal::push_term(this->terms_, GDTB_INT_LAPLACE, &diffusion, (decltype(&diffusion)) nullptr, 0., GDTB_HI_DIAMETER);
    this->terms_.keep.push_back(std::make_shared<XT::Functions::GridFunction<E>>(diffusion));
  }
};

// local/integrands/product.hh:56-65, 389-404; with_ansatz: integrands/interfaces.hh:225-229 + conversion.hh:42-124
template <class E, std::size_t r = 1, class TF = double, class F = double, class AF = double>
class LocalElementProductIntegrand : public LocalBinaryElementIntegrandInterface<E>
{
public:
  LocalElementProductIntegrand(XT::Functions::GridFunction<E> weight = 1.)
    : weight_(weight)
  {
    internal::push_term(this->terms_, GDTB_INT_PRODUCT, &weight_, (decltype(&weight_)) nullptr, 0., GDTB_HI_DIAMETER);
    this->terms_.keep.push_back(std::make_shared<XT::Functions::GridFunction<E>>(weight_));
  }
  LocalUnaryElementIntegrandInterface<E> with_ansatz(const XT::Functions::GridFunction<E>& function) const
  {
    internal::IntegrandTerms t;
    internal::push_term(t, GDTB_INT_PRODUCT, &weight_, &function, 0., GDTB_HI_DIAMETER);
    t.keep.push_back(std::make_shared<XT::Functions::GridFunction<E>>(weight_));
    t.keep.push_back(std::make_shared<XT::Functions::GridFunction<E>>(function));
    return LocalUnaryElementIntegrandInterface<E>(t);
  }

private:
  XT::Functions::GridFunction<E> weight_;
};
template <class E, std::size_t r = 1, class TF = double, class F = double, class AF = double>
using LocalProductIntegrand = LocalElementProductIntegrand<E, r, TF, F, AF>;

template <class I>
class LocalQuaternaryIntersectionIntegrandInterface
{
public:
  explicit LocalQuaternaryIntersectionIntegrandInterface(internal::IntegrandTerms terms = {})
    : terms_(std::move(terms))
  {}
  const internal::IntegrandTerms& terms() const
  {
    return terms_;
  }
  LocalQuaternaryIntersectionIntegrandInterface
  operator+(const LocalQuaternaryIntersectionIntegrandInterface& other) const // interfaces.hh:609-612
  {
    return LocalQuaternaryIntersectionIntegrandInterface(internal::concat(terms_, other.terms_));
  }

protected:
  internal::IntegrandTerms terms_;
};

template <class I>
class LocalBinaryIntersectionIntegrandInterface
{
public:
  explicit LocalBinaryIntersectionIntegrandInterface(internal::IntegrandTerms terms = {})
    : terms_(std::move(terms))
  {}
  const internal::IntegrandTerms& terms() const
  {
    return terms_;
  }
  LocalBinaryIntersectionIntegrandInterface
  operator+(const LocalBinaryIntersectionIntegrandInterface& other) const // interfaces.hh:480-483
  {
    return LocalBinaryIntersectionIntegrandInterface(internal::concat(terms_, other.terms_));
  }

protected:
  internal::IntegrandTerms terms_;
};

enum class IntersectionDiameter
{
  diameter = GDTB_HI_DIAMETER, // internal::default_intersection_diameter (ipdg.hh:27-38)
  volume = GDTB_HI_VOLUME      // [](const auto& intersection) { return intersection.geometry().volume(); }
};

namespace LocalLaplaceIPDGIntegrands {
// local/integrands/laplace-ipdg.hh:47-61
template <class I>
class InnerCoupling : public LocalQuaternaryIntersectionIntegrandInterface<I>
{
  using E = typename I::Entity;
  static constexpr std::size_t d = E::dimension;

public:
  InnerCoupling(const double& symmetry_prefactor,
                XT::Functions::GridFunction<E, d, d> diffusion,
                XT::Functions::GridFunction<E, d, d> weight_function = 1.)
  {
    internal::push_term(this->terms_, 
        GDTB_INT_IPDG_INNER_COUPLING, &diffusion, &weight_function, symmetry_prefactor, GDTB_HI_DIAMETER);
  }
};
// local/integrands/laplace-ipdg.hh:237-252 (binary use: the bilinear form on Dirichlet faces)
template <class I>
class DirichletCoupling : public LocalBinaryIntersectionIntegrandInterface<I>
{
  using E = typename I::Entity;
  static constexpr std::size_t d = E::dimension;

public:
  DirichletCoupling(const double& symmetry_prefactor, XT::Functions::GridFunction<E, d, d> diffusion)
  {
    internal::push_term(this->terms_, GDTB_INT_IPDG_DIRICHLET_COUPLING,
                                                     &diffusion,
                                                     (decltype(&diffusion)) nullptr,
                                                     symmetry_prefactor,
                                                     GDTB_HI_DIAMETER);
  }
};
} // namespace LocalLaplaceIPDGIntegrands

namespace LocalIPDGIntegrands {
// local/integrands/ipdg.hh:59-72
template <class I>
class InnerPenalty : public LocalQuaternaryIntersectionIntegrandInterface<I>
{
  using E = typename I::Entity;
  static constexpr std::size_t d = E::dimension;

public:
  InnerPenalty(const double& penalty,
               XT::Functions::GridFunction<E, d, d> weight_function = 1.,
               const IntersectionDiameter intersection_diameter = IntersectionDiameter::diameter)
  {
    internal::push_term(this->terms_, GDTB_INT_IPDG_INNER_PENALTY,
                                                     (decltype(&weight_function)) nullptr,
                                                     &weight_function,
                                                     penalty,
                                                     int(intersection_diameter));
  }
};
// local/integrands/ipdg.hh:201-213
template <class I>
class BoundaryPenalty : public LocalBinaryIntersectionIntegrandInterface<I>
{
  using E = typename I::Entity;
  static constexpr std::size_t d = E::dimension;

public:
  BoundaryPenalty(const double& penalty,
                  XT::Functions::GridFunction<E, d, d> weight_function = 1.,
                  const IntersectionDiameter intersection_diameter = IntersectionDiameter::diameter)
  {
    internal::push_term(this->terms_, GDTB_INT_IPDG_BOUNDARY_PENALTY,
                                                     (decltype(&weight_function)) nullptr,
                                                     &weight_function,
                                                     penalty,
                                                     int(intersection_diameter));
  }
};
} // namespace LocalIPDGIntegrands

// ---- local forms (local/bilinear-forms/integrals.hh, local/functionals/integrals.hh) ---------------------------
template <class E>
class LocalElementBilinearFormInterface
{
public:
  virtual ~LocalElementBilinearFormInterface() = default;
  virtual gdtb_form descriptor(double scaling) const = 0;
  // the descriptor with every GenericFunction lambda sampled at the rule of this form on `space`
  virtual internal::LoweredForm lowered(double scaling, gdtb_space* space, const gdtb_grid_desc& grid) const = 0;
};
template <class I>
class LocalCouplingIntersectionBilinearFormInterface
{
public:
  virtual ~LocalCouplingIntersectionBilinearFormInterface() = default;
  virtual gdtb_form descriptor(double scaling) const = 0;
  // the descriptor with every GenericFunction lambda sampled at the rule of this form on `space`
  virtual internal::LoweredForm lowered(double scaling, gdtb_space* space, const gdtb_grid_desc& grid) const = 0;
};
template <class I>
class LocalIntersectionBilinearFormInterface
{
public:
  virtual ~LocalIntersectionBilinearFormInterface() = default;
  virtual gdtb_form descriptor(double scaling) const = 0;
  // the descriptor with every GenericFunction lambda sampled at the rule of this form on `space`
  virtual internal::LoweredForm lowered(double scaling, gdtb_space* space, const gdtb_grid_desc& grid) const = 0;
};
template <class E>
class LocalElementFunctionalInterface
{
public:
  virtual ~LocalElementFunctionalInterface() = default;
  virtual gdtb_form descriptor() const = 0;
  virtual internal::LoweredForm lowered(gdtb_space* space, const gdtb_grid_desc& grid) const = 0;
};

// integrals.hh:52-63
template <class E, std::size_t r = 1>
class LocalElementIntegralBilinearForm : public LocalElementBilinearFormInterface<E>
{
public:
  LocalElementIntegralBilinearForm(const LocalBinaryElementIntegrandInterface<E>& integrand,
                                   const int over_integrate = 0)
    : terms_(integrand.terms())
    , over_integrate_(over_integrate)
  {}
  gdtb_form descriptor(double scaling) const override
  {
    return internal::make_form(terms_, over_integrate_, scaling);
  }
  internal::LoweredForm lowered(double scaling, gdtb_space* space, const gdtb_grid_desc& grid) const override
  {
    return internal::lower_form(terms_, over_integrate_, scaling, GDTB_ROLE_ELEMENT, space, grid);
  }

private:
  internal::IntegrandTerms terms_;
  int over_integrate_;
};

// integrals.hh:169-180
template <class I, std::size_t r = 1>
class LocalCouplingIntersectionIntegralBilinearForm : public LocalCouplingIntersectionBilinearFormInterface<I>
{
public:
  LocalCouplingIntersectionIntegralBilinearForm(const LocalQuaternaryIntersectionIntegrandInterface<I>& integrand,
                                                const int over_integrate = 0)
    : terms_(integrand.terms())
    , over_integrate_(over_integrate)
  {}
  gdtb_form descriptor(double scaling) const override
  {
    return internal::make_form(terms_, over_integrate_, scaling);
  }
  internal::LoweredForm lowered(double scaling, gdtb_space* space, const gdtb_grid_desc& grid) const override
  {
    return internal::lower_form(terms_, over_integrate_, scaling, GDTB_ROLE_COUPLING, space, grid);
  }

private:
  internal::IntegrandTerms terms_;
  int over_integrate_;
};

// integrals.hh:305-316
template <class I, std::size_t r = 1>
class LocalIntersectionIntegralBilinearForm : public LocalIntersectionBilinearFormInterface<I>
{
public:
  LocalIntersectionIntegralBilinearForm(const LocalBinaryIntersectionIntegrandInterface<I>& integrand,
                                        const int over_integrate = 0)
    : terms_(integrand.terms())
    , over_integrate_(over_integrate)
  {}
  gdtb_form descriptor(double scaling) const override
  {
    return internal::make_form(terms_, over_integrate_, scaling);
  }
  internal::LoweredForm lowered(double scaling, gdtb_space* space, const gdtb_grid_desc& grid) const override
  {
    return internal::lower_form(terms_, over_integrate_, scaling, GDTB_ROLE_BOUNDARY, space, grid);
  }

private:
  internal::IntegrandTerms terms_;
  int over_integrate_;
};

// local/functionals/integrals.hh:41-52
template <class E, std::size_t r = 1>
class LocalElementIntegralFunctional : public LocalElementFunctionalInterface<E>
{
public:
  LocalElementIntegralFunctional(const LocalUnaryElementIntegrandInterface<E>& integrand, const int over_integrate = 0)
    : terms_(integrand.terms())
    , over_integrate_(over_integrate)
  {}
  gdtb_form descriptor() const override
  {
    return internal::make_form(terms_, over_integrate_, 1.);
  }
  internal::LoweredForm lowered(gdtb_space* space, const gdtb_grid_desc& grid) const override
  {
    return internal::lower_form(terms_, over_integrate_, 1., GDTB_ROLE_FUNCTIONAL, space, grid);
  }

private:
  internal::IntegrandTerms terms_;
  int over_integrate_;
};

// ---- VectorBasedFunctional (functionals/vector-based.hh:133-286) ----------------------------------------------
template <class V, class GV>
class VectorBasedFunctional
{
public:
  using ElementType = typename GV::Element;

  explicit VectorBasedFunctional(const SpaceInterface<GV>& space)
    : space_(space)
    , vector_(std::make_shared<V>(space.mapper().size()))
  {
    gdtb_vecfun* raw = nullptr;
    internal::check(gdtb_vecfun_create(internal::context(), space_.handle(), &raw));
    handle_ = internal::Handle<gdtb_vecfun, gdtb_vecfun_destroy>(raw);
    vector_->bind_device(raw);
  }
  // vector-based.hh:214-222
  VectorBasedFunctional& append(const LocalElementFunctionalInterface<ElementType>& local_functional,
                                const XT::Common::Parameter& /*param*/ = {},
                                const XT::Grid::ElementFilter<GV>& /*filter*/ = XT::Grid::ApplyOn::AllElements<GV>())
  {
    // cloned by the library (samples of GenericFunction lambdas included) before `f` goes out of scope
    const internal::LoweredForm f = local_functional.lowered(space_.handle(), space_.grid_view().desc());
    internal::check(gdtb_vecfun_append_element(handle_.get(), &f.form));
    return *this;
  }
  // vector-based.hh:276-279
  void assemble(const bool /*use_tbb*/ = false)
  {
    internal::check(gdtb_assemble(nullptr, handle_.get(), mode()));
    finalize();
  }
  V& vector()
  {
    return *vector_;
  }
  const V& vector() const
  {
    return *vector_;
  }
  const SpaceInterface<GV>& source_space() const
  {
    return space_;
  }
  // -- used by the walker / MatrixOperator to fuse the functional into their grid walk
  gdtb_vecfun* handle() const
  {
    return handle_.get();
  }
  int mode() const
  {
    return fresh_ ? GDTB_ASSEMBLE_OVERWRITE : GDTB_ASSEMBLE_ACCUMULATE;
  }
  void finalize()
  {
    fresh_ = false;
    internal::check(gdtb_vecfun_download(handle_.get(), vector_->data()));
    internal::check(gdtb_vecfun_clear_forms(handle_.get()));
  }

private:
  SpaceInterface<GV> space_;
  std::shared_ptr<V> vector_;
  internal::Handle<gdtb_vecfun, gdtb_vecfun_destroy> handle_;
  bool fresh_ = true;
};

template <class V, class GV>
VectorBasedFunctional<V, GV> make_vector_functional(const SpaceInterface<GV>& space)
{
  return VectorBasedFunctional<V, GV>(space);
}

// ---- MatrixOperator (operators/matrix-based.hh:245-508) --------------------------------------------------------
template <class M, class GV>
class MatrixOperator
{
public:
  using MatrixType = M;
  using E = typename GV::Element;
  using I = typename GV::Intersection;
  using FieldType = double;

  // ctor which creates an appropriate matrix from a sparsity pattern (matrix-based.hh:298-311)
  MatrixOperator(const GV& /*assembly_grid_view*/,
                 const SpaceInterface<GV>& source_space,
                 const SpaceInterface<GV>& range_space,
                 const XT::LA::SparsityPatternDefault& pattern)
    : scaling(1.)
    , source_space_(source_space)
    , range_space_(range_space)
    , pattern_(pattern)
    , matrix_(std::make_shared<M>(range_space.mapper().size(), source_space.mapper().size(), pattern))
  {
    gdtb_matop* raw = nullptr;
    // rows = range (test) space, cols = source (ansatz) space (matrix-based.hh:73-80, 361-366)
    internal::check(
        gdtb_matop_create(internal::context(), range_space_.handle(), source_space_.handle(), pattern_.handle(), &raw));
    handle_ = internal::Handle<gdtb_matop, gdtb_matop_destroy>(raw);
    matrix_->bind_device(raw);
  }

  FieldType scaling; // captured by value at append time (matrix-based.hh:342,365)

  // matrix-based.hh:346-369
  MatrixOperator& append(const LocalElementBilinearFormInterface<E>& local_bilinear_form,
                         const XT::Common::Parameter& /*param*/ = {},
                         const XT::Grid::ElementFilter<GV>& /*filter*/ = XT::Grid::ApplyOn::AllElements<GV>())
  {
    const internal::LoweredForm f =
        local_bilinear_form.lowered(scaling, range_space_.handle(), range_space_.grid_view().desc());
    internal::check(gdtb_matop_append_element(handle_.get(), &f.form));
    return *this;
  }
  MatrixOperator& operator+=(const LocalElementBilinearFormInterface<E>& local_bilinear_form) // :450-455
  {
    return append(local_bilinear_form);
  }
  // matrix-based.hh:371-393
  MatrixOperator& append(const LocalCouplingIntersectionBilinearFormInterface<I>& local_bilinear_form,
                         const XT::Common::Parameter& /*param*/ = {},
                         const XT::Grid::IntersectionFilter<GV>& filter = XT::Grid::ApplyOn::AllIntersections<GV>())
  {
    const gdtb_form f =
        local_bilinear_form.lowered(scaling, range_space_.handle(), range_space_.grid_view().desc()).form;
    const int flt = filter.gdtb_filter();
    if (flt != GDTB_FILTER_INNER_ONCE && flt != GDTB_FILTER_INNER_AND_PERIODIC_ONCE)
      throw Dune::NotImplemented("coupling forms are supported with ApplyOn::InnerIntersectionsOnce (and the "
                                 "periodic variant) only");
    internal::check(gdtb_matop_append_coupling(handle_.get(), &f, flt));
    return *this;
  }
  // matrix-based.hh:395-408
  MatrixOperator& append(const LocalIntersectionBilinearFormInterface<I>& local_bilinear_form,
                         const XT::Common::Parameter& /*param*/ = {},
                         const XT::Grid::IntersectionFilter<GV>& filter = XT::Grid::ApplyOn::AllIntersections<GV>())
  {
    const gdtb_form f =
        local_bilinear_form.lowered(scaling, range_space_.handle(), range_space_.grid_view().desc()).form;
    if (filter.gdtb_filter() != GDTB_FILTER_ALL_BOUNDARY)
      throw Dune::NotImplemented("boundary forms are supported with CustomBoundaryIntersections(AllDirichlet...) only");
    internal::check(gdtb_matop_append_boundary(handle_.get(), &f, GDTB_FILTER_ALL_BOUNDARY));
    return *this;
  }
  // operator-as-walker: assemble the functional in the same grid walk (examples/adaptive_elliptic_swipdg.cc:250)
  template <class V>
  MatrixOperator& append(VectorBasedFunctional<V, GV>& functional)
  {
    riders_.push_back([&functional]() { return functional.handle(); });
    modes_.push_back([&functional]() { return functional.mode(); });
    finalizers_.push_back([&functional]() { functional.finalize(); });
    return *this;
  }

  // matrix-based.hh:496-500: one grid walk; afterwards the functor list is empty (dune-xt clears it after walk)
  void assemble(const bool /*use_tbb*/ = false)
  {
    walk();
  }
  void walk(const bool /*thread_parallel*/ = false)
  {
    if (riders_.size() > 1)
      throw Dune::NotImplemented("at most one functional can ride along with a matrix operator");
    gdtb_vecfun* fun = riders_.empty() ? nullptr : riders_[0]();
    if (fun && modes_[0]() != mode()) {
      internal::check(gdtb_assemble(handle_.get(), nullptr, mode()));
      internal::check(gdtb_assemble(nullptr, fun, modes_[0]()));
    } else
      internal::check(gdtb_assemble(handle_.get(), fun, mode()));
    finalize();
    for (auto& fin : finalizers_)
      fin();
    riders_.clear();
    modes_.clear();
    finalizers_.clear();
  }

  M& matrix()
  {
    return *matrix_;
  }
  const M& matrix() const
  {
    return *matrix_;
  }
  const SpaceInterface<GV>& source_space() const
  {
    return source_space_;
  }
  const SpaceInterface<GV>& range_space() const
  {
    return range_space_;
  }
  // device-side access for callers that keep the matrix on the GPU
  double* device_values() const
  {
    double* p = nullptr;
    internal::check(gdtb_matop_values_device(handle_.get(), &p));
    return p;
  }
  gdtb_matop* handle() const
  {
    return handle_.get();
  }
  int mode() const
  {
    return fresh_ ? GDTB_ASSEMBLE_OVERWRITE : GDTB_ASSEMBLE_ACCUMULATE;
  }
  void finalize()
  {
    fresh_ = false;
    internal::check(gdtb_matop_values_download(handle_.get(), matrix_->values().data()));
    internal::check(gdtb_matop_clear_forms(handle_.get()));
  }

private:
  SpaceInterface<GV> source_space_, range_space_;
  XT::LA::SparsityPatternDefault pattern_;
  std::shared_ptr<M> matrix_;
  internal::Handle<gdtb_matop, gdtb_matop_destroy> handle_;
  std::vector<std::function<gdtb_vecfun*()>> riders_;
  std::vector<std::function<int()>> modes_;
  std::vector<std::function<void()>> finalizers_;
  bool fresh_ = true;
};

// make_matrix_operator<M>(space, stencil) (matrix-based.hh:650-658)
template <class M, class GV>
MatrixOperator<M, GV> make_matrix_operator(const SpaceInterface<GV>& space, const Stencil stencil = Stencil::element)
{
  return MatrixOperator<M, GV>(
      space.grid_view(), space, space, make_sparsity_pattern(space, space, space.grid_view(), stencil));
}
// make_matrix_operator<M>(view, source_space, range_space, pattern) (matrix-based.hh:514-560)
template <class M, class GV>
MatrixOperator<M, GV> make_matrix_operator(const GV& grid_view,
                                           const SpaceInterface<GV>& source_space,
                                           const SpaceInterface<GV>& range_space,
                                           const XT::LA::SparsityPatternDefault& pattern)
{
  return MatrixOperator<M, GV>(grid_view, source_space, range_space, pattern);
}

// ---- DirichletConstraints (tools/dirichlet-constraints.hh:44-230) -------------------------------------------------
// The Dirichlet DoFs are collected on the device when the object is created (the reference collects them during the
// grid walk: appending the object to a walker is accepted and does nothing); apply() works on the device operator /
// functional behind the host containers and refreshes the host copies.
template <class GV>
class DirichletConstraints
{
public:
  using I = typename GV::Intersection;
  DirichletConstraints(const SpaceInterface<GV>& space, const XT::Grid::AllDirichletBoundaryInfo<I>& /*boundary_info*/)
    : space_(space)
  {
    gdtb_dirichlet* raw = nullptr;
    internal::check(gdtb_dirichlet_create(internal::context(), space_.handle(), GDTB_BOUNDARY_ALL, &raw));
    handle_ = internal::Handle<gdtb_dirichlet, gdtb_dirichlet_destroy>(raw);
  }
  // dirichlet_DoFs() (:112-115)
  std::set<std::size_t> dirichlet_DoFs() const
  {
    std::vector<std::int64_t> dofs(std::size_t(gdtb_dirichlet_size(handle_.get())));
    internal::check(gdtb_dirichlet_dofs_download(handle_.get(), dofs.data()));
    return std::set<std::size_t>(dofs.begin(), dofs.end());
  }
  // apply(matrix, vector, only_clear, ensure_symmetry) (:122-184)
  template <class M, class V>
  void apply(M& matrix, V& vector, const bool only_clear = false, const bool ensure_symmetry = true) const
  {
    if (!matrix.device() || !vector.device())
      throw Exceptions::device_error("matrix / vector are not attached to a device operator / functional");
    if (matrix.rows() != space_.mapper().size() || vector.size() != space_.mapper().size())
      throw XT::Common::Exceptions::shapes_do_not_match("matrix / vector do not match the constrained space"); // :125-140
    internal::check(gdtb_dirichlet_apply(handle_.get(), matrix.device(), vector.device(), only_clear ? 1 : 0,
                                         ensure_symmetry ? 1 : 0));
    internal::check(gdtb_matop_values_download(matrix.device(), matrix.values().data()));
    internal::check(gdtb_vecfun_download(vector.device(), vector.data()));
  }

private:
  SpaceInterface<GV> space_;
  internal::Handle<gdtb_dirichlet, gdtb_dirichlet_destroy> handle_;
};

template <class GV, class Info>
DirichletConstraints<GV> make_dirichlet_constraints(const SpaceInterface<GV>& space, const Info& boundary_info)
{
  return DirichletConstraints<GV>(space, boundary_info);
}

// ---- DiscreteFunction / error norms (discretefunction/default.hh, operators/bilinear-form.hh:35-470) ---------------
template <class V, class GV>
class DiscreteFunction
{
public:
  explicit DiscreteFunction(const SpaceInterface<GV>& space)
    : space_(space)
    , vector_(space.mapper().size())
  {}
  struct Dofs
  {
    V& v;
    V& vector()
    {
      return v;
    }
  };
  Dofs dofs()
  {
    return Dofs{vector_};
  }
  const V& dof_vector() const
  {
    return vector_;
  }
  const SpaceInterface<GV>& space() const
  {
    return space_;
  }

private:
  SpaceInterface<GV> space_;
  V vector_;
};

template <class V, class GV>
DiscreteFunction<V, GV> make_discrete_function(const SpaceInterface<GV>& space)
{
  return DiscreteFunction<V, GV>(space);
}

// `solution - exact_solution` (examples/stationary-heat-equation.cc:114): u_h - f with f a GridFunction
template <class V, class GV>
struct DifferenceFunction
{
  const DiscreteFunction<V, GV>& u_h;
  XT::Functions::GridFunction<typename GV::Element> f;
};
template <class V, class GV>
DifferenceFunction<V, GV> operator-(const DiscreteFunction<V, GV>& u_h, const XT::Functions::GridFunction<typename GV::Element>& f)
{
  return DifferenceFunction<V, GV>{u_h, f};
}

// make_bilinear_form(grid_view, error, error) += LocalElementIntegralBilinearForm(...); result() (bilinear-form.hh:
// 35-470): sum over the elements of the local form applied to (e, e), evaluated on the device when the walker runs
template <class V, class GV>
class BilinearForm
{
public:
  using E = typename GV::Element;
  explicit BilinearForm(const DifferenceFunction<V, GV>& e)
    : e_(e)
  {}
  BilinearForm& operator+=(const LocalElementBilinearFormInterface<E>& local_form)
  {
    forms_.push_back(local_form.descriptor(1.));
    return *this;
  }
  void assemble(const bool /*use_tbb*/ = false)
  {
    result_ = 0.;
    for (const gdtb_form& f : forms_) {
      double r = 0.;
      gdtb_function fn = e_.f.descriptor();
      std::vector<double> samples;
      if (e_.f.generic()) {
        // an exact solution given by lambdas (value + jacobian, examples/stationary-heat-equation.cc:71-85): sampled at
        // the points of the rule this form is integrated with
        const auto& g = *e_.f.generic();
        if (g.comps != 1 || !g.jacobian)
          throw Dune::NotImplemented("norms against a GenericFunction need a scalar function with a jacobian lambda");
        const gdtb_grid_desc& grid = e_.u_h.space().grid_view().desc();
        std::int32_t order = 0, m = 0;
        internal::check(gdtb_bilinear_form_quadrature_order(e_.u_h.space().handle(), 1, g.order, &f, &order));
        double xh[8], w[8];
        internal::check(gdtb_gauss_rule(order, &m, xh, w));
        internal::sample_generic(g, grid, m, xh, true, samples);
        fn = gdtb_function{};
        fn.kind = GDTB_FN_QP_VALUE_GRAD;
        fn.order = g.order;
        fn.qp_per_element = 1;
        for (int k = 0; k < grid.dim; ++k)
          fn.qp_per_element *= m;
        fn.data = samples.data();
      }
      internal::check(gdtb_bilinear_form_apply2_host(internal::context(), e_.u_h.space().handle(),
                                                     e_.u_h.dof_vector().data(), &fn, &f, &r));
      result_ += r;
    }
    assembled_ = true;
  }
  double result() const
  {
    if (!assembled_)
      throw Dune::InvalidStateException("BilinearForm::result() before the grid walk");
    return result_;
  }

private:
  DifferenceFunction<V, GV> e_;
  std::vector<gdtb_form> forms_;
  double result_ = 0.;
  bool assembled_ = false;
};

template <class V, class GV>
BilinearForm<V, GV> make_bilinear_form(const GV& /*grid_view*/, const DifferenceFunction<V, GV>& source,
                                       const DifferenceFunction<V, GV>& /*range: the same error function*/)
{
  return BilinearForm<V, GV>(source);
}

} // namespace GDT

namespace XT {
namespace Grid {

// XT::Grid::Walker: collects operators / functionals and assembles them in ONE grid walk
// (examples/stationary-heat-equation.cc:102-106)
template <class GV>
class Walker
{
public:
  explicit Walker(const GV& grid_view)
    : grid_view_(grid_view)
  {}
  template <class M>
  Walker& append(GDT::MatrixOperator<M, GV>& op)
  {
    if (op_walk_)
      throw Dune::NotImplemented("one matrix operator per walk");
    op_handle_ = [&op]() { return op.handle(); };
    op_mode_ = [&op]() { return op.mode(); };
    op_walk_ = [&op]() { op.finalize(); };
    return *this;
  }
  template <class V>
  Walker& append(GDT::VectorBasedFunctional<V, GV>& fun)
  {
    if (fun_walk_)
      throw Dune::NotImplemented("one functional per walk");
    fun_handle_ = [&fun]() { return fun.handle(); };
    fun_mode_ = [&fun]() { return fun.mode(); };
    fun_walk_ = [&fun]() { fun.finalize(); };
    return *this;
  }
  // DirichletConstraints collect their DoFs on the device at construction: nothing left to do during the walk
  Walker& append(GDT::DirichletConstraints<GV>& /*constraints*/)
  {
    return *this;
  }
  template <class V>
  Walker& append(GDT::BilinearForm<V, GV>& form)
  {
    others_.push_back([&form]() { form.assemble(); });
    return *this;
  }
  void walk(const bool /*thread_parallel*/ = false)
  {
    for (auto& other : others_)
      other();
    others_.clear();
    gdtb_matop* op = op_walk_ ? op_handle_() : nullptr;
    gdtb_vecfun* fun = fun_walk_ ? fun_handle_() : nullptr;
    if (op && fun && op_mode_() != fun_mode_()) {
      GDT::internal::check(gdtb_assemble(op, nullptr, op_mode_()));
      GDT::internal::check(gdtb_assemble(nullptr, fun, fun_mode_()));
    } else if (op || fun)
      GDT::internal::check(gdtb_assemble(op, fun, op ? op_mode_() : fun_mode_()));
    if (op_walk_)
      op_walk_();
    if (fun_walk_)
      fun_walk_();
    op_walk_ = nullptr; // walk(thread_parallel, clear_functors = true)
    fun_walk_ = nullptr;
  }

private:
  GV grid_view_;
  std::function<gdtb_matop*()> op_handle_;
  std::function<gdtb_vecfun*()> fun_handle_;
  std::function<int()> op_mode_, fun_mode_;
  std::function<void()> op_walk_, fun_walk_;
  std::vector<std::function<void()>> others_;
};

template <class GV>
Walker<GV> make_walker(const GV& grid_view)
{
  return Walker<GV>(grid_view);
}

} // namespace Grid
} // namespace XT

namespace GDT {

// ---- finite volumes (local/numerical-fluxes/upwind.hh, operators/advection-fv.hh) ------------------------------
// stand-ins for the GenericFunction flux lambdas of the reference drivers
struct LinearFlux // f(u) = a u (test/linear-transport/base.hh:50-57)
{
  std::array<double, 3> direction{{1., 0., 0.}};
};
struct BurgersFlux // f(u) = u^2/2 (1,...,1) (test/burgers/base.hh:38-44)
{};

template <class I, std::size_t d, std::size_t m = 1>
class NumericalFluxInterface
{
public:
  const gdtb_flux& descriptor() const
  {
    return flux_;
  }

protected:
  gdtb_flux flux_{};
};

// local/numerical-fluxes/upwind.hh:44-50
template <class I, std::size_t d, std::size_t m = 1>
class NumericalUpwindFlux : public NumericalFluxInterface<I, d, m>
{
  static_assert(m == 1, "the upwind flux is only available for scalar equations (upwind.hh:24-28)");

public:
  NumericalUpwindFlux(const LinearFlux& f)
  {
    this->flux_.kind = GDTB_FLUX_LINEAR;
    this->flux_.numflux = GDTB_NUMFLUX_UPWIND;
    for (std::size_t k = 0; k < d; ++k)
      this->flux_.p[k] = f.direction[k];
  }
  NumericalUpwindFlux(const BurgersFlux&)
  {
    this->flux_.kind = GDTB_FLUX_BURGERS;
    this->flux_.numflux = GDTB_NUMFLUX_UPWIND;
  }
};

// local/numerical-fluxes/lax-friedrichs.hh:33-58
template <class I, std::size_t d, std::size_t m = 1>
class NumericalLaxFriedrichsFlux : public NumericalFluxInterface<I, d, m>
{
public:
  NumericalLaxFriedrichsFlux(const LinearFlux& f)
  {
    this->flux_.kind = GDTB_FLUX_LINEAR;
    this->flux_.numflux = GDTB_NUMFLUX_LAX_FRIEDRICHS;
    for (std::size_t k = 0; k < d; ++k)
      this->flux_.p[k] = f.direction[k];
  }
  NumericalLaxFriedrichsFlux(const BurgersFlux&)
  {
    this->flux_.kind = GDTB_FLUX_BURGERS;
    this->flux_.numflux = GDTB_NUMFLUX_LAX_FRIEDRICHS;
  }
  NumericalLaxFriedrichsFlux() = default;
  void set_euler(const double gamma, const double lambda)
  {
    this->flux_.kind = GDTB_FLUX_EULER;
    this->flux_.numflux = GDTB_NUMFLUX_LAX_FRIEDRICHS;
    this->flux_.p[0] = gamma;
    this->flux_.p[1] = lambda;
  }
};

// tools/euler.hh: EulerTools<d> -- conversions on the host; flux, jacobian and eigendecomposition live in the kernels
// (the GenericFunction lambdas `euler_tools.flux(w)` / `flux_jacobian(w)` and the flux_eigen_decomposition lambda of the
// reference's 2d_euler driver are exactly these, examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:387-409)
template <std::size_t d, class R = double>
class EulerTools
{
  static_assert(d == 1 || d == 2, "tools/euler.hh: flux jacobian and eigendecomposition exist for d = 1, 2");

public:
  static constexpr std::size_t m = d + 2;
  explicit EulerTools(const double gmma)
    : gamma_(gmma)
  {}
  int flux_order() const
  {
    return 4;
  }
  double gamma() const
  {
    return gamma_;
  }
  // (rho, rho v, E), E = p / (gamma - 1) + rho |v|^2 / 2 (euler.hh:127-131, 153-157); scalar velocity: every component
  FieldVector<R, int(m)> conservative(const R rho, const FieldVector<R, int(d)>& v, const R p) const
  {
    FieldVector<R, int(m)> w{};
    R v2 = 0;
    w[0] = rho;
    for (std::size_t i = 0; i < d; ++i) {
      w[1 + i] = rho * v[i];
      v2 += v[i] * v[i];
    }
    w[m - 1] = p / (gamma_ - 1.) + 0.5 * rho * v2;
    return w;
  }
  FieldVector<R, int(m)> conservative(const R rho, const R velocity, const R p) const
  {
    FieldVector<R, int(d)> v{};
    for (auto& c : v)
      c = velocity;
    return conservative(rho, v, p);
  }
  R density(const FieldVector<R, int(m)>& w) const
  {
    return w[0];
  }
  R pressure(const FieldVector<R, int(m)>& w) const
  {
    R v2 = 0;
    for (std::size_t i = 0; i < d; ++i)
      v2 += (w[1 + i] / w[0]) * (w[1 + i] / w[0]);
    return (gamma_ - 1.) * (w[m - 1] - 0.5 * w[0] * v2);
  }

private:
  double gamma_;
};

// the Euler flux as a value (stands for the GenericFunction<m, d, m> built from EulerTools::flux / flux_jacobian)
struct EulerFlux
{
  double gamma;
};
template <std::size_t d>
EulerFlux make_euler_flux(const EulerTools<d>& tools)
{
  return EulerFlux{tools.gamma()};
}

// local/numerical-fluxes/vijayasundaram.hh:29-153, eigendecomposition from EulerTools
template <class I, std::size_t d, std::size_t m>
class NumericalVijayasundaramFlux : public NumericalFluxInterface<I, d, m>
{
  static_assert(m == d + 2, "the Euler equations have m = d + 2 components");

public:
  NumericalVijayasundaramFlux(const EulerFlux& f)
  {
    this->flux_.kind = GDTB_FLUX_EULER;
    this->flux_.numflux = GDTB_NUMFLUX_VIJAYASUNDARAM;
    this->flux_.p[0] = f.gamma;
  }
};

// local/numerical-fluxes/lax-friedrichs.hh:33-58 for systems: lambda has to be provided (:40-41)
template <class I, std::size_t d, std::size_t m>
NumericalLaxFriedrichsFlux<I, d, m> make_numerical_lax_friedrichs_flux(const EulerFlux& f, const double lambda)
{
  NumericalLaxFriedrichsFlux<I, d, m> g;
  g.set_euler(f.gamma, lambda);
  return g;
}

// Boundary treatments (local/operators/advection-fv.hh:188-457) as value types: v = a u + b / g = a (f(u) . n) + b
struct BoundaryTreatmentByCustomExtrapolation
{
  double a, b;
};
struct BoundaryTreatmentByCustomNumericalFlux
{
  double a, b;
};

// operators/advection-fv.hh:44-128
template <class M, class GV>
class AdvectionFvOperator
{
public:
  using V = XT::LA::IstlDenseVector<double>;
  static constexpr std::size_t d = GV::dimension;

  template <class I, std::size_t m>
  AdvectionFvOperator(const GV& /*assembly_grid_view*/,
                      const NumericalFluxInterface<I, d, m>& numerical_flux,
                      const SpaceInterface<GV>& source_space,
                      const SpaceInterface<GV>& range_space)
    : source_space_(source_space)
    , range_space_(range_space)
  {
    if (source_space.type() != SpaceType::finite_volume || range_space.type() != SpaceType::finite_volume)
      throw Exceptions::operator_error("Use LocalAdvectionDgCouplingOperator instead!"); // advection-fv.hh:131-134
    gdtb_fvop* raw = nullptr;
    internal::check(gdtb_fvop_create(internal::context(), source_space_.handle(), &numerical_flux.descriptor(), &raw));
    handle_ = internal::Handle<gdtb_fvop, gdtb_fvop_destroy>(raw);
  }
  // LocalizableOperator::apply(source, range, param) (operators/localizable-operator.hh:382-387)
  void apply(const V& source, V& range, const XT::Common::Parameter& /*param*/ = {}) const
  {
    if (source.size() != source_space_.mapper().size() || range.size() != range_space_.mapper().size())
      throw XT::Common::Exceptions::shapes_do_not_match("vector sizes do not match the finite volume spaces");
    internal::check(gdtb_fvop_apply_host(handle_.get(), source.data(), range.data()));
  }
  // OperatorInterface::apply(source, param) (operators/interfaces.hh:645-650)
  V apply(const V& source, const XT::Common::Parameter& param = {}) const
  {
    V range(range_space_.mapper().size());
    apply(source, range, param);
    return range;
  }
  // append(boundary treatment, param_type, filter) (operators/advection-fv.hh:96-123).  The reference takes a C++
  // lambda and an intersection filter; across the C ABI the treatment is one of the affine families of
  // gdtb_fv_boundary and the filter a mask of domain sides (bit 2k+s: outer normal -e_k / +e_k).
  AdvectionFvOperator& append(const BoundaryTreatmentByCustomExtrapolation& t, const unsigned side_mask = ~0u)
  {
    gdtb_fv_boundary b{GDTB_FVBND_EXTRAPOLATION, side_mask & ((1u << (2 * d)) - 1u), t.a, t.b};
    internal::check(gdtb_fvop_append_boundary(handle_.get(), &b));
    return *this;
  }
  AdvectionFvOperator& append(const BoundaryTreatmentByCustomNumericalFlux& t, const unsigned side_mask = ~0u)
  {
    gdtb_fv_boundary b{GDTB_FVBND_NUMERICAL_FLUX, side_mask & ((1u << (2 * d)) - 1u), t.a, t.b};
    internal::check(gdtb_fvop_append_boundary(handle_.get(), &b));
    return *this;
  }
  const SpaceInterface<GV>& source_space() const
  {
    return source_space_;
  }
  const SpaceInterface<GV>& range_space() const
  {
    return range_space_;
  }
  gdtb_fvop* handle() const
  {
    return handle_.get();
  }

private:
  SpaceInterface<GV> source_space_, range_space_;
  internal::Handle<gdtb_fvop, gdtb_fvop_destroy> handle_;
};

// estimate_dt_for_hyperbolic_system(grid_view, state, flux, boundary_data_range) (tools/hyperbolic.hh:38-86); grid view
// and flux are the operator's, the state is a finite-volume DoF vector
template <class M, class GV>
double estimate_dt_for_hyperbolic_system(const AdvectionFvOperator<M, GV>& op,
                                         const XT::LA::IstlDenseVector<double>& state,
                                         const std::pair<double, double>* boundary_data_range = nullptr)
{
  if (state.size() != op.source_space().mapper().size())
    throw XT::Common::Exceptions::shapes_do_not_match("state vector does not match the finite volume space");
  double range[2] = {0., 0.}, dt = 0.;
  if (boundary_data_range) {
    range[0] = boundary_data_range->first;
    range[1] = boundary_data_range->second;
  }
  internal::check(gdtb_fv_estimate_dt_host(op.handle(), state.data(), boundary_data_range ? range : nullptr, &dt));
  return dt;
}

// tools/timestepper/interface.hh: TimeStepperMethods (explicit Runge-Kutta members)
enum class TimeStepperMethods
{
  explicit_euler = GDTB_RK_EULER,
  explicit_rungekutta_second_order_ssp = GDTB_RK_SSP2,
  explicit_rungekutta_third_order_ssp = GDTB_RK_SSP3,
  explicit_rungekutta_classic_fourth_order = GDTB_RK_CLASSIC4,
  explicit_rungekutta_other = GDTB_RK_OTHER
};

// tools/timestepper/explicit-rungekutta.hh:158-270: u_t = r L(u).  Like the reference the stepper works on the
// caller's initial_values vector in place (current_solution() is that vector).
template <class OperatorImp, TimeStepperMethods method = TimeStepperMethods::explicit_euler>
class ExplicitRungeKuttaTimeStepper
{
public:
  using V = XT::LA::IstlDenseVector<double>;

  ExplicitRungeKuttaTimeStepper(const OperatorImp& op,
                                V& initial_values,
                                const double r = 1.0,
                                const double t_0 = 0.0,
                                const std::vector<std::vector<double>>& A = {},
                                const std::vector<double>& b = {},
                                const std::vector<double>& c = {})
    : u_(initial_values)
  {
    if (initial_values.size() != op.source_space().mapper().size())
      throw XT::Common::Exceptions::shapes_do_not_match("initial values do not match the operator's source space");
    std::vector<double> flat;
    for (const auto& row : A) {
      if (row.size() != A.size())
        throw XT::Common::Exceptions::shapes_do_not_match("A has to be a square matrix"); // :210
      flat.insert(flat.end(), row.begin(), row.end());
    }
    if (b.size() != A.size() || c.size() != A.size())
      throw XT::Common::Exceptions::shapes_do_not_match("b and c must have as many entries as A has rows"); // :211-212
    gdtb_rk* raw = nullptr;
    internal::check(gdtb_rk_create(op.handle(), int(method), int(A.size()), A.empty() ? nullptr : flat.data(),
                                   A.empty() ? nullptr : b.data(), A.empty() ? nullptr : c.data(), r, t_0, &raw));
    handle_ = internal::Handle<gdtb_rk, gdtb_rk_destroy>(raw);
  }
  double current_time() const
  {
    return gdtb_rk_current_time(handle_.get());
  }
  V& current_solution()
  {
    return u_;
  }
  // step(dt, max_dt) (:237-270)
  double step(const double dt, const double max_dt)
  {
    double ret = dt;
    internal::check(gdtb_rk_step_host(handle_.get(), u_.data(), dt, max_dt, &ret));
    return ret;
  }
  // TimeStepperInterface::solve(t_end, initial_dt) (tools/timestepper/interface.hh:191-263) with nothing saved,
  // visualized or printed; returns the dt for the next step
  double solve(const double t_end, const double initial_dt)
  {
    std::int64_t steps = 0;
    double next = initial_dt;
    internal::check(gdtb_rk_solve_host(handle_.get(), u_.data(), t_end, initial_dt, &steps, &next));
    num_steps_ = steps;
    return next;
  }
  std::int64_t num_steps() const
  {
    return num_steps_;
  }

private:
  V& u_;
  internal::Handle<gdtb_rk, gdtb_rk_destroy> handle_;
  std::int64_t num_steps_ = 0;
};

// operators/advection-fv.hh:130-141
template <class M, class GV, class I, std::size_t d, std::size_t m>
AdvectionFvOperator<M, GV> make_advection_fv_operator(const GV& assembly_grid_view,
                                                      const NumericalFluxInterface<I, d, m>& numerical_flux,
                                                      const SpaceInterface<GV>& source_space,
                                                      const SpaceInterface<GV>& range_space)
{
  return AdvectionFvOperator<M, GV>(assembly_grid_view, numerical_flux, source_space, range_space);
}

// default_interpolation<V>(order, f, fv_space) (interpolations/default.hh:76-83): the FV "interpolation" is the cell
// average by a Gauss rule of the declared order (spaces/basis/finite-volume.hh:244-252); `f` carries its order
template <class V, class GV>
V default_interpolation(const XT::Functions::GridFunction<typename GV::Element>& f, const SpaceInterface<GV>& fv_space)
{
  V u(std::size_t(fv_space.mapper().size()), 0.);
  internal::check(gdtb_fv_interpolate_host(internal::context(), fv_space.handle(), &f.descriptor(), u.data()));
  return u;
}

// default_interpolation<V>(order, lambda, fv_space) for a vector-valued lambda x -> FieldVector<double, m>
// (interpolations/default.hh:76-83 with spaces/basis/finite-volume.hh:244-252: cell averages by the Gauss rule of the
// given order; order 0 = the value at the cell centre).  Evaluated on the host (a lambda cannot cross the C ABI), the
// DoF vector is [element][component].
template <class V, class GV, class F>
auto default_interpolation(const int order, F&& f, const SpaceInterface<GV>& fv_space)
    -> decltype(f(std::declval<FieldVector<double, int(GV::dimension)>>(), XT::Common::Parameter{}), V())
{
  constexpr int d = int(GV::dimension);
  const gdtb_grid_desc& g = fv_space.grid_view().desc();
  const int m = fv_space.range_dim();
  std::int32_t nq = 0;
  std::vector<double> qx(16), qw(16);
  internal::check(gdtb_gauss_rule(order, &nq, qx.data(), qw.data()));
  std::int64_t ne = 1;
  for (int k = 0; k < d; ++k)
    ne *= g.n[k];
  V u(std::size_t(ne * m), 0.);
  for (std::int64_t e = 0; e < ne; ++e) {
    std::int64_t idx[3] = {0, 0, 0}, r = e;
    for (int k = 0; k < d; ++k) {
      idx[k] = r % g.n[k];
      r /= g.n[k];
    }
    int total = 1;
    for (int k = 0; k < d; ++k)
      total *= nq;
    for (int q = 0; q < total; ++q) {
      FieldVector<double, d> x{};
      double w = 1.;
      int qq = q;
      for (int k = 0; k < d; ++k) {
        const double h = (g.upper[k] - g.lower[k]) / double(g.n[k]);
        x[k] = g.lower[k] + (double(idx[k]) + qx[qq % nq]) * h;
        w *= qw[qq % nq];
        qq /= nq;
      }
      const auto val = f(x, XT::Common::Parameter{});
      for (int c = 0; c < m; ++c)
        u[std::size_t(e * m + c)] += w * val[c];
    }
  }
  return u;
}

// explicit_euler of examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:141-159, run on the device:
// returns u after the loop `while (time < T_end + dt)`
template <class M, class GV>
XT::LA::IstlDenseVector<double> explicit_euler(const XT::LA::IstlDenseVector<double>& initial_values,
                                               const AdvectionFvOperator<M, GV>& spatial_op,
                                               const double T_end,
                                               const double dt)
{
  std::int64_t steps = 0;
  double time = 0.;
  while (time < T_end + dt) {
    time += dt;
    ++steps;
  }
  XT::LA::IstlDenseVector<double> u(initial_values);
  internal::check(gdtb_fvop_euler_host(spatial_op.handle(), u.data(), dt, steps));
  return u;
}

} // namespace GDT
} // namespace Dune

// grid type aliases of dune/xt/grid/grids.hh
using YASP_1D_EQUIDISTANT_OFFSET = Dune::XT::Grid::YaspEquidistantOffset<1>;
using YASP_2D_EQUIDISTANT_OFFSET = Dune::XT::Grid::YaspEquidistantOffset<2>;
using YASP_3D_EQUIDISTANT_OFFSET = Dune::XT::Grid::YaspEquidistantOffset<3>;

#endif // DUNE_GDT_B200_HH

// casadi-lite: the subset of the CasADi C++ API that the CUDA_SQP path of
// LockedFlysher/optimal_control_problem touches, written from scratch.
//
// CasADi itself is not in this image.  The reference reaches it at
//   src/sqp_solver/AutoDifferentiator.cpp:14-27   (Function, gradient, hessian, jacobian)
//   src/sqp_solver/SQPOptimizationSolver.cpp:47-77 (vertcat, sym, Function of SX)
//   include/.../sqp_solver/CuCaQP.h:114-122        (Sparsity::colind/row, DM::nonzeros)
//   src/OCP_config/OCPConfig.cpp, src/OptimalControlProblem.cpp (SX/DM containers, Slice)
// so this header keeps those names, argument meanings and value semantics:
//   * SX is a column-compressed sparse matrix of scalar expression handles,
//     DM the same container of doubles; vectors are dense n-by-1 matrices.
//   * jacobian()/hessian() return the STRUCTURAL pattern (dependency of each
//     output on each input after construction-time simplification), CCS with
//     strictly increasing row indices per column, zeros kept when structural.
//   * gradient() is dense and shaped like its argument.
// Beyond the API, the expression arena is exposed (SXElem::op/dep/value) so the
// stage code generator can walk the DAG.
#pragma once

#include <cmath>
#include <cstdint>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace casadi {

typedef long long casadi_int;
const double inf = std::numeric_limits<double>::infinity();
const double nan = std::numeric_limits<double>::quiet_NaN();

class CasadiException : public std::runtime_error {
 public:
  explicit CasadiException(const std::string& m) : std::runtime_error(m) {}
};
#define casadi_assert(cond, msg) \
  do { if (!(cond)) throw ::casadi::CasadiException(std::string("casadi-lite: ") + (msg)); } while (0)

// ---------------------------------------------------------------------------
// Scalar expression nodes
// ---------------------------------------------------------------------------
enum Operation : uint8_t {
  OP_CONST = 0, OP_PARAMETER,
  OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_ATAN2, OP_FMIN, OP_FMAX, OP_LT,   // binary
  OP_NEG, OP_SQ, OP_SQRT, OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN,   // unary
  OP_EXP, OP_LOG, OP_FABS, OP_SIGN, OP_TANH, OP_SINH, OP_COSH,
  OP_NUM_OPS
};
inline bool op_is_binary(int op) { return op >= OP_ADD && op <= OP_LT; }
inline bool op_is_unary(int op) { return op >= OP_NEG && op < OP_NUM_OPS; }
const char* op_name(int op);
double op_eval(int op, double a, double b);

// A handle into the global, append-only, hash-consed expression arena.  Node ids
// grow with construction time, so children always have smaller ids than their
// parents: ascending id order is a topological order.
class SXElem {
 public:
  SXElem() : id_(zero_id()) {}
  SXElem(double v);  // NOLINT: implicit by design (CasADi allows SXElem = 2.0)
  SXElem(int v) : SXElem(static_cast<double>(v)) {}
  static SXElem sym(const std::string& name);
  static SXElem from_id(int id) { SXElem e; e.id_ = id; return e; }
  static SXElem unary(int op, const SXElem& a);
  static SXElem binary(int op, const SXElem& a, const SXElem& b);

  int id() const { return id_; }
  int op() const;
  SXElem dep(int i) const;
  bool is_constant() const { return op() == OP_CONST; }
  bool is_symbolic() const { return op() == OP_PARAMETER; }
  double to_double() const;          // value of a constant node
  const std::string& name() const;   // name of a symbol node
  bool is_zero() const { return is_constant() && to_double() == 0.0; }
  bool is_one() const { return is_constant() && to_double() == 1.0; }
  bool is_minus_one() const { return is_constant() && to_double() == -1.0; }
  bool is_equal(const SXElem& o) const { return id_ == o.id_; }

  SXElem operator-() const { return unary(OP_NEG, *this); }
  friend SXElem operator+(const SXElem& a, const SXElem& b) { return binary(OP_ADD, a, b); }
  friend SXElem operator-(const SXElem& a, const SXElem& b) { return binary(OP_SUB, a, b); }
  friend SXElem operator*(const SXElem& a, const SXElem& b) { return binary(OP_MUL, a, b); }
  friend SXElem operator/(const SXElem& a, const SXElem& b) { return binary(OP_DIV, a, b); }
  SXElem& operator+=(const SXElem& b) { *this = *this + b; return *this; }
  SXElem& operator-=(const SXElem& b) { *this = *this - b; return *this; }
  SXElem& operator*=(const SXElem& b) { *this = *this * b; return *this; }

  static size_t arena_size();
  std::string str() const;

 private:
  static int zero_id();
  int id_;
};
std::ostream& operator<<(std::ostream& os, const SXElem& e);

// Scalar helpers used by the Matrix element-wise functions
template <typename T> struct ScalarOps;
template <> struct ScalarOps<double> {
  static double unary(int op, double a) { return op_eval(op, a, 0.0); }
  static double binary(int op, double a, double b) { return op_eval(op, a, b); }
  static bool is_zero(double a) { return a == 0.0; }
};
template <> struct ScalarOps<SXElem> {
  static SXElem unary(int op, const SXElem& a) { return SXElem::unary(op, a); }
  static SXElem binary(int op, const SXElem& a, const SXElem& b) { return SXElem::binary(op, a, b); }
  static bool is_zero(const SXElem& a) { return a.is_zero(); }
};

// ---------------------------------------------------------------------------
// Slice, Sparsity
// ---------------------------------------------------------------------------
class Slice {
 public:
  casadi_int start, stop, step;
  Slice() : start(0), stop(std::numeric_limits<casadi_int>::max()), step(1) {}
  Slice(casadi_int i) : start(i), stop(i + 1), step(1) {}  // NOLINT
  Slice(int i) : start(i), stop(i + 1), step(1) {}         // NOLINT
  Slice(casadi_int a, casadi_int b, casadi_int s = 1) : start(a), stop(b), step(s) {}
  Slice(int a, int b, int s = 1) : start(a), stop(b), step(s) {}
  Slice(int a, casadi_int b, int s = 1) : start(a), stop(b), step(s) {}
  Slice(casadi_int a, int b, int s = 1) : start(a), stop(b), step(s) {}
  std::vector<casadi_int> all(casadi_int len) const;
};

// Column-compressed pattern.  Immutable and shared.
class Sparsity {
 public:
  Sparsity() : Sparsity(0, 0) {}
  Sparsity(casadi_int nrow, casadi_int ncol);  // all structural zeros
  Sparsity(casadi_int nrow, casadi_int ncol, const std::vector<casadi_int>& colind,
           const std::vector<casadi_int>& row);
  static Sparsity dense(casadi_int nrow, casadi_int ncol = 1);

  casadi_int size1() const { return d_->nrow; }
  casadi_int size2() const { return d_->ncol; }
  casadi_int numel() const { return d_->nrow * d_->ncol; }
  casadi_int nnz() const { return static_cast<casadi_int>(d_->row.size()); }
  const casadi_int* colind() const { return d_->colind.data(); }
  const casadi_int* row() const { return d_->row.data(); }
  const std::vector<casadi_int>& get_colind() const { return d_->colind; }
  const std::vector<casadi_int>& get_row() const { return d_->row; }
  bool is_dense() const { return nnz() == numel(); }
  bool operator==(const Sparsity& o) const;
  bool operator!=(const Sparsity& o) const { return !(*this == o); }
  // index of (rr, cc) in the nonzero array, or -1
  casadi_int get_nz(casadi_int rr, casadi_int cc) const;
  Sparsity T() const;

 private:
  struct Data {
    casadi_int nrow, ncol;
    std::vector<casadi_int> colind, row;
  };
  std::shared_ptr<const Data> d_;
};

// ---------------------------------------------------------------------------
// GenericType / Dict
// ---------------------------------------------------------------------------
class GenericType {
 public:
  GenericType() : kind_(K_NONE), i_(0), d_(0) {}
  GenericType(bool b) : kind_(K_BOOL), i_(b), d_(b) {}                    // NOLINT
  GenericType(int i) : kind_(K_INT), i_(i), d_(i) {}                      // NOLINT
  GenericType(casadi_int i) : kind_(K_INT), i_(i), d_(double(i)) {}       // NOLINT
  GenericType(double d) : kind_(K_DOUBLE), i_(casadi_int(d)), d_(d) {}    // NOLINT
  GenericType(const std::string& s) : kind_(K_STRING), i_(0), d_(0), s_(s) {}  // NOLINT
  GenericType(const char* s) : kind_(K_STRING), i_(0), d_(0), s_(s) {}    // NOLINT
  bool is_bool() const { return kind_ == K_BOOL; }
  bool is_int() const { return kind_ == K_INT; }
  bool is_double() const { return kind_ == K_DOUBLE; }
  bool is_string() const { return kind_ == K_STRING; }
  casadi_int as_int() const;
  double as_double() const;
  bool as_bool() const;
  const std::string& as_string() const;
  casadi_int to_int() const { return as_int(); }
  double to_double() const { return as_double(); }
  bool to_bool() const { return as_bool(); }
  operator bool() const { return as_bool(); }  // NOLINT: CasADi converts option values implicitly
 private:
  enum Kind { K_NONE, K_BOOL, K_INT, K_DOUBLE, K_STRING } kind_;
  casadi_int i_;
  double d_;
  std::string s_;
};
typedef std::map<std::string, GenericType> Dict;

// ---------------------------------------------------------------------------
// Matrix<T>
// ---------------------------------------------------------------------------
template <typename T> class Matrix;
typedef Matrix<SXElem> SX;
typedef Matrix<double> DM;
typedef std::vector<SX> SXVector;
typedef std::vector<DM> DMVector;
typedef std::map<std::string, SX> SXDict;
typedef std::map<std::string, DM> DMDict;

// Assignable view, as in CasADi: xs(0) = ..., lbx(Slice(0, nf)) = frame.
template <typename M, typename I>
class SubIndex : public M {
 public:
  SubIndex(M& mat, const I& i) : M(mat.get_sub(i)), mat_(mat), i_(i) {}
  const M& operator=(const M& y) { mat_.set_sub(y, i_); M::operator=(mat_.get_sub(i_)); return y; }
  const M& operator=(const SubIndex& y) { return (*this) = static_cast<const M&>(y); }
  M operator+=(const M& y) { M s = static_cast<const M&>(*this) + y; (*this) = s; return s; }
  M operator-=(const M& y) { M s = static_cast<const M&>(*this) - y; (*this) = s; return s; }
 private:
  M& mat_;
  I i_;
};

template <typename T>
class Matrix {
 public:
  Matrix() : sp_(0, 0) {}
  Matrix(casadi_int nrow, casadi_int ncol) : sp_(nrow, ncol) {}
  Matrix(double v) : sp_(Sparsity::dense(1, 1)), nz_(1, T(v)) {}  // NOLINT
  Matrix(int v) : Matrix(static_cast<double>(v)) {}               // NOLINT
  Matrix(float v) : Matrix(static_cast<double>(v)) {}             // NOLINT
  Matrix(casadi_int v) : Matrix(static_cast<double>(v)) {}        // NOLINT
  Matrix(const std::vector<double>& v)                            // NOLINT
      : sp_(Sparsity::dense(static_cast<casadi_int>(v.size()), 1)) {
    nz_.reserve(v.size());
    for (double x : v) nz_.push_back(T(x));
  }
  Matrix(std::initializer_list<double> v) : Matrix(std::vector<double>(v)) {}  // NOLINT
  Matrix(const Sparsity& sp, const std::vector<T>& nz) : sp_(sp), nz_(nz) {
    casadi_assert(static_cast<casadi_int>(nz_.size()) == sp_.nnz(), "Matrix: nnz mismatch");
  }
  Matrix(const Sparsity& sp, const T& v) : sp_(sp), nz_(sp.nnz(), v) {}
  // SX from a scalar expression; DM(SX) (constant expressions only) -- see specialisations
  template <typename U, typename = typename std::enable_if<
                            std::is_same<U, SXElem>::value && std::is_same<T, SXElem>::value>::type>
  Matrix(const U& e) : sp_(Sparsity::dense(1, 1)), nz_(1, e) {}  // NOLINT
  template <typename U, typename = typename std::enable_if<
                            std::is_same<U, SXElem>::value && std::is_same<T, double>::value>::type,
            typename = void>
  explicit Matrix(const Matrix<U>& x) : sp_(x.sparsity()) {
    nz_.reserve(x.nonzeros().size());
    for (const U& e : x.nonzeros()) nz_.push_back(e.to_double());  // throws if symbolic
  }

  // ---- creation
  static Matrix zeros(casadi_int nrow = 1, casadi_int ncol = 1) {
    return Matrix(Sparsity::dense(nrow, ncol), T(0.0));
  }
  static Matrix ones(casadi_int nrow = 1, casadi_int ncol = 1) {
    return Matrix(Sparsity::dense(nrow, ncol), T(1.0));
  }
  static Matrix eye(casadi_int n);
  static Matrix sym(const std::string& name, casadi_int nrow = 1, casadi_int ncol = 1);
  static Matrix vertcat(const std::vector<Matrix>& v);
  static Matrix horzcat(const std::vector<Matrix>& v);
  static Matrix repmat(const Matrix& a, casadi_int n, casadi_int m = 1);

  // ---- shape
  casadi_int size1() const { return sp_.size1(); }
  casadi_int size2() const { return sp_.size2(); }
  casadi_int numel() const { return sp_.numel(); }
  casadi_int nnz() const { return sp_.nnz(); }
  std::pair<casadi_int, casadi_int> size() const { return {size1(), size2()}; }
  bool is_empty(bool both = false) const {
    return both ? (size1() == 0 && size2() == 0) : (size1() == 0 || size2() == 0);
  }
  bool is_scalar() const { return size1() == 1 && size2() == 1; }
  bool is_dense() const { return sp_.is_dense(); }
  bool is_column() const { return size2() == 1; }
  const Sparsity& sparsity() const { return sp_; }
  std::vector<T>& nonzeros() { return nz_; }
  const std::vector<T>& nonzeros() const { return nz_; }
  T scalar() const {
    casadi_assert(is_scalar(), "scalar(): not a 1-by-1 matrix");
    return nz_.empty() ? T(0.0) : nz_[0];
  }
  // element (rr,cc); a structural zero reads as 0
  T elem(casadi_int rr, casadi_int cc = 0) const {
    casadi_int k = sp_.get_nz(rr, cc);
    return k < 0 ? T(0.0) : nz_[k];
  }
  Matrix T_() const;

  // ---- indexing (read): a single index / Slice addresses a column vector
  Matrix get_sub(const Slice& s) const;
  Matrix get_sub(const std::pair<Slice, Slice>& rc) const;
  void set_sub(const Matrix& y, const Slice& s);
  void set_sub(const Matrix& y, const std::pair<Slice, Slice>& rc);
  Matrix operator()(const Slice& s) const { return get_sub(s); }
  Matrix operator()(casadi_int i) const { return get_sub(Slice(i)); }
  Matrix operator()(int i) const { return get_sub(Slice(i)); }
  Matrix operator()(const Slice& r, const Slice& c) const { return get_sub(std::make_pair(r, c)); }
  SubIndex<Matrix, Slice> operator()(const Slice& s) { return SubIndex<Matrix, Slice>(*this, s); }
  SubIndex<Matrix, Slice> operator()(casadi_int i) { return SubIndex<Matrix, Slice>(*this, Slice(i)); }
  SubIndex<Matrix, Slice> operator()(int i) { return SubIndex<Matrix, Slice>(*this, Slice(i)); }
  SubIndex<Matrix, std::pair<Slice, Slice>> operator()(const Slice& r, const Slice& c) {
    return SubIndex<Matrix, std::pair<Slice, Slice>>(*this, std::make_pair(r, c));
  }

  // ---- arithmetic (element-wise; a 1-by-1 operand broadcasts)
  static Matrix unary(int op, const Matrix& a);
  static Matrix binary(int op, const Matrix& a, const Matrix& b);
  Matrix operator-() const { return unary(OP_NEG, *this); }
  friend Matrix operator+(const Matrix& a, const Matrix& b) { return binary(OP_ADD, a, b); }
  friend Matrix operator-(const Matrix& a, const Matrix& b) { return binary(OP_SUB, a, b); }
  friend Matrix operator*(const Matrix& a, const Matrix& b) { return binary(OP_MUL, a, b); }
  friend Matrix operator/(const Matrix& a, const Matrix& b) { return binary(OP_DIV, a, b); }
  Matrix& operator+=(const Matrix& b) { *this = *this + b; return *this; }
  Matrix& operator-=(const Matrix& b) { *this = *this - b; return *this; }
  Matrix& operator*=(const Matrix& b) { *this = *this * b; return *this; }
  friend Matrix pow(const Matrix& a, const Matrix& b) { return binary(OP_POW, a, b); }
  friend Matrix atan2(const Matrix& a, const Matrix& b) { return binary(OP_ATAN2, a, b); }
  friend Matrix fmin(const Matrix& a, const Matrix& b) { return binary(OP_FMIN, a, b); }
  friend Matrix fmax(const Matrix& a, const Matrix& b) { return binary(OP_FMAX, a, b); }
  friend Matrix sq(const Matrix& a) { return unary(OP_SQ, a); }
  friend Matrix sqrt(const Matrix& a) { return unary(OP_SQRT, a); }
  friend Matrix sin(const Matrix& a) { return unary(OP_SIN, a); }
  friend Matrix cos(const Matrix& a) { return unary(OP_COS, a); }
  friend Matrix tan(const Matrix& a) { return unary(OP_TAN, a); }
  friend Matrix asin(const Matrix& a) { return unary(OP_ASIN, a); }
  friend Matrix acos(const Matrix& a) { return unary(OP_ACOS, a); }
  friend Matrix atan(const Matrix& a) { return unary(OP_ATAN, a); }
  friend Matrix exp(const Matrix& a) { return unary(OP_EXP, a); }
  friend Matrix log(const Matrix& a) { return unary(OP_LOG, a); }
  friend Matrix fabs(const Matrix& a) { return unary(OP_FABS, a); }
  friend Matrix tanh(const Matrix& a) { return unary(OP_TANH, a); }
  friend Matrix sinh(const Matrix& a) { return unary(OP_SINH, a); }
  friend Matrix cosh(const Matrix& a) { return unary(OP_COSH, a); }
  friend Matrix mtimes(const Matrix& a, const Matrix& b) { return Matrix::mtimes_(a, b); }
  friend Matrix dot(const Matrix& a, const Matrix& b) { return Matrix::dot_(a, b); }
  friend Matrix cross(const Matrix& a, const Matrix& b) { return Matrix::cross_(a, b); }
  friend Matrix sum1(const Matrix& a) { return Matrix::sum1_(a); }
  friend Matrix norm_2(const Matrix& a) { return sqrt(Matrix::dot_(a, a)); }
  friend Matrix norm_inf(const Matrix& a) { return Matrix::norm_inf_(a); }
  friend Matrix densify(const Matrix& a) { return Matrix::densify_(a); }
  friend Matrix transpose(const Matrix& a) { return a.T_(); }
  friend Matrix vertcat(const std::vector<Matrix>& v) { return Matrix::vertcat(v); }

  // ---- calculus (SX only; src/sqp_solver/AutoDifferentiator.cpp:16-27)
  static Matrix gradient(const Matrix& ex, const Matrix& arg);
  static Matrix jacobian(const Matrix& ex, const Matrix& arg);
  static Matrix hessian(const Matrix& ex, const Matrix& arg);
  static Matrix hessian(const Matrix& ex, const Matrix& arg, Matrix& g);
  // forward directional derivative  J(ex, arg) * v  (v shaped like arg)
  static Matrix jtimes(const Matrix& ex, const Matrix& arg, const Matrix& v);

  std::string str() const;

 private:
  static Matrix mtimes_(const Matrix& a, const Matrix& b);
  static Matrix dot_(const Matrix& a, const Matrix& b);
  static Matrix cross_(const Matrix& a, const Matrix& b);
  static Matrix sum1_(const Matrix& a);
  static Matrix norm_inf_(const Matrix& a);
  static Matrix densify_(const Matrix& a);
  Sparsity sp_;
  std::vector<T> nz_;
};

template <typename T>
std::ostream& operator<<(std::ostream& os, const Matrix<T>& m) { return os << m.str(); }
std::ostream& operator<<(std::ostream& os, const DMDict& d);

// ---------------------------------------------------------------------------
// Function: a frozen evaluation tape over purely symbolic inputs
// ---------------------------------------------------------------------------
class Function {
 public:
  Function() {}
  Function(const std::string& name, const SXVector& in, const SXVector& out);
  Function(const std::string& name, std::initializer_list<SX> in, std::initializer_list<SX> out)
      : Function(name, SXVector(in), SXVector(out)) {}

  bool is_null() const { return !d_; }
  const std::string& name() const;
  casadi_int n_in() const;
  casadi_int n_out() const;
  const Sparsity& sparsity_in(casadi_int i) const;
  const Sparsity& sparsity_out(casadi_int i) const;
  casadi_int nnz_in(casadi_int i) const { return sparsity_in(i).nnz(); }
  casadi_int nnz_out(casadi_int i) const { return sparsity_out(i).nnz(); }
  const SXVector& sx_in() const;
  const SXVector& sx_out() const;
  casadi_int n_instructions() const;

  // numeric evaluation (the reference's SX virtual machine, SQPOptimizationSolver.cpp:116-117)
  DMVector operator()(const DMVector& arg) const;
  DMVector operator()(const DM& arg0) const { return (*this)(DMVector{arg0}); }
  // raw-buffer evaluation: arg[i] -> nnz_in(i) doubles, res[i] -> nnz_out(i) doubles, w -> sz_w()
  void eval(const double* const* arg, double* const* res, double* w) const;
  size_t sz_w() const;
  // symbolic evaluation (substitution)
  SXVector operator()(const SXVector& arg) const;
  SXVector operator()(const SX& arg0) const { return (*this)(SXVector{arg0}); }

  // text serialisation of the tape (stand-in for casadi::Function::save,
  // src/OptimalControlProblem.cpp:412)
  void save(const std::string& filename) const;
  // C source in the layout of casadi::Function::generate / casadi::CodeGenerator: casadi_real / casadi_int,
  // `int name(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem)`, compact CCS
  // `name_sparsity_in/out`, `name_n_in/out`, `name_name_in/out`, `name_work`.  Intermediates live in w[] (what CasADi
  // does for MX functions; its SX functions use locals -- both honour the same signature).
  void generate(const std::string& filename) const;
  void generate_body(std::ostream& os, int index) const;   // one function of a multi-function file (CodeGenerator)

 private:
  struct Data;
  std::shared_ptr<const Data> d_;
};

// casadi::CodeGenerator subset: several functions in one C file
class CodeGenerator {
 public:
  explicit CodeGenerator(const std::string& name) : name_(name) {}
  void add(const Function& f) { fs_.push_back(f); }
  std::string generate(const std::string& prefix = "") const;   // writes <prefix><name>, returns the path
 private:
  std::string name_;
  std::vector<Function> fs_;
};

// ---------------------------------------------------------------------------
// DAG utilities shared by jacobian() and the stage code generator
// ---------------------------------------------------------------------------
namespace dag {
// ids of all nodes reachable from `roots`, ascending (= a topological order)
std::vector<int> reachable(const std::vector<SXElem>& roots);

// For every node reachable from `roots`: the sorted set of positions in `vars`
// it depends on.  Sets are interned; set_of[node id] indexes `sets`.
struct DepSets {
  std::vector<int> order;                 // reachable node ids, ascending
  std::vector<int> set_of;                // by position in `order`
  std::vector<std::vector<int>> sets;     // interned sorted sets; sets[0] is empty
  const std::vector<int>& of_root(size_t root_idx) const { return sets[root_set[root_idx]]; }
  std::vector<int> root_set;              // set id of each root
};
DepSets dependency_sets(const std::vector<SXElem>& roots, const std::vector<SXElem>& vars);

// Forward-mode tangents: d roots / d (sum_j vars[j] * seeds[j]) with symbolic or
// numeric seeds.  Nodes that do not depend on any seeded variable get tangent 0.
std::vector<SXElem> forward(const std::vector<SXElem>& roots, const std::vector<SXElem>& vars,
                            const std::vector<SXElem>& seeds);
// Reverse-mode: gradient of a scalar root with respect to vars.
std::vector<SXElem> reverse(const SXElem& root, const std::vector<SXElem>& vars);
}  // namespace dag

}  // namespace casadi

// casadi-lite implementation: expression arena, Matrix<T>, AD, Function.
// See host/casadi/casadi.hpp for the reference call sites this stands in for.
#include "casadi/casadi.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <unordered_map>

namespace casadi {

// ===========================================================================
// Arena
// ===========================================================================
namespace {
struct Node {
  int32_t a, b;
  double v;
  uint8_t op;
};

struct Arena {
  std::vector<Node> nodes;
  std::vector<std::string> names;   // symbol names, indexed by Node::a
  std::vector<int32_t> table;       // open addressing, -1 = empty
  size_t used = 0;
  int zero_id = -1;

  Arena() {
    table.assign(1 << 16, -1);
    nodes.reserve(1 << 16);
    zero_id = constant(0.0);
  }
  static uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
  }
  static uint64_t hash(const Node& n) {
    uint64_t bits;
    std::memcpy(&bits, &n.v, 8);
    uint64_t h = mix(bits + 0x9e3779b97f4a7c15ULL * (uint64_t(n.op) + 1));
    h = mix(h ^ (uint64_t(uint32_t(n.a)) << 32 | uint32_t(n.b)));
    return h;
  }
  static bool same(const Node& x, const Node& y) {
    return x.op == y.op && x.a == y.a && x.b == y.b && std::memcmp(&x.v, &y.v, 8) == 0;
  }
  void grow() {
    std::vector<int32_t> nt(table.size() * 2, -1);
    size_t mask = nt.size() - 1;
    for (int32_t id : table) {
      if (id < 0) continue;
      size_t h = hash(nodes[id]) & mask;
      while (nt[h] >= 0) h = (h + 1) & mask;
      nt[h] = id;
    }
    table.swap(nt);
  }
  int intern(const Node& n) {
    if (used * 2 >= table.size()) grow();
    size_t mask = table.size() - 1;
    size_t h = hash(n) & mask;
    while (table[h] >= 0) {
      if (same(nodes[table[h]], n)) return table[h];
      h = (h + 1) & mask;
    }
    int id = static_cast<int>(nodes.size());
    casadi_assert(nodes.size() < size_t(0x7fffffff), "expression arena exhausted");
    nodes.push_back(n);
    table[h] = id;
    ++used;
    return id;
  }
  int constant(double v) {
    Node n{-1, -1, v, OP_CONST};
    return intern(n);
  }
  int symbol(const std::string& name) {
    // symbols are never shared: two sym("x") calls give two variables
    Node n{static_cast<int32_t>(names.size()), -1, 0.0, OP_PARAMETER};
    names.push_back(name);
    int id = static_cast<int>(nodes.size());
    nodes.push_back(n);
    return id;
  }
  int make(int op, int a, int b) {
    Node n{a, b, 0.0, static_cast<uint8_t>(op)};
    return intern(n);
  }
};
Arena& arena() {
  static Arena* A = new Arena();  // intentionally leaked: handles may outlive static teardown
  return *A;
}
}  // namespace

const char* op_name(int op) {
  static const char* names[] = {"const", "param", "add", "sub", "mul", "div", "pow", "atan2", "fmin",
                                "fmax", "lt", "neg", "sq", "sqrt", "sin", "cos", "tan", "asin", "acos",
                                "atan", "exp", "log", "fabs", "sign", "tanh", "sinh", "cosh"};
  return (op >= 0 && op < OP_NUM_OPS) ? names[op] : "?";
}

double op_eval(int op, double a, double b) {
  switch (op) {
    case OP_ADD: return a + b;
    case OP_SUB: return a - b;
    case OP_MUL: return a * b;
    case OP_DIV: return a / b;
    case OP_POW: return std::pow(a, b);
    case OP_ATAN2: return std::atan2(a, b);
    case OP_FMIN: return std::fmin(a, b);
    case OP_FMAX: return std::fmax(a, b);
    case OP_LT: return a < b ? 1.0 : 0.0;
    case OP_NEG: return -a;
    case OP_SQ: return a * a;
    case OP_SQRT: return std::sqrt(a);
    case OP_SIN: return std::sin(a);
    case OP_COS: return std::cos(a);
    case OP_TAN: return std::tan(a);
    case OP_ASIN: return std::asin(a);
    case OP_ACOS: return std::acos(a);
    case OP_ATAN: return std::atan(a);
    case OP_EXP: return std::exp(a);
    case OP_LOG: return std::log(a);
    case OP_FABS: return std::fabs(a);
    case OP_SIGN: return a > 0 ? 1.0 : (a < 0 ? -1.0 : 0.0);
    case OP_TANH: return std::tanh(a);
    case OP_SINH: return std::sinh(a);
    case OP_COSH: return std::cosh(a);
    default: throw CasadiException("op_eval: bad op");
  }
}

// ===========================================================================
// SXElem
// ===========================================================================
int SXElem::zero_id() { return arena().zero_id; }
SXElem::SXElem(double v) : id_(arena().constant(v)) {}
SXElem SXElem::sym(const std::string& name) { return from_id(arena().symbol(name)); }
int SXElem::op() const { return arena().nodes[id_].op; }
SXElem SXElem::dep(int i) const {
  const Node& n = arena().nodes[id_];
  return from_id(i == 0 ? n.a : n.b);
}
double SXElem::to_double() const {
  const Node& n = arena().nodes[id_];
  casadi_assert(n.op == OP_CONST, "to_double(): expression is not constant");
  return n.v;
}
const std::string& SXElem::name() const {
  const Node& n = arena().nodes[id_];
  casadi_assert(n.op == OP_PARAMETER, "name(): not a symbol");
  return arena().names[n.a];
}
size_t SXElem::arena_size() { return arena().nodes.size(); }

SXElem SXElem::unary(int op, const SXElem& a) {
  if (a.is_constant()) return SXElem(op_eval(op, a.to_double(), 0.0));
  if (op == OP_NEG && a.op() == OP_NEG) return a.dep(0);
  if (op == OP_FABS && (a.op() == OP_FABS || a.op() == OP_SQ)) return a;
  return from_id(arena().make(op, a.id(), -1));
}

SXElem SXElem::binary(int op, const SXElem& a, const SXElem& b) {
  if (a.is_constant() && b.is_constant()) return SXElem(op_eval(op, a.to_double(), b.to_double()));
  switch (op) {
    case OP_ADD:
      if (a.is_zero()) return b;
      if (b.is_zero()) return a;
      if (b.op() == OP_NEG) return binary(OP_SUB, a, b.dep(0));
      if (a.op() == OP_NEG) return binary(OP_SUB, b, a.dep(0));
      if (a.is_equal(b)) return binary(OP_MUL, SXElem(2.0), a);
      break;
    case OP_SUB:
      if (b.is_zero()) return a;
      if (a.is_zero()) return unary(OP_NEG, b);
      if (a.is_equal(b)) return SXElem(0.0);
      if (b.op() == OP_NEG) return binary(OP_ADD, a, b.dep(0));
      break;
    case OP_MUL:
      if (a.is_zero() || b.is_zero()) return SXElem(0.0);
      if (a.is_one()) return b;
      if (b.is_one()) return a;
      if (a.is_minus_one()) return unary(OP_NEG, b);
      if (b.is_minus_one()) return unary(OP_NEG, a);
      if (a.is_equal(b)) return unary(OP_SQ, a);
      break;
    case OP_DIV:
      if (a.is_zero()) return SXElem(0.0);
      if (b.is_one()) return a;
      if (b.is_minus_one()) return unary(OP_NEG, a);
      if (a.is_equal(b)) return SXElem(1.0);
      break;
    case OP_POW:
      if (b.is_constant()) {
        double e = b.to_double();
        if (e == 0.0) return SXElem(1.0);
        if (e == 1.0) return a;
        if (e == 2.0) return unary(OP_SQ, a);
        if (e == 0.5) return unary(OP_SQRT, a);
        if (e == -1.0) return binary(OP_DIV, SXElem(1.0), a);
      }
      break;
    default: break;
  }
  int ia = a.id(), ib = b.id();
  if ((op == OP_ADD || op == OP_MUL || op == OP_FMIN || op == OP_FMAX) && ia > ib) std::swap(ia, ib);
  return from_id(arena().make(op, ia, ib));
}

static void print_expr(std::ostream& os, const SXElem& e, int depth) {
  if (depth > 24) { os << "..."; return; }
  int op = e.op();
  if (op == OP_CONST) { os << e.to_double(); return; }
  if (op == OP_PARAMETER) { os << e.name(); return; }
  if (op_is_binary(op)) {
    static const char* infix[] = {"+", "-", "*", "/"};
    if (op <= OP_DIV) {
      os << "("; print_expr(os, e.dep(0), depth + 1); os << infix[op - OP_ADD];
      print_expr(os, e.dep(1), depth + 1); os << ")";
    } else {
      os << op_name(op) << "("; print_expr(os, e.dep(0), depth + 1); os << ",";
      print_expr(os, e.dep(1), depth + 1); os << ")";
    }
    return;
  }
  if (op == OP_NEG) { os << "(-"; print_expr(os, e.dep(0), depth + 1); os << ")"; return; }
  os << op_name(op) << "("; print_expr(os, e.dep(0), depth + 1); os << ")";
}
std::string SXElem::str() const { std::ostringstream ss; print_expr(ss, *this, 0); return ss.str(); }
std::ostream& operator<<(std::ostream& os, const SXElem& e) { return os << e.str(); }

// ===========================================================================
// Slice / Sparsity / GenericType
// ===========================================================================
std::vector<casadi_int> Slice::all(casadi_int len) const {
  casadi_int a = start < 0 ? start + len : start;
  casadi_int b = stop == std::numeric_limits<casadi_int>::max() ? len : (stop < 0 ? stop + len : stop);
  casadi_assert(step > 0, "Slice: only positive steps are supported");
  casadi_assert(a >= 0 && b <= len, "Slice out of range");
  std::vector<casadi_int> r;
  for (casadi_int i = a; i < b; i += step) r.push_back(i);
  return r;
}

Sparsity::Sparsity(casadi_int nrow, casadi_int ncol) {
  auto d = std::make_shared<Data>();
  d->nrow = nrow; d->ncol = ncol;
  d->colind.assign(ncol + 1, 0);
  d_ = d;
}
Sparsity::Sparsity(casadi_int nrow, casadi_int ncol, const std::vector<casadi_int>& colind,
                   const std::vector<casadi_int>& row) {
  casadi_assert(static_cast<casadi_int>(colind.size()) == ncol + 1, "Sparsity: colind size");
  casadi_assert(colind.back() == static_cast<casadi_int>(row.size()), "Sparsity: row size");
  auto d = std::make_shared<Data>();
  d->nrow = nrow; d->ncol = ncol; d->colind = colind; d->row = row;
  d_ = d;
}
Sparsity Sparsity::dense(casadi_int nrow, casadi_int ncol) {
  std::vector<casadi_int> colind(ncol + 1), row(nrow * ncol);
  for (casadi_int j = 0; j <= ncol; ++j) colind[j] = j * nrow;
  for (casadi_int j = 0; j < ncol; ++j)
    for (casadi_int i = 0; i < nrow; ++i) row[j * nrow + i] = i;
  return Sparsity(nrow, ncol, colind, row);
}
bool Sparsity::operator==(const Sparsity& o) const {
  if (d_ == o.d_) return true;
  return d_->nrow == o.d_->nrow && d_->ncol == o.d_->ncol && d_->colind == o.d_->colind &&
         d_->row == o.d_->row;
}
casadi_int Sparsity::get_nz(casadi_int rr, casadi_int cc) const {
  casadi_assert(rr >= 0 && rr < d_->nrow && cc >= 0 && cc < d_->ncol, "index out of range");
  auto b = d_->row.begin() + d_->colind[cc], e = d_->row.begin() + d_->colind[cc + 1];
  auto it = std::lower_bound(b, e, rr);
  return (it != e && *it == rr) ? (it - d_->row.begin()) : -1;
}
Sparsity Sparsity::T() const {
  std::vector<casadi_int> cnt(d_->nrow + 1, 0), row(d_->row.size());
  for (casadi_int r : d_->row) cnt[r + 1]++;
  for (casadi_int i = 0; i < d_->nrow; ++i) cnt[i + 1] += cnt[i];
  std::vector<casadi_int> next(cnt.begin(), cnt.end() - 1);
  for (casadi_int j = 0; j < d_->ncol; ++j)
    for (casadi_int k = d_->colind[j]; k < d_->colind[j + 1]; ++k) row[next[d_->row[k]]++] = j;
  return Sparsity(d_->ncol, d_->nrow, cnt, row);
}

casadi_int GenericType::as_int() const {
  casadi_assert(kind_ == K_INT || kind_ == K_BOOL, "GenericType: not an int");
  return i_;
}
double GenericType::as_double() const {
  casadi_assert(kind_ == K_DOUBLE || kind_ == K_INT, "GenericType: not a double");
  return d_;
}
bool GenericType::as_bool() const {
  casadi_assert(kind_ == K_BOOL || kind_ == K_INT, "GenericType: not a bool");
  return i_ != 0;
}
const std::string& GenericType::as_string() const {
  casadi_assert(kind_ == K_STRING, "GenericType: not a string");
  return s_;
}

// ===========================================================================
// Matrix<T>
// ===========================================================================
template <typename T>
Matrix<T> Matrix<T>::eye(casadi_int n) {
  std::vector<casadi_int> colind(n + 1), row(n);
  for (casadi_int j = 0; j <= n; ++j) colind[j] = j;
  for (casadi_int j = 0; j < n; ++j) row[j] = j;
  return Matrix(Sparsity(n, n, colind, row), T(1.0));
}

template <>
SX SX::sym(const std::string& name, casadi_int nrow, casadi_int ncol) {
  Sparsity sp = Sparsity::dense(nrow, ncol);
  std::vector<SXElem> nz;
  nz.reserve(sp.nnz());
  if (nrow == 1 && ncol == 1) {
    nz.push_back(SXElem::sym(name));
  } else {
    for (casadi_int k = 0; k < sp.nnz(); ++k) nz.push_back(SXElem::sym(name + "_" + std::to_string(k)));
  }
  return SX(sp, nz);
}
template <>
DM DM::sym(const std::string&, casadi_int, casadi_int) {
  throw CasadiException("DM::sym is not defined");
}

template <typename T>
Matrix<T> Matrix<T>::vertcat(const std::vector<Matrix>& v) {
  casadi_int ncol = -1, nrow = 0;
  for (const Matrix& m : v) {
    if (m.is_empty(true)) continue;  // 0-by-0 placeholders (an empty SX()) are skipped
    if (ncol < 0) ncol = m.size2();
    casadi_assert(m.size2() == ncol, "vertcat: column count mismatch");
    nrow += m.size1();
  }
  if (ncol < 0) return Matrix();
  std::vector<casadi_int> colind(ncol + 1, 0), row;
  std::vector<T> nz;
  for (casadi_int j = 0; j < ncol; ++j) {
    casadi_int off = 0;
    for (const Matrix& m : v) {
      if (m.is_empty(true)) continue;
      const casadi_int* ci = m.sp_.colind();
      const casadi_int* ri = m.sp_.row();
      for (casadi_int k = ci[j]; k < ci[j + 1]; ++k) {
        row.push_back(ri[k] + off);
        nz.push_back(m.nz_[k]);
      }
      off += m.size1();
    }
    colind[j + 1] = static_cast<casadi_int>(row.size());
  }
  return Matrix(Sparsity(nrow, ncol, colind, row), nz);
}

template <typename T>
Matrix<T> Matrix<T>::horzcat(const std::vector<Matrix>& v) {
  casadi_int nrow = -1, ncol = 0;
  std::vector<casadi_int> colind(1, 0), row;
  std::vector<T> nz;
  for (const Matrix& m : v) {
    if (m.is_empty(true)) continue;
    if (nrow < 0) nrow = m.size1();
    casadi_assert(m.size1() == nrow, "horzcat: row count mismatch");
    const casadi_int* ci = m.sp_.colind();
    const casadi_int* ri = m.sp_.row();
    for (casadi_int j = 0; j < m.size2(); ++j) {
      for (casadi_int k = ci[j]; k < ci[j + 1]; ++k) { row.push_back(ri[k]); nz.push_back(m.nz_[k]); }
      colind.push_back(static_cast<casadi_int>(row.size()));
    }
    ncol += m.size2();
  }
  if (nrow < 0) return Matrix();
  return Matrix(Sparsity(nrow, ncol, colind, row), nz);
}

template <typename T>
Matrix<T> Matrix<T>::repmat(const Matrix& a, casadi_int n, casadi_int m) {
  Matrix col = vertcat(std::vector<Matrix>(n, a));
  if (n == 0) col = Matrix(0, a.size2());
  if (m == 1) return col;
  return horzcat(std::vector<Matrix>(m, col));
}

template <typename T>
Matrix<T> Matrix<T>::T_() const {
  casadi_int nr = size1(), nc = size2();
  std::vector<casadi_int> cnt(nr + 1, 0), row(nz_.size());
  std::vector<T> nz(nz_.size(), T(0.0));
  const casadi_int* ci = sp_.colind();
  const casadi_int* ri = sp_.row();
  for (casadi_int k = 0; k < nnz(); ++k) cnt[ri[k] + 1]++;
  for (casadi_int i = 0; i < nr; ++i) cnt[i + 1] += cnt[i];
  std::vector<casadi_int> next(cnt.begin(), cnt.end() - 1);
  for (casadi_int j = 0; j < nc; ++j)
    for (casadi_int k = ci[j]; k < ci[j + 1]; ++k) {
      casadi_int p = next[ri[k]]++;
      row[p] = j;
      nz[p] = nz_[k];
    }
  return Matrix(Sparsity(nc, nr, cnt, row), nz);
}

template <typename T>
Matrix<T> Matrix<T>::densify_(const Matrix& a) {
  if (a.is_dense()) return a;
  Matrix r = zeros(a.size1(), a.size2());
  const casadi_int* ci = a.sp_.colind();
  const casadi_int* ri = a.sp_.row();
  for (casadi_int j = 0; j < a.size2(); ++j)
    for (casadi_int k = ci[j]; k < ci[j + 1]; ++k) r.nz_[j * a.size1() + ri[k]] = a.nz_[k];
  return r;
}

template <typename T>
Matrix<T> Matrix<T>::get_sub(const Slice& s) const {
  // linear (column-major) indexing; for the n-by-1 vectors of the OCP path this is row indexing
  std::vector<casadi_int> idx = s.all(numel());
  Matrix r = zeros(static_cast<casadi_int>(idx.size()), 1);
  casadi_int nr = size1();
  for (size_t k = 0; k < idx.size(); ++k) {
    if (is_dense()) r.nz_[k] = nz_[idx[k]];
    else r.nz_[k] = elem(idx[k] % nr, idx[k] / nr);
  }
  return r;
}

template <typename T>
Matrix<T> Matrix<T>::get_sub(const std::pair<Slice, Slice>& rc) const {
  std::vector<casadi_int> rows = rc.first.all(size1()), cols = rc.second.all(size2());
  std::vector<casadi_int> newrow(size1(), -1);
  for (size_t i = 0; i < rows.size(); ++i) newrow[rows[i]] = static_cast<casadi_int>(i);
  std::vector<casadi_int> colind(1, 0), row;
  std::vector<T> nz;
  const casadi_int* ci = sp_.colind();
  const casadi_int* ri = sp_.row();
  for (casadi_int c : cols) {
    for (casadi_int k = ci[c]; k < ci[c + 1]; ++k)
      if (newrow[ri[k]] >= 0) { row.push_back(newrow[ri[k]]); nz.push_back(nz_[k]); }
    colind.push_back(static_cast<casadi_int>(row.size()));
  }
  return Matrix(Sparsity(static_cast<casadi_int>(rows.size()), static_cast<casadi_int>(cols.size()),
                         colind, row), nz);
}

template <typename T>
void Matrix<T>::set_sub(const Matrix& y, const Slice& s) {
  std::vector<casadi_int> idx = s.all(numel());
  if (!is_dense()) *this = densify_(*this);
  Matrix yd = densify_(y);
  casadi_assert(yd.numel() == 1 || yd.numel() == static_cast<casadi_int>(idx.size()),
                "set_sub: dimension mismatch");
  for (size_t k = 0; k < idx.size(); ++k) nz_[idx[k]] = yd.numel() == 1 ? yd.nz_[0] : yd.nz_[k];
}

template <typename T>
void Matrix<T>::set_sub(const Matrix& y, const std::pair<Slice, Slice>& rc) {
  std::vector<casadi_int> rows = rc.first.all(size1()), cols = rc.second.all(size2());
  if (!is_dense()) *this = densify_(*this);
  Matrix yd = densify_(y);
  casadi_assert(yd.numel() == 1 || (yd.size1() == static_cast<casadi_int>(rows.size()) &&
                                    yd.size2() == static_cast<casadi_int>(cols.size())),
                "set_sub: dimension mismatch");
  for (size_t j = 0; j < cols.size(); ++j)
    for (size_t i = 0; i < rows.size(); ++i)
      nz_[cols[j] * size1() + rows[i]] = yd.numel() == 1 ? yd.nz_[0] : yd.nz_[j * rows.size() + i];
}

static bool unary_keeps_zero(int op) {
  switch (op) {
    case OP_NEG: case OP_SQ: case OP_SQRT: case OP_SIN: case OP_TAN: case OP_ASIN: case OP_ATAN:
    case OP_FABS: case OP_SIGN: case OP_TANH: case OP_SINH: return true;
    default: return false;
  }
}

template <typename T>
Matrix<T> Matrix<T>::unary(int op, const Matrix& a) {
  Matrix x = (a.is_dense() || unary_keeps_zero(op)) ? a : densify_(a);
  Matrix r(x.sp_, T(0.0));
  for (size_t k = 0; k < x.nz_.size(); ++k) r.nz_[k] = ScalarOps<T>::unary(op, x.nz_[k]);
  return r;
}

template <typename T>
Matrix<T> Matrix<T>::binary(int op, const Matrix& a, const Matrix& b) {
  if (a.is_scalar() && !b.is_scalar()) {
    T s = a.scalar();
    bool keep = b.is_dense() || op == OP_MUL;
    Matrix y = keep ? b : densify_(b);
    Matrix r(y.sp_, T(0.0));
    for (size_t k = 0; k < y.nz_.size(); ++k) r.nz_[k] = ScalarOps<T>::binary(op, s, y.nz_[k]);
    return r;
  }
  if (b.is_scalar() && !a.is_scalar()) {
    T s = b.scalar();
    bool keep = a.is_dense() || op == OP_MUL || op == OP_DIV;
    Matrix x = keep ? a : densify_(a);
    Matrix r(x.sp_, T(0.0));
    for (size_t k = 0; k < x.nz_.size(); ++k) r.nz_[k] = ScalarOps<T>::binary(op, x.nz_[k], s);
    return r;
  }
  casadi_assert(a.size1() == b.size1() && a.size2() == b.size2(),
                "element-wise operation: dimension mismatch (" + std::to_string(a.size1()) + "x" +
                    std::to_string(a.size2()) + " vs " + std::to_string(b.size1()) + "x" +
                    std::to_string(b.size2()) + ")");
  if (a.sp_ == b.sp_) {
    Matrix r(a.sp_, T(0.0));
    for (size_t k = 0; k < a.nz_.size(); ++k) r.nz_[k] = ScalarOps<T>::binary(op, a.nz_[k], b.nz_[k]);
    return r;
  }
  return binary(op, densify_(a), densify_(b));
}

template <typename T>
Matrix<T> Matrix<T>::mtimes_(const Matrix& a, const Matrix& b) {
  if (a.is_scalar() || b.is_scalar()) return a * b;
  casadi_assert(a.size2() == b.size1(), "mtimes: inner dimension mismatch");
  casadi_int nr = a.size1(), nc = b.size2();
  std::vector<casadi_int> colind(1, 0), row;
  std::vector<T> nz;
  std::vector<T> acc(nr, T(0.0));
  std::vector<char> flag(nr, 0);
  const casadi_int *aci = a.sp_.colind(), *ari = a.sp_.row(), *bci = b.sp_.colind(), *bri = b.sp_.row();
  for (casadi_int j = 0; j < nc; ++j) {
    std::vector<casadi_int> touched;
    for (casadi_int kb = bci[j]; kb < bci[j + 1]; ++kb) {
      casadi_int k = bri[kb];
      for (casadi_int ka = aci[k]; ka < aci[k + 1]; ++ka) {
        casadi_int i = ari[ka];
        T prod = ScalarOps<T>::binary(OP_MUL, a.nz_[ka], b.nz_[kb]);
        if (!flag[i]) { flag[i] = 1; acc[i] = prod; touched.push_back(i); }
        else acc[i] = ScalarOps<T>::binary(OP_ADD, acc[i], prod);
      }
    }
    std::sort(touched.begin(), touched.end());
    for (casadi_int i : touched) { row.push_back(i); nz.push_back(acc[i]); flag[i] = 0; }
    colind.push_back(static_cast<casadi_int>(row.size()));
  }
  return Matrix(Sparsity(nr, nc, colind, row), nz);
}

template <typename T>
Matrix<T> Matrix<T>::dot_(const Matrix& a, const Matrix& b) {
  casadi_assert(a.numel() == b.numel(), "dot: dimension mismatch");
  Matrix x = densify_(a), y = densify_(b);
  T s(0.0);
  for (size_t k = 0; k < x.nz_.size(); ++k)
    s = ScalarOps<T>::binary(OP_ADD, s, ScalarOps<T>::binary(OP_MUL, x.nz_[k], y.nz_[k]));
  return Matrix(Sparsity::dense(1, 1), std::vector<T>(1, s));
}

template <typename T>
Matrix<T> Matrix<T>::cross_(const Matrix& a, const Matrix& b) {
  casadi_assert(a.numel() == 3 && b.numel() == 3, "cross: 3-vectors only");
  Matrix x = densify_(a), y = densify_(b);
  auto mul = [](const T& p, const T& q) { return ScalarOps<T>::binary(OP_MUL, p, q); };
  auto sub = [](const T& p, const T& q) { return ScalarOps<T>::binary(OP_SUB, p, q); };
  std::vector<T> r = {sub(mul(x.nz_[1], y.nz_[2]), mul(x.nz_[2], y.nz_[1])),
                      sub(mul(x.nz_[2], y.nz_[0]), mul(x.nz_[0], y.nz_[2])),
                      sub(mul(x.nz_[0], y.nz_[1]), mul(x.nz_[1], y.nz_[0]))};
  return Matrix(Sparsity::dense(3, 1), r);
}

template <typename T>
Matrix<T> Matrix<T>::sum1_(const Matrix& a) {
  Matrix r = zeros(1, a.size2());
  const casadi_int* ci = a.sp_.colind();
  for (casadi_int j = 0; j < a.size2(); ++j)
    for (casadi_int k = ci[j]; k < ci[j + 1]; ++k)
      r.nz_[j] = ScalarOps<T>::binary(OP_ADD, r.nz_[j], a.nz_[k]);
  return r;
}

template <typename T>
Matrix<T> Matrix<T>::norm_inf_(const Matrix& a) {
  T s(0.0);
  for (const T& e : a.nz_) s = ScalarOps<T>::binary(OP_FMAX, s, ScalarOps<T>::unary(OP_FABS, e));
  return Matrix(Sparsity::dense(1, 1), std::vector<T>(1, s));
}

static void print_scalar(std::ostream& os, double v) {
  if (std::isinf(v)) os << (v > 0 ? "inf" : "-inf");
  else if (std::isnan(v)) os << "nan";
  else os << v;
}
static void print_scalar(std::ostream& os, const SXElem& v) { os << v; }

template <typename T>
std::string Matrix<T>::str() const {
  std::ostringstream os;
  os << std::setprecision(std::is_same<T, double>::value ? 10 : 6);
  if (is_empty()) { os << "[](" << size1() << "x" << size2() << ")"; return os.str(); }
  if (is_scalar()) {
    if (nnz() == 0) os << "00"; else print_scalar(os, nz_[0]);
    return os.str();
  }
  auto put = [&](casadi_int i, casadi_int j) {
    casadi_int k = sp_.get_nz(i, j);
    if (k < 0) os << "00"; else print_scalar(os, nz_[k]);
  };
  if (size2() == 1) {
    os << "[";
    for (casadi_int i = 0; i < size1(); ++i) { if (i) os << ", "; put(i, 0); }
    os << "]";
    return os.str();
  }
  os << "\n[";
  for (casadi_int i = 0; i < size1(); ++i) {
    os << (i ? " [" : "[");
    for (casadi_int j = 0; j < size2(); ++j) { if (j) os << ", "; put(i, j); }
    os << (i + 1 < size1() ? "], \n" : "]]");
  }
  return os.str();
}

std::ostream& operator<<(std::ostream& os, const DMDict& d) {
  os << "{";
  bool first = true;
  for (const auto& kv : d) { if (!first) os << ", "; first = false; os << "\"" << kv.first << "\": " << kv.second; }
  return os << "}";
}

// ===========================================================================
// DAG utilities
// ===========================================================================
namespace dag {

std::vector<int> reachable(const std::vector<SXElem>& roots) {
  const std::vector<Node>& nodes = arena().nodes;
  std::vector<char> seen(nodes.size(), 0);
  std::vector<int> stack, out;
  for (const SXElem& r : roots)
    if (!seen[r.id()]) { seen[r.id()] = 1; stack.push_back(r.id()); }
  while (!stack.empty()) {
    int id = stack.back(); stack.pop_back();
    out.push_back(id);
    const Node& n = nodes[id];
    if (n.op == OP_CONST || n.op == OP_PARAMETER) continue;
    if (n.a >= 0 && !seen[n.a]) { seen[n.a] = 1; stack.push_back(n.a); }
    if (n.b >= 0 && !seen[n.b]) { seen[n.b] = 1; stack.push_back(n.b); }
  }
  std::sort(out.begin(), out.end());
  return out;
}

namespace {
// partial derivatives of node f = op(a, b)
void partials(int op, const SXElem& a, const SXElem& b, const SXElem& f, SXElem& da, SXElem& db) {
  const SXElem one(1.0), zero(0.0), two(2.0);
  da = zero; db = zero;
  switch (op) {
    case OP_ADD: da = one; db = one; break;
    case OP_SUB: da = one; db = SXElem(-1.0); break;
    case OP_MUL: da = b; db = a; break;
    case OP_DIV: da = one / b; db = -(f / b); break;
    case OP_POW:
      if (b.is_constant()) { da = b * SXElem::binary(OP_POW, a, SXElem(b.to_double() - 1.0)); }
      else { da = b * SXElem::binary(OP_POW, a, b - one); db = f * SXElem::unary(OP_LOG, a); }
      break;
    case OP_ATAN2: { SXElem d = a * a + b * b; da = b / d; db = -(a / d); break; }
    case OP_FMIN: { SXElem c = SXElem::binary(OP_LT, a, b); da = c; db = one - c; break; }
    case OP_FMAX: { SXElem c = SXElem::binary(OP_LT, b, a); da = c; db = one - c; break; }
    case OP_LT: break;
    case OP_NEG: da = SXElem(-1.0); break;
    case OP_SQ: da = two * a; break;
    case OP_SQRT: da = one / (two * f); break;
    case OP_SIN: da = SXElem::unary(OP_COS, a); break;
    case OP_COS: da = -SXElem::unary(OP_SIN, a); break;
    case OP_TAN: da = one + f * f; break;
    case OP_ASIN: da = one / SXElem::unary(OP_SQRT, one - a * a); break;
    case OP_ACOS: da = -(one / SXElem::unary(OP_SQRT, one - a * a)); break;
    case OP_ATAN: da = one / (one + a * a); break;
    case OP_EXP: da = f; break;
    case OP_LOG: da = one / a; break;
    case OP_FABS: da = SXElem::unary(OP_SIGN, a); break;
    case OP_SIGN: break;
    case OP_TANH: da = one - f * f; break;
    case OP_SINH: da = SXElem::unary(OP_COSH, a); break;
    case OP_COSH: da = SXElem::unary(OP_SINH, a); break;
    default: throw CasadiException("partials: bad op");
  }
}
}  // namespace

DepSets dependency_sets(const std::vector<SXElem>& roots, const std::vector<SXElem>& vars) {
  DepSets D;
  D.order = reachable(roots);
  const size_t N = D.order.size();
  D.set_of.assign(N, 0);
  D.sets.push_back({});
  std::map<std::vector<int>, int> intern;
  intern[{}] = 0;
  std::unordered_map<uint64_t, int> ucache;
  std::unordered_map<int, int> var_pos;  // node id -> position in vars
  for (size_t j = 0; j < vars.size(); ++j) {
    casadi_assert(vars[j].is_symbolic(), "differentiation variables must be purely symbolic");
    var_pos[vars[j].id()] = static_cast<int>(j);
  }
  std::unordered_map<int, int> pos_of;  // node id -> position in order
  pos_of.reserve(N * 2);
  for (size_t i = 0; i < N; ++i) pos_of[D.order[i]] = static_cast<int>(i);
  auto get_id = [&](std::vector<int>&& s) {
    auto it = intern.find(s);
    if (it != intern.end()) return it->second;
    int id = static_cast<int>(D.sets.size());
    intern[s] = id;
    D.sets.push_back(std::move(s));
    return id;
  };
  auto unite = [&](int x, int y) {
    if (x == y || y == 0) return x;
    if (x == 0) return y;
    if (x > y) std::swap(x, y);
    uint64_t key = (uint64_t(uint32_t(x)) << 32) | uint32_t(y);
    auto it = ucache.find(key);
    if (it != ucache.end()) return it->second;
    std::vector<int> u;
    std::set_union(D.sets[x].begin(), D.sets[x].end(), D.sets[y].begin(), D.sets[y].end(),
                   std::back_inserter(u));
    int id = get_id(std::move(u));
    ucache[key] = id;
    return id;
  };
  const std::vector<Node>& nodes = arena().nodes;
  for (size_t i = 0; i < N; ++i) {
    const Node& n = nodes[D.order[i]];
    if (n.op == OP_CONST) continue;
    if (n.op == OP_PARAMETER) {
      auto it = var_pos.find(D.order[i]);
      if (it != var_pos.end()) D.set_of[i] = get_id(std::vector<int>{it->second});
      continue;
    }
    int sa = D.set_of[pos_of[n.a]];
    int sb = n.b >= 0 ? D.set_of[pos_of[n.b]] : 0;
    D.set_of[i] = unite(sa, sb);
  }
  D.root_set.resize(roots.size());
  for (size_t r = 0; r < roots.size(); ++r) D.root_set[r] = D.set_of[pos_of[roots[r].id()]];
  return D;
}

std::vector<SXElem> forward(const std::vector<SXElem>& roots, const std::vector<SXElem>& vars,
                            const std::vector<SXElem>& seeds) {
  casadi_assert(vars.size() == seeds.size(), "forward: seed count mismatch");
  std::vector<int> order = reachable(roots);
  std::unordered_map<int, int> pos_of;
  pos_of.reserve(order.size() * 2);
  for (size_t i = 0; i < order.size(); ++i) pos_of[order[i]] = static_cast<int>(i);
  std::vector<SXElem> tan(order.size(), SXElem(0.0));
  for (size_t j = 0; j < vars.size(); ++j) {
    auto it = pos_of.find(vars[j].id());
    if (it != pos_of.end()) tan[it->second] = seeds[j];
  }
  const std::vector<Node>* nodes = &arena().nodes;
  for (size_t i = 0; i < order.size(); ++i) {
    Node n = (*nodes)[order[i]];
    if (n.op == OP_CONST || n.op == OP_PARAMETER) continue;
    const SXElem ta = tan[pos_of[n.a]];
    const SXElem tb = n.b >= 0 ? tan[pos_of[n.b]] : SXElem(0.0);
    if (ta.is_zero() && tb.is_zero()) continue;
    SXElem a = SXElem::from_id(n.a), b = n.b >= 0 ? SXElem::from_id(n.b) : SXElem(0.0);
    SXElem da, db;
    partials(n.op, a, b, SXElem::from_id(order[i]), da, db);
    nodes = &arena().nodes;  // the arena may have grown
    SXElem t(0.0);
    if (!ta.is_zero()) t = t + da * ta;
    if (!tb.is_zero()) t = t + db * tb;
    nodes = &arena().nodes;
    tan[i] = t;
  }
  std::vector<SXElem> out(roots.size());
  for (size_t r = 0; r < roots.size(); ++r) out[r] = tan[pos_of[roots[r].id()]];
  return out;
}

std::vector<SXElem> reverse(const SXElem& root, const std::vector<SXElem>& vars) {
  std::vector<int> order = reachable({root});
  std::unordered_map<int, int> pos_of;
  pos_of.reserve(order.size() * 2);
  for (size_t i = 0; i < order.size(); ++i) pos_of[order[i]] = static_cast<int>(i);
  std::vector<SXElem> adj(order.size(), SXElem(0.0));
  adj[pos_of[root.id()]] = SXElem(1.0);
  for (size_t ii = order.size(); ii-- > 0;) {
    Node n = arena().nodes[order[ii]];
    if (n.op == OP_CONST || n.op == OP_PARAMETER) continue;
    const SXElem w = adj[ii];
    if (w.is_zero()) continue;
    SXElem a = SXElem::from_id(n.a), b = n.b >= 0 ? SXElem::from_id(n.b) : SXElem(0.0);
    SXElem da, db;
    partials(n.op, a, b, SXElem::from_id(order[ii]), da, db);
    int pa = pos_of[n.a];
    adj[pa] = adj[pa] + w * da;
    if (n.b >= 0) {
      int pb = pos_of[n.b];
      adj[pb] = adj[pb] + w * db;
    }
  }
  std::vector<SXElem> g(vars.size(), SXElem(0.0));
  for (size_t j = 0; j < vars.size(); ++j) {
    auto it = pos_of.find(vars[j].id());
    if (it != pos_of.end()) g[j] = adj[it->second];
  }
  return g;
}

}  // namespace dag

// ===========================================================================
// Calculus
// ===========================================================================
static std::vector<SXElem> symbolic_vars(const SX& arg) {
  for (const SXElem& e : arg.nonzeros())
    casadi_assert(e.is_symbolic(), "differentiation argument must be purely symbolic");
  return arg.nonzeros();
}

template <>
SX SX::gradient(const SX& ex, const SX& arg) {
  casadi_assert(ex.is_scalar(), "gradient: expression must be scalar");
  std::vector<SXElem> g = dag::reverse(ex.scalar(), symbolic_vars(arg));
  return SX(arg.sparsity(), g);  // dense, shaped like arg
}

template <>
SX SX::jtimes(const SX& ex, const SX& arg, const SX& v) {
  casadi_assert(v.nnz() == arg.nnz(), "jtimes: seed shape mismatch");
  std::vector<SXElem> t = dag::forward(ex.nonzeros(), symbolic_vars(arg), v.nonzeros());
  return SX(ex.sparsity(), t);
}

// Structural Jacobian: entry (r, j) is present iff output nonzero r depends on
// input nonzero j; its value is the forward-mode tangent, kept even when it
// simplifies to a constant (including 0).
template <>
SX SX::jacobian(const SX& ex, const SX& arg) {
  const std::vector<SXElem> vars = symbolic_vars(arg);
  const std::vector<SXElem>& roots = ex.nonzeros();
  const casadi_int nrow = ex.numel(), ncol = arg.numel();
  // linear (column-major) index of each nonzero of ex / arg
  auto linear_index = [](const SX& m) {
    std::vector<casadi_int> li(m.nnz());
    const casadi_int* ci = m.sparsity().colind();
    const casadi_int* ri = m.sparsity().row();
    for (casadi_int j = 0; j < m.size2(); ++j)
      for (casadi_int k = ci[j]; k < ci[j + 1]; ++k) li[k] = ri[k] + j * m.size1();
    return li;
  };
  std::vector<casadi_int> out_row = linear_index(ex), in_col = linear_index(arg);

  dag::DepSets D = dag::dependency_sets(roots, vars);
  const size_t N = D.order.size();
  std::unordered_map<int, int> pos_of;
  pos_of.reserve(N * 2);
  for (size_t i = 0; i < N; ++i) pos_of[D.order[i]] = static_cast<int>(i);

  // cone[j]: positions (ascending) of the nodes that depend on variable j
  std::vector<std::vector<int>> cone(vars.size());
  for (size_t i = 0; i < N; ++i)
    for (int j : D.sets[D.set_of[i]]) cone[j].push_back(static_cast<int>(i));
  // rows_of[j]: output nonzeros depending on variable j (ascending)
  std::vector<std::vector<int>> rows_of(vars.size());
  for (size_t r = 0; r < roots.size(); ++r)
    for (int j : D.of_root(r)) rows_of[j].push_back(static_cast<int>(r));

  std::vector<int> stamp(N, -1);
  std::vector<SXElem> tan(N, SXElem(0.0));
  std::vector<casadi_int> colind(ncol + 1, 0), row;
  std::vector<SXElem> nz;
  std::vector<std::vector<std::pair<casadi_int, SXElem>>> cols(ncol);
  for (size_t j = 0; j < vars.size(); ++j) {
    if (rows_of[j].empty()) continue;
    const int sj = static_cast<int>(j);
    for (int i : cone[j]) {
      Node n = arena().nodes[D.order[i]];
      SXElem t(0.0);
      if (n.op == OP_PARAMETER) {
        t = SXElem(1.0);  // the only symbol in cone[j] is vars[j] itself
      } else {
        int pa = pos_of[n.a], pb = n.b >= 0 ? pos_of[n.b] : -1;
        SXElem ta = stamp[pa] == sj ? tan[pa] : SXElem(0.0);
        SXElem tb = (pb >= 0 && stamp[pb] == sj) ? tan[pb] : SXElem(0.0);
        if (!ta.is_zero() || !tb.is_zero()) {
          SXElem a = SXElem::from_id(n.a), b = n.b >= 0 ? SXElem::from_id(n.b) : SXElem(0.0);
          SXElem da, db;
          dag::partials(n.op, a, b, SXElem::from_id(D.order[i]), da, db);
          if (!ta.is_zero()) t = t + da * ta;
          if (!tb.is_zero()) t = t + db * tb;
        }
      }
      tan[i] = t;
      stamp[i] = sj;
    }
    for (int r : rows_of[j]) {
      int pr = pos_of[roots[r].id()];
      cols[in_col[j]].push_back({out_row[r], stamp[pr] == sj ? tan[pr] : SXElem(0.0)});
    }
  }
  for (casadi_int c = 0; c < ncol; ++c) {
    std::sort(cols[c].begin(), cols[c].end(),
              [](const std::pair<casadi_int, SXElem>& x, const std::pair<casadi_int, SXElem>& y) {
                return x.first < y.first;
              });
    for (auto& e : cols[c]) { row.push_back(e.first); nz.push_back(e.second); }
    colind[c + 1] = static_cast<casadi_int>(row.size());
  }
  return SX(Sparsity(nrow, ncol, colind, row), nz);
}

template <>
SX SX::hessian(const SX& ex, const SX& arg, SX& g) {
  g = gradient(ex, arg);
  return jacobian(g, arg);
}
template <>
SX SX::hessian(const SX& ex, const SX& arg) {
  SX g;
  return hessian(ex, arg, g);
}

template <> DM DM::gradient(const DM&, const DM&) { throw CasadiException("DM::gradient undefined"); }
template <> DM DM::jacobian(const DM&, const DM&) { throw CasadiException("DM::jacobian undefined"); }
template <> DM DM::hessian(const DM&, const DM&) { throw CasadiException("DM::hessian undefined"); }
template <> DM DM::hessian(const DM&, const DM&, DM&) { throw CasadiException("DM::hessian undefined"); }
template <> DM DM::jtimes(const DM&, const DM&, const DM&) { throw CasadiException("DM::jtimes undefined"); }

template class Matrix<double>;
template class Matrix<SXElem>;

// ===========================================================================
// Function
// ===========================================================================
struct Function::Data {
  std::string name;
  SXVector in, out;
  struct Instr { uint8_t op; int res, a, b; };
  std::vector<Instr> tape;                 // only real operations
  std::vector<std::pair<int, double>> consts;  // work index, value
  std::vector<std::vector<int>> in_w;      // work index of each input nonzero (-1: unused)
  std::vector<std::vector<int>> out_w;     // work index of each output nonzero
  size_t nwork = 0;
};

Function::Function(const std::string& name, const SXVector& in, const SXVector& out) {
  auto d = std::make_shared<Data>();
  d->name = name; d->in = in; d->out = out;
  std::vector<SXElem> roots;
  for (const SX& o : out) roots.insert(roots.end(), o.nonzeros().begin(), o.nonzeros().end());
  std::vector<int> order = dag::reachable(roots);
  std::unordered_map<int, int> pos_of;
  pos_of.reserve(order.size() * 2);
  for (size_t i = 0; i < order.size(); ++i) pos_of[order[i]] = static_cast<int>(i);
  d->nwork = order.size();
  std::unordered_map<int, char> is_input;
  d->in_w.resize(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    for (const SXElem& e : in[i].nonzeros()) {
      casadi_assert(e.is_symbolic(), "Function '" + name + "': inputs must be purely symbolic");
      auto it = pos_of.find(e.id());
      d->in_w[i].push_back(it == pos_of.end() ? -1 : it->second);
      is_input[e.id()] = 1;
    }
  }
  const std::vector<Node>& nodes = arena().nodes;
  for (size_t i = 0; i < order.size(); ++i) {
    const Node& n = nodes[order[i]];
    if (n.op == OP_CONST) { d->consts.push_back({static_cast<int>(i), n.v}); continue; }
    if (n.op == OP_PARAMETER) {
      casadi_assert(is_input.count(order[i]), "Function '" + name + "': free variable '" +
                                                  arena().names[n.a] + "'");
      continue;
    }
    d->tape.push_back({n.op, static_cast<int>(i), pos_of[n.a], n.b >= 0 ? pos_of[n.b] : -1});
  }
  d->out_w.resize(out.size());
  for (size_t i = 0; i < out.size(); ++i)
    for (const SXElem& e : out[i].nonzeros()) d->out_w[i].push_back(pos_of[e.id()]);
  d_ = d;
}

const std::string& Function::name() const { return d_->name; }
casadi_int Function::n_in() const { return static_cast<casadi_int>(d_->in.size()); }
casadi_int Function::n_out() const { return static_cast<casadi_int>(d_->out.size()); }
const Sparsity& Function::sparsity_in(casadi_int i) const { return d_->in.at(i).sparsity(); }
const Sparsity& Function::sparsity_out(casadi_int i) const { return d_->out.at(i).sparsity(); }
const SXVector& Function::sx_in() const { return d_->in; }
const SXVector& Function::sx_out() const { return d_->out; }
casadi_int Function::n_instructions() const { return static_cast<casadi_int>(d_->tape.size()); }
size_t Function::sz_w() const { return d_->nwork; }

void Function::eval(const double* const* arg, double* const* res, double* w) const {
  const Data& d = *d_;
  for (const auto& c : d.consts) w[c.first] = c.second;
  for (size_t i = 0; i < d.in_w.size(); ++i)
    for (size_t k = 0; k < d.in_w[i].size(); ++k)
      if (d.in_w[i][k] >= 0) w[d.in_w[i][k]] = arg[i][k];
  for (const Data::Instr& t : d.tape) w[t.res] = op_eval(t.op, w[t.a], t.b >= 0 ? w[t.b] : 0.0);
  for (size_t i = 0; i < d.out_w.size(); ++i)
    for (size_t k = 0; k < d.out_w[i].size(); ++k) res[i][k] = w[d.out_w[i][k]];
}

DMVector Function::operator()(const DMVector& arg) const {
  casadi_assert(d_, "call of a null Function");
  const Data& d = *d_;
  casadi_assert(arg.size() == d.in.size(), "Function '" + d.name + "': wrong number of inputs");
  std::vector<DM> dense_arg(arg.size());
  std::vector<const double*> ap(arg.size());
  for (size_t i = 0; i < arg.size(); ++i) {
    casadi_assert(arg[i].numel() == d.in[i].numel() || (arg[i].is_empty() && d.in[i].is_empty()),
                  "Function '" + d.name + "': input " + std::to_string(i) + " has " +
                      std::to_string(arg[i].numel()) + " elements, expected " +
                      std::to_string(d.in[i].numel()));
    casadi_assert(d.in[i].is_dense(), "Function: sparse symbolic inputs are not supported");
    dense_arg[i] = densify(arg[i]);
    ap[i] = dense_arg[i].nonzeros().data();
  }
  DMVector res(d.out.size());
  std::vector<double*> rp(d.out.size());
  for (size_t i = 0; i < d.out.size(); ++i) {
    res[i] = DM(d.out[i].sparsity(), 0.0);
    rp[i] = res[i].nonzeros().data();
  }
  std::vector<double> w(d.nwork);
  eval(ap.data(), rp.data(), w.data());
  return res;
}

SXVector Function::operator()(const SXVector& arg) const {
  casadi_assert(d_, "call of a null Function");
  const Data& d = *d_;
  casadi_assert(arg.size() == d.in.size(), "Function '" + d.name + "': wrong number of inputs");
  std::vector<SXElem> w(d.nwork, SXElem(0.0));
  for (const auto& c : d.consts) w[c.first] = SXElem(c.second);
  for (size_t i = 0; i < d.in_w.size(); ++i) {
    SX a = densify(arg[i]);
    casadi_assert(a.numel() == d.in[i].numel() || (a.is_empty() && d.in[i].is_empty()),
                  "Function '" + d.name + "': input dimension mismatch");
    for (size_t k = 0; k < d.in_w[i].size(); ++k)
      if (d.in_w[i][k] >= 0) w[d.in_w[i][k]] = a.nonzeros()[k];
  }
  for (const Data::Instr& t : d.tape)
    w[t.res] = t.b >= 0 ? SXElem::binary(t.op, w[t.a], w[t.b]) : SXElem::unary(t.op, w[t.a]);
  SXVector res(d.out.size());
  for (size_t i = 0; i < d.out.size(); ++i) {
    std::vector<SXElem> nz(d.out_w[i].size());
    for (size_t k = 0; k < nz.size(); ++k) nz[k] = w[d.out_w[i][k]];
    res[i] = SX(d.out[i].sparsity(), nz);
  }
  return res;
}

void Function::save(const std::string& filename) const {
  casadi_assert(d_, "save of a null Function");
  const Data& d = *d_;
  std::ofstream f(filename);
  casadi_assert(f.good(), "Function::save: cannot open " + filename);
  f << std::setprecision(17);
  f << "casadi-lite-function 1\nname " << d.name << "\n";
  auto put_sp = [&](const char* tag, const Sparsity& sp) {
    f << tag << " " << sp.size1() << " " << sp.size2() << " " << sp.nnz() << "\n";
    for (casadi_int c : sp.get_colind()) f << c << " ";
    f << "\n";
    for (casadi_int r : sp.get_row()) f << r << " ";
    f << "\n";
  };
  f << "n_in " << d.in.size() << "\n";
  for (size_t i = 0; i < d.in.size(); ++i) {
    put_sp("in", d.in[i].sparsity());
    for (int w : d.in_w[i]) f << w << " ";
    f << "\n";
  }
  f << "n_out " << d.out.size() << "\n";
  for (size_t i = 0; i < d.out.size(); ++i) {
    put_sp("out", d.out[i].sparsity());
    for (int w : d.out_w[i]) f << w << " ";
    f << "\n";
  }
  f << "work " << d.nwork << "\nconsts " << d.consts.size() << "\n";
  for (const auto& c : d.consts) f << c.first << " " << c.second << "\n";
  f << "tape " << d.tape.size() << "\n";
  for (const Data::Instr& t : d.tape) f << op_name(t.op) << " " << t.res << " " << t.a << " " << t.b << "\n";
}


// ---------------------------------------------------------------------------------------------
// C code generation in CasADi's format (casadi/core/code_generator.cpp, function_internal.cpp of upstream
// define the layout; written from the published structure of generated files, not from their sources)
// ---------------------------------------------------------------------------------------------
namespace {
const char* kCHeader =
    "/* This file was automatically generated by casadi-lite (optimal_control_problem_b200) in the format of\n"
    "   CasADi's C code generator (Function::generate): same preamble, signatures and sparsity encoding. */\n"
    "#ifdef __cplusplus\nextern \"C\" {\n#endif\n\n"
    "/* How to prefix internal symbols */\n#ifdef CASADI_CODEGEN_PREFIX\n"
    "  #define CASADI_NAMESPACE_CONCAT(NS, ID) _CASADI_NAMESPACE_CONCAT(NS, ID)\n"
    "  #define _CASADI_NAMESPACE_CONCAT(NS, ID) NS ## ID\n"
    "  #define CASADI_PREFIX(ID) CASADI_NAMESPACE_CONCAT(CODEGEN_PREFIX, ID)\n#else\n"
    "  #define CASADI_PREFIX(ID) %s_ ## ID\n#endif\n\n#include <math.h>\n\n"
    "#ifndef casadi_real\n#define casadi_real double\n#endif\n\n#ifndef casadi_int\n#define casadi_int long long int\n#endif\n\n"
    "/* Add prefix to internal symbols */\n#define casadi_f0 CASADI_PREFIX(f0)\n#define casadi_sq CASADI_PREFIX(sq)\n"
    "#define casadi_sign CASADI_PREFIX(sign)\n#define casadi_fmin CASADI_PREFIX(fmin)\n#define casadi_fmax CASADI_PREFIX(fmax)\n\n"
    "/* Symbol visibility in DLLs */\n#ifndef CASADI_SYMBOL_EXPORT\n"
    "  #if defined(_WIN32) || defined(__WIN32__) || defined(__CYGWIN__)\n    #if defined(STATIC_LINKED)\n"
    "      #define CASADI_SYMBOL_EXPORT\n    #else\n      #define CASADI_SYMBOL_EXPORT __declspec(dllexport)\n    #endif\n"
    "  #elif defined(__GNUC__) && defined(GCC_HASCLASSVISIBILITY)\n"
    "    #define CASADI_SYMBOL_EXPORT __attribute__ ((visibility (\"default\")))\n  #else\n    #define CASADI_SYMBOL_EXPORT\n  #endif\n#endif\n\n"
    "static casadi_real casadi_sq(casadi_real x) { return x*x;}\n\n"
    "static casadi_real casadi_sign(casadi_real x) { return x<0 ? -1 : x>0 ? 1 : x;}\n\n"
    "static casadi_real casadi_fmin(casadi_real x, casadi_real y) { return x<y ? x : y;}\n\n"
    "static casadi_real casadi_fmax(casadi_real x, casadi_real y) { return x>y ? x : y;}\n\n";

std::string c_expr(int op, const std::string& a, const std::string& b) {
  switch (op) {
    case OP_ADD: return "(" + a + "+" + b + ")";
    case OP_SUB: return "(" + a + "-" + b + ")";
    case OP_MUL: return "(" + a + "*" + b + ")";
    case OP_DIV: return "(" + a + "/" + b + ")";
    case OP_POW: return "pow(" + a + "," + b + ")";
    case OP_ATAN2: return "atan2(" + a + "," + b + ")";
    case OP_FMIN: return "casadi_fmin(" + a + "," + b + ")";
    case OP_FMAX: return "casadi_fmax(" + a + "," + b + ")";
    case OP_LT: return "(" + a + "<" + b + ")";
    case OP_NEG: return "(-" + a + ")";
    case OP_SQ: return "casadi_sq(" + a + ")";
    case OP_SIGN: return "casadi_sign(" + a + ")";
    default: return std::string(op_name(op)) + "(" + a + ")";   // sqrt sin cos tan asin acos atan exp log fabs tanh sinh cosh
  }
}
}  // namespace

void Function::generate_body(std::ostream& os, int index) const {
  casadi_assert(d_, "generate of a null Function");
  const Data& d = *d_;
  const std::string& nm = d.name;
  os << std::setprecision(17);
  // compact CCS: {nrow, ncol, colind[0..ncol], row[0..nnz)}; dense patterns as {nrow, ncol, 1}
  auto sp_array = [&](const Sparsity& sp, const std::string& id) {
    std::vector<casadi_int> v{sp.size1(), sp.size2()};
    if (sp.is_dense() && sp.numel() > 0) {
      v.push_back(1);
    } else {
      for (casadi_int c : sp.get_colind()) v.push_back(c);
      for (casadi_int r : sp.get_row()) v.push_back(r);
    }
    os << "static const casadi_int " << id << "[" << v.size() << "] = {";
    for (size_t i = 0; i < v.size(); ++i) os << (i ? ", " : "") << v[i];
    os << "};\n";
  };
  for (size_t i = 0; i < d.in.size(); ++i) sp_array(d.in[i].sparsity(), nm + "_s_in" + std::to_string(i));
  for (size_t i = 0; i < d.out.size(); ++i) sp_array(d.out[i].sparsity(), nm + "_s_out" + std::to_string(i));
  os << "\n/* " << nm << ":(";
  for (size_t i = 0; i < d.in.size(); ++i) os << (i ? "," : "") << "i" << i << "[" << d.in[i].size1() << "x" << d.in[i].size2() << "]";
  os << ")->(";
  for (size_t i = 0; i < d.out.size(); ++i) os << (i ? "," : "") << "o" << i << "[" << d.out[i].size1() << "x" << d.out[i].size2() << "]";
  os << ") */\n";
  os << "static int " << nm << "_f" << index << "(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem) {\n";
  for (const auto& c : d.consts) os << "  w[" << c.first << "] = " << c.second << ";\n";
  for (size_t i = 0; i < d.in_w.size(); ++i)
    for (size_t k = 0; k < d.in_w[i].size(); ++k)
      if (d.in_w[i][k] >= 0) os << "  w[" << d.in_w[i][k] << "] = arg[" << i << "]? arg[" << i << "][" << k << "] : 0;\n";
  for (const Data::Instr& t : d.tape) {
    const std::string a = "w[" + std::to_string(t.a) + "]", b = t.b >= 0 ? "w[" + std::to_string(t.b) + "]" : std::string();
    os << "  w[" << t.res << "] = " << c_expr(t.op, a, b) << ";\n";
  }
  for (size_t i = 0; i < d.out_w.size(); ++i)
    for (size_t k = 0; k < d.out_w[i].size(); ++k)
      os << "  if (res[" << i << "]!=0) res[" << i << "][" << k << "]=w[" << d.out_w[i][k] << "];\n";
  os << "  return 0;\n}\n\n";
  os << "CASADI_SYMBOL_EXPORT int " << nm << "(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem){\n"
     << "  return " << nm << "_f" << index << "(arg, res, iw, w, mem);\n}\n\n";
  os << "CASADI_SYMBOL_EXPORT int " << nm << "_alloc_mem(void) {\n  return 0;\n}\n\n";
  os << "CASADI_SYMBOL_EXPORT int " << nm << "_init_mem(int mem) {\n  return 0;\n}\n\n";
  os << "CASADI_SYMBOL_EXPORT void " << nm << "_free_mem(int mem) {\n}\n\n";
  os << "CASADI_SYMBOL_EXPORT int " << nm << "_checkout(void) {\n  return 0;\n}\n\n";
  os << "CASADI_SYMBOL_EXPORT void " << nm << "_release(int mem) {\n}\n\n";
  os << "CASADI_SYMBOL_EXPORT void " << nm << "_incref(void) {\n}\n\n";
  os << "CASADI_SYMBOL_EXPORT void " << nm << "_decref(void) {\n}\n\n";
  os << "CASADI_SYMBOL_EXPORT casadi_int " << nm << "_n_in(void) { return " << d.in.size() << ";}\n\n";
  os << "CASADI_SYMBOL_EXPORT casadi_int " << nm << "_n_out(void) { return " << d.out.size() << ";}\n\n";
  os << "CASADI_SYMBOL_EXPORT casadi_real " << nm << "_default_in(casadi_int i) {\n  switch (i) {\n    default: return 0;\n  }\n}\n\n";
  auto names = [&](const char* what, size_t count, char letter) {
    os << "CASADI_SYMBOL_EXPORT const char* " << nm << "_name_" << what << "(casadi_int i) {\n  switch (i) {\n";
    for (size_t i = 0; i < count; ++i) os << "    case " << i << ": return \"" << letter << i << "\";\n";
    os << "    default: return 0;\n  }\n}\n\n";
  };
  names("in", d.in.size(), 'i');
  names("out", d.out.size(), 'o');
  auto sps = [&](const char* what, size_t count) {
    os << "CASADI_SYMBOL_EXPORT const casadi_int* " << nm << "_sparsity_" << what << "(casadi_int i) {\n  switch (i) {\n";
    for (size_t i = 0; i < count; ++i) os << "    case " << i << ": return " << nm << "_s_" << what << i << ";\n";
    os << "    default: return 0;\n  }\n}\n\n";
  };
  sps("in", d.in.size());
  sps("out", d.out.size());
  os << "CASADI_SYMBOL_EXPORT int " << nm << "_work(casadi_int *sz_arg, casadi_int* sz_res, casadi_int *sz_iw, casadi_int *sz_w) {\n"
     << "  if (sz_arg) *sz_arg = " << d.in.size() << ";\n  if (sz_res) *sz_res = " << d.out.size() << ";\n"
     << "  if (sz_iw) *sz_iw = 0;\n  if (sz_w) *sz_w = " << d.nwork << ";\n  return 0;\n}\n\n";
}

void Function::generate(const std::string& filename) const {
  casadi_assert(d_, "generate of a null Function");
  CodeGenerator cg(filename);
  cg.add(*this);
  cg.generate();
}

std::string CodeGenerator::generate(const std::string& prefix) const {
  std::string path = prefix + name_;
  if (path.size() < 2 || path.substr(path.size() - 2) != ".c") path += ".c";
  std::ofstream f(path);
  casadi_assert(f.good(), "CodeGenerator: cannot open " + path);
  std::string stem = name_;
  const size_t slash = stem.find_last_of('/');
  if (slash != std::string::npos) stem = stem.substr(slash + 1);
  if (stem.size() > 2 && stem.substr(stem.size() - 2) == ".c") stem = stem.substr(0, stem.size() - 2);
  char buf[4096];
  std::snprintf(buf, sizeof(buf), kCHeader, stem.c_str());
  f << buf;
  for (size_t i = 0; i < fs_.size(); ++i) fs_[i].generate_body(f, static_cast<int>(i));
  f << "\n#ifdef __cplusplus\n} /* extern \"C\" */\n#endif\n";
  casadi_assert(f.good(), "CodeGenerator: write to " + path + " failed");
  return path;
}

}  // namespace casadi

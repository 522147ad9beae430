// AutoDifferentiator implementation (reference: src/sqp_solver/AutoDifferentiator.cpp).
// Scalar expression -> gradient + Hessian + Jacobian functions; vector expression ->
// Jacobian only (:16-27).  getLinearization(SX) = {J, -F} (:132-140).
#include "optimal_control_problem/sqp_solver/AutoDifferentiator.h"

using casadi::DM;
using casadi::Function;
using casadi::SX;

AutoDifferentiator::AutoDifferentiator(const SX& variables, const SX& expression)
    : x_(variables), expr_(expression), dim_(static_cast<size_t>(variables.size1())) {
  try {
    F_ = Function("F", {x_}, {expr_});
    if (expression.size1() == 1) {
      G_ = Function("G", {x_}, {SX::gradient(expr_, x_)});
      H_ = Function("H", {x_}, {SX::hessian(expr_, x_)});
      J_ = Function("J", {x_}, {SX::jacobian(expr_, x_)});
    } else if (expression.size1() > 1) {
      J_ = Function("J", {x_}, {SX::jacobian(expr_, x_)});
    }
  } catch (const std::exception& e) {
    throw AutoDifferentiatorException(std::string("construction failed: ") + e.what());
  }
}

template <typename M>
void AutoDifferentiator::checkPoint(const M& point) const {
  if (point.size2() != 1) throw AutoDifferentiatorException("exactly one evaluation point is expected");
  if (point.size1() != x_.size1())
    throw AutoDifferentiatorException("input dimension mismatch: expected " + std::to_string(x_.size1()) +
                                      ", got " + std::to_string(point.size1()));
}

SX AutoDifferentiator::getExpression() const { return F_(x_)[0]; }

namespace {
// runs fn, re-labelling foreign exceptions like the reference's try/catch blocks do
template <typename Fn>
auto guarded(const char* what, Fn fn) -> decltype(fn()) {
  try {
    return fn();
  } catch (const AutoDifferentiatorException&) {
    throw;
  } catch (const std::exception& e) {
    throw AutoDifferentiatorException(std::string(what) + e.what());
  }
}
}  // namespace

DM AutoDifferentiator::getJacobian(const DM& point) const {
  if (expr_.size1() == 1) throw AutoDifferentiatorException("scalar expression: use getGradient");
  return guarded("Jacobian evaluation failed: ", [&] { checkPoint(point); return J_(point)[0]; });
}
SX AutoDifferentiator::getJacobian(const SX& point) const {
  if (expr_.size1() == 1) throw AutoDifferentiatorException("scalar expression: use getGradient");
  return guarded("Jacobian evaluation failed: ", [&] { checkPoint(point); return J_(point)[0]; });
}
DM AutoDifferentiator::getGradient(const DM& point) const {
  if (expr_.size1() > 1) throw AutoDifferentiatorException("vector expression: use getJacobian");
  return guarded("gradient evaluation failed: ", [&] { checkPoint(point); return G_(point)[0]; });
}
SX AutoDifferentiator::getGradient(const SX& point) const {
  if (expr_.size1() > 1) throw AutoDifferentiatorException("vector expression: use getJacobian");
  return guarded("gradient evaluation failed: ", [&] { checkPoint(point); return G_(point)[0]; });
}
DM AutoDifferentiator::getHessian(const DM& point) const {
  return guarded("Hessian evaluation failed: ", [&] { checkPoint(point); return H_(point)[0]; });
}
SX AutoDifferentiator::getHessian(const SX& point) const {
  return guarded("Hessian evaluation failed: ", [&] { checkPoint(point); return H_(point)[0]; });
}

casadi::SXVector AutoDifferentiator::getLinearization(const SX& point) {
  return guarded("linearisation failed: ", [&] {
    SX J = getJacobian(point);
    SX b = -F_(point)[0];
    return casadi::SXVector{J, b};
  });
}
casadi::DMVector AutoDifferentiator::getLinearization(const DM& point) {
  return guarded("linearisation failed: ", [&] {
    DM J = getJacobian(point);
    DM b = F_(point)[0];
    return casadi::DMVector{J, b};
  });
}

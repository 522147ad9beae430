// AutoDifferentiator: gradient / Hessian / Jacobian functions of one expression.
// Public surface of the reference class (include/optimal_control_problem/sqp_solver/
// AutoDifferentiator.h:15-69); note getLinearization(SX) returns {J, -F}
// (src/sqp_solver/AutoDifferentiator.cpp:132-140) while the DM overload returns {J, +F}.
#pragma once

#include <stdexcept>
#include <string>

#include "casadi/casadi.hpp"

class AutoDifferentiatorException : public std::runtime_error {
 public:
  explicit AutoDifferentiatorException(const std::string& message) : std::runtime_error(message) {}
};

class AutoDifferentiator {
 public:
  explicit AutoDifferentiator(const casadi::SX& variables, const casadi::SX& expression);
  AutoDifferentiator(const AutoDifferentiator&) = delete;
  AutoDifferentiator& operator=(const AutoDifferentiator&) = delete;
  AutoDifferentiator(AutoDifferentiator&&) noexcept = default;
  AutoDifferentiator& operator=(AutoDifferentiator&&) noexcept = default;
  ~AutoDifferentiator() = default;

  const casadi::SX& getSymbolicVar() const noexcept { return x_; }
  casadi::SX getExpression() const;
  const casadi::Function& getJacobianFunction() const noexcept { return J_; }
  const casadi::Function& getHessianFunction() const noexcept { return H_; }
  const casadi::Function& getGradientFunction() const noexcept { return G_; }

  casadi::SX getJacobian(const casadi::SX& point) const;
  casadi::SX getGradient(const casadi::SX& point) const;
  casadi::SX getHessian(const casadi::SX& point) const;
  casadi::DM getJacobian(const casadi::DM& point) const;
  casadi::DM getGradient(const casadi::DM& point) const;
  casadi::DM getHessian(const casadi::DM& point) const;

  casadi::SXVector getLinearization(const casadi::SX& point);
  casadi::DMVector getLinearization(const casadi::DM& point);

  void setVerbose(bool verbose) noexcept { verbose_ = verbose; }
  bool isVerbose() const noexcept { return verbose_; }

 protected:
  template <typename M> void checkPoint(const M& point) const;
  casadi::SX x_;
  casadi::SX expr_;
  casadi::Function F_, G_, H_, J_;
  size_t dim_;
  bool verbose_{false};
};

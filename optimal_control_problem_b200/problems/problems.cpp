// See problems.h.  Dynamics are transcribed by multiple shooting with one explicit RK4
// step per stage: x_{k+1} - F(x_k, u_k) = 0 for k = 0..H-2 (addEquationConstraint), tracking
// costs through addVectorCost, box limits through the frame bounds in the YAML.
#include "problems.h"

#include <cmath>
#include <sstream>

using casadi::DM;
using casadi::Slice;
using casadi::SX;

namespace ocp_problems {
namespace {

SX rk4(SX (*f)(const SX&, const SX&), const SX& x, const SX& u, double dt) {
  SX k1 = f(x, u);
  SX k2 = f(x + (0.5 * dt) * k1, u);
  SX k3 = f(x + (0.5 * dt) * k2, u);
  SX k4 = f(x + dt * k3, u);
  return x + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
}

// Euler-rate matrix of ZYX roll/pitch/yaw times a body-rate vector
SX euler_rates(const SX& rpy, const SX& w) {
  SX sr = sin(rpy(0)), cr = cos(rpy(0)), tp = tan(rpy(1)), cp = cos(rpy(1));
  return SX::vertcat({w(0) + sr * tp * w(1) + cr * tp * w(2), cr * w(1) - sr * w(2),
                      (sr / cp) * w(1) + (cr / cp) * w(2)});
}

// ---------------------------------------------------------------- quadrotor
const double kQuadMass = 1.0, kQuadArm = 0.17, kQuadKappa = 0.016, kGravity = 9.81;
const double kQuadJ[3] = {0.01, 0.01, 0.02};

SX quadrotor_ode(const SX& x, const SX& u) {
  SX rpy = x(Slice(3, 6)), v = x(Slice(6, 9)), w = x(Slice(9, 12));
  SX sr = sin(rpy(0)), cr = cos(rpy(0)), sp = sin(rpy(1)), cp = cos(rpy(1)), sy = sin(rpy(2)), cy = cos(rpy(2));
  SX T = u(0) + u(1) + u(2) + u(3);
  // third column of R = Rz(yaw) Ry(pitch) Rx(roll)
  SX ax = (cy * sp * cr + sy * sr) * T / kQuadMass;
  SX ay = (sy * sp * cr - cy * sr) * T / kQuadMass;
  SX az = (cp * cr) * T / kQuadMass - kGravity;
  SX tx = kQuadArm * (u(1) - u(3)), ty = kQuadArm * (u(2) - u(0)), tz = kQuadKappa * (u(0) - u(1) + u(2) - u(3));
  SX wx = w(0), wy = w(1), wz = w(2);
  SX dwx = (tx - (kQuadJ[2] - kQuadJ[1]) * wy * wz) / kQuadJ[0];
  SX dwy = (ty - (kQuadJ[0] - kQuadJ[2]) * wz * wx) / kQuadJ[1];
  SX dwz = (tz - (kQuadJ[1] - kQuadJ[0]) * wx * wy) / kQuadJ[2];
  return SX::vertcat({v, euler_rates(rpy, w), ax, ay, az, dwx, dwy, dwz});
}

class QuadrotorOCP : public OptimalControlProblem {
 public:
  explicit QuadrotorOCP(YAML::Node node) : OptimalControlProblem(node) { setProblemName("quadrotor"); }
  void deployConstraintsAndAddCost() override {
    const int H = OCPConfigPtr_->getHorizon();
    const double dt = OCPConfigPtr_->getDt();
    const std::vector<double> Q = {10, 10, 10, 5, 5, 5, 1, 1, 1, 0.5, 0.5, 0.5};
    const std::vector<double> R = {0.1, 0.1, 0.1, 0.1};
    const double hover = kQuadMass * kGravity / 4.0;
    auto state = [&](int k) {
      return SX::vertcat({OCPConfigPtr_->getVariable(k, "pos"), OCPConfigPtr_->getVariable(k, "rpy"),
                          OCPConfigPtr_->getVariable(k, "vel"), OCPConfigPtr_->getVariable(k, "omega")});
    };
    for (int k = 0; k < H; ++k) {
      SX xk = state(k), uk = OCPConfigPtr_->getVariable(k, "thrust");
      addVectorCost(Q, xk - reference_);
      addVectorCost(R, uk - hover);
      if (k + 1 < H) addEquationConstraint("dynamics", state(k + 1), rk4(quadrotor_ode, xk, uk, dt));
    }
  }
};

const char* kQuadVars = R"(
  OCP_variables:
    - name: pos
      size: 3
      lower_bound: [-.inf, -.inf, -.inf]
      upper_bound: [.inf, .inf, .inf]
    - name: rpy
      size: 3
      lower_bound: [-0.8, -0.8, -.inf]
      upper_bound: [0.8, 0.8, .inf]
    - name: vel
      size: 3
      lower_bound: [-.inf, -.inf, -.inf]
      upper_bound: [.inf, .inf, .inf]
    - name: omega
      size: 3
      lower_bound: [-.inf, -.inf, -.inf]
      upper_bound: [.inf, .inf, .inf]
    - name: thrust
      size: 4
      lower_bound: [0.0, 0.0, 0.0, 0.0]
      upper_bound: [4.905, 4.905, 4.905, 4.905]
)";

// ---------------------------------------------------------------- cart-pole
const double kCartM = 1.0, kPoleM = 0.1, kPoleL = 0.5;

SX cartpole_ode(const SX& x, const SX& u) {
  // theta measured from the upright position
  SX th = x(1), ds = x(2), dth = x(3);
  SX s = sin(th), c = cos(th);
  const double total = kCartM + kPoleM;
  SX temp = (u(0) + kPoleM * kPoleL * dth * dth * s) / total;
  SX ddth = (kGravity * s - c * temp) / (kPoleL * (4.0 / 3.0 - kPoleM * c * c / total));
  SX dds = temp - kPoleM * kPoleL * ddth * c / total;
  return SX::vertcat({ds, dth, dds, ddth});
}

class CartPoleOCP : public OptimalControlProblem {
 public:
  explicit CartPoleOCP(YAML::Node node) : OptimalControlProblem(node) { setProblemName("cartpole"); }
  void deployConstraintsAndAddCost() override {
    const int H = OCPConfigPtr_->getHorizon();
    const double dt = OCPConfigPtr_->getDt();
    const std::vector<double> Q = {1.0, 10.0, 0.1, 0.1};
    const std::vector<double> R = {0.01};
    for (int k = 0; k < H; ++k) {
      SX xk = OCPConfigPtr_->getVariable(k, "state"), uk = OCPConfigPtr_->getVariable(k, "force");
      addVectorCost(Q, xk - reference_);
      addVectorCost(R, uk);
      if (k + 1 < H)
        addEquationConstraint("dynamics", OCPConfigPtr_->getVariable(k + 1, "state"), rk4(cartpole_ode, xk, uk, dt));
    }
  }
};

const char* kCartVars = R"(
  OCP_variables:
    - name: state
      size: 4
      lower_bound: [-2.4, -.inf, -.inf, -.inf]
      upper_bound: [2.4, .inf, .inf, .inf]
    - name: force
      size: 1
      lower_bound: [-20.0]
      upper_bound: [20.0]
)";

// ---------------------------------------------------------------- centroidal legged robot
const double kLegMass = 12.0, kLegMu = 0.6;
const double kLegIinv[3] = {1.0 / 0.25, 1.0 / 0.5, 1.0 / 0.6};

SX centroidal_ode(const SX& x, const SX& u) {
  SX com = x(Slice(0, 3)), rpy = x(Slice(3, 6)), lin = x(Slice(6, 9)), ang = x(Slice(9, 12));
  SX omega = SX::vertcat({kLegIinv[0] * ang(0), kLegIinv[1] * ang(1), kLegIinv[2] * ang(2)});
  SX dlin = SX::vertcat({SX(0.0), SX(0.0), SX(-kLegMass * kGravity)});
  SX dang = SX::zeros(3);
  for (int i = 0; i < 4; ++i) {
    SX f = u(Slice(3 * i, 3 * i + 3));
    SX r = x(Slice(12 + 3 * i, 15 + 3 * i)) - com;
    dlin = dlin + f;
    dang = dang + cross(r, f);
  }
  return SX::vertcat({lin / kLegMass, euler_rates(rpy, omega), dlin, dang, SX::zeros(12)});
}

class CentroidalOCP : public OptimalControlProblem {
 public:
  explicit CentroidalOCP(YAML::Node node) : OptimalControlProblem(node) { setProblemName("centroidal"); }
  void deployConstraintsAndAddCost() override {
    const int H = OCPConfigPtr_->getHorizon();
    const double dt = OCPConfigPtr_->getDt();
    std::vector<double> Q = {50, 50, 100, 20, 20, 10, 1, 1, 1, 2, 2, 2};
    for (int i = 0; i < 12; ++i) Q.push_back(100.0);  // stance feet stay where they are
    const std::vector<double> R(12, 1e-3);
    const double fz0 = kLegMass * kGravity / 4.0;
    SX unom = SX::vertcat({SX(0.0), SX(0.0), SX(fz0), SX(0.0), SX(0.0), SX(fz0), SX(0.0), SX(0.0), SX(fz0),
                           SX(0.0), SX(0.0), SX(fz0)});
    auto state = [&](int k) {
      return SX::vertcat({OCPConfigPtr_->getVariable(k, "com"), OCPConfigPtr_->getVariable(k, "rpy"),
                          OCPConfigPtr_->getVariable(k, "lin_mom"), OCPConfigPtr_->getVariable(k, "ang_mom"),
                          OCPConfigPtr_->getVariable(k, "feet")});
    };
    const double ninf = -casadi::inf, pinf = casadi::inf;
    for (int k = 0; k < H; ++k) {
      SX xk = state(k), uk = OCPConfigPtr_->getVariable(k, "grf");
      addVectorCost(Q, xk - reference_);
      addVectorCost(R, uk - unom);
      if (k + 1 < H) addEquationConstraint("dynamics", state(k + 1), rk4(centroidal_ode, xk, uk, dt));
    }
    // friction pyramids after the dynamics rows: 4 feet x 5 rows per stage
    for (int k = 0; k < H; ++k) {
      SX uk = OCPConfigPtr_->getVariable(k, "grf");
      for (int i = 0; i < 4; ++i) {
        SX fx = uk(3 * i), fy = uk(3 * i + 1), fz = uk(3 * i + 2);
        addInequalityConstraint("friction", DM({ninf, ninf, ninf, ninf, 0.0}),
                                SX::vertcat({fx - kLegMu * fz, -fx - kLegMu * fz, fy - kLegMu * fz,
                                             -fy - kLegMu * fz, fz}),
                                DM({0.0, 0.0, 0.0, 0.0, pinf}));
      }
    }
  }
};

std::string centroidal_vars() {
  std::ostringstream s;
  auto field = [&](const char* name, int size, const char* lo, const char* hi) {
    s << "    - name: " << name << "\n      size: " << size << "\n      lower_bound: [";
    for (int i = 0; i < size; ++i) s << (i ? ", " : "") << lo;
    s << "]\n      upper_bound: [";
    for (int i = 0; i < size; ++i) s << (i ? ", " : "") << hi;
    s << "]\n";
  };
  s << "  OCP_variables:\n";
  field("com", 3, "-.inf", ".inf");
  field("rpy", 3, "-.inf", ".inf");
  field("lin_mom", 3, "-.inf", ".inf");
  field("ang_mom", 3, "-.inf", ".inf");
  field("feet", 12, "-.inf", ".inf");
  // fx, fy free; fz <= 200 N
  s << "    - name: grf\n      size: 12\n      lower_bound: [";
  for (int i = 0; i < 12; ++i) s << (i ? ", " : "") << "-.inf";
  s << "]\n      upper_bound: [";
  for (int i = 0; i < 12; ++i) s << (i ? ", " : "") << (i % 3 == 2 ? "200.0" : ".inf");
  s << "]\n";
  return s.str();
}

// ---------------------------------------------------------------- random inputs
struct SplitMix64 {
  unsigned long long s;
  explicit SplitMix64(unsigned long long seed) : s(seed) {}
  unsigned long long next() {
    unsigned long long z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  double uniform(double lo, double hi) { return lo + (hi - lo) * ((next() >> 11) * (1.0 / 9007199254740992.0)); }
};

}  // namespace

std::string default_yaml(const std::string& name, int horizon, double alpha, int step_num, bool verbose) {
  double dt;
  int H;
  std::string vars;
  if (name == "quadrotor") { dt = 0.005; H = 20; vars = kQuadVars; }
  else if (name == "cartpole") { dt = 0.01; H = 200; vars = kCartVars; }
  else if (name == "centroidal") { dt = 0.01; H = 50; vars = centroidal_vars(); }
  else throw std::invalid_argument("unknown problem: " + name);
  if (horizon > 0) H = horizon;
  std::ostringstream s;
  s.precision(17);
  s << "  discretization_settings:\n    dt: " << dt << "\n    horizon: " << H << "\n";
  s << "  solver_settings:\n    max_iter: 1000\n    warm_start: true\n    verbose: " << (verbose ? "true" : "false")
    << "\n    gen_code: false\n    recompile: false\n    load_lib: false\n    solve_method: CUDA_SQP\n"
    << "    SQP_settings:\n      alpha: " << alpha << "\n      step_num: " << step_num << "\n";
  s << vars;
  return s.str();
}

int state_size(const std::string& name) {
  if (name == "quadrotor") return 12;
  if (name == "cartpole") return 4;
  if (name == "centroidal") return 24;
  throw std::invalid_argument("unknown problem: " + name);
}

std::unique_ptr<OptimalControlProblem> make_problem(const std::string& name, const std::string& yaml_text) {
  YAML::Node node = YAML::Load(yaml_text);
  if (node["optimal_control_problem"]) node = node["optimal_control_problem"];
  std::unique_ptr<OptimalControlProblem> ocp;
  if (name == "quadrotor") ocp.reset(new QuadrotorOCP(node));
  else if (name == "cartpole") ocp.reset(new CartPoleOCP(node));
  else if (name == "centroidal") ocp.reset(new CentroidalOCP(node));
  else throw std::invalid_argument("unknown problem: " + name);
  ocp->setReference(SX::sym("ref", state_size(name)));
  ocp->deployConstraintsAndAddCost();
  return ocp;
}

void sample_inputs(const std::string& name, int B, unsigned long long seed, std::vector<double>& frames,
                   std::vector<double>& refs) {
  SplitMix64 rng(seed);
  frames.clear();
  refs.clear();
  for (int b = 0; b < B; ++b) {
    if (name == "quadrotor") {
      for (int i = 0; i < 3; ++i) frames.push_back(rng.uniform(-1.0, 1.0));
      for (int i = 0; i < 3; ++i) frames.push_back(rng.uniform(-0.3, 0.3));
      for (int i = 0; i < 3; ++i) frames.push_back(rng.uniform(-0.5, 0.5));
      for (int i = 0; i < 3; ++i) frames.push_back(rng.uniform(-0.5, 0.5));
      for (int i = 0; i < 4; ++i) frames.push_back(kQuadMass * kGravity / 4.0);
      for (int i = 0; i < 12; ++i) refs.push_back(0.0);
    } else if (name == "cartpole") {
      frames.push_back(rng.uniform(-0.1, 0.1));
      frames.push_back(rng.uniform(M_PI - 0.5, M_PI + 0.5));
      frames.push_back(rng.uniform(-0.1, 0.1));
      frames.push_back(rng.uniform(-0.1, 0.1));
      frames.push_back(0.0);
      for (int i = 0; i < 4; ++i) refs.push_back(0.0);
    } else if (name == "centroidal") {
      const double stance[4][3] = {{0.25, 0.15, 0.0}, {0.25, -0.15, 0.0}, {-0.25, 0.15, 0.0}, {-0.25, -0.15, 0.0}};
      const double com0[3] = {0.0, 0.0, 0.35};
      for (int i = 0; i < 3; ++i) frames.push_back(com0[i] + rng.uniform(-0.05, 0.05));
      for (int i = 0; i < 3; ++i) frames.push_back(rng.uniform(-0.05, 0.05));
      for (int i = 0; i < 6; ++i) frames.push_back(rng.uniform(-0.2, 0.2));
      for (int f = 0; f < 4; ++f) for (int i = 0; i < 3; ++i) frames.push_back(stance[f][i]);
      for (int f = 0; f < 4; ++f) { frames.push_back(0.0); frames.push_back(0.0); frames.push_back(kLegMass * kGravity / 4.0); }
      for (int i = 0; i < 3; ++i) refs.push_back(com0[i]);
      for (int i = 0; i < 9; ++i) refs.push_back(0.0);
      for (int f = 0; f < 4; ++f) for (int i = 0; i < 3; ++i) refs.push_back(stance[f][i]);
    } else {
      throw std::invalid_argument("unknown problem: " + name);
    }
  }
}

// ---------------------------------------------------------------- test/test.cpp cases
KatCase make_kat(int id) {
  const double INF = std::numeric_limits<float>::infinity();  // test/test.cpp:11
  KatCase k;
  switch (id) {
    case 1: {  // test/test.cpp:13-36
      SX xs = SX::sym("x", 2);
      k.nlp = {{"x", xs}, {"f", pow(xs(0), 2) + pow(xs(1), 2)}, {"g", SX::vertcat({xs(0) + xs(1) - 1})}, {"p", SX()}};
      k.arg = {{"lbx", {-50, -100}}, {"ubx", {50, 100}}, {"lbg", {-0.00}}, {"ubg", {0.00}}, {"p", {}}};
      k.expected = {0.5, 0.5};
      k.description = "equality-constrained QP";
      break;
    }
    case 2: {  // :38-59
      SX xs = SX::sym("x", 2);
      k.nlp = {{"x", xs}, {"f", pow(xs(0) - 3, 2) + pow(xs(1) + 2, 2)}, {"g", SX()}, {"p", SX()}};
      k.arg = {{"lbx", {-50, -100}}, {"ubx", {50, 100}}, {"lbg", {}}, {"ubg", {}}, {"p", {}}};
      k.expected = {3, -2};
      k.description = "unconstrained QP";
      break;
    }
    case 3: {  // :61-84
      SX xs = SX::sym("x", 2);
      k.nlp = {{"x", xs}, {"f", pow(xs(0) - 2, 2) + pow(xs(1) - 3, 2)}, {"g", SX::vertcat({xs(0) + xs(1) - 1})}, {"p", SX()}};
      k.arg = {{"lbx", {-100, -100}}, {"ubx", {100, 100}}, {"lbg", {1}}, {"ubg", {INF}}, {"p", {}}};
      k.expected = {2, 3};
      k.description = "inactive inequality";
      break;
    }
    case 4: {  // :86-110
      SX xs = SX::sym("x", 2);
      k.nlp = {{"x", xs}, {"f", pow(xs(0), 2) + pow(xs(1), 2)}, {"g", SX::vertcat({xs(0), xs(1)})}, {"p", SX()}};
      k.arg = {{"lbx", {-100, -100}}, {"ubx", {100, 100}}, {"lbg", {1, 2}}, {"ubg", {INF, INF}}, {"p", {}}};
      k.expected = {1, 2};
      k.description = "two active inequalities";
      break;
    }
    case 5: {  // :112-136
      SX xs = SX::sym("x", 3);
      k.nlp = {{"x", xs},
               {"f", pow(xs(0) - 1, 2) + pow(xs(1) - 2, 2) + pow(xs(2) - 3, 2)},
               {"g", SX::vertcat({xs(0) + xs(1) + xs(2) - 5})},
               {"p", SX()}};
      k.arg = {{"lbx", {0, 0, 0}}, {"ubx", {INF, INF, INF}}, {"lbg", {0}}, {"ubg", {0}}, {"p", {}}};
      k.expected = {2.0 / 3.0, 5.0 / 3.0, 8.0 / 3.0};
      k.description = "equality + non-negativity";
      break;
    }
    case 6: {  // :138-161
      SX xs = SX::sym("x", 2);
      SX p = SX::sym("p", 1);
      k.nlp = {{"x", xs}, {"f", pow(xs(0) - p, 2) + pow(xs(1), 2)}, {"g", SX()}, {"p", p}};
      k.arg = {{"lbx", {-100, -100}}, {"ubx", {100, 100}}, {"lbg", {}}, {"ubg", {}}, {"p", {5.0}}};
      k.expected = {5, 0};
      k.description = "parametric objective";
      break;
    }
    case 7: {  // :163-185
      SX xs = SX::sym("x", 2);
      k.nlp = {{"x", xs}, {"f", pow(xs(0) - 3, 2) + pow(xs(1) - 4, 2)}, {"g", SX()}, {"p", SX()}};
      k.arg = {{"lbx", {0, 0}}, {"ubx", {2, 3}}, {"lbg", {}}, {"ubg", {}}, {"p", {}}};
      k.expected = {2, 3};
      k.description = "box-constrained QP";
      break;
    }
    case 8: {  // :187-211, non-convex: no known answer
      SX xs = SX::sym("x", 2);
      k.nlp = {{"x", xs}, {"f", pow(xs(0), 2) - pow(xs(1), 2)}, {"g", SX::vertcat({pow(xs(0), 2) + pow(xs(1), 2) - 1})}, {"p", SX()}};
      k.arg = {{"lbx", {-100, -100}}, {"ubx", {100, 100}}, {"lbg", {-INF}}, {"ubg", {1}}, {"p", {}}};
      k.description = "non-convex (solver dependent)";
      break;
    }
    default: throw std::invalid_argument("KAT id must be 1..8");
  }
  return k;
}

}  // namespace ocp_problems

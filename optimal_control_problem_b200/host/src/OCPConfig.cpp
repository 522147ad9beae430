// OCPConfig implementation.  Behaviour follows the reference
// (src/OCP_config/OCPConfig.cpp): which YAML keys are read (:90-92, :119-218), how
// ".inf" strings map to +-infinity, one frame of bounds replicated `horizon`
// times (:293-328), X = SX::sym("X", horizon * frameSize) stage-major (:102, :37).
#include "optimal_control_problem/OCP_config/OCPConfig.h"

#include <cstdlib>

bool ocp_b200_log_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("OCP_B200_LOG");
    return e && *e && std::string(e) != "0";
  }();
  return on;
}

namespace {
// one bound entry: ".inf"/"-.inf" spellings or a number (reference OCPConfig.cpp:147-165)
double parseBound(const YAML::Node& item, const std::string& var, size_t i) {
  if (!item.IsScalar()) {
    OCP_ERROR("bound " << i << " of variable " << var << " is not a scalar; 0 is used");
    return 0.0;
  }
  const std::string s = item.as<std::string>();
  if (s == ".inf" || s == ".Inf" || s == ".INF") return casadi::inf;
  if (s == "-.inf" || s == "-.Inf" || s == "-.INF") return -casadi::inf;
  return std::stod(s);  // throws std::invalid_argument like the reference
}

casadi::SX parseBoundVector(const YAML::Node& var, const char* key, const std::string& name, int size) {
  if (!var[key]) throw std::invalid_argument(std::string("Missing ") + key + " for variable: " + name);
  const YAML::Node seq = var[key];
  casadi::SX out = casadi::SX::zeros(size);
  if (!seq.IsSequence()) {
    OCP_WARN(key << " of variable " << name << " is not a sequence; zeros are used");
    return out;
  }
  if (static_cast<int>(seq.size()) != size)
    OCP_WARN(key << " of variable " << name << " has " << seq.size() << " entries, expected " << size);
  for (size_t i = 0; i < seq.size() && static_cast<int>(i) < size; ++i)
    out(static_cast<int>(i)) = parseBound(seq[i], name, i);
  return out;
}
}  // namespace

OCPConfig::OCPConfig(YAML::Node node) {
  horizon_ = 10;
  dt_ = 0.1;
  dt_ = node["discretization_settings"]["dt"].as<double>();
  horizon_ = node["discretization_settings"]["horizon"].as<int>();
  verbose_ = node["solver_settings"]["verbose"].as<bool>();
  OCP_INFO("dt: " << dt_ << ", horizon: " << horizon_);
  parseOCPBounds(node);
  variables_ = casadi::SX::sym("X", horizon_ * variableFrame_.totalSize, 1);
  OCP_INFO("frame size: " << variableFrame_.totalSize << ", decision variables: " << variables_.size1());
}

void OCPConfig::initializeFrame(Frame& frame, const YAML::Node& config) {
  frame.totalSize = 0;
  frame.fields.clear();
  frame.fieldOffsets.clear();
  for (const auto& field : config) {
    if (!field["name"]) throw std::invalid_argument("Field name not found in frame");
    const std::string name = field["name"].as<std::string>();
    if (!field["size"]) throw std::invalid_argument("Field size not found in frame");
    const int size = field["size"].as<int>();
    if (size <= 0) throw std::invalid_argument("Field size must be positive: " + name);
    frame.fields.emplace_back(name, size);
    frame.fieldOffsets[name] = frame.totalSize;
    frame.totalSize += size;
  }
}

void OCPConfig::parseOCPBounds(YAML::Node node) {
  if (!node["OCP_variables"]) throw std::invalid_argument("node [OCP_variables] not found in YAML file");
  const YAML::Node frame = node["OCP_variables"];
  if (!frame.IsSequence()) throw std::invalid_argument("status_frame should be a sequence");
  initializeFrame(variableFrame_, frame);
  casadi::SXVector lower, upper;
  for (size_t v = 0; v < frame.size(); ++v) {
    const YAML::Node var = frame[v];
    const std::string name = var["name"].as<std::string>();
    const int size = var["size"].as<int>();
    lower.push_back(parseBoundVector(var, "lower_bound", name, size));
    upper.push_back(parseBoundVector(var, "upper_bound", name, size));
  }
  coverLowerBounds(casadi::SX::vertcat(lower));
  coverUpperBounds(casadi::SX::vertcat(upper));
}

void OCPConfig::coverLowerBounds(const casadi::SX& oneFrameLowerBound) {
  lowerBounds_.assign(horizon_, casadi::DM(oneFrameLowerBound));
}

void OCPConfig::coverUpperBounds(const casadi::SX& oneFrameUpperBound) {
  upperBounds_.assign(horizon_, casadi::DM(oneFrameUpperBound));
}

casadi::SX OCPConfig::getVariable(int stepID, const std::string& variableName) const {
  if (stepID < 0 || stepID >= horizon_) throw std::out_of_range("Frame ID out of range");
  auto off = variableFrame_.fieldOffsets.find(variableName);
  if (off == variableFrame_.fieldOffsets.end()) throw std::invalid_argument("Field name not found in frame");
  int size = 0;
  for (const auto& f : variableFrame_.fields)
    if (f.first == variableName) { size = f.second; break; }
  const int start = stepID * variableFrame_.totalSize + off->second;
  return variables_(casadi::Slice(start, start + size));
}

casadi::SX OCPConfig::getVariables() const { return variables_; }
std::vector<casadi::DM> OCPConfig::getLowerBounds() const {
  if (lowerBounds_.empty()) OCP_WARN("lower bounds requested but empty");
  return lowerBounds_;
}
std::vector<casadi::DM> OCPConfig::getUpperBounds() const {
  if (upperBounds_.empty()) OCP_WARN("upper bounds requested but empty");
  return upperBounds_;
}
int OCPConfig::getHorizon() const { return horizon_; }
double OCPConfig::getDt() const { return dt_; }
int OCPConfig::getFrameSize() const { return variableFrame_.totalSize; }

void OCPConfig::setInitialGuess(const casadi::DM& initialGuess) {
  const int expected = horizon_ * variableFrame_.totalSize;
  if (initialGuess.size1() != expected)
    throw std::invalid_argument("initial guess has " + std::to_string(initialGuess.size1()) +
                                " entries, expected " + std::to_string(expected));
  initialGuess_ = initialGuess;
}
casadi::DM OCPConfig::getInitialGuess() { return initialGuess_; }

// CasadiCInterop -- a CasADi-generated C file as the stage library.
//
// The reference's local system is ONE casadi::Function, localSystemFunction(p, x, l, u) -> (H, grad f, J, l - c,
// u - c) (src/sqp_solver/SQPOptimizationSolver.cpp:74-77), which `gen_code: true` serialises
// (src/OptimalControlProblem.cpp:403-425).  CasADi can emit such a function as C (`Function::generate`,
// `CodeGenerator`): `casadi_real` / `casadi_int`, the entry point
//     int f(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem)
// and compact-CCS `f_sparsity_in / f_sparsity_out`, `f_work`.  This file turns such a C file -- from the real
// CasADi, or from casadi-lite's emitter, which writes the same layout -- into a library implementing
// include/ocp_b200_model.h:
//   1. the C file is compiled for the HOST (cc -shared) and its sparsity / work functions are called to get
//      np, N, m and the CCS patterns of H and J;
//   2. a CUDA wrapper #includes the very same C file with every function turned into a __device__ function
//      (CASADI_SYMBOL_EXPORT and `static` are re-defined around the #include) and adds one kernel in which each
//      thread evaluates the function for one instance, its work vector in a global workspace, writing straight
//      into the batched H / q / A / l / u buffers;
//   3. nvcc (sm_100a) -> shared library -> ocp_b200_create(model_library = ...).
// One thread per instance over the whole horizon is the compatibility path: the warp-per-(instance, stage)
// templates of StageCodegen are 20-50x faster and remain the default; parity between the two is tested.
#include <dlfcn.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "optimal_control_problem/codegen/StageCodegen.h"

namespace ocp_codegen {
namespace {

std::string sh_quote(const std::string& v) {
  std::string out = "'";
  for (char c : v) {
    if (c == '\'') out += "'\\''";
    else out += c;
  }
  return out + "'";
}

std::string run(const std::string& cmd, int* rc) {
  FILE* pipe = popen((cmd + " 2>&1").c_str(), "r");
  if (!pipe) { *rc = -1; return "cannot start: " + cmd; }
  std::string log;
  char buf[512];
  while (fgets(buf, sizeof(buf), pipe)) log += buf;
  *rc = pclose(pipe);
  return log;
}

bool valid_symbol(const std::string& s) {
  return !s.empty() && s.size() <= 96 &&
         s.find_first_not_of("ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789_") == std::string::npos;
}

struct Ccs { long long nrow = 0, ncol = 0; std::vector<int> colptr, rowidx; };

// CasADi's compact CCS: {nrow, ncol, colind[ncol + 1], row[nnz]}, or {nrow, ncol, 1} for a dense pattern
Ccs decode(const long long* sp) {
  Ccs c;
  c.nrow = sp[0]; c.ncol = sp[1];
  const bool dense = sp[2] != 0;   // colind[0] is always 0 in the sparse form
  c.colptr.resize(c.ncol + 1);
  if (dense) {
    for (long long j = 0; j <= c.ncol; ++j) c.colptr[j] = static_cast<int>(j * c.nrow);
    for (long long j = 0; j < c.ncol; ++j)
      for (long long i = 0; i < c.nrow; ++i) c.rowidx.push_back(static_cast<int>(i));
  } else {
    for (long long j = 0; j <= c.ncol; ++j) c.colptr[j] = static_cast<int>(sp[2 + j]);
    for (int k = 0; k < c.colptr[c.ncol]; ++k) c.rowidx.push_back(static_cast<int>(sp[2 + c.ncol + 1 + k]));
  }
  return c;
}

unsigned long long fnv(const std::string& s) {
  unsigned long long h = 1469598103934665603ULL;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
  return h;
}

}  // namespace

std::string compile_casadi_c(const std::string& c_file, const std::string& local_system_fn, const std::string& objective_fn,
                             const std::string& name, int nf, int horizon, const std::string& code_dir, bool verbose) {
  namespace fs = std::filesystem;
  if (!valid_symbol(local_system_fn) || !valid_symbol(objective_fn) || !valid_symbol(name))
    throw std::invalid_argument("casadi C interop: function / problem names must match [A-Za-z0-9_]+");
  if (!fs::exists(c_file)) throw std::invalid_argument("casadi C interop: no such file: " + c_file);
  fs::create_directories(code_dir);
  std::string text;
  {
    std::ifstream in(c_file);
    std::stringstream ss;
    ss << in.rdbuf();
    text = ss.str();
  }
  char hx[32];
  std::snprintf(hx, sizeof(hx), "%016llx", fnv(text + local_system_fn + objective_fn + std::to_string(nf) + "x" + std::to_string(horizon)));
  const std::string stem = (fs::path(code_dir) / (name + "_casadi_" + hx)).string();
  const std::string so = stem + ".so";
  if (fs::exists(so)) return so;
  const std::string pid = std::to_string(static_cast<long>(::getpid()));
  const std::string abs_c = fs::absolute(c_file).string();

  // 1. host build: sparsity patterns and work sizes
  const std::string host_so = stem + ".host.tmp." + pid + ".so";
  const char* cc_env = std::getenv("OCP_B200_CC");
  int rc = 0;
  std::string log = run(sh_quote(cc_env ? cc_env : "cc") + " -shared -fPIC -O1 -x c " + sh_quote(abs_c) + " -o " + sh_quote(host_so) + " -lm", &rc);
  if (rc != 0) throw std::runtime_error("casadi C interop: host compilation of " + c_file + " failed\n" + log);
  void* h = dlopen(host_so.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) { std::error_code ec; fs::remove(host_so, ec); throw std::runtime_error(std::string("casadi C interop: dlopen: ") + dlerror()); }
  typedef const long long* (*sp_fn)(long long);
  typedef long long (*n_fn)(void);
  typedef int (*work_fn)(long long*, long long*, long long*, long long*);
  auto sym = [&](const std::string& s) {
    void* p = dlsym(h, s.c_str());
    if (!p) { const std::string msg = "casadi C interop: symbol '" + s + "' is missing in " + c_file; dlclose(h); std::error_code ec; fs::remove(host_so, ec); throw std::runtime_error(msg); }
    return p;
  };
  Ccs sp_in[4], sp_out[5];
  long long sz_arg = 0, sz_res = 0, sz_iw = 0, sz_w = 0, osz_arg = 0, osz_res = 0, osz_iw = 0, osz_w = 0;
  {
    if (reinterpret_cast<n_fn>(sym(local_system_fn + "_n_in"))() != 4 || reinterpret_cast<n_fn>(sym(local_system_fn + "_n_out"))() != 5) {
      dlclose(h); std::error_code ec; fs::remove(host_so, ec);
      throw std::runtime_error("casadi C interop: " + local_system_fn + " must map (p, x, l, u) -> (H, grad, J, l', u')");
    }
    sp_fn si = reinterpret_cast<sp_fn>(sym(local_system_fn + "_sparsity_in")), so_ = reinterpret_cast<sp_fn>(sym(local_system_fn + "_sparsity_out"));
    for (int i = 0; i < 4; ++i) sp_in[i] = decode(si(i));
    for (int i = 0; i < 5; ++i) sp_out[i] = decode(so_(i));
    reinterpret_cast<work_fn>(sym(local_system_fn + "_work"))(&sz_arg, &sz_res, &sz_iw, &sz_w);
    if (reinterpret_cast<n_fn>(sym(objective_fn + "_n_in"))() != 2 || reinterpret_cast<n_fn>(sym(objective_fn + "_n_out"))() != 1) {
      dlclose(h); std::error_code ec; fs::remove(host_so, ec);
      throw std::runtime_error("casadi C interop: " + objective_fn + " must map (p, x) -> f");
    }
    reinterpret_cast<work_fn>(sym(objective_fn + "_work"))(&osz_arg, &osz_res, &osz_iw, &osz_w);
  }
  dlclose(h);
  { std::error_code ec; fs::remove(host_so, ec); }
  const long long np = sp_in[0].nrow * sp_in[0].ncol, N = sp_in[1].nrow * sp_in[1].ncol, m = sp_in[2].nrow * sp_in[2].ncol;
  const long long n = np + N, ng = m - n;
  if (N != static_cast<long long>(nf) * horizon || ng < 0 || sp_out[0].nrow != n || sp_out[0].ncol != n || sp_out[2].nrow != m ||
      sp_out[2].ncol != n || sp_out[1].nrow * sp_out[1].ncol != n || sp_out[3].nrow * sp_out[3].ncol != m)
    throw std::runtime_error("casadi C interop: dimensions of " + local_system_fn + " do not describe an augmented local system "
                             "(w = [p; x], c = [p; x; g]) with the given stage layout");
  for (int k : {1, 3, 4})
    if (static_cast<long long>(sp_out[k].rowidx.size()) != sp_out[k].nrow * sp_out[k].ncol)
      throw std::runtime_error("casadi C interop: gradient and bounds must be dense outputs");

  // 2. CUDA wrapper around the same C file
  std::ostringstream s;
  auto arr = [&](const char* nm, const std::vector<int>& v) {
    s << "static const int " << nm << "[" << (v.empty() ? 1 : v.size()) << "] = {";
    for (size_t i = 0; i < v.size(); ++i) s << (i ? "," : "") << v[i];
    if (v.empty()) s << "0";
    s << "};\n";
  };
  s << "// generated by CasadiCInterop.cpp: stage library (include/ocp_b200_model.h) around a CasADi-format C file\n"
    << "#include <cuda_runtime.h>\n#include <math.h>\n#include <stdio.h>\n#include \"ocp_b200_model.h\"\n\n"
    << "// every function of the C file becomes a __device__ function; its headers were included above (guards)\n"
    << "#define CASADI_SYMBOL_EXPORT __device__\n#define static __device__ static\n"
    << "#include " << '"' << abs_c << '"' << "\n#undef static\n#undef CASADI_SYMBOL_EXPORT\n\n";
  arr("H_COLPTR", sp_out[0].colptr); arr("H_ROWIDX", sp_out[0].rowidx); arr("A_COLPTR", sp_out[2].colptr); arr("A_ROWIDX", sp_out[2].rowidx);
  s << "#define NP " << np << "\n#define NF " << nf << "\n#define HORIZON " << horizon << "\n#define NX " << N << "\n#define NG " << ng
    << "\n#define NN " << n << "\n#define MM " << m << "\n#define SZ_W " << std::max<long long>(1, std::max(sz_w, osz_w))
    << "\n#define SZ_IW " << std::max<long long>(1, std::max(sz_iw, osz_iw)) << "\n#define SZ_ARG " << std::max<long long>(4, std::max(sz_arg, osz_arg))
    << "\n#define SZ_RES " << std::max<long long>(5, std::max(sz_res, osz_res)) << "\n\n";
  s << "static const ocp_b200_model_info INFO = {OCP_B200_MODEL_ABI_VERSION, NP, NF, HORIZON, NG, NN, MM, " << sp_out[0].rowidx.size() << ", "
    << sp_out[2].rowidx.size() << ", H_COLPTR, H_ROWIDX, A_COLPTR, A_ROWIDX, 1, 1, \"" << name << "\", 0x" << hx << "ULL};\n"
    << "extern \"C\" const ocp_b200_model_info* ocp_b200_model_get_info(void) { return &INFO; }\n\n";
  s << R"(
// per-instance workspace: [l_full (MM) | u_full (MM) | w (SZ_W)] doubles and SZ_IW casadi_ints, grown on demand
static double* g_work = nullptr; static long long* g_iwork = nullptr; static size_t g_cap = 0;
static int reserve(int B) {
  if (static_cast<size_t>(B) <= g_cap) return 0;
  if (g_work) cudaFree(g_work);
  if (g_iwork) cudaFree(g_iwork);
  g_work = nullptr; g_iwork = nullptr; g_cap = 0;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&g_work), sizeof(double) * size_t(B) * (2 * size_t(MM) + SZ_W));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&g_iwork), sizeof(long long) * size_t(B) * SZ_IW);
  if (e != cudaSuccess) return static_cast<int>(e);
  g_cap = B;
  return 0;
}

__global__ void casadi_assemble_kernel(int B, const double* __restrict__ x, const double* __restrict__ p,
                                       const double* __restrict__ frames, const double* __restrict__ lbx,
                                       const double* __restrict__ ubx, const double* __restrict__ lbg,
                                       const double* __restrict__ ubg, double* h_vals, int ld_h, double* q, int ld_n,
                                       double* a_vals, int ld_a, double* l, double* u, int ld_m, double* work, long long* iwork) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double* lf = work + size_t(b) * (2 * size_t(MM) + SZ_W);
  double* uf = lf + MM;
  double* w = uf + MM;
  const double* pb = p + size_t(b) * NP;
  // l = [p; lbx; lbg], u = [p; ubx; ubg] with the first frame pinned (OptimalControlProblem.cpp:93-99)
  for (int i = 0; i < NP; ++i) { lf[i] = pb[i]; uf[i] = pb[i]; }
  for (int i = 0; i < NX; ++i) {
    const bool pin = frames != nullptr && i < NF;
    lf[NP + i] = pin ? frames[size_t(b) * NF + i] : lbx[i];
    uf[NP + i] = pin ? frames[size_t(b) * NF + i] : ubx[i];
  }
  for (int i = 0; i < NG; ++i) { lf[NN + i] = lbg[i]; uf[NN + i] = ubg[i]; }
  const casadi_real* arg[SZ_ARG] = {pb, x + size_t(b) * NX, lf, uf};
  casadi_real* res[SZ_RES] = {h_vals + size_t(b) * ld_h, q + size_t(b) * ld_n, a_vals + size_t(b) * ld_a, l + size_t(b) * ld_m,
                              u + size_t(b) * ld_m};
  LOCAL_SYSTEM_FN(arg, res, reinterpret_cast<casadi_int*>(iwork + size_t(b) * SZ_IW), w, 0);
}

__global__ void casadi_objective_kernel(int B, const double* __restrict__ x, const double* __restrict__ p, double* f,
                                        double* work, long long* iwork) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double* w = work + size_t(b) * (2 * size_t(MM) + SZ_W) + 2 * size_t(MM);
  const casadi_real* arg[SZ_ARG] = {p + size_t(b) * NP, x + size_t(b) * NX};
  casadi_real* res[SZ_RES] = {f + b};
  OBJECTIVE_FN(arg, res, reinterpret_cast<casadi_int*>(iwork + size_t(b) * SZ_IW), w, 0);
}

extern "C" int ocp_b200_model_assemble(int B, const double* x, const double* p, const double* frames, const double* lbx,
                                       const double* ubx, const double* lbg, const double* ubg, double* h_vals, int ld_h,
                                       double* q, int ld_n, double* a_vals, int ld_a, double* l, double* u, int ld_m,
                                       void* stream) {
  if (B <= 0) return 0;
  if (int rc = reserve(B)) return rc;
  casadi_assemble_kernel<<<(B + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
      B, x, p, frames, lbx, ubx, lbg, ubg, h_vals, ld_h, q, ld_n, a_vals, ld_a, l, u, ld_m, g_work, g_iwork);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ocp_b200_model_objective(int B, const double* x, const double* p, double* f, void* stream) {
  if (B <= 0) return 0;
  if (int rc = reserve(B)) return rc;
  casadi_objective_kernel<<<(B + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(B, x, p, f, g_work, g_iwork);
  return static_cast<int>(cudaGetLastError());
}
)";
  std::string cu_text = s.str();
  auto replace_all = [&](const std::string& from, const std::string& to) {
    for (size_t pos = 0; (pos = cu_text.find(from, pos)) != std::string::npos; pos += to.size()) cu_text.replace(pos, from.size(), to);
  };
  replace_all("LOCAL_SYSTEM_FN", local_system_fn);
  replace_all("OBJECTIVE_FN", objective_fn);
  const std::string cu_tmp = stem + ".tmp." + pid + ".cu", so_tmp = so + ".tmp." + pid;
  {
    std::ofstream f(cu_tmp);
    f << cu_text;
    f.close();
    if (!f.good()) throw std::runtime_error("casadi C interop: cannot write " + cu_tmp);
  }
  const char* nvcc_env = std::getenv("OCP_B200_NVCC");
  const std::string cmd = sh_quote(nvcc_env ? nvcc_env : "nvcc") + " -gencode arch=compute_100a,code=sm_100a -O1 -lineinfo -std=c++17 -shared "
                          "-Xcompiler -fPIC -I" + sh_quote(include_dir()) + " -o " + sh_quote(so_tmp) + " " + sh_quote(cu_tmp);
  if (verbose) std::printf("compiling CasADi-C stage library: %s\n", cmd.c_str());
  log = run(cmd, &rc);
  std::error_code ec;
  if (rc != 0 || !fs::exists(so_tmp)) {
    fs::rename(cu_tmp, stem + ".failed.cu", ec);
    throw std::runtime_error("casadi C interop: nvcc failed for " + stem + ".failed.cu\n" + log);
  }
  fs::rename(cu_tmp, stem + ".cu", ec);
  fs::rename(so_tmp, so);
  return so;
}

}  // namespace ocp_codegen

"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Builds oracle/_build/liboracle.so: the CPU restatement of the reference's SQP + OSQP path
(oracle/sqp_oracle.hpp, oracle/osqp_restate.hpp) behind the C entry points of
oracle/oracle_capi.cpp.  It links the device-free half of the C++ front-end (problem
definitions, casadi-lite symbolic layer) and nothing CUDA; the product never loads it.

The reference itself (CasADi + OSQP v1.0.0.beta1 + OSQP-Eigen 0.9.0, none of them in
/root/reference or in this image) cannot be compiled here, so there is no oracle/_ref:
PARITY IS UNPINNED against the real OSQP; see the header of osqp_restate.hpp for what pins it and
tools/pin_reference.py for the recipe that produces real-library fixtures elsewhere.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
OUT = HERE / "_build"

sys.path.insert(0, str(ROOT))
from optimal_control_problem_b200 import _build as B  # noqa: E402  (shared compile helpers only)

FLAGS = ["-std=c++17", "-O3", "-march=native", "-fPIC", "-I" + str(B.PKG / "host"), "-I" + str(B.PKG),
         "-I" + str(B.INCLUDE), "-I" + str(HERE)]


def build(force: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    out = OUT / "liboracle.so"
    objs = B.compile_objects(B.FRONTEND_SOURCES, FLAGS, OUT / "obj", tag="oracle")
    capi = OUT / "obj" / "oracle_capi.o"
    deps = [HERE / "oracle_capi.cpp", HERE / "sqp_oracle.hpp", HERE / "osqp_restate.hpp"] + B._headers()
    if force or B._stale(capi, deps):
        B._run([B.CXX, *FLAGS, "-c", str(HERE / "oracle_capi.cpp"), "-o", str(capi)])
    if force or B._stale(out, objs + [capi]):
        B._run([B.CXX, "-shared", "-o", str(out), str(capi), *map(str, objs), "-pthread"])
    res = subprocess.run(["nm", "-D", "--undefined-only", str(out)], stdout=subprocess.PIPE, text=True)
    if "ocp_b200_" in res.stdout or "cuda" in res.stdout.lower():
        raise RuntimeError("liboracle.so must not depend on the device library")
    return out


if __name__ == "__main__":
    print(build("--force" in sys.argv))

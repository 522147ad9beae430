#!/bin/bash
# Builds A/B variants of the kernels into optimal_control_problem_b200/lib_ab/<name>/ (both libraries, so that
# OCP_B200_LIB_DIR=<that directory> selects the variant).  usage: tools/ab_build.sh name "-DMACRO=1 ..." [name flags]...
# AB_SOURCES: the translation units that see the macros (default: direct_compact).
set -e
cd "$(dirname "$0")/.."
PKG=optimal_control_problem_b200
SRCS=${AB_SOURCES:-direct_compact}
python -c "from optimal_control_problem_b200 import _build; _build.build_host()"
NV="nvcc -gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=$PKG/lib_ab/$name; mkdir -p $d
  (
    pids=""
    for s in $SRCS; do
      $NV -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -I$PKG/csrc $flags -c $PKG/csrc/$s.cu -o $d/$s.cu.o & pids="$pids $!"
    done
    for p in $pids; do wait $p; done
    objs=""
    for o in $PKG/lib/obj/*.cu.o; do
      b=$(basename $o .cu.o); if [ -f $d/$b.cu.o ]; then objs="$objs $d/$b.cu.o"; else objs="$objs $o"; fi
    done
    $NV -shared -o $d/libocp_b200.so $objs -ldl && cp $PKG/lib/libocp_b200_host.so $d/ && rm $d/*.cu.o && echo built $name
  ) &
done
wait

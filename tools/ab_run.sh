python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compact or repeated" 2>&1 | tail -15
for v in $VARIANTS; do
  OCP_B200_LIB_DIR=/root/repo/optimal_control_problem_b200/lib_ab/$v python tools/probe_plan.py compact 4 2>&1 | tail -1 | sed "s/^/$v /"
done

# usage: VARIANTS="a b" CMD="python ..." bash tools/ab_run.sh  -- runs CMD once per variant library (lib_ab/<name>)
for v in $VARIANTS; do
  OCP_B200_LIB_DIR=/root/repo/optimal_control_problem_b200/lib_ab/$v bash -c "$CMD" 2>&1 | tail -1 | sed "s/^/$v /"
done

# usage: VARIANTS="main a b" CMD="python ..." bash tools/ab_run.sh  -- runs CMD once per variant library (lib_ab/<name>;
# "main" = the regular build in lib/)
for v in $VARIANTS; do
  if [ "$v" = main ]; then bash -c "$CMD" 2>&1 | tail -1 | sed "s/^/$v /"
  else OCP_B200_LIB_DIR=/root/repo/optimal_control_problem_b200/lib_ab/$v bash -c "$CMD" 2>&1 | tail -1 | sed "s/^/$v /"; fi
done

#!/bin/bash
# Round profiling pass (run under gpurun, one GPU):  bash tools/profile_all.sh
# 1. launch list of the default bench command line (shares of the step, not absolutes)
# 2. one `--set full` capture of a timed-region launch of the step kernel for every BASELINE.json workload
set -u
mkdir -p gpurun_out
P="python bench.py --steps 20 --warmup 3 --mode-steps 5 --cpu-seconds 0"
[ "${SKIP_LAUNCH_LIST:-0}" = 1 ] || $P > gpurun_out/plain_bench.log 2>&1 &&
[ "${SKIP_LAUNCH_LIST:-0}" = 1 ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_launches.log 2>&1
STEPK='regex:st_main_kernel<\(int\)[12], \(int\)[012], \(int\)0, unsigned [a-z ]*, \(bool\)[01]>'
TPEK='regex:st_step_tpe_kernel'
for spec in C2:85 C3:9:tpe C4:4 C5a:4 C5b:5:tpe; do
  W=${spec%%:*}; rest=${spec#*:}; [ -n "${ONLY:-}" ] && [ "$ONLY" != "$W" ] && continue; S=${rest%%:*}; K="$STEPK"; [ "${rest##*:}" = tpe ] && K="$TPEK"
  Q="python bench.py --workload $W --steps 4 --warmup 3 --modes none --cpu-seconds 0 --no-e2e"
  $Q > gpurun_out/plain_$W.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s $S -c 1 -f -o gpurun_out/prof_$W $Q > gpurun_out/ncu_$W.log 2>&1
  tail -1 gpurun_out/ncu_$W.log
done

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
COLK='regex:st_step_cols_kernel'
# workload : launches to skip (burn-ins, warm-up) : kernel family [: steps per launch]
for spec in C2:90:cols C3:12:tpe C4:4:main C5a:4:main C5b:6:tpe C2_T32:8:cols:32 C3_T32:5:tpe:32; do
  IFS=: read -r WN S FAM T <<< "$spec"; T=${T:-1}; W=${WN%%_*}
  [ -n "${ONLY:-}" ] && [ "$ONLY" != "$WN" ] && continue
  K="$STEPK"; [ "$FAM" = tpe ] && K="$TPEK"; [ "$FAM" = cols ] && K="$COLK"
  Q="python bench.py --workload $W --steps 4 --warmup 3 --modes none --cpu-seconds 0 --no-e2e --steps-per-launch $T"
  W=$WN
  $Q > gpurun_out/plain_$W.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s $S -c 1 -f -o gpurun_out/prof_$W $Q > gpurun_out/ncu_$W.log 2>&1
  tail -1 gpurun_out/ncu_$W.log
  # the summaries are what travels back (gpurun merges at most 64 MiB): the reports of the image modes are dropped
  (echo "# gpurun_out/prof_$W.ncu-rep  (ncu --set full --clock-control none; per launch)"; echo; python tools/ncu_summary.py gpurun_out/prof_$W.ncu-rep; echo; python tools/ncu_stalls.py gpurun_out/prof_$W.ncu-rep) > gpurun_out/sum_$W.txt 2>&1
  case $W in C4|C5a|C5b|C3_T32) rm -f gpurun_out/prof_$W.ncu-rep;; esac
done

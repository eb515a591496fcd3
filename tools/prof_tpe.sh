Q="python bench.py --workload C3 --steps 4 --warmup 3 --modes none --cpu-seconds 0 --no-e2e"
export ST_B200_RAM_PATH=thread
$Q > gpurun_out/plain_tpe.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:st_step_tpe_kernel -s 4 -c 1 -f -o gpurun_out/prof_tpe_C3 $Q > gpurun_out/ncu_tpe.log 2>&1
tail -2 gpurun_out/ncu_tpe.log

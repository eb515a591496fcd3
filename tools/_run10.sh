mkdir -p gpurun_out
export ST_B200_RAM_PATH=thread
ST_B200_LIB=$PWD/gym_simpletetris_b200/libst_r2a.so python tools/knob_sweep.py C2:1048576,C2:262144,C3:262144 ST_B200_TPE_EPW=16 > gpurun_out/sweep10.log 2>&1
python tools/knob_sweep.py C2:1048576,C2:262144 ST_B200_TPE_EPW=16 ST_B200_TPE_WPC=4,8 ST_B200_TPE_STAGED=0,1 >> gpurun_out/sweep10.log 2>&1
ST_B200_NO_PDL=1 python tools/knob_sweep.py C2:1048576 ST_B200_TPE_EPW=16 >> gpurun_out/sweep10.log 2>&1
cat > /tmp/one.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench
torch.cuda.set_device(0)
bench.WORKLOADS["X"] = dict(n=1048576, kw=bench.WORKLOADS["C2"]["kw"], desc="x")
r = bench.time_workload("X", 4, 3, 0, 1, None, burn_in=60)
print(r["ms_per_step"])
P
ST_B200_TPE_EPW=16 ncu --set full --clock-control none --import-source on -k regex:st_step_tpe_kernel -s 70 -c 1 -f -o gpurun_out/prof_1M python /tmp/one.py > gpurun_out/ncu_1M.log 2>&1
tail -2 gpurun_out/ncu_1M.log
ST_B200_LIB=$PWD/gym_simpletetris_b200/libst_r2a.so ST_B200_TPE_EPW=16 ncu --set full --clock-control none --import-source on -k regex:st_step_tpe_kernel -s 70 -c 1 -f -o gpurun_out/prof_1M_old python /tmp/one.py > gpurun_out/ncu_1M_old.log 2>&1
tail -2 gpurun_out/ncu_1M_old.log

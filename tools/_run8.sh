mkdir -p gpurun_out
export ST_B200_RAM_PATH=thread
python tools/knob_sweep.py C2:8192,C2:12288,C2:16384,C2:24576,C2:32768,C2:49152,C3,C2:131072,C2:1048576 ST_B200_TPE_EPW=4,8,16,32 ST_B200_TPE_WPC=4 > gpurun_out/sweep8.log 2>&1
python tools/knob_sweep.py C3 ST_B200_TPE_EPW=16,32 ST_B200_TPE_WPC=2,8 >> gpurun_out/sweep8.log 2>&1
python tools/knob_sweep.py C5b:8192,C5b:16384,C5b:32768,C5b,C5b:262144 ST_B200_TPE_EPW=4,8,16 ST_B200_TPE_WPC=4 >> gpurun_out/sweep8.log 2>&1
ST_B200_RAM_PATH=warp python tools/knob_sweep.py C2:8192,C2:12288,C2:16384,C2:24576,C5b:8192,C5b:16384,C5b:32768 ST_B200_X=0 >> gpurun_out/sweep8.log 2>&1
python tools/knob_sweep.py C3,C2:4096,C2:16384 ST_B200_TPE_EPW=4,8,16 T=32 >> gpurun_out/sweep8.log 2>&1
ST_B200_RAM_PATH=warp python tools/knob_sweep.py C3,C2:4096,C2:16384 T=32 >> gpurun_out/sweep8.log 2>&1
unset ST_B200_RAM_PATH
ONLY=C3 SKIP_LAUNCH_LIST=1 bash tools/profile_all.sh

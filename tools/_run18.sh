mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
(ST_B200_TPE_STAGED=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thread" 2>&1 | tail -4)
export ST_B200_RAM_PATH=thread
python tools/knob_sweep.py C3,C2:16384,C2:32768,C5b,C3:262144 ST_B200_TPE_EPW=4,8,16,32 ST_B200_TPE_STAGED=0,1 > gpurun_out/sweep18.log 2>&1
python tools/knob_sweep.py C2:4096,C2:8192 ST_B200_TPE_EPW=4,8,16 ST_B200_TPE_WPC=1,4 ST_B200_TPE_STAGED=1 >> gpurun_out/sweep18.log 2>&1
python tools/knob_sweep.py C3,C2:8192,C5b ST_B200_TPE_EPW=8,16 ST_B200_TPE_STAGED=0,1 T=32 >> gpurun_out/sweep18.log 2>&1
export ST_B200_LIB=$PWD/gym_simpletetris_b200/libst_trace.so
( echo "=== C3 epw=16 wpc=4 staged"; ST_B200_TPE_EPW=16 ST_B200_TPE_WPC=4 ST_B200_TPE_STAGED=1 python tools/tpe_trace.py C3 2>&1 | tail -11 ) > gpurun_out/trace18.log 2>&1

mkdir -p gpurun_out
export ST_B200_RAM_PATH=thread
python tools/knob_sweep.py C2:655360,C2:1048576,C2:2097152 ST_B200_TPE_EPW=16 ST_B200_TPE_L2=0,1,2,3 > gpurun_out/sweep13.log 2>&1
echo "== old" >> gpurun_out/sweep13.log
ST_B200_LIB=$PWD/gym_simpletetris_b200/libst_r2a.so python tools/knob_sweep.py C2:1179648,C2:1310720,C2:1572864 ST_B200_TPE_EPW=16 >> gpurun_out/sweep13.log 2>&1

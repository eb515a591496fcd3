mkdir -p gpurun_out
export ST_B200_RAM_PATH=thread
echo "== product" > gpurun_out/sweep12.log
python tools/knob_sweep.py C2:524288,C2:655360,C2:786432,C2:917504,C2:1000000,C2:1048576,C2:1500000,C2:2097152 ST_B200_TPE_EPW=16 >> gpurun_out/sweep12.log 2>&1
echo "== old" >> gpurun_out/sweep12.log
ST_B200_LIB=$PWD/gym_simpletetris_b200/libst_r2a.so python tools/knob_sweep.py C2:524288,C2:786432,C2:1048576,C2:2097152 ST_B200_TPE_EPW=16 >> gpurun_out/sweep12.log 2>&1

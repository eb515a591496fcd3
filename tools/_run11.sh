mkdir -p gpurun_out
export ST_B200_RAM_PATH=thread
for v in plainst nocpa both; do echo "== $v"; ST_B200_LIB=$PWD/gym_simpletetris_b200/libst_$v.so python tools/knob_sweep.py C2:1048576,C2:524288,C3,C3:262144,C5b ST_B200_TPE_EPW=16; done > gpurun_out/sweep11.log 2>&1
echo "== product" >> gpurun_out/sweep11.log
python tools/knob_sweep.py C2:1048576,C2:524288,C3,C3:262144,C5b ST_B200_TPE_EPW=16 >> gpurun_out/sweep11.log 2>&1

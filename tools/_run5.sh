mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
export ST_B200_RAM_PATH=thread
python tools/knob_sweep.py C3,C2:4096,C2:16384,C5b,C3:262144 ST_B200_TPE_EPW=8,16 ST_B200_TPE_L2=0,1,2,3 > gpurun_out/sweep5.log 2>&1
python tools/knob_sweep.py C3,C2:4096 ST_B200_TPE_EPW=4,8,16 ST_B200_TPE_L2=0,3 T=32 >> gpurun_out/sweep5.log 2>&1
export ST_B200_LIB=$PWD/gym_simpletetris_b200/libst_trace.so
for cfg in "16 4 0 0 0" "16 4 0 0 3" "8 4 0 0 3"; do
  set -- $cfg
  echo "=== epw=$1 wpc=$2 cap=$3 staged=$4 l2=$5"
  ST_B200_TPE_EPW=$1 ST_B200_TPE_WPC=$2 ST_B200_TPE_CTAS_PER_SM=$3 ST_B200_TPE_STAGED=$4 ST_B200_TPE_L2=$5 python tools/tpe_trace.py C3 2>&1 | tail -11
done > gpurun_out/trace5.log 2>&1
echo "=== C2 4096 epw=16 wpc=1" >> gpurun_out/trace5.log
ST_B200_TPE_EPW=16 ST_B200_TPE_WPC=1 python tools/tpe_trace.py C2 >> gpurun_out/trace5.log 2>&1

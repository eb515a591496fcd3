mkdir -p gpurun_out
export ST_B200_LIB=$PWD/gym_simpletetris_b200/libst_trace.so
for cfg in "16 4 0 0" "8 4 0 0" "32 4 0 0" "16 4 0 1" "8 2 13 1"; do
  set -- $cfg
  echo "=== epw=$1 wpc=$2 cap=$3 staged=$4"
  ST_B200_TPE_EPW=$1 ST_B200_TPE_WPC=$2 ST_B200_TPE_CTAS_PER_SM=$3 ST_B200_TPE_STAGED=$4 python tools/tpe_trace.py C3 2>&1 | tail -12
done

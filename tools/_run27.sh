mkdir -p gpurun_out
(tools/ab.sh "product" C3:65536 ST_B200_TPE_EPW=32 ST_B200_TPE_WPC=1,2,4;
 tools/ab.sh "product" C3:65536 ST_B200_TPE_EPW=16 ST_B200_TPE_R72=0,1 ) > gpurun_out/ab27.log 2>&1
cat gpurun_out/ab27.log

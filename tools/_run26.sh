mkdir -p gpurun_out
(tools/ab.sh "product" C2:66304,C2:70000,C2:81920,C2:98304,C2:114688,C2:131072,C2:196608,C2:262144 ST_B200_TPE_EPW=16,32 ) > gpurun_out/ab26.log 2>&1
cat gpurun_out/ab26.log

mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t1.log 2>&1; tail -3 gpurun_out/t1.log
(tools/ab.sh "old product" C3:65536,C5b:65536,C2:16384,C2:32768,C3:262144 ST_B200_TPE_EPW=8,16,32; tools/ab.sh "old product" C2:4096,C2:8192 ST_B200_RAM_PATH=thread ST_B200_TPE_EPW=4,16; tools/ab.sh "old product" C3:65536,C2:16384 T=32 ) > gpurun_out/ab1.log 2>&1
cat gpurun_out/ab1.log

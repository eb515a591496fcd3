mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t2.log 2>&1; tail -3 gpurun_out/t2.log
(tools/ab.sh "product mb3" C3:65536,C5b:65536,C2:16384,C2:32768,C3:262144 ST_B200_TPE_SPEC=0,1 ST_B200_TPE_EPW=8,16; tools/ab.sh "product" C2:4096,C2:8192 ST_B200_RAM_PATH=thread ST_B200_TPE_SPEC=0,1 ST_B200_TPE_EPW=4,16; tools/ab.sh "product" C3:65536,C2:16384 ST_B200_TPE_SPEC=0,1 T=32 ) > gpurun_out/ab2.log 2>&1
cat gpurun_out/ab2.log

mkdir -p gpurun_out
# per-CTA smem at C3 epw16 wpc4 = 12.2 KB; pad so that k CTAs fit in 227 KB: k=2 -> 100000, 3 -> 62000, 4 -> 44000, 5 -> 33000, 6 -> 25500
(tools/ab.sh "product" C3:65536 ST_B200_TPE_SMEM_PAD=0,25500,33000,44000,62000,100000 ST_B200_TPE_EPW=16;
 tools/ab.sh "product" C3:65536 ST_B200_TPE_SMEM_PAD=20000,30000,40000,50000,70000 ST_B200_TPE_EPW=8;
 tools/ab.sh "product" C3:65536 ST_B200_TPE_SMEM_PAD=60000,100000,200000 ST_B200_TPE_EPW=16 ST_B200_TPE_WPC=8 ) > gpurun_out/ab12.log 2>&1
cat gpurun_out/ab12.log

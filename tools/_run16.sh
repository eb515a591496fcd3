mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
python tools/knob_sweep.py C2:4096,C2:8192,C2:12288,C2:16384,C2:32768,C3 ST_B200_RAM_PATH=warp,thread T=32 > gpurun_out/sweep16.log 2>&1
ST_B200_RAM_PATH=thread python tools/knob_sweep.py C2:8192,C2:32768,C3,C5b ST_B200_TPE_EPW=4,8,16,32 T=32 >> gpurun_out/sweep16.log 2>&1
(time python bench.py) > gpurun_out/bench16.log 2> gpurun_out/bench16.err; tail -3 gpurun_out/bench16.err

"""SM clock / throttle reasons / power while a large ram batch is stepped (run on the GPU box).
usage: python tools/clock_probe.py C2:1048576 C2:524288 ..."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml, torch, bench
torch.cuda.set_device(0)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for spec in sys.argv[1:]:
    name, n = spec.split(":")
    bench.WORKLOADS["X"] = dict(n=int(n), kw=bench.WORKLOADS[name]["kw"], desc="x")
    samples, stop = [], threading.Event()

    def poll():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                            pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
            stop.wait(0.005)
    th = threading.Thread(target=poll, daemon=True)
    th.start()
    r = bench.time_workload("X", 40, 5, 0, 1, None, burn_in=60)
    stop.set(); th.join()
    busy = samples[-max(1, len(samples) // 10):]  # the timed region is the end of the call
    print(f"{spec}: {r['ms_per_step'] * 1e3:.1f} us frac {r['roofline']['frac']:.3f} | last samples: sm {sorted(s[0] for s in busy)[len(busy) // 2]} MHz "
          f"mem {busy[-1][1]} MHz power max {max(s[2] for s in busy):.0f} W reasons {sorted({hex(s[3]) for s in busy})} | all: sm min {min(s[0] for s in samples)} "
          f"power max {max(s[2] for s in samples):.0f} W", flush=True)

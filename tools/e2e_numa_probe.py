"""Does NUMA placement of the pinned host buffers limit the 8-GPU host-buffer path?  Run under torchrun on the GPU box:
python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_numa_probe.py"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml, torch, torch.distributed as dist
import bench

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
    print(subprocess.run(["bash", "-c", "lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"], capture_output=True, text=True).stdout, flush=True)


def d2h(label):
    if world > 1:
        dist.barrier()
    v = bench.pcie_d2h_peak_gbs()
    vs = [None] * world
    if world > 1:
        dist.all_gather_object(vs, v)
    else:
        vs = [v]
    if rank == 0:
        print(f"{label}: concurrent 256 MiB D2H per rank GB/s {vs}  sum {sum(vs):.1f}", flush=True)


def e2e(label):
    r = bench.time_e2e("C2", 200, 5, rank, world, dist if world > 1 else None)
    if rank == 0:
        print(f"{label}: e2e C2 {r['value'] / 1e6:.1f} M steps/s whole job, pipelined {r['pipelined']['d2h_gbs_per_gpu']} GB/s per GPU, "
              f"sync {r['sync']['d2h_gbs_per_gpu']} GB/s", flush=True)


print(f"rank {rank}: affinity before {sorted(os.sched_getaffinity(0))[:4]}.. ({len(os.sched_getaffinity(0))} cpus)", flush=True)
d2h("unbound")
e2e("unbound")
pynvml.nvmlDeviceSetCpuAffinity(h)
print(f"rank {rank}: affinity after nvmlDeviceSetCpuAffinity {sorted(os.sched_getaffinity(0))[:4]}.. ({len(os.sched_getaffinity(0))} cpus)", flush=True)
d2h("bound")
e2e("bound")
if world > 1:
    dist.destroy_process_group()

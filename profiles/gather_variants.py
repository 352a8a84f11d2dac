"""How should the per-step all-gather be issued?  Times, at N ranks, K back-to-back cfg2 steps with the gather
  a) absent,  b) captured in the step's CUDA graph (one graph exec),  c) eager after every replay,
  d) captured, two graph execs used alternately,  e) captured in its OWN small graph replayed after the step graph.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/gather_variants.py
"""
import os
import sys
import threading

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from wild_deep_mvs_b200 import ops, shard  # noqa: E402

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dist.barrier()
net, feats, cams, projs, depth = bench.make_workload(seed=rank)
net = net.to(dev)
dfeats = [ops.to_nhwc(f.to(dev)) for f in feats]
dprojs = list(torch.unbind(projs.to(dev), 1))
ddepth = depth.to(dev)
h, w = bench.CFG["h"], bench.CFG["w"]
K = 100


def timed(fn, name):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / K], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-70s %.4f ms/step" % (name, t.item()), flush=True)


g_plain = net.graphed(dfeats, dprojs, ddepth)
timed(lambda: g_plain(), "a) no gather")
ga = shard.DepthGather(1, (h, w), dev)
g_in = net.graphed(dfeats, dprojs, ddepth, gather=ga)
timed(lambda: g_in(), "b) gather captured in the step graph, one exec back to back")
gb = shard.DepthGather(1, (h, w), dev)
g_out = net.graphed(dfeats, dprojs, ddepth)


def eager():
    g_out()
    gb.local.copy_(g_out.depth)
    gb.all_gather()


timed(eager, "c) eager all_gather_into_tensor after every replay")
gc2 = shard.DepthGather(1, (h, w), dev)
g_in2 = net.graphed(dfeats, dprojs, ddepth, gather=gc2)
flip = [0]


def alt():
    (g_in if flip[0] else g_in2)()
    flip[0] ^= 1


timed(alt, "d) gather captured, two graph execs alternating")
gd = shard.DepthGather(1, (h, w), dev)
g_step = net.graphed(dfeats, dprojs, ddepth)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    gd.all_gather()
torch.cuda.current_stream().wait_stream(side)
g_gather = torch.cuda.CUDAGraph()
with torch.cuda.graph(g_gather):
    gd.local.copy_(g_step.depth)
    gd.all_gather()


def two():
    g_step()
    g_gather.replay()


timed(two, "e) step graph + a separate small gather graph")
# ---- does clock sampling during the timed region disturb the in-graph gather? (rank 0 samples) ----
for period in (0.01, 0.05):
    cs = bench.ClockSampler(local, enabled=(rank == 0), period_s=period)
    with cs:
        timed(lambda: g_in(), "b) + NVML sampler thread on rank 0, %.0f Hz" % (1 / period))
import subprocess
proc = None
if rank == 0:
    proc = subprocess.Popen(["nvidia-smi", "--query-gpu=index,clocks.sm,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "50", "-i", str(local)],
                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
timed(lambda: g_in(), "b) + nvidia-smi -lms 50 in a separate process")
if proc is not None:
    proc.terminate()
    out = proc.communicate()[0].strip().splitlines()
    print("   nvidia-smi samples: %d, first: %s" % (len(out), out[0] if out else None), flush=True)
g_plain = g_in = g_in2 = g_out = g_step = g_gather = None
import gc
gc.collect()
torch.cuda.synchronize()
dist.barrier()
th = threading.Thread(target=lambda: (shard.DepthGather.destroy_communicators(), dist.destroy_process_group()), daemon=True)
th.start()
th.join(20.0)
os._exit(0)

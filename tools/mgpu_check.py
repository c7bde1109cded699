#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/mgpu_check.py : a sharded render + the library's own exchange must equal the single-GPU frame bit
for bit -- peer-memory exchange (stores fused into the last resolve kernel) and NCCL all-gather + un-permute, accumulators and display
words, two consecutive frames each, plus the NCCL gather on a side stream ("gather_async").  Prints BITWISE_OK on rank 0 when everything
matches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vpt_b200 as V
from vpt_b200.scene import synthetic_env
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
W, H, P = 1920, 1080, 4
vol = V.Volume.load_vdb(V.find_asset("dragon.vdb"))
scene = V.Scene([vol.instance()], device=f"cuda:{local}", env=synthetic_env(512, 256))
def kp():
    k = V.default_kernel_params(); k.environment_type = 1; k.ray_depth = 100; k.max_interactions = 1000; return k
ok = True
one = None
def single_gpu_frames():
    """rank 0: the same two consecutive frames on one GPU"""
    r1 = V.Renderer(scene, W, H, kp=kp(), cam=cam)
    scene.reset_blue_noise(); r1.render(P); torch.cuda.synchronize()
    a1, d1 = r1.buffers.accum.clone(), r1.buffers.display.clone()
    r1.kp.iteration = 0; r1.render(P); torch.cuda.synchronize()          # second frame: blue-noise state advanced by P, as on the ranks
    a2 = r1.buffers.accum.clone(); r1.close()
    return a1, d1, a2
cam = None
for exchange in ("p2p", "nccl"):
    # 1. exchange inside the render call (accumulators + display words), two frames back to back
    scene.reset_blue_noise()
    dr = V.DistributedRenderer(scene, W, H, kp=kp(), cam=cam, stripe_rows=8, gather_display=True, exchange=exchange)
    cam = dr.r.cam
    dr.render(P); torch.cuda.synchronize()
    full = dr.full_accum().reshape(-1, 3).clone(); disp = dr.full_display.clone()
    dr.r.kp.iteration = 0; dr.render(P); torch.cuda.synchronize()
    full2 = dr.full_accum().reshape(-1, 3).clone()
    lost = dr.abandoned_waits()
    if rank == 0:
        if one is None: one = single_gpu_frames()
        a = torch.equal(full, one[0]); d = torch.equal(disp, one[1]); b = torch.equal(full2, one[2])
        ok = ok and a and b and d and lost == 0
        print(f"[mgpu_check] world={world} exchange={exchange}: gathered == single-GPU frame: accum {a}, display {d}, second frame {b}; abandoned flag waits {lost}; mean {float(full.mean()):.6f}", flush=True)
    dist.barrier(); dr.close()
# 2. NCCL on a side stream: two frames back to back, the second overwrites the accumulator the first gather reads
scene.reset_blue_noise()
da = V.DistributedRenderer(scene, W, H, kp=kp(), cam=cam, stripe_rows=8, options={"gather_async": 1}, exchange="nccl")
da.render(P); da.r.kp.iteration = 0; da.render(P)
V.lib.vpt_comm_wait(da.r.ctx, None); torch.cuda.synchronize()
full_async = da.full_accum().reshape(-1, 3).clone()
if rank == 0:
    b = torch.equal(full_async, one[2]); ok = ok and b
    print(f"[mgpu_check] world={world} exchange=nccl, side stream: second frame {b}", flush=True)
    if ok: print("BITWISE_OK", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)

#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/mgpu_check.py : sharded render + the library's own NCCL all-gather + un-permute must equal
the single-GPU frame bit for bit -- in-stream gather, side-stream ("gather_async") gather over two consecutive frames, and the
display-word gather.  Prints BITWISE_OK on rank 0 when everything matches."""
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
# 1. in-stream gather (the default), accumulators + display words
dr = V.DistributedRenderer(scene, W, H, kp=kp(), stripe_rows=8, gather_display=True)
dr.render(P); torch.cuda.synchronize()
full = dr.full_accum().reshape(-1, 3).clone(); disp = dr.full_display.clone()
# 2. side-stream gather: two frames back to back, the second overwrites the accumulator the first gather reads
da = V.DistributedRenderer(scene, W, H, kp=kp(), cam=dr.r.cam, stripe_rows=8, options={"gather_async": 1})
scene.reset_blue_noise()
da.render(P); da.r.kp.iteration = 0; da.render(P)
V.lib.vpt_comm_wait(da.r.ctx, None); torch.cuda.synchronize()
full_async = da.full_accum().reshape(-1, 3).clone()
if rank == 0:
    one = V.Renderer(scene, W, H, kp=kp(), cam=dr.r.cam)
    scene.reset_blue_noise(); one.render(P); torch.cuda.synchronize()
    a = torch.equal(full, one.buffers.accum); d = torch.equal(disp, one.buffers.display)
    one.kp.iteration = 0; one.render(P); torch.cuda.synchronize()          # second frame: blue-noise state advanced by P, as on the ranks
    b = torch.equal(full_async, one.buffers.accum)
    ok = a and b and d
    print(f"[mgpu_check] world={world} gathered == single-GPU frame: accum {a}, display {d}, async second frame {b}; mean {float(full.mean()):.6f}", flush=True)
    if ok: print("BITWISE_OK", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)

#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/mgpu_check.py : sharded render + NCCL all-gather + un-permute == single-GPU frame (bitwise)."""
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
dr = V.DistributedRenderer(scene, W, H, kp=kp(), stripe_rows=8)
dr.render(P); full = dr.full_accum(); torch.cuda.synchronize()
ok = True
if rank == 0:
    scene.reset_blue_noise()
    one = V.Renderer(scene, W, H, kp=kp(), cam=dr.r.cam)
    one.render(P); torch.cuda.synchronize()
    ok = torch.equal(full.view(-1, 3), one.buffers.accum)
    print(f"[mgpu_check] world={world} gathered == single-GPU frame: {ok}; mean {float(full.mean()):.6f}", flush=True)
    if ok: print("BITWISE_OK", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)

#!/usr/bin/env python3
"""Per-rank kernel times of the headline frame WITHOUT a multi-GPU box: one GPU renders shard 0 of N (interleaved 8-row stripes, no
exchange) for N = 1, 2, 4, 8 and prints the per-kernel milliseconds next to the ideal 1/N of the full frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vpt_b200 as V
vol = V.Volume.load_vdb(V.find_asset("dragon.vdb"))
scene = V.Scene([vol.instance()], env="Barce_Rooftop_C_3k.hdr")
def kp():
    k = V.default_kernel_params(); k.environment_type = 1; k.ray_depth = 100; k.max_interactions = 1000; return k
base = None
cam = None
for n in (1, 2, 4, 8):
    r = V.Renderer(scene, 1920, 1080, kp=kp(), cam=cam, rank=0, n_ranks=n, stripe_rows=8)
    cam = r.cam
    for _ in range(3): r.kp.iteration = 0; r.render(64)
    torch.cuda.synchronize()
    r.set_option("profile", 1); r.kernel_times()
    for _ in range(5): r.kp.iteration = 0; r.render(64)
    torch.cuda.synchronize()
    kt = {k: v["ms"] / 5 for k, v in r.kernel_times().items()}
    if base is None: base = kt
    print(f"N={n}: " + ", ".join(f"{k} {v:.3f} (x{v / (base[k] / n):.3f} of ideal)" for k, v in kt.items()) + f"; sum {sum(kt.values()):.3f} ms, ideal {sum(base.values()) / n:.3f}", flush=True)
    r.close()

#!/usr/bin/env python3
"""Dev aid: integrator-1 parity breakdown by light class (runs ours vs the reference kernel on the GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import vpt_b200 as V
from vpt_b200.scene import synthetic_env
import oracle_ref
from test_parity_gpu import _sky_power_table, make_kp

dragon = V.Volume.load_vdb(V.find_asset("dragon.vdb"))
orc = oracle_ref.RefOracle()

def run(name, W=256, H=160, passes=2, lights=None, **kw):
    scene = V.Scene([dragon.instance()], env=synthetic_env(512, 256), lights=lights)
    orc.atmosphere_init(scene.atmos)
    tables = V.EnvTables(_sky_power_table())
    kpk = dict(integrator=1, environment_type=1); kpk.update(kw)
    cam = scene.frame_camera(W, H)
    mine = V.Renderer(scene, W, H, kp=make_kp(**kpk), cam=cam); ref = V.Renderer(scene, W, H, kp=make_kp(**kpk), cam=cam)
    tables.apply(mine.kp); tables.apply(ref.kp)
    scene.reset_blue_noise(); orc.render(ref, passes)
    scene.reset_blue_noise(); mine.render(passes); torch.cuda.synchronize()
    a = mine.buffers.accum.cpu().numpy().astype(np.float64); b = ref.buffers.accum.cpu().numpy().astype(np.float64)
    rel = (np.abs(a - b) / (1e-5 + np.abs(b))).max(axis=1)
    bad = rel > 1e-4
    qs = np.quantile(rel, [0.5, 0.9, 0.99, 0.999, 1.0])
    print(f"{name:28s} flipped {bad.mean():.5f}  rel q50 {qs[0]:.2e} q90 {qs[1]:.2e} q99 {qs[2]:.2e} q999 {qs[3]:.2e} max {qs[4]:.2e}", flush=True)
    idx = np.argsort(-rel)[:4]
    for i in idx: print(f"      px {i}: ours {a[i]} ref {b[i]} rel {rel[i]:.3e}")

L2 = [((9.0, 6.0, 2.0), (1.0, 0.8, 0.6), 40.0), ((-2.0, 3.0, 8.0), (0.5, 0.7, 1.0), 25.0)]
run("sun only d8", ray_depth=8, sky_mult=0.0)
run("sun only d100 dense", ray_depth=100, sky_mult=0.0, density_mult=3.0, phase_g1=0.6)
run("none (no lights) d100", ray_depth=100, sky_mult=0.0, sun_mult=0.0, density_mult=3.0, phase_g1=0.6)
run("sky(hdri) only d8", ray_depth=8, sun_mult=0.0)
run("sky(hdri) only d100 dense", ray_depth=100, sun_mult=0.0, density_mult=3.0, phase_g1=0.6)
run("sky(cdf) only d8", ray_depth=8, sun_mult=0.0, environment_type=0)
run("points only d8", ray_depth=8, sun_mult=0.0, sky_mult=0.0, lights=L2)
run("all d1", ray_depth=1)
run("all d2", ray_depth=2)

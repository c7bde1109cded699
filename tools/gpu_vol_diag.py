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

def run(name, W=256, H=160, passes=2, lights=None, probes=(), **kw):
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
    for flag in probes:
        pr = V.Renderer(scene, W, H, kp=make_kp(**kpk), cam=cam, options={"debug_flags": flag}); tables.apply(pr.kp)
        scene.reset_blue_noise(); pr.render(passes); torch.cuda.synchronize()
        c = pr.buffers.accum.cpu().numpy().astype(np.float64)
        relc = (np.abs(c - b) / (1e-5 + np.abs(b))).max(axis=1)
        print(f"      probe {flag}: flipped {np.mean(relc > 1e-4):.5f}; of the {int(bad.sum())} baseline-bad pixels {int((relc[bad] <= 1e-4).sum())} now match")


def horizon(name, y_cam, look=(4.0, None, 0.0), fov=30.0, up=(0, 1, 0), **kw):
    """integrator 0, env 0, camera looking horizontally: miss pixels evaluate the sky along known directions."""
    W, H = 256, 256
    scene = V.Scene([dragon.instance()], env=synthetic_env(512, 256))
    orc.atmosphere_init(scene.atmos)
    cam = V.camera()
    tgt = (look[0], y_cam if look[1] is None else look[1], look[2])
    V.lib.vpt_camera_look_at(C.byref(cam), N.fvec((4.0, y_cam, 40.0)), N.fvec(tgt), N.fvec(up), fov, 1.0, 0.0)
    kpk = dict(integrator=0, environment_type=0, ray_depth=1); kpk.update(kw)
    mine = V.Renderer(scene, W, H, kp=make_kp(**kpk), cam=cam); ref = V.Renderer(scene, W, H, kp=make_kp(**kpk), cam=cam)
    scene.reset_blue_noise(); orc.render(ref, 1)
    scene.reset_blue_noise(); mine.render(1); torch.cuda.synchronize()
    a = mine.buffers.accum.cpu().numpy().astype(np.float64).reshape(H, W, 3); b = ref.buffers.accum.cpu().numpy().astype(np.float64).reshape(H, W, 3)
    rel = (np.abs(a - b) / (1e-5 + np.abs(b))).max(axis=2)
    bad = rel > 1e-4
    rows = np.where(bad.any(axis=1))[0]
    print(f"{name}: flipped {bad.mean():.5f}; bad rows {rows.tolist()[:20]} counts {bad.sum(axis=1)[rows].tolist()[:20]} max rel {rel.max():.3e}", flush=True)
    for r in rows[:3]:
        c = int(np.argmax(rel[r])); print(f"     row {r} col {c}: ours {a[r, c]} ref {b[r, c]}")


import ctypes as C
from vpt_b200 import _native as N
def dir_probe(name, **kw):
    """HDRI that encodes direction finely: equal outputs <=> equal final directions (integrator 0, L = 0)."""
    W, H = 256, 160
    hh, ww = 2048, 4096
    v = (np.arange(hh, dtype=np.float32) + 0.5)[:, None] / hh; u = (np.arange(ww, dtype=np.float32) + 0.5)[None, :] / ww
    env = np.zeros((hh, ww, 4), dtype=np.float32)
    env[..., 0] = 1.0 + np.sin(40 * np.pi * u) * 0.5 + 0 * v; env[..., 1] = 1.0 + np.cos(64 * np.pi * v) * 0.5 + 0 * u; env[..., 2] = 1.0 + np.sin(300 * u + 200 * v) * 0.5; env[..., 3] = 1
    scene = V.Scene([dragon.instance()], env=env)
    kpk = dict(integrator=0, environment_type=1, sun_mult=0.0); kpk.update(kw)
    cam = scene.frame_camera(W, H)
    mine = V.Renderer(scene, W, H, kp=make_kp(**kpk), cam=cam); ref = V.Renderer(scene, W, H, kp=make_kp(**kpk), cam=cam)
    scene.reset_blue_noise(); orc.render(ref, 2)
    scene.reset_blue_noise(); mine.render(2); torch.cuda.synchronize()
    a = mine.buffers.accum.cpu().numpy().astype(np.float64); b = ref.buffers.accum.cpu().numpy().astype(np.float64)
    rel = (np.abs(a - b) / (1e-5 + np.abs(b))).max(axis=1)
    print(f"{name}: exact-equal pixels {np.mean(rel == 0):.5f}; rel>1e-6 {np.mean(rel > 1e-6):.5f} rel>1e-5 {np.mean(rel > 1e-5):.5f} rel>1e-4 {np.mean(rel > 1e-4):.5f} max {rel.max():.3e}", flush=True)

dir_probe("dir probe d1", ray_depth=1)
run("i0 env0 d1 sunless", passes=2, integrator=0, environment_type=0, ray_depth=1, sun_mult=0.0)
run("dark d1 1pass", passes=1, ray_depth=1, sky_mult=0.0, sun_mult=0.0)
run("dark d100 dense 1pass", passes=1, ray_depth=100, sky_mult=0.0, sun_mult=0.0, density_mult=3.0, phase_g1=0.6)
run("sun only d100 dense", ray_depth=100, sky_mult=0.0, density_mult=3.0, phase_g1=0.6)
run("sky(cdf) only d8", ray_depth=8, sun_mult=0.0, environment_type=0)
run("i1 hdri all d8", ray_depth=8)

#!/usr/bin/env python3
"""Does the reference's image depend on the MAJORANT?  It should not for a correct delta tracker, but the reference advances the walk by the
CUMULATIVE free-flight distance (`ray_pos += ray_dir * t` with `t` never reset, render_kernel.cu:1653-1656), so the step
law -- and with it the expectation of the image -- is a function of the majorant.  This script renders the dragon with the volume's own
max_density and with a looser (still valid) majorant of 2x, 256 spp each, plus a second run on other random streams as the noise floor.
Uses this repo's renderer, which is per-seed identical to the reference kernel (tests/test_parity_gpu.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vpt_b200 as V
from vpt_b200.scene import synthetic_env
W, H, P = 320, 200, 256
def render(scale, it0=0):
    vol = V.Volume.load_vdb(V.find_asset("dragon.vdb"))
    vol.rec.vdb_info.max_density = vol.rec.vdb_info.max_density * scale
    scene = V.Scene([vol.instance()], env=synthetic_env(512, 256), keep=[vol])
    kp = V.default_kernel_params(); kp.environment_type = 1; kp.ray_depth = 4; kp.max_interactions = 100000
    r = V.Renderer(scene, W, H, kp=kp)
    if it0 == 0:
        r.render(P); torch.cuda.synchronize()
        return r.buffers.accum.cpu().numpy().copy()
    acc = np.zeros((W * H, 3), dtype=np.float64)
    for p in range(P):                                           # other Philox streams: each pass into a zeroed accumulator, averaged by hand
        r.buffers.accum.zero_(); r.kp.iteration = it0 + p
        r.render(1); torch.cuda.synchronize()
        acc += r.buffers.accum.cpu().numpy().astype(np.float64) * (it0 + p + 1)
    return (acc / P).astype(np.float32)
a = render(1.0); b = render(1.0, it0=5000); c = render(2.0); d = render(0.75)
rm = lambda x, y: float(np.sqrt(np.mean((x.astype(np.float64) - y) ** 2)))
lit = a.sum(axis=1) > 0
print(f"dragon {W}x{H}, {P} spp, ray_depth 4")
print(f"   same majorant, other streams : RMSE {rm(a, b):.5f}   mean {a.mean():.5f} vs {b.mean():.5f}")
print(f"   majorant x2                  : RMSE {rm(a, c):.5f}   mean {c.mean():.5f}  ({100 * (c.mean() / a.mean() - 1):+.2f} %)")
print(f"   majorant x0.75 (not a bound) : RMSE {rm(a, d):.5f}   mean {d.mean():.5f}  ({100 * (d.mean() / a.mean() - 1):+.2f} %)")

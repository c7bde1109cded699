#!/usr/bin/env python3
"""What does the texture unit use as linear-filter weight?  A 2x2x2 texture holding 0 at x = 0 and 1 at x = 1 returns the weight itself.
Prints how many of 4096 probe coordinates agree with candidate quantisation rules (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import vpt_b200 as V
from vpt_b200.scene import texture_3d
for N in (2, 96, 1024):
    data = np.zeros((2, 2, N), dtype=np.float32); data[:, :, 1::2] = 1.0          # 0,1,0,1,... along x
    t = texture_3d(data)
    n = 8192
    rng = np.random.RandomState(1)
    cell = rng.randint(0, N - 1, size=n) & ~1                                          # even cell: left texel 0, right texel 1
    frac = rng.rand(n).astype(np.float32)
    xB = cell + frac                                                                   # texel-space coordinate (after the -0.5 shift)
    u = ((xB + 0.5) / N).astype(np.float32)
    pts = np.stack([u, np.full(n, 0.25, np.float32), np.full(n, 0.25, np.float32)], axis=1).astype(np.float32)
    out = np.empty(n, dtype=np.float32)
    V._native.check(V.lib.vpt_debug_texture_sample(t.tex, pts.ctypes.data_as(C.POINTER(C.c_float)), n, out.ctypes.data_as(C.POINTER(C.c_float))), None, "sample")
    x = (u.astype(np.float32) * np.float32(N) - np.float32(0.5)).astype(np.float32)
    f = (x - np.floor(x)).astype(np.float32)
    print(f"N = {N}: returned weights are multiples of 1/256: {np.all(out * 256 == np.round(out * 256))}, of 1/512: {np.all(out * 512 == np.round(out * 512))}")
    for name, w in (("round(f*256)/256", np.floor(f * 256 + 0.5) / 256), ("floor(f*256)/256", np.floor(f * 256) / 256), ("f", f),
                    ("round(f*512)/512", np.floor(f * 512 + 0.5) / 512), ("round((x*256))/256 frac", (np.floor(x.astype(np.float64) * 256 + 0.5) / 256) % 1.0)):
        d = np.abs(out - w.astype(np.float32))
        print(f"   {name:26s}: equal {100 * np.mean(d == 0):6.2f} %, max |d| {d.max():.3g}, mean |d| {d.mean():.3g}")
    worst = np.argsort(-np.abs(out - np.floor(f * 256 + 0.5) / 256))[:5]
    for i in worst: print(f"      u {u[i]!r} x {x[i]!r} f*256 {f[i] * 256:.4f} hw*256 {out[i] * 256:.4f}")
    t.destroy()

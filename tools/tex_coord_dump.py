#!/usr/bin/env python3
"""Dump the texture unit's x / y / z filter weight as a function of the normalised coordinate for texture sizes that are NOT powers of
two (where u * N is not exact in fp32): 0/1 ramp along one axis, the other two coordinates on texel centres of a size-2 axis.  Writes
gpurun_out/tex_coord.npz for tools/tex_coord_fit.py (offline)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import vpt_b200 as V
from vpt_b200.scene import texture_3d

def hw_sample(t, pts):
    out = np.empty(len(pts), dtype=np.float32)
    V._native.check(V.lib.vpt_debug_texture_sample(t.tex, pts.ctypes.data_as(C.POINTER(C.c_float)), len(pts), out.ctypes.data_as(C.POINTER(C.c_float))), None, "sample")
    return out

rng = np.random.RandomState(5)
n = 200000
res = {}
for N in (3, 5, 7, 31, 49, 70, 72, 80, 96, 100, 333, 1000, 1023, 2047):
    for ax in (0, 1, 2):
        if ax != 0 and N not in (49, 96, 1000): continue
        shape = [2, 2, 2]; shape[2 - ax] = N                     # data index order (z, y, x)
        d = np.zeros(shape, dtype=np.float32)
        idx = [None, None, None]; idx[2 - ax] = slice(None)
        d += (np.arange(N) % 2).astype(np.float32)[tuple(idx)]
        t = texture_3d(d)
        u = rng.rand(n).astype(np.float32)
        pts = np.full((n, 3), 0.25, dtype=np.float32); pts[:, ax] = u
        res[f"u_{N}_{ax}"] = u; res[f"hw_{N}_{ax}"] = hw_sample(t, pts)
        t.destroy()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/tex_coord.npz", **res)
print("saved", len(res) // 2, "series")

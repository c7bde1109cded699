#!/usr/bin/env python3
"""Dump the texture unit's eight trilinear corner weights: 2x2x2 one-hot textures (a single 1 at one corner) sampled at the same random
points return exactly the weight the hardware gives that corner.  Writes gpurun_out/tex_weights.npz (points, weights[8][n], and a
random-valued 64^3 texture with its hardware samples) for offline fitting of the blend rule (tools/tex_weight_fit.py)."""
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

rng = np.random.RandomState(11)
n = 60000
pts = (0.25 + 0.5 * rng.rand(n, 3)).astype(np.float32)          # inside the single interior cell of a 2x2x2 texture
# a third of the points on exact 1/256 weight positions (+ ties), a third on 1/512 positions
q = rng.randint(0, 257, size=(n // 3, 3)); pts[:n // 3] = (0.25 + q / 512.0).astype(np.float32)
q = rng.randint(0, 513, size=(n // 3, 3)); pts[n // 3:2 * (n // 3)] = (0.25 + q / 1024.0).astype(np.float32)
w = np.zeros((8, n), dtype=np.float32)
for c in range(8):
    d = np.zeros((2, 2, 2), dtype=np.float32); d[(c >> 2) & 1, (c >> 1) & 1, c & 1] = 1.0     # index = (z, y, x)
    t = texture_3d(d); w[c] = hw_sample(t, pts); t.destroy()
data = rng.rand(64, 64, 64).astype(np.float32)
t = texture_3d(data); pr = rng.rand(n, 3).astype(np.float32); hw = hw_sample(t, pr); t.destroy()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/tex_weights.npz", pts=pts, w=w, data=data, pr=pr, hw=hw)
s = w.sum(axis=0)
print("corner weights are multiples of 1/256:", bool(np.all(w * 256 == np.round(w * 256))), " of 1/65536:", bool(np.all(w * 65536 == np.round(w * 65536))))
print("sum of the eight weights: min", s.min(), "max", s.max(), " == 1 on", 100 * np.mean(s == 1.0), "%")
for i in range(5): print(pts[i], (w[:, i] * 256))

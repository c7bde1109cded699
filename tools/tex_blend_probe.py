#!/usr/bin/env python3
"""How does the texture unit BLEND fp32 texels?  Random-valued 3-D textures are sampled at random points and the result is compared
with numpy emulations (float64) of candidate rules: per-axis weights rounded to 1/256, weight PRODUCTS quantised, texels narrowed.
Run on the GPU box; the output is committed under profiles/."""
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

def axis(u, n, bits=8):
    x = u.astype(np.float64) * n - 0.5
    fl = np.floor(x)
    q = float(1 << bits)
    w = np.floor((x - fl) * q + 0.5) / q
    c = fl.astype(np.int64)
    up = w >= 1.0
    c = np.where(up, c + 1, c); w = np.where(up, 0.0, w)
    lo = c < 0
    w = np.where(lo, 0.0, w); c = np.where(lo, 0, c)
    c0 = np.clip(c, 0, n - 1); c1 = np.clip(c + 1, 0, n - 1)
    return c0, c1, w

def emulate(data, pts, bits=8, narrow=None, wq=None, order="sum"):
    nz, ny, nx = data.shape
    d = data.astype(np.float64)
    if narrow == "tf32":
        b = data.view(np.uint32); d = ((b + 0x1000) & 0xffffe000).astype(np.uint32).view(np.float32).astype(np.float64)
    if narrow == "fp16": d = data.astype(np.float16).astype(np.float64)
    if narrow == "bf16":
        b = data.view(np.uint32); d = ((b + 0x8000) & 0xffff0000).astype(np.uint32).view(np.float32).astype(np.float64)
    i0, i1, a = axis(pts[:, 0], nx, bits); j0, j1, b_ = axis(pts[:, 1], ny, bits); k0, k1, c = axis(pts[:, 2], nz, bits)
    if order == "lerp32":
        f = np.float32
        g = lambda k, j, i: data[k, j, i].astype(f)
        a32, b32, c32 = a.astype(f), b_.astype(f), c.astype(f)
        lx = lambda k, j: g(k, j, i0) + a32 * (g(k, j, i1) - g(k, j, i0))
        ly = lambda k: lx(k, j0) + b32 * (lx(k, j1) - lx(k, j0))
        return (ly(k0) + c32 * (ly(k1) - ly(k0))).astype(np.float64)
    res = np.zeros(len(pts))
    for kk, wc in ((k0, 1 - c), (k1, c)):
        for jj, wb in ((j0, 1 - b_), (j1, b_)):
            for ii, wa in ((i0, 1 - a), (i1, a)):
                w = wa * wb * wc
                if wq: w = np.floor(w * wq + 0.5) / wq
                res += w * d[kk, jj, ii]
    return res

rng = np.random.RandomState(7)
for shape, kind in (((10, 12, 16), "uniform"), ((72, 80, 96), "smooth"), ((64, 64, 64), "smooth")):
    nz, ny, nx = shape
    if kind == "uniform": data = rng.rand(*shape).astype(np.float32)
    else:
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        data = (0.4 * np.sin(x * 0.21) * np.cos(y * 0.17) + 0.3 * np.sin(z * 0.13 + x * 0.05) + 0.05 * rng.rand(*shape)).astype(np.float32)
    t = texture_3d(data)
    n = 20000
    pts = rng.rand(n, 3).astype(np.float32)
    hw = hw_sample(t, pts).astype(np.float64)
    print(f"texture {nx}x{ny}x{nz} ({kind}), {n} points, |value| mean {np.abs(hw).mean():.3g}")
    for name, kw in (("per-axis 8-bit weights, fp64 blend", {}),
                     ("per-axis 8-bit weights, fp32 nested lerps", dict(order="lerp32")),
                     ("per-axis 9-bit weights", dict(bits=9)),
                     ("per-axis 7-bit weights", dict(bits=7)),
                     ("full-precision weights", dict(bits=30)),
                     ("8-bit weights, products rounded to 1/2^8", dict(wq=256.0)),
                     ("8-bit weights, products rounded to 1/2^12", dict(wq=4096.0)),
                     ("8-bit weights, products rounded to 1/2^16", dict(wq=65536.0)),
                     ("8-bit weights, texels narrowed to tf32", dict(narrow="tf32")),
                     ("8-bit weights, texels narrowed to fp16", dict(narrow="fp16")),
                     ("8-bit weights, texels narrowed to bf16", dict(narrow="bf16"))):
        e = emulate(data, pts, **kw)
        d = np.abs(e - hw)
        print(f"   {name:46s}: max |d| {d.max():.3g}, mean |d| {d.mean():.3g}, within 1e-6: {100 * np.mean(d <= 1e-6):6.2f} %, bit-equal as fp32: {100 * np.mean(e.astype(np.float32) == hw.astype(np.float32)):6.2f} %")
    # one axis at a time: data varying along a single axis exposes that axis' weight
    for ax, nm in ((2, "x"), (1, "y"), (0, "z")):
        ramp = np.zeros(shape, dtype=np.float32)
        idx = [None, None, None]; idx[ax] = slice(None)
        ramp += (np.arange(shape[ax]) % 2).astype(np.float32)[tuple(idx)]
        tr = texture_3d(ramp)
        h = hw_sample(tr, pts).astype(np.float64)
        e = emulate(ramp, pts)
        print(f"   axis {nm}: 0/1 ramp, emulation equal to hardware on {100 * np.mean(e == h):6.2f} %, hardware values multiples of 1/256: {bool(np.all(h * 256 == np.round(h * 256)))}")
        tr.destroy()
    t.destroy()

#!/usr/bin/env python3
"""Offline fit: how does the texture unit turn a normalised fp32 coordinate u into (cell, 8-bit weight) when N is not a power of two?
Input: gpurun_out/tex_coord.npz (tools/tex_coord_dump.py).  The 0/1 ramp returns w on even cells and 1 - w on odd cells."""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
full = os.path.join(root, "gpurun_out", "tex_coord.npz")      # scratch; a 5-series sample is committed under profiles/
z = np.load(full if os.path.exists(full) else os.path.join(root, "profiles", "r02g_tex_coord_sample.npz"))
def decode(cell, A, N):
    """ramp value the hardware would return for (cell, A) with clamp addressing"""
    c0 = np.clip(cell, 0, N - 1); c1 = np.clip(cell + 1, 0, N - 1)
    return ((256 - A) * (c0 % 2) + A * (c1 % 2)) / 256.0
def rule_exact(u, N):
    x = u.astype(np.float64) * N - 0.5; fl = np.floor(x); return fl.astype(np.int64), np.floor((x - fl) * 256 + 0.5).astype(np.int64)
def rule_fixed(u, N, ubits, mode):
    """u quantised to `ubits` fractional bits (trunc / nearest) before the multiplication"""
    q = float(1 << ubits); uu = u.astype(np.float64) * q
    uu = np.floor(uu) if mode == "trunc" else np.floor(uu + 0.5)
    x = uu * N / q - 0.5; fl = np.floor(x); return fl.astype(np.int64), np.floor((x - fl) * 256 + 0.5).astype(np.int64)
def rule_prod(u, N, pbits, mode):
    """u * N exact, then quantised to `pbits` fractional bits, then - 0.5"""
    q = float(1 << pbits); p = u.astype(np.float64) * N * q
    p = np.floor(p) if mode == "trunc" else np.floor(p + 0.5)
    x = p / q - 0.5; fl = np.floor(x); return fl.astype(np.int64), np.floor((x - fl) * 256 + 0.5).astype(np.int64)
def rule_f32(u, N, mode):
    p = (u.astype(np.float32) * np.float32(N)) if mode == "rn" else None
    x = p.astype(np.float64) - 0.5; fl = np.floor(x); return fl.astype(np.int64), np.floor((x - fl) * 256 + 0.5).astype(np.int64)
series = sorted({k[3:] for k in z.files if k.startswith("hw_")}, key=lambda s: (int(s.split("_")[0]), int(s.split("_")[1])))
cands = [("exact", lambda u, N: rule_exact(u, N)), ("fp32 product RN", lambda u, N: rule_f32(u, N, "rn"))]
for b in (16, 20, 21, 22, 23, 24): 
    for m in ("trunc", "near"): cands.append((f"u to {b} bits {m}", lambda u, N, b=b, m=m: rule_fixed(u, N, b, m)))
for b in (8, 9, 10, 12, 16):
    for m in ("trunc", "near"): cands.append((f"u*N to {b} frac bits {m}", lambda u, N, b=b, m=m: rule_prod(u, N, b, m)))
print(f"{'rule':28s}" + "".join(f"{s:>10s}" for s in series))
for name, fn in cands:
    row = []
    for s in series:
        N, ax = [int(v) for v in s.split("_")]
        u = z[f"u_{s}"]; hw = z[f"hw_{s}"].astype(np.float64)
        cell, A = fn(u, N)
        up = A >= 256; cell = np.where(up, cell + 1, cell); A = np.where(up, 0, A)
        row.append(100 * np.mean(decode(cell, A, N) == hw))
    print(f"{name:28s}" + "".join(f"{v:10.3f}" for v in row))

#!/usr/bin/env python3
"""Dev tool (GPU box): render the same scene with the reference kernel (oracle/_ref) and with libvpt_b200,
compare every output buffer.  Usage: python tools/gpu_compare.py [W H passes] [key=value ...]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import vpt_b200 as V
from oracle_ref import RefOracle

def stats(name, a, b, rtol=1e-4, atol=1e-5):
    a = a.double().cpu().numpy(); b = b.double().cpu().numpy()
    d = np.abs(a - b); tol = atol + rtol * np.abs(b)
    bad = d > tol
    if a.ndim > 1: badpx = bad.any(axis=-1)
    else: badpx = bad
    print(f"  {name:8s} max|d|={d.max():.3e} mean|d|={d.mean():.3e} bad_px={int(badpx.sum())}/{badpx.size} ({100.0*badpx.mean():.4f}%) exact={100.0*(d==0).mean():.3f}%  ref_mean={b.mean():.5f} mine_mean={a.mean():.5f}")
    return badpx

def main():
    args = [a for a in sys.argv[1:] if "=" not in a]
    kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
    W = int(args[0]) if len(args) > 0 else 256; H = int(args[1]) if len(args) > 1 else 256; P = int(args[2]) if len(args) > 2 else 2
    vol = V.Volume.load_vdb(V.find_asset("dragon.vdb"))
    scene = V.Scene([vol.instance()], env="Barce_Rooftop_C_3k.hdr")
    print("data notes:", scene.data_notes)
    def mk():
        kp = V.default_kernel_params(); kp.environment_type = 1; kp.ray_depth = int(kv.get("ray_depth", 1)); kp.volume_depth = int(kv.get("volume_depth", 1))
        kp.max_interactions = 1000; kp.phase_g1 = float(kv.get("g", 0.0)); kp.density_mult = float(kv.get("density_mult", 1.0)); kp.tr_depth = float(kv.get("tr_depth", 1.0))
        return kp
    opts = {k: int(v) for k, v in kv.items() if k in ("passes_per_chunk", "sched_min_lanes", "ctas_per_sm", "debug_flags")}
    mine = V.Renderer(scene, W, H, kp=mk(), options=opts)
    ref = V.Renderer(scene, W, H, kp=mk(), cam=mine.cam)
    orc = RefOracle()
    # octree cross-check: run the reference kernel on the reference-built octree as well
    use_ref_octree = kv.get("ref_octree", "1") == "1"
    if use_ref_octree:
        root = orc.build_octree(scene.h_volumes, len(scene.instances))
        import ctypes as C
        ref.params.p_oct.value = root
    scene.reset_blue_noise()
    t0 = time.time(); orc.render(ref, P); t_ref = time.time() - t0
    bn_ref = scene.d_blue_noise.clone()
    scene.reset_blue_noise()
    torch.cuda.synchronize(); t0 = time.time()
    if kv.get("fused", "1") == "1": mine.render(P)
    else:
        for _ in range(P): mine.render_pass()
    torch.cuda.synchronize(); t_mine = time.time() - t0
    print(f"{W}x{H} x {P} passes: ref {t_ref*1e3:.1f} ms, mine {t_mine*1e3:.1f} ms (wall, incl. first-launch overheads); stats {mine.stats(True)}")
    bad = stats("accum", mine.buffers.accum, ref.buffers.accum)
    stats("depth", mine.buffers.depth, ref.buffers.depth)
    stats("cost", mine.buffers.cost, ref.buffers.cost)
    stats("raw", mine.buffers.raw, ref.buffers.raw, rtol=1e-3, atol=1e-4)
    dm = mine.buffers.display.cpu().numpy().view(np.uint8).reshape(-1, 4).astype(int); dr = ref.buffers.display.cpu().numpy().view(np.uint8).reshape(-1, 4).astype(int)
    print(f"  display  max channel diff {np.abs(dm-dr).max()}  px differing {(np.abs(dm-dr).max(axis=1)>0).mean()*100:.4f}%")
    stats("bluenoise", scene.d_blue_noise, bn_ref, rtol=0, atol=0)
    idx = np.nonzero(bad)[0][:8]
    for i in idx:
        print(f"   px {i} (x={i%W}, y={i//W}): mine {mine.buffers.accum[i].tolist()} ref {ref.buffers.accum[i].tolist()}")
    os.makedirs("gpurun_out", exist_ok=True)
    np.save("gpurun_out/cmp_mine.npy", mine.buffers.accum.cpu().numpy().reshape(H, W, 3)); np.save("gpurun_out/cmp_ref.npy", ref.buffers.accum.cpu().numpy().reshape(H, W, 3))

if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running THE REFERENCE'S OWN KERNEL (oracle/_ref, compiled from
/root/reference/source/render_kernel.cu) on a B200:  `gpurun -- python tests/golden/make_golden.py`
writes gpurun_out/golden/*.npz, which are then committed under tests/golden/.

Protocol: the UNMODIFIED reference kernel, launched exactly as main.cpp:1823-1829 does.  The frames are 256 pixels wide and at
most 256 high, the one shape at which the kernel's blue-noise read / update is race-free (every thread touches only its own
entry); octree built by the reference builder; environment = the deterministic procedural map
`synthetic_env(512, 256)` so that the fixtures need no 13 MB HDRI.
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import vpt_b200 as V
from vpt_b200.scene import synthetic_env
from oracle_ref import RefOracle

CASES = {
    "dragon_single": dict(W=256, H=128, passes=2, kp=dict(ray_depth=2, volume_depth=1), lights=None),
    "dragon_multi":  dict(W=256, H=128, passes=2, kp=dict(ray_depth=3, volume_depth=3, phase_g1=0.6, density_mult=2.0, tr_depth=1.5),
                          lights=[((9.0, 6.0, 2.0), (1.0, 0.8, 0.6), 40.0), ((-2.0, 3.0, 8.0), (0.5, 0.7, 1.0), 25.0)]),
}

def build(case):
    vol = V.Volume.load_vdb(V.find_asset("dragon.vdb"))
    scene = V.Scene([vol.instance()], env=synthetic_env(512, 256), lights=case["lights"])
    kp = V.default_kernel_params(); kp.environment_type = 1; kp.max_interactions = 1000
    for k, v in case["kp"].items(): setattr(kp, k, v)
    return scene, kp

def main():
    out = os.path.join(ROOT, "gpurun_out", "golden"); os.makedirs(out, exist_ok=True)
    orc = RefOracle()
    for name, case in CASES.items():
        scene, kp = build(case)
        r = V.Renderer(scene, case["W"], case["H"], kp=kp)
        r.params.p_oct.value = orc.build_octree(scene.h_volumes, len(scene.instances))      # reference-built octree
        scene.reset_blue_noise()
        assert case["W"] == 256 and case["H"] <= 256, "the unmodified kernel is race-free only at this shape"
        orc.render(r, case["passes"], race_free=False)
        np.savez_compressed(os.path.join(out, name + ".npz"),
                            accum=r.buffers.accum.cpu().numpy().reshape(case["H"], case["W"], 3),
                            depth=r.buffers.depth.cpu().numpy().reshape(case["H"], case["W"]),
                            raw=r.buffers.raw.cpu().numpy().reshape(case["H"], case["W"], 4),
                            display=r.buffers.display.cpu().numpy().reshape(case["H"], case["W"]),
                            blue_noise=scene.d_blue_noise.cpu().numpy(),
                            camera=np.frombuffer(bytes(r.cam), dtype=np.uint8),
                            meta=json.dumps(dict(case=name, W=case["W"], H=case["H"], passes=case["passes"], kp=case["kp"], lights=case["lights"],
                                                 generator="UNMODIFIED reference volume_rt_kernel (sm_100a rebuild), its own launch protocol, race-free frame shape")))
        print("wrote", name, "mean", float(r.buffers.accum.mean()))

if __name__ == "__main__":
    main()

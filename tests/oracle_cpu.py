"""Test-side wrapper of the CPU oracle (oracle/vpt_oracle.c -> oracle/liboracle_cpu.so): the plain-C
restatement of the reference render pass.  Only tests/, smoke() and bench.py's cpu_baseline leg use it."""
import ctypes as C
import os
import time

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(_REPO, "oracle", "liboracle_cpu.so")


class OrcVolume(C.Structure):
    _fields_ = [("dim", C.c_int * 3), ("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("xform", C.c_float * 16),
                ("max_density", C.c_float), ("min_density", C.c_float), ("voxelsize", C.c_float),
                ("density", C.POINTER(C.c_float)), ("emission", C.POINTER(C.c_float)), ("edim", C.c_int * 3),
                ("color4", C.POINTER(C.c_float)), ("cdim", C.c_int * 3), ("inv", C.c_float * 12)]


class OrcScene(C.Structure):
    _fields_ = [("n_volumes", C.c_int), ("volumes", C.POINTER(OrcVolume)),
                ("env_rgba", C.POINTER(C.c_float)), ("env_w", C.c_int), ("env_h", C.c_int),
                ("emission_lut", C.POINTER(C.c_float)), ("density_color_lut", C.POINTER(C.c_float)),
                ("sph_center", C.c_float * 3), ("sph_radius", C.c_float), ("sph_color", C.c_float * 3), ("sph_roughness", C.c_float),
                ("n_lights", C.c_int), ("lights", C.POINTER(C.c_float)),
                ("node_min", (C.c_float * 3) * 585), ("node_max", (C.c_float * 3) * 585), ("node_nvol", C.c_int * 585), ("node_exists", C.c_int * 585),
                ("leaf_lists", C.POINTER(C.c_int)), ("root_max_ext", C.c_float), ("root_min_ext", C.c_float)]


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class CpuOracle:
    """Scene description in host memory + orc_render_pass().

    volumes: list of dicts {density (z,y,x) f32, bbox_min, bbox_max, xform (4,4) memory image, max_density,
    min_density, voxelsize, emission (optional), color (optional (z,y,x,3|4))}."""

    def __init__(self, volumes, env_rgba, blue_noise, emission_lut=None, density_color_lut=None, lights=None,
                 sphere=((0.0, 1000.0, 0.0), 1.0, (10.0, 0.0, 0.0), 1.0)):
        if not os.path.exists(LIB):
            raise RuntimeError("oracle/liboracle_cpu.so missing: run `make -C oracle cpu`")
        from vpt_b200 import _native as N
        self.N = N
        self.lib = C.CDLL(LIB)
        self.lib.orc_prepare.argtypes = [C.POINTER(OrcScene)]; self.lib.orc_prepare.restype = C.c_int
        self.lib.orc_release.argtypes = [C.POINTER(OrcScene)]
        self.lib.orc_render_pass.argtypes = [C.POINTER(OrcScene), C.POINTER(N.camera), C.POINTER(N.Kernel_params), C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        self.lib.orc_render_pass.restype = C.c_int
        self.lib.orc_bn_advance.argtypes = [C.POINTER(C.c_float), C.c_int]
        self.lib.orc_sizeof_scene.restype = C.c_size_t; self.lib.orc_sizeof_volume.restype = C.c_size_t
        assert self.lib.orc_sizeof_scene() == C.sizeof(OrcScene) and self.lib.orc_sizeof_volume() == C.sizeof(OrcVolume), "oracle struct layout drifted"
        self._keep = []
        vs = (OrcVolume * len(volumes))()
        for i, v in enumerate(volumes):
            d = np.ascontiguousarray(v["density"], dtype=np.float32); self._keep.append(d)
            dz, dy, dx = d.shape
            vs[i].dim[:] = [dx, dy, dz]
            vs[i].bmin[:] = [float(x) for x in v["bbox_min"]]; vs[i].bmax[:] = [float(x) for x in v["bbox_max"]]
            vs[i].xform[:] = [float(x) for x in np.asarray(v["xform"], dtype=np.float32).reshape(-1)]
            vs[i].max_density = float(v["max_density"]); vs[i].min_density = float(v["min_density"]); vs[i].voxelsize = float(v["voxelsize"])
            vs[i].density = _fp(d)
            if v.get("emission") is not None:
                e = np.ascontiguousarray(v["emission"], dtype=np.float32); self._keep.append(e)
                vs[i].emission = _fp(e); vs[i].edim[:] = [e.shape[2], e.shape[1], e.shape[0]]
            if v.get("color") is not None:
                c = v["color"]; c4 = np.zeros(c.shape[:3] + (4,), dtype=np.float32); c4[..., :3] = c[..., :3]; c4[..., 3] = 1.0
                self._keep.append(c4); vs[i].color4 = _fp(c4); vs[i].cdim[:] = [c4.shape[2], c4.shape[1], c4.shape[0]]
        self.vs = vs
        s = OrcScene(); s.n_volumes = len(volumes); s.volumes = vs
        self.env = np.ascontiguousarray(env_rgba, dtype=np.float32); s.env_rgba = _fp(self.env); s.env_h, s.env_w = self.env.shape[:2]
        self.elut = np.ascontiguousarray(emission_lut if emission_lut is not None else np.zeros((256, 3)), dtype=np.float32)
        self.dlut = np.ascontiguousarray(density_color_lut if density_color_lut is not None else np.ones((256, 3)), dtype=np.float32)
        s.emission_lut = _fp(self.elut); s.density_color_lut = _fp(self.dlut)
        s.sph_center[:] = list(sphere[0]); s.sph_radius = sphere[1]; s.sph_color[:] = list(sphere[2]); s.sph_roughness = sphere[3]
        if lights:
            self.lights = np.ascontiguousarray([list(p) + list(c) + [pw] for p, c, pw in lights], dtype=np.float32)
            s.n_lights = len(lights); s.lights = _fp(self.lights)
        self.scene = s
        rc = self.lib.orc_prepare(C.byref(s))
        if rc: raise RuntimeError(f"orc_prepare -> {rc}")
        self.bn0 = np.ascontiguousarray(blue_noise, dtype=np.float32).reshape(-1).copy()
        self.bn = self.bn0.copy()

    @staticmethod
    def from_scene_assets(vdb_path, env_rgba, blue_noise, **kw):
        import vpt_b200 as V
        from vpt_b200.scene import load_vdb_grid
        dens, meta = load_vdb_grid(vdb_path, "density")
        vol = dict(density=dens, bbox_min=meta["bbox_min"], bbox_max=meta["bbox_max"], xform=meta["xform"], max_density=meta["max_value"],
                   min_density=meta["min_density"], voxelsize=meta["voxel_size"])
        return CpuOracle([vol], env_rgba, blue_noise, **kw)

    def render(self, cam, kp, n_passes, rect=None, want_aux=False):
        """n progressive passes from kp.iteration (kp is not modified).  Returns accum (H, W, 3) [, depth, raw, display]."""
        N = self.N
        W, H = int(kp.resolution.x), int(kp.resolution.y)
        x0, y0, x1, y1 = rect if rect else (0, 0, W, H)
        accum = np.zeros((H, W, 3), dtype=np.float32); depth = np.zeros((H, W), dtype=np.float32)
        raw = np.zeros((H, W, 4), dtype=np.float32); disp = np.zeros((H, W), dtype=np.uint32)
        k = N.Kernel_params(); C.memmove(C.byref(k), C.byref(kp), C.sizeof(k))
        for _ in range(n_passes):
            rc = self.lib.orc_render_pass(C.byref(self.scene), C.byref(cam), C.byref(k), x0, y0, x1, y1, _fp(accum), _fp(depth), _fp(raw),
                                          disp.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(self.bn))
            if rc: raise RuntimeError(f"orc_render_pass -> {rc}")
            self.lib.orc_bn_advance(_fp(self.bn), min(W * H, 65536))
            k.iteration += 1
        return (accum, depth, raw, disp) if want_aux else accum

    def reset_blue_noise(self):
        self.bn[:] = self.bn0


def timed_sample(scene, cam, kp, width, height, seconds_target=12.0):
    """bench.py cpu_baseline leg: time the CPU port on a bounded tile of the same 1920x1080 workload."""
    import vpt_b200 as V
    from vpt_b200.scene import load_vdb_grid, load_hdr, find_asset
    dens, meta = load_vdb_grid(find_asset("dragon.vdb"), "density")
    p = find_asset("Barce_Rooftop_C_3k.hdr")
    env = load_hdr(p) if p else V.scene.synthetic_env()
    vol = dict(density=dens, bbox_min=meta["bbox_min"], bbox_max=meta["bbox_max"], xform=meta["xform"], max_density=meta["max_value"],
               min_density=meta["min_density"], voxelsize=meta["voxel_size"])
    orc = CpuOracle([vol], env, scene.bn_host.reshape(256, 256, 3))
    k = V.Kernel_params(); C.memmove(C.byref(k), C.byref(kp), C.sizeof(k)); k.iteration = 0
    k.resolution = V.u2(width, height)
    cores = os.cpu_count() or 1
    try:                                    # torchrun exports OMP_NUM_THREADS=1: ask the OpenMP runtime for all host cores and report what it grants
        gomp = C.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(cores)); cores = int(gomp.omp_get_max_threads())
    except OSError:
        cores = int(os.environ.get("OMP_NUM_THREADS", cores))
    # the whole frame, as many passes as fit the time target (64 at most: the workload's own count)
    tw, th, spp = 1920, 1080, 1
    rect = ((width - tw) // 2, (height - th) // 2, (width + tw) // 2, (height + th) // 2)
    t0 = time.perf_counter(); orc.render(cam, k, spp, rect=rect); dt = time.perf_counter() - t0
    reps = max(1, min(64, int(seconds_target / max(dt, 1e-3))))
    t0 = time.perf_counter(); orc.render(cam, k, reps, rect=rect); dt = time.perf_counter() - t0
    samples = tw * th * reps
    return {"value": samples / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": f"oracle/vpt_oracle.c (OpenMP, {cores} threads): {tw}x{th} window of the 1920x1080 frame x {reps} spp, {dt:.1f} s"}

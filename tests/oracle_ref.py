"""Test-side wrapper of the ORACLE: the reference's own `volume_rt_kernel`, octree builder and
blue-noise update, compiled from /root/reference by oracle/Makefile into oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py's reference arm may import this module.
"""
import ctypes as C
import os

import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(_REPO, "oracle", "_ref")
LIB = os.path.join(REF_DIR, "libvpt_ref.so")


_ATMO_CACHE = {}


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in
               ("libvpt_ref.so", "render_kernel_ref.cubin", "render_kernel_ref_nobn.cubin", "bn_advance_ref.cubin"))


class RefOracle:
    """Drives the reference kernel with a byte-identical parameter block (LaunchParams.array)."""
    UNMODIFIED, NOBN, CANDIDATE = 0, 1, 2

    def __init__(self):
        if not available():
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle ref` where /root/reference exists)")
        self.lib = C.CDLL(LIB)
        L = self.lib
        L.vptref_load_kernel.argtypes = [C.c_char_p, C.c_int]; L.vptref_load_kernel.restype = C.c_int
        L.vptref_load_bn_kernel.argtypes = [C.c_char_p]; L.vptref_load_bn_kernel.restype = C.c_int
        L.vptref_launch.argtypes = [C.POINTER(C.c_void_p), C.c_uint, C.c_uint, C.c_int, C.c_int]; L.vptref_launch.restype = C.c_int
        L.vptref_bn_advance.argtypes = [C.c_void_p, C.c_int, C.c_uint]; L.vptref_bn_advance.restype = C.c_int
        L.vptref_build_octree.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]; L.vptref_build_octree.restype = C.c_int
        L.vptref_build_bvh.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_float)]
        L.vptref_build_bvh.restype = C.c_int
        L.vptref_fill_volume.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]; L.vptref_fill_volume.restype = C.c_int
        L.vptref_sizes.argtypes = [C.POINTER(C.c_size_t), C.c_int]; L.vptref_sizes.restype = C.c_int
        L.vptref_update_camera.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float]
        L.vptref_bounds.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.vptref_instance_xform.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float)]
        L.vptref_atmosphere_init.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.vptref_atmosphere_init.restype = C.c_int
        self._loaded = False

    def atmosphere_init(self, atmos, use_constant_solar_spectrum=True, use_ozone=True, luminance=0, white_balance=True, exposure=1.0):
        """Run the reference's own Bruneton precompute (atmosphere::init, defaults of main.cpp:1433-1436) and
        write the resulting AtmosphereParameters (scalars + the four look-up textures) into `atmos` in place."""
        if not os.path.exists(os.path.join(REF_DIR, "atmo", "atmosphere_kernels.ptx")):
            raise RuntimeError("oracle/_ref/atmo is not built")
        assert C.sizeof(atmos) == 464
        key = (bool(use_constant_solar_spectrum), bool(use_ozone), int(luminance), bool(white_balance), float(exposure))
        if key not in _ATMO_CACHE:                       # one precompute per model per process (the textures stay alive)
            buf = (C.c_ubyte * 464)()
            rc = self.lib.vptref_atmosphere_init(os.path.join(REF_DIR, "atmo").encode(), int(key[0]), int(key[1]),
                                                 key[2], int(key[3]), key[4], C.cast(buf, C.c_void_p))
            if rc: raise RuntimeError(f"vptref_atmosphere_init -> {rc}")
            _ATMO_CACHE[key] = bytes(buf)
        C.memmove(C.byref(atmos), _ATMO_CACHE[key], 464)

    def fill_volume(self, d_buffer_ptr, dims, scale=0.1, noise_type=0):
        """The reference's fill_volume_buffer into a device buffer of dims[0]*dims[1]*dims[2] floats."""
        p = os.path.join(REF_DIR, "texture_kernels_ref.cubin")
        rc = self.lib.vptref_fill_volume(p.encode(), C.c_void_p(d_buffer_ptr), int(dims[0]), int(dims[1]), int(dims[2]), float(scale), int(noise_type))
        if rc: raise RuntimeError(f"vptref_fill_volume -> {rc}")

    def load_kernels(self):
        if self._loaded:
            return
        for which, name in ((0, "render_kernel_ref.cubin"), (1, "render_kernel_ref_nobn.cubin")):
            rc = self.lib.vptref_load_kernel(os.path.join(REF_DIR, name).encode(), which)
            if rc: raise RuntimeError(f"vptref_load_kernel({name}) -> {rc}")
        rc = self.lib.vptref_load_bn_kernel(os.path.join(REF_DIR, "bn_advance_ref.cubin").encode())
        if rc: raise RuntimeError(f"vptref_load_bn_kernel -> {rc}")
        self._loaded = True

    def load_candidate(self, cubin_path):
        """Load another module through the reference loader sequence (cuModuleLoad + cuModuleGetFunction("volume_rt_kernel"))."""
        rc = self.lib.vptref_load_kernel(os.fspath(cubin_path).encode(), self.CANDIDATE)
        if rc: raise RuntimeError(f"vptref_load_kernel({cubin_path}) -> {rc}")

    def sizes(self):
        out = (C.c_size_t * 12)(); n = self.lib.vptref_sizes(out, 12)
        return list(out)[:n]

    def build_octree(self, h_volumes, n):
        """Reference octree (device heap).  Returns the device root pointer."""
        root = C.c_void_p(0)
        rc = self.lib.vptref_build_octree(C.cast(h_volumes, C.c_void_p), n, C.byref(root))
        if rc: raise RuntimeError(f"vptref_build_octree -> {rc}")
        return root.value

    def launch(self, params_array, width, height, which=NOBN, sync=True):
        """One progressive pass exactly as main.cpp:1823-1829 issues it (caller bumps kp.iteration)."""
        self.load_kernels()
        rc = self.lib.vptref_launch(params_array, width, height, which, 1 if sync else 0)
        if rc: raise RuntimeError(f"vptref_launch -> {rc}")

    def bn_advance(self, kp, sync=True):
        """The reference's own update statements for the entries a pass really advances: min(W*H, 65536)."""
        self.load_kernels()
        rc = self.lib.vptref_bn_advance(C.cast(C.byref(kp), C.c_void_p), 1 if sync else 0, min(int(kp.resolution.x) * int(kp.resolution.y), 65536))
        if rc: raise RuntimeError(f"vptref_bn_advance -> {rc}")

    def render(self, renderer, n_passes, race_free=True):
        """Run n reference passes on `renderer`'s parameter block and buffers (race-free protocol: the
        'nobn' build + the reference's own blue-noise statements as a separate launch)."""
        for _ in range(n_passes):
            if race_free:
                self.launch(renderer.params.array, renderer.width, renderer.height, self.NOBN)
                self.bn_advance(renderer.kp)
            else:
                self.launch(renderer.params.array, renderer.width, renderer.height, self.UNMODIFIED)
            renderer.kp.iteration += 1
        torch.cuda.synchronize()

"""Render loop on top of the C ABI: the headless counterpart of the reference frame loop
(source/main.cpp:1527-1860) -- fill Kernel_params, issue passes, keep `iteration`.

`Renderer.render_pass()` is one reference launch (`vpt_render_pass`); `Renderer.render(n)` is n of them
fused (`vpt_render_passes`).  `DistributedRenderer` shards the frame by interleaved row stripes over the
ranks of a torch.distributed job, replicates the scene, and all-gathers the rank-local accumulators
(one NCCL all-gather per call) before un-permuting them into a full frame.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from ._native import lib, check
from .scene import Scene, default_kernel_params


class FrameBuffers:
    """accum/cost (float3), depth (float), raw (float4), display (uint32) for `n_pixels` pixels."""

    def __init__(self, n_pixels, device):
        self.n = int(n_pixels)
        self.accum = torch.zeros(self.n, 3, dtype=torch.float32, device=device)
        self.cost = torch.zeros(self.n, 3, dtype=torch.float32, device=device)
        self.depth = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.raw = torch.zeros(self.n, 4, dtype=torch.float32, device=device)
        self.display = torch.zeros(self.n, dtype=torch.int32, device=device)

    def zero_(self):
        for t in (self.accum, self.cost, self.depth, self.raw, self.display):
            t.zero_()

    def bind(self, kp: N.Kernel_params):
        kp.accum_buffer = self.accum.data_ptr(); kp.cost_buffer = self.cost.data_ptr(); kp.depth_buffer = self.depth.data_ptr()
        kp.raw_buffer = self.raw.data_ptr(); kp.display_buffer = self.display.data_ptr()


class LaunchParams:
    """The `void* params[9]` array of main.cpp:1826, kept alive together with the values it points at."""

    def __init__(self, scene: Scene, cam: N.camera, kp: N.Kernel_params):
        self.cam, self.kp, self.scene = cam, kp, scene
        self.lights = scene.lights
        self.p_vol = C.c_uint64(scene.d_volumes.data_ptr())
        self.p_sphere = C.c_uint64(scene.d_sphere.data_ptr())
        self.p_geo = C.c_uint64(scene.d_geo_list.data_ptr())
        self.p_bvh = C.c_uint64(scene.bvh[0] if getattr(scene, "bvh", None) else scene.d_bvh.data_ptr())
        self.p_oct = C.c_uint64(scene.d_oct_root)
        self.atmos = scene.atmos
        self.array = (C.c_void_p * 9)(
            C.cast(C.byref(self.cam), C.c_void_p), C.cast(C.byref(self.lights), C.c_void_p),
            C.cast(C.byref(self.p_vol), C.c_void_p), C.cast(C.byref(self.p_sphere), C.c_void_p),
            C.cast(C.byref(self.p_geo), C.c_void_p), C.cast(C.byref(self.p_bvh), C.c_void_p),
            C.cast(C.byref(self.p_oct), C.c_void_p), C.cast(C.byref(self.atmos), C.c_void_p),
            C.cast(C.byref(self.kp), C.c_void_p))


class Renderer:
    def __init__(self, scene: Scene, width, height, cam: N.camera = None, kp: N.Kernel_params = None,
                 rank=0, n_ranks=1, stripe_rows=16, options=None):
        self.scene, self.width, self.height = scene, int(width), int(height)
        self.ctx = C.c_void_p(0)
        check(lib.vpt_create(C.byref(self.ctx)), None, "vpt_create")
        check(lib.vpt_set_partition(self.ctx, rank, n_ranks, stripe_rows), self.ctx, "vpt_set_partition")
        for k, v in (options or {}).items():
            check(lib.vpt_set_option(self.ctx, k.encode(), int(v)), self.ctx, f"vpt_set_option({k})")
        self.rank, self.n_ranks = rank, n_ranks
        self.n_local = int(lib.vpt_local_pixels(self.ctx, self.width, self.height))
        self.kp = kp if kp is not None else default_kernel_params()
        self.kp.resolution = N.u2(self.width, self.height)
        self.cam = cam if cam is not None else scene.frame_camera(self.width, self.height)
        self.buffers = FrameBuffers(self.n_local, scene.device)
        self.buffers.bind(self.kp)
        self.kp.blue_noise_buffer = scene.d_blue_noise.data_ptr()
        self.kp.emission_texture = scene.d_emission_lut.data_ptr()
        self.kp.density_color_texture = scene.d_density_color.data_ptr()
        if scene.env_tex is not None:
            self.kp.env_tex = scene.env_tex.tex
        self.params = LaunchParams(scene, self.cam, self.kp)

    # -- the reference frame loop body: launch, ++iteration --
    def render_pass(self, stream=None):
        check(lib.vpt_render_pass(self.ctx, self.params.array, C.c_void_p(stream or 0)), self.ctx, "vpt_render_pass")
        self.kp.iteration += 1

    def render(self, n_passes, stream=None):
        check(lib.vpt_render_passes(self.ctx, self.params.array, int(n_passes), C.c_void_p(stream or 0)), self.ctx, "vpt_render_passes")
        self.kp.iteration += int(n_passes)

    def reset(self):
        self.kp.iteration = 0
        self.buffers.zero_()
        self.scene.reset_blue_noise()

    def stats(self, with_queue=False):
        n = C.c_ulonglong(0); q = C.c_uint(0)
        check(lib.vpt_get_stats(self.ctx, C.byref(n), C.byref(q) if with_queue else None), self.ctx, "vpt_get_stats")
        return int(n.value), int(q.value)

    def set_option(self, key, value):
        check(lib.vpt_set_option(self.ctx, key.encode(), int(value)), self.ctx, f"vpt_set_option({key})")

    def set_brick_volume(self, volume):
        """Fast mode: trace volume 0 from `volume`'s brick pool (None: back to the tex3D parity path)."""
        if volume is None:
            check(lib.vpt_set_brick_volume(self.ctx, 0, 0, 0, 0), self.ctx, "vpt_set_brick_volume")
        else:
            if not getattr(volume, "brick_pool", 0): volume.build_bricks()
            check(lib.vpt_set_brick_volume(self.ctx, volume.brick_pool, *volume.dims), self.ctx, "vpt_set_brick_volume")

    def set_cell_volume(self, volume):
        """Cell mode: trace volume 0 from `volume`'s cell table (None: back to the texture path)."""
        if volume is None:
            check(lib.vpt_set_cell_volume(self.ctx, 0, 0, 0, 0), self.ctx, "vpt_set_cell_volume")
        else:
            if not getattr(volume, "cell_table", 0): volume.build_cells()
            check(lib.vpt_set_cell_volume(self.ctx, volume.cell_table, *volume.dims), self.ctx, "vpt_set_cell_volume")

    def counters(self, reset=True):
        out = (C.c_ulonglong * 8)()
        check(lib.vpt_get_counters(self.ctx, out, 1 if reset else 0), self.ctx, "vpt_get_counters")
        v = list(out)
        return dict(lookups=v[0], lane_steps=v[1], warp_step_iters=v[2], lane_services=v[3], warp_service_rounds=v[4], rays=v[5], brick_fetches=v[6])

    def kernel_times(self):
        ms = (C.c_float * 4)(); n = (C.c_int * 4)()
        check(lib.vpt_get_kernel_times(self.ctx, ms, n), self.ctx, "vpt_get_kernel_times")
        names = ("generate", "trace", "resolve", "bn_advance")
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(names)}

    def accum_image(self):
        """(H, W, 3) float32 linear radiance (single rank only)."""
        assert self.n_ranks == 1
        return self.buffers.accum.view(self.height, self.width, 3)

    def unpermute(self, gathered: torch.Tensor, elem_floats: int):
        full = torch.empty(self.height * self.width, elem_floats, dtype=gathered.dtype, device=gathered.device)
        check(lib.vpt_unpermute(self.ctx, C.c_void_p(gathered.data_ptr()), C.c_void_p(full.data_ptr()), self.width, self.height,
                                4 * elem_floats, C.c_void_p(0)), self.ctx, "vpt_unpermute")
        return full

    def close(self):
        if self.ctx:
            lib.vpt_destroy(self.ctx); self.ctx = C.c_void_p(0)


class _DeviceView:
    """Zero-copy torch view of library-owned device memory (__cuda_array_interface__)."""
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class DistributedRenderer:
    """Frame sharded over the ranks of a torch.distributed job (one process per GPU); scene replicated on each.

    torch.distributed is only the BOOTSTRAP here: it carries the 64-byte IPC handles (exchange="p2p") or the 128-byte NCCL id
    (exchange="nccl") between the ranks.  The exchange itself is the C++ library's:
      * "p2p" (default, up to 8 GPUs of one NVSwitch domain): the last resolve kernel of a render call stores every finished pixel
        straight into every rank's full frame over peer mappings -- gather and stripe un-permutation fused into the producer;
      * "nccl": one ncclAllGather of the rank-local accumulators + an un-permutation kernel behind the last kernel of the call.
    Pixels are independent and every Philox stream is keyed by the GLOBAL pixel index, so the gathered frame is bit-identical to a
    single-GPU render whatever the partition."""

    def __init__(self, scene: Scene, width, height, cam=None, kp=None, stripe_rows=16, options=None, gather_display=False, exchange="p2p"):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > 8 and exchange == "p2p": exchange = "nccl"
        self.exchange = exchange if self.world > 1 else "none"
        self.r = Renderer(scene, width, height, cam=cam, kp=kp, rank=self.rank, n_ranks=self.world,
                          stripe_rows=stripe_rows, options=options)
        n_px = self.r.height * self.r.width
        self.full = None; self.full_display = None
        self.collective_note = "none (1 rank)"
        if self.world > 1 and self.exchange == "p2p":
            # set-up can fail on one rank only (no peer access between two devices, IPC disabled in a container): every rank reports, and
            # all of them fall back to NCCL together
            ok, err = 1, ""
            raw = (C.c_ubyte * 64)()
            try:
                check(lib.vpt_comm_p2p_export(self.r.ctx, self.rank, self.world, stripe_rows, self.r.width, self.r.height, 1 if gather_display else 0, raw),
                      self.r.ctx, "vpt_comm_p2p_export")
            except N.VptError as e:
                ok, err = 0, str(e)
            mine = torch.tensor(list(raw), dtype=torch.uint8, device=scene.device)
            allh = [torch.zeros(64, dtype=torch.uint8, device=scene.device) for _ in range(self.world)]
            dist.all_gather(allh, mine)                                        # bootstrap only
            if ok:
                blob = (C.c_ubyte * (64 * self.world))(*torch.cat(allh).cpu().tolist())
                try:
                    check(lib.vpt_comm_p2p_import(self.r.ctx, blob), self.r.ctx, "vpt_comm_p2p_import")
                except N.VptError as e:
                    ok, err = 0, str(e)
            flag = torch.tensor([ok], dtype=torch.int32, device=scene.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)                        # also: every rank has mapped every block before the first exchange
            if int(flag.item()) == 0:
                if ok: lib.vpt_comm_p2p_enable(self.r.ctx, 0)
                self.exchange = "nccl"; self.p2p_error = err or "another rank could not set up the peer mappings"
        if self.world > 1 and self.exchange == "p2p":
            pa, pd = C.c_uint64(0), C.c_uint64(0)
            check(lib.vpt_comm_p2p_frame(self.r.ctx, C.byref(pa), C.byref(pd)), self.r.ctx, "vpt_comm_p2p_frame")
            self.full = torch.as_tensor(_DeviceView(pa.value, (n_px, 3), "<f4"), device=scene.device)
            if gather_display: self.full_display = torch.as_tensor(_DeviceView(pd.value, (n_px,), "<i4"), device=scene.device)
            self.collective_note = ("peer-memory exchange: the last resolve kernel stores each pixel into every rank's frame over NVLink peer mappings "
                                    "(all-gather + un-permutation fused into the producer), two flag exchanges per call; no NCCL on the data path")
        else:
            self.full = torch.zeros(n_px, 3, dtype=torch.float32, device=scene.device)
            self.full_display = torch.zeros(n_px, dtype=torch.int32, device=scene.device) if gather_display else None
        if self.world > 1 and self.exchange == "nccl":
            idb = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                raw = (C.c_ubyte * 128)()
                check(lib.vpt_comm_get_unique_id(raw), None, "vpt_comm_get_unique_id")
                idb = torch.tensor(list(raw), dtype=torch.uint8)
            idb = idb.to(scene.device); dist.broadcast(idb, src=0)          # bootstrap only
            raw = (C.c_ubyte * 128)(*idb.cpu().tolist())
            check(lib.vpt_comm_init(self.r.ctx, raw, self.rank, self.world, stripe_rows), self.r.ctx, "vpt_comm_init")
            self.set_gather(True)
            v = C.c_int(0); lib.vpt_comm_info(self.r.ctx, C.byref(v), None, None)
            self.collective_note = (f"1 ncclAllGather (NCCL {v.value}) of the float3 accumulators per step, issued by libvpt_b200.so behind the last "
                                    "resolve kernel, + stripe un-permutation kernel")

    def render(self, n_passes, stream=None):
        self.r.render(n_passes, stream=stream)                               # the exchange is part of the call
        if self.world == 1:
            self.full.copy_(self.r.buffers.accum)
            if self.full_display is not None: self.full_display.copy_(self.r.buffers.display)

    def set_gather(self, enabled: bool):
        """Switch the per-call exchange off (rank-local work only: profiling runs that not every rank takes part in) or back on."""
        if self.world == 1: return
        if self.exchange == "p2p":
            check(lib.vpt_comm_p2p_enable(self.r.ctx, 1 if enabled else 0), self.r.ctx, "vpt_comm_p2p_enable")
        elif enabled:
            check(lib.vpt_comm_set_gather(self.r.ctx, C.c_void_p(self.full.data_ptr()),
                                          C.c_void_p(self.full_display.data_ptr()) if self.full_display is not None else None), self.r.ctx, "vpt_comm_set_gather")
        else:
            check(lib.vpt_comm_set_gather(self.r.ctx, None, None), self.r.ctx, "vpt_comm_set_gather")

    def abandoned_waits(self):
        if self.exchange != "p2p": return 0
        n = C.c_uint64(0); check(lib.vpt_comm_p2p_status(self.r.ctx, C.byref(n)), self.r.ctx, "vpt_comm_p2p_status"); return int(n.value)

    def full_accum(self):
        return self.full.view(self.r.height, self.r.width, 3)

    def close(self):
        self.full = None; self.full_display = None
        self.r.close()


class LevelAModule:
    """Level (A) of the drop-in boundary driven the way the reference application drives its kernel (main.cpp:1221-1236, 1823-1829):
    cuModuleLoad -> cuModuleGetFunction("volume_rt_kernel") -> cuLaunchKernel(grid (w/16+1, h/16+1), block (16,16), params[9]),
    through the CUDA Driver API (cuda-python).  `Renderer` only supplies the parameter block."""

    def __init__(self, cubin_path=None):
        import os
        from cuda.bindings import driver as drv
        self.drv = drv
        torch.cuda.current_stream()                                   # make sure the primary context exists and is current
        path = cubin_path or os.path.join(os.path.dirname(N.LIB_PATH), "volume_rt_kernel_b200.cubin")
        err, self.module = drv.cuModuleLoad(path.encode())
        if int(err): raise N.VptError(f"cuModuleLoad({path}) -> {err}")
        err, self.fn = drv.cuModuleGetFunction(self.module, b"volume_rt_kernel")
        if int(err): raise N.VptError(f"cuModuleGetFunction(volume_rt_kernel) -> {err}")

    def launch(self, r: "Renderer", sync=True):
        """One progressive pass; the caller keeps doing ++iteration like the reference loop."""
        gx, gy = int(r.width // 16) + 1, int(r.height // 16) + 1
        err, = self.drv.cuLaunchKernel(self.fn, gx, gy, 1, 16, 16, 1, 0, 0, C.addressof(r.params.array), 0)
        if int(err): raise N.VptError(f"cuLaunchKernel(volume_rt_kernel) -> {err}")
        r.kp.iteration += 1
        if sync: torch.cuda.synchronize()

    def close(self):
        if getattr(self, "module", None) is not None:
            self.drv.cuModuleUnload(self.module); self.module = None


def stripe_rows_of_rank(height, rank, n_ranks, stripe_rows):
    """Host-side mirror of the kernel's local-row -> global-row map (used by the CPU/gloo tests)."""
    if n_ranks == 1:
        return np.arange(height)
    stripes = (height + stripe_rows - 1) // stripe_rows
    per_rank = (stripes + n_ranks - 1) // n_ranks
    rows = []
    for lr in range(per_rank * stripe_rows):
        s = lr // stripe_rows
        rows.append((s * n_ranks + rank) * stripe_rows + lr % stripe_rows)
    return np.array(rows)

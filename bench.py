#!/usr/bin/env python3
"""bench.py -- Msamples/s (pixels x spp) of the volumetric render pass on dragon.vdb 1920x1080.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one 64-spp frame of BASELINE.json configs[1] (dragon.vdb, 1920x1080, ray_depth 100,
volume_depth 1, direct integrator, sun + HDRI environment): 132.7 M samples.
  * ours      : vpt_render_passes(64) through the C ABI (libvpt_b200.so); N > 1 shards the frame by
                interleaved row stripes over N ranks + one NCCL all-gather of the accumulators per step.
  * reference : the reference's own `volume_rt_kernel` (oracle/_ref, compiled from /root/reference),
                launched as source/main.cpp:1823-1829 does: one launch + cudaDeviceSynchronize per spp.
                The reference has no CPU path and no multi-GPU path: rank 0 runs it on one GPU.
Rank 0 prints ONE JSON line.  Timing: CUDA events on the launching stream around each step, L2 flushed
between steps (excluded), barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

WIDTH, HEIGHT, SPP = 1920, 1080, 64
METRIC = "Msamples/sec (pixels x spp) on dragon.vdb 1920x1080"


def workload_params(V):
    kp = V.default_kernel_params()
    kp.environment_type = 1          # HDRI environment
    kp.ray_depth = 100
    kp.volume_depth = 1
    kp.integrator = 0
    kp.max_interactions = 1000
    return kp


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  The sampler is started before the
    warm-up (nvidia-smi needs ~1 s to start) and samples are kept by timestamp inside the timed window."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index; self.proc = None; self.lines = []; self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True); self.t.start()
            time.sleep(1.5)
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try: self.proc.wait(timeout=2)
        except Exception: self.proc.kill()
        def parse(lines):
            sm, smax, reasons = [], None, set()
            for _, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9: continue
                try:
                    sm.append(float(f[1])); smax = float(f[2])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"): reasons.add(name)
            return sm, smax, reasons
        inside = [x for x in self.lines if self.t0 is not None and self.t0 - 0.02 <= x[0] <= self.t1 + 0.05]
        sm, smax, reasons = parse(inside if inside else self.lines[-10:])
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "window_s": None if self.t0 is None else round(self.t1 - self.t0, 4)}


def flush_l2(buf):
    buf.add_(1)      # 256 MiB read+write > 126 MB L2


def timed_steps(step_fn, steps, warmup, dist, flush_buf, sampler=None):
    for _ in range(warmup):
        step_fn()
    torch.cuda.synchronize()
    if dist is not None: dist.barrier()
    torch.cuda.synchronize()
    evs = []
    wall0 = time.perf_counter(); tw0 = time.time()
    for _ in range(steps):
        flush_l2(flush_buf)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); step_fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if dist is not None: dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    if sampler is not None: sampler.window(tw0, time.time())
    ms = sum(a.elapsed_time(b) for a, b in evs)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, wall


def cpu_baseline_sample(V, scene_args, cam, kp, seconds_target=12.0):
    """CPU restatement (oracle/vpt_oracle.c, 'port') timed on the host cores over a bounded tile of the same workload."""
    try:
        import oracle_cpu
    except Exception as e:                                     # pragma: no cover
        return {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}
    return oracle_cpu.timed_sample(scene_args, cam, kp, WIDTH, HEIGHT, seconds_target)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunk", type=int, default=0, help="passes fused per kernel round (0 = library default)")
    ap.add_argument("--sched-min-lanes", type=int, default=0, help="trace scheduler threshold (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--generic-kernel", action="store_true", help="A/B: force the generic trace kernel instantiation")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local}"))
        dist = dist_mod
    else:
        torch.cuda.set_device(0)
    if args.impl == "reference" and rank != 0:
        if dist is not None: dist.destroy_process_group()
        return 0

    import vpt_b200 as V
    dev = f"cuda:{local if world > 1 else 0}"
    vol = V.Volume.load_vdb(V.find_asset("dragon.vdb"))
    scene = V.Scene([vol.instance()], device=dev, env="Barce_Rooftop_C_3k.hdr")
    data = "dragon.vdb (reference asset) + " + ("Barce_Rooftop_C_3k.hdr" if not scene.data_notes else "; ".join(scene.data_notes))
    kp = workload_params(V)
    flush_buf = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    samples_per_step = WIDTH * HEIGHT * SPP
    sampler = ClockSampler(local if world > 1 else 0)
    config = {"workload": "dragon.vdb 1920x1080 64spp ray_depth=100 volume_depth=1 direct integrator, sun + HDRI env (BASELINE configs[1])",
              "width": WIDTH, "height": HEIGHT, "spp_per_step": SPP, "l2": "flushed between steps (256 MiB rewrite, untimed)"}

    if args.impl == "reference":
        import oracle_ref
        orc = oracle_ref.RefOracle(); orc.load_kernels()
        r = V.Renderer(scene, WIDTH, HEIGHT, kp=kp)
        def step():
            r.kp.iteration = 0
            for _ in range(SPP):                               # main.cpp:1823-1829: launch, ++iteration, cudaDeviceSynchronize
                orc.launch(r.params.array, WIDTH, HEIGHT, orc.UNMODIFIED, sync=True)
                r.kp.iteration += 1
        sampler.start()
        ms, wall = timed_steps(step, args.steps, args.warmup, None, flush_buf, sampler)
        clocks = sampler.stop()
        val = samples_per_step * args.steps / (ms * 1e-3) / 1e6
        config.update({"launch_protocol": "reference main loop: 1 launch + cudaDeviceSynchronize per spp, grid (W/16+1,H/16+1)x(16,16)",
                       "kernel": "source/render_kernel.cu compiled unmodified with -O3 --use_fast_math --maxrregcount=128 for sm_100a"})
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": data, "config": config, "clocks": clocks, "gpu_launches": SPP * args.steps,
                "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": 1, "kind": "reference",
                                 "sample": "the reference has no CPU path (BASELINE.json): its own CUDA kernel on 1 B200, 1 host thread driving it, full workload"},
                "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": wall}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------------------------------ ours
    opts = {}
    if args.chunk: opts["passes_per_chunk"] = args.chunk
    if args.sched_min_lanes: opts["sched_min_lanes"] = args.sched_min_lanes
    if args.generic_kernel: opts["generic_kernel"] = 1
    stream = torch.cuda.current_stream().cuda_stream
    if world > 1:
        dr = V.DistributedRenderer(scene, WIDTH, HEIGHT, kp=kp, stripe_rows=8, options=opts)
        r = dr.r
        def step():
            r.kp.iteration = 0
            r.render(SPP, stream=stream)
            dist.all_gather_into_tensor(dr.gathered, r.buffers.accum)
            dr.full = dr.full_accum()
    else:
        r = V.Renderer(scene, WIDTH, HEIGHT, kp=kp, options=opts)
        def step():
            r.kp.iteration = 0
            r.render(SPP, stream=stream)

    l0, _ = r.stats()
    sampler.start()
    ms, wall = timed_steps(step, args.steps, args.warmup, dist, flush_buf, sampler)
    clocks = sampler.stop()
    l1, _ = r.stats()
    launches_per_step = (l1 - l0) // (args.steps + args.warmup)
    value = samples_per_step * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the public API with host buffers: parameter blocks from host memory in,
    #      the finished frame (accum float3 + display u32) copied back to pinned host memory, every step
    pin_accum = torch.empty(WIDTH * HEIGHT, 3, dtype=torch.float32).pin_memory()
    pin_disp = torch.empty(WIDTH * HEIGHT, dtype=torch.int32).pin_memory()
    def step_e2e():
        step()
        if world > 1:
            if rank == 0: pin_accum.copy_(dr.full.view(-1, 3), non_blocking=True)
        else:
            pin_accum.copy_(r.buffers.accum, non_blocking=True)
            pin_disp.copy_(r.buffers.display, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_e2e, _ = timed_steps(step_e2e, max(3, args.steps // 2), 1, dist, flush_buf)
    e2e_steps = max(3, args.steps // 2)
    e2e_val = samples_per_step * e2e_steps / (ms_e2e * 1e-3) / 1e6
    h2d = 104 + 16 + 5 * 8 + 464 + 312                     # the by-value launch parameter block, per render call
    d2h = WIDTH * HEIGHT * (12 + 4) if world == 1 else (WIDTH * HEIGHT * 12 if rank == 0 else 0)

    # ---- roofline of the dominant kernel (k_trace), measured live with CUDA events on its stream
    roofline = None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak = 6650.0; peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
        def render_only():                                  # rank-local work only: no collective here, the other ranks have moved on
            r.kp.iteration = 0
            r.render(SPP, stream=stream)
        r.set_option("profile", 1); r.kernel_times()         # per-kernel CUDA events, production kernels
        for _ in range(2): render_only()
        torch.cuda.synchronize()
        kt = r.kernel_times()
        r.set_option("profile", 0); r.set_option("count_stats", 1); r.counters(reset=True)   # work counters: generic instantiation, same algorithm
        for _ in range(2): render_only()
        torch.cuda.synchronize()
        cnt = r.counters()
        r.set_option("count_stats", 0)
        n_steps_prof = 2
        samples = r.n_local * SPP * n_steps_prof
        launches = max(1, kt["trace"]["launches"])
        t_trace = kt["trace"]["ms"] / launches
        lookups_per_sample = cnt["lookups"] / max(1, samples)
        # k_trace's own algorithmic bytes: 32 B per density lookup + per traced ray a 24-B queue record
        # (+16 B start position) read and a 48-B sample record written
        bytes_trace = 32.0 * cnt["lookups"] + (24.0 + 16.0 + 48.0) * cnt["rays"]
        achieved = bytes_trace / launches / (t_trace * 1e-3) / 1e9
        simt = cnt["lane_steps"] / max(1, 32 * cnt["warp_step_iters"])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("passes_per_launch") == opts.get("passes_per_chunk", 32) and world == 1: traffic = tj["k_trace_dram_bytes_per_launch"]
        bytes_per_sample = 88.0 + 32.0 * lookups_per_sample          # framebuffer stream + 32 B per density lookup
        step_gbs = value * bytes_per_sample / 1e3
        roofline = {"bound": "hbm", "kernel": "k_trace", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bytes_trace / launches, "avg_launch_ms": t_trace, "launches_per_step": launches / n_steps_prof,
                    "density_lookups_per_sample": lookups_per_sample, "rays_traced_per_sample": cnt["rays"] / max(1, samples),
                    "step_loop_simt_efficiency": simt, "service_round_lanes": cnt["lane_services"] / max(1, cnt["warp_service_rounds"]),
                    "kernel_ms_per_step": {k: v["ms"] / n_steps_prof for k, v in kt.items()},
                    "step": {"bytes_per_sample": bytes_per_sample, "achieved": step_gbs, "frac": step_gbs / peak,
                             "note": "whole step charged with SURVEY 8(d)'s 88 B + 32 B x lookups per sample"},
                    "note": "dragon.vdb is 425 KB: volume lookups are served by L1/TEX/L2, DRAM only sees the ray queue and the sample planes; "
                            "the path is bound by latency / instruction issue, not HBM (SURVEY 8(d)); see profiles/"}

    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_sample(V, scene, r.cam, kp)

    if rank == 0:
        config.update({"passes_per_chunk": opts.get("passes_per_chunk", 32), "partition": f"{world} rank(s), interleaved 8-row stripes" if world > 1 else "single GPU",
                       "collective": "1 NCCL all_gather_into_tensor of float3 accumulators per step" if world > 1 else "none"})
        line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": data, "config": config, "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches_per_step * args.steps), "roofline": roofline, "cpu_baseline": cpu_base, "wall_s": wall}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

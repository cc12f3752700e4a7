#!/usr/bin/env python
"""bench.py -- SRmeetsPS outer loop on B200 (BASELINE.json metric: ms per outer iteration).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload 4k|1080p|mitten-size]

A *step* is one outer iteration (lighting -> albedo -> depth CG (101 passes) -> normals/energy,
SRPS.cu:276-317) of the BASELINE.json workload: synthetic 4096x4096 HR scene, sf=4, 32 images
(config 4), which fits one GPU.  Prints ONE JSON line.

 value     device time per outer iteration, state resident in HBM (cudaEvents on the context's stream)
 e2e       the same solve through the C ABI from pinned HOST buffers: upload + K iterations + download
 roofline  the dominant CG kernel, algorithmic bytes / time vs MEASURED_PEAKS.json: the persistent driver's one launch per
           solve timed in the loop (events around it), else the per-pass kernel timed alone
 cpu_baseline  oracle C/OpenMP transcription on a bounded sample (rank 0, N=1 only)

--impl reference times the reference's own CUDA build (oracle/_ref/ref_replay: unmodified
devicecalls.cu + legacy-cuSPARSE shim) on a bounded sample of the workload (it cannot run
4096^2 x 32: int32 nnz overflow) and reports the per-pixel-linear extrapolation.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (h, w, sf, n, seed)        BASELINE.md §3
    "4k": (4096, 4096, 4, 32, 2000),         # config 4 (the metric's configuration)
    "1080p": (1080, 1920, 4, 20, 1000),      # config 3
    "1k": (1024, 1024, 4, 32, 2000),         # the largest square both arms can run: the same-size pair of the reference ratio
    "small": (512, 512, 4, 8, 7),            # CI-sized
    "n8sim": (4096, 1024, 4, 32, 2000),      # at 2 GPUs: the per-GPU strip of config 4 at 8 GPUs (512 lines x 4096 pixels);
                                             # at 1 GPU: the strip of config 4 at 4 GPUs without any exchange
    "slab8": (4096, 512, 4, 32, 2000),       # at 1 GPU: the strip of config 4 at 8 GPUs without any exchange
}
PARITY_SCENE = (1024, 1024, 4, 8, 77, 3)     # strip_parity: h, w, sf, n, seed, outer iterations
METRIC = "ms per outer iteration (4096x4096 HR, sf=4, 32 images)"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------
def cpu_baseline(h, w, sf, n, seed, full_pixels):
    """oracle port (C/OpenMP transcription of the same iteration) on a bounded sample."""
    import numpy as np
    from oracle import srps_oracle as o
    from oracle.port import Port
    sc = o.synth_scene(h, w, sf, n, seed=seed)
    st = o.init_state(sc["I"], sc["z"], sc["z0s"], sc["ops"], sc["K"], np.float32)
    pt = Port(sc["ops"], sc["n"], sc["c"], st["fx"], st["fy"], st["xx"], st["yy"])
    stp = {k: np.ascontiguousarray(st[k]).copy() for k in ("s", "rho", "z", "N", "dz", "I", "z0s")}
    pt.outer_iteration(stp)                       # warm-up (page faults, OpenMP pool)
    t0 = time.perf_counter()
    iters = 2
    for _ in range(iters):
        pt.outer_iteration(stp)
    ms = (time.perf_counter() - t0) * 1e3 / iters
    scale = full_pixels / float(h * w)
    return {"value": ms * scale, "unit": "ms per outer iteration", "cores": pt.threads(), "kind": "port",
            "sample": f"{h}x{w} sf={sf} n={n} full mask, {iters} outer iterations after 1 warm-up, reference-CG albedo, "
                      f"measured {ms:.1f} ms/iter, scaled x{scale:g} (work is linear in pixels) to the {full_pixels}-pixel workload"}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CUDA build (unmodified devicecalls.cu) on a bounded sample."""
    if rank != 0:
        return
    h, w, sf, n, seed = WORKLOADS[args.workload]
    full_pixels = h * w
    replay = os.path.join(ROOT, "oracle", "_ref", "ref_replay")
    line = {"impl": "reference", "metric": METRIC if args.workload == "4k" else f"ms per outer iteration ({args.workload})",
            "unit": "ms", "higher_is_better": False, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "dtype": "f32", "data": "synthetic", "vs_baseline": None, "scaling": "strong"}
    import numpy as np
    from oracle import datasets as ds
    from oracle import srps_oracle as o
    from srmeetsps_cuda_b200.snapshot import write_snapshot
    if os.path.exists(replay) and not args.ref_cpu:
        sh, sw = args.ref_sample, args.ref_sample
        sc = o.synth_scene(sh, sw, sf, n, seed=seed)
        with tempfile.TemporaryDirectory() as td:
            snap = os.path.join(td, "in.snap")
            write_snapshot(snap, ds.replay_snapshot_arrays(sc))
            res = subprocess.run([replay, snap, os.path.join(td, "o"), "--iters", str(args.steps + args.warmup), "--no-dump"],
                                 capture_output=True, text=True)
        rows = [json.loads(ln.replace(": nan", ": NaN")) for ln in res.stdout.splitlines() if ln.startswith('{"iteration"')]
        if res.returncode == 0 and len(rows) >= args.steps + args.warmup:
            timed = [float(r["ms_total"]) for r in rows[args.warmup:]]
            ms = float(np.mean(timed))
            scale = full_pixels / float(sh * sw)
            line.update({"value": ms * scale, "ms_per_step": ms * scale,
                         # the reference cannot run the 4096^2 x 32 workload: `value` is a per-pixel-linear
                         # extrapolation of the sample; `same_size` is the un-scaled measurement, to be set against the
                         # `same_size` block of this repo's own line (bench.py --workload 1k reproduces both)
                         "same_config": bool(scale == 1.0), "extrapolated": bool(scale != 1.0),
                         "same_size": {"workload": f"{sh}x{sw} HR, sf={sf}, {n} images, full mask", "value": ms, "unit": "ms",
                                       "min": float(np.min(timed)), "median": float(np.median(timed)), "max": float(np.max(timed)),
                                       "steps": len(timed)},
                         "config": {"workload": f"{h}x{w} sf={sf} n={n}", "sample": f"{sh}x{sw}", "scaled_by": scale},
                         "cpu_baseline": {"value": ms * scale, "unit": "ms per outer iteration", "cores": 1, "kind": "reference",
                                          "sample": f"reference CUDA build (cuSPARSE/cuBLAS path on the GPU, no CPU path exists) on "
                                                    f"{sh}x{sw} sf={sf} n={n}: {ms:.1f} ms/iter measured, x{scale:g} per-pixel-linear "
                                                    f"extrapolation; it cannot run {h}x{w}x{n} (int32 nnz overflow, SURVEY F7)"},
                         "e2e": {"value": ms * scale, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                         "phases_ms_sample": {k: float(np.mean([r[k] for r in rows[args.warmup:]])) for k in
                                              ("ms_lighting", "ms_albedo", "ms_depth", "ms_normals")}})
            print(json.dumps(line))
            return
        sys.stderr.write(f"ref_replay failed (rc={res.returncode}); falling back to the CPU port\n{res.stderr[-2000:]}\n")
    cb = cpu_baseline(args.cpu_sample, args.cpu_sample, sf, n, seed, full_pixels)
    line.update({"value": cb["value"], "ms_per_step": cb["value"], "config": {"workload": f"{h}x{w} sf={sf} n={n}"},
                 "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def batch_scenes(local, world, rank, S, albedo, iters=10):
    """BASELINE config 5: this GPU processes S independent 1080p scenes back to back, fixed 10 outer
    iterations each; the upload of scene k+1 (other context, other stream, a host thread -- ctypes releases the GIL
    inside the library) overlaps the iterations of scene k.  Returns scenes/s of this process and the counters."""
    import threading

    import torch
    from srmeetsps_cuda_b200 import Context
    from srmeetsps_cuda_b200.synth import synth_scene_torch
    h, w, sf, n, seed0 = WORKLOADS["1080p"]
    torch.cuda.set_device(local)
    scenes = [synth_scene_torch(h, w, sf, n, seed0 + rank * S + k, device=f"cuda:{local}") for k in range(S)]   # pinned host arrays
    torch.cuda.empty_cache()
    ctxs = [Context(scenes[0]["mask"], n, sf, scenes[0]["K"], device=local, albedo_mode=albedo) for _ in range(2)]
    zout = torch.empty(ctxs[0].npix, dtype=torch.float32, pin_memory=True).numpy()

    def upload(k):
        ctxs[k & 1].upload_state(scenes[k]["I"], scenes[k]["z"], scenes[k]["z0s"])

    def one_round():
        energies = []
        upload(0)
        for k in range(S):
            nxt = None
            if k + 1 < S:
                nxt = threading.Thread(target=upload, args=(k + 1,))
                nxt.start()
            energies.append(float(ctxs[k & 1].run(fixed_iters=iters)[-1]))
            ctxs[k & 1].download("z", out=zout)
            if nxt is not None:
                nxt.join()
        return energies

    one_round()                                                            # warm-up: graphs, allocator, clocks
    sampler = ClockSampler(local)
    barrier(world)
    sampler.start()
    t0 = time.perf_counter()
    energies = one_round()
    for c in ctxs:
        c.synchronize()
    barrier(world)
    dt = max_over_ranks(time.perf_counter() - t0, world)
    clocks = sampler.stop()
    launches = sum(c.timings()["launches"] for c in ctxs) // 2             # one of the two rounds
    for c in ctxs:
        c.close()
    sc = scenes[0]
    res = {"workload": f"BASELINE config 5: {S} scenes per GPU x {world} GPU(s), {h}x{w} HR, sf={sf}, {n} images, full mask, "
                       f"{iters} outer iterations each, double-buffered upload",
           "value": world * S / dt, "unit": "scenes/s", "seconds": dt, "scenes": world * S,
           "h2d_bytes_per_scene": float(sc["I"].nbytes + sc["z"].nbytes + sc["z0s"].nbytes), "d2h_bytes_per_scene": float(zout.nbytes),
           "gpu_launches": int(launches), "clocks": clocks, "energy_last": float(energies[-1]),
           "energies_finite": bool(all(e == e and abs(e) < 1e30 for e in energies))}
    del scenes
    torch.cuda.empty_cache()
    return res


def run_batch(args, rank, world, local):
    """bench.py --batch-scenes S [--gpus N]: BASELINE config 5 as its own line (scenes/s for the box, uploads and the
    download of z included); the default bench line stays BASELINE's headline metric."""
    S = args.batch_scenes
    r = batch_scenes(local, world, rank, S, args.albedo)
    if rank != 0:
        return
    print(json.dumps({
        "metric": f"scenes/s ({world * S} synthetic 1080p scenes, sf=4, 20 images, 10 outer iterations each, uploads included)",
        "value": r["value"], "unit": "scenes/s", "n_gpus": world, "steps": S, "warmup": S, "ms_per_step": r["seconds"] / S * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": r["workload"], "albedo": args.albedo,
                   "parallelism": f"{world} GPU(s), independent scenes, double-buffered upload",
                   "l2": "image stack 0.5 GB per scene: exceeds the 126 MB L2"},
        "e2e": {"value": r["value"], "unit": "scenes/s", "h2d_bytes_per_step": r["h2d_bytes_per_scene"],
                "d2h_bytes_per_step": r["d2h_bytes_per_scene"]},
        "gpu_launches": r["gpu_launches"], "clocks": r["clocks"], "energy_last": r["energy_last"]}))


# ------------------------------------------------------------------------------------------------
def strip_parity_check(rank, world, local, albedo):
    """N > 1, before the timed region: the strip partition against ONE GPU on the same host arrays.

    A 1024x1024x8 full-mask scene (sf 4) is generated once per rank from the same counter-based noise (identical
    arrays everywhere), rank 0 runs it in a single-GPU context, all ranks run their strips of it; after each of 3
    outer iterations the gathered z / rho / s are compared.  Only the fp64 summation order of the dot products differs
    between the two runs.  The bench fails when the partition changes the result beyond fp32 round-off."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from srmeetsps_cuda_b200 import Context
    from srmeetsps_cuda_b200.dist import make_strip_context, strip_bounds
    from srmeetsps_cuda_b200.synth import synth_scene_torch
    h, w, sf, n, seed, iters = PARITY_SCENE
    full = synth_scene_torch(h, w, sf, n, seed, device=f"cuda:{local}", pin=False)
    ref = []
    if rank == 0:
        with Context(full["mask"], n, sf, full["K"], device=local, albedo_mode=albedo) as c1:
            c1.upload_state(full["I"], full["z"], full["z0s"])
            for _ in range(iters):
                e, k = c1.outer_iteration()
                ref.append((e, k, c1.download("z"), c1.download("rho"), c1.download("s")))
    dist.barrier()
    j0, j1 = strip_bounds(w, world)[rank]
    p0, p1, q0, q1 = j0 * h, j1 * h, (j0 // sf) * (h // sf), (j1 // sf) * (h // sf)
    ctx = make_strip_context(full["mask"], n, sf, full["K"], rank, world, local, albedo_mode=albedo)
    assert ctx.pixel_range() == (p0, p1, q0, q1)
    ctx.upload_state_strided(np.ascontiguousarray(full["I"]).reshape(-1)[p0:], h * w, full["z"][p0:p1], full["z0s"][q0:q1])
    out = {"scene": f"{h}x{w} HR, sf={sf}, {n} images, full mask, {iters} outer iterations, strips vs 1 GPU from the same host arrays",
           "z_rel_rmse": 0.0, "rho_maxabs": 0.0, "s_maxabs": 0.0, "energy_rel": 0.0, "cg_iters": [], "cg_iters_1gpu": [],
           "bit_identical_full_mask": True, "ranks_agree": True}
    for it in range(iters):
        e, k = ctx.outer_iteration()
        parts = [None] * world
        dist.all_gather_object(parts, (ctx.download("z"), ctx.download("rho"), ctx.download("s"), e, k))
        if rank == 0:
            z = np.concatenate([p[0] for p in parts]); rho = np.concatenate([p[1] for p in parts], axis=1)
            e1, k1, z1, rho1, s1 = ref[it]
            out["ranks_agree"] &= all(p[3] == parts[0][3] and p[4] == parts[0][4] and np.array_equal(p[2], parts[0][2]) for p in parts)
            zr = float(np.sqrt(np.mean((z.astype(np.float64) - z1) ** 2)) / np.sqrt(np.mean(z1.astype(np.float64) ** 2)))
            out["z_rel_rmse"] = max(out["z_rel_rmse"], zr)
            out["rho_maxabs"] = max(out["rho_maxabs"], float(np.abs(rho - rho1).max()))
            out["s_maxabs"] = max(out["s_maxabs"], float(np.abs(parts[0][2] - s1).max()))
            out["energy_rel"] = max(out["energy_rel"], abs(e - e1) / abs(e1))
            out["cg_iters"].append(int(k)); out["cg_iters_1gpu"].append(int(k1))
            out["bit_identical_full_mask"] &= bool(np.array_equal(z, z1) and np.array_equal(rho, rho1))
    ctx.close()
    ok = True
    if rank == 0:
        # north-star bound / 10: the two runs differ by the summation order of fp64 dot products only
        ok = (out["ranks_agree"] and out["z_rel_rmse"] <= 1e-5 and out["rho_maxabs"] <= 1e-4 and out["energy_rel"] <= 1e-4
              and all(abs(a - b) <= 1 for a, b in zip(out["cg_iters"], out["cg_iters_1gpu"])))
        out["ok"] = bool(ok)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    del full
    torch.cuda.empty_cache()
    if int(flag.item()) != 1:
        if rank == 0:
            sys.stderr.write("strip_parity FAILED: " + json.dumps(out) + "\n")
        dist.barrier()
        sys.exit(3)
    return out


def quantise_stack_u8(I, local):
    """The synthetic fp32 stack as the 8-bit samples an image dataset would hold (round(255 I)), pinned host memory;
    converted plane by plane on the GPU."""
    import torch
    I8 = torch.empty(I.shape, dtype=torch.uint8, pin_memory=True)
    src = torch.from_numpy(I)
    for i in range(I.shape[0]):
        I8[i].copy_((src[i].to(f"cuda:{local}", non_blocking=True) * 255.0).round_().clamp_(0, 255).to(torch.uint8))
    torch.cuda.synchronize()
    return I8


def time_workload(name, local, albedo, steps, warmup, stack="f32"):
    """One single-GPU workload timed the way the headline is (device-resident value, e2e from pinned host): the
    same-size leg of the reference ratio, the config-3 line and the 8-bit-stack line (stack="u8": the images are
    quantised to 8 bits and stay 8-bit in HBM, srps_upload_images_u8)."""
    import numpy as np
    import torch
    from srmeetsps_cuda_b200 import Context
    from srmeetsps_cuda_b200.synth import synth_scene_torch
    h, w, sf, n, seed = WORKLOADS[name]
    sc = synth_scene_torch(h, w, sf, n, seed, device=f"cuda:{local}")
    I8t = quantise_stack_u8(sc["I"], local) if stack == "u8" else None
    I8 = I8t.numpy() if I8t is not None else None

    def upload(ctx):
        if I8 is not None:
            ctx.upload_images_u8(I8)
            ctx.upload_state(None, sc["z"], sc["z0s"])
        else:
            ctx.upload_state(sc["I"], sc["z"], sc["z0s"])

    with Context(sc["mask"], n, sf, sc["K"], device=local, albedo_mode=albedo) as ctx:
        upload(ctx)
        for _ in range(warmup):
            ctx.outer_iteration()
        ctx.timer_start()
        ctx.run(fixed_iters=steps)
        ms_value = ctx.timer_stop() / steps
        per, cg = [], []
        cg_k = 0
        for _ in range(3):
            _, cg_k = ctx.outer_iteration()
            t = ctx.timings()
            per.append(t["ms_total"]); cg.append(t["ms_depth_cg"])
        out = {k: torch.empty(s_, dtype=torch.float32, pin_memory=True).numpy() for k, s_ in
               (("z", (ctx.npix,)), ("rho", (3, ctx.npix)), ("N", (4, ctx.npix)), ("s", (n, 3, 4)))}
        ctx.timer_start()
        upload(ctx)
        ctx.run(fixed_iters=steps)
        for k in out:
            ctx.download(k, out=out[k])
        e2e = ctx.timer_stop() / steps
        phases = {key: float(t[key]) for key in ("ms_lighting", "ms_albedo", "ms_depth")}
        h2d = (I8.nbytes if I8 is not None else sc["I"].nbytes) + sc["z"].nbytes + sc["z0s"].nbytes
    del sc
    torch.cuda.empty_cache()
    return {"workload": f"{h}x{w} HR, sf={sf}, {n} images, full mask", "value": float(ms_value), "unit": "ms",
            "per_separately_synchronised_iteration": float(np.mean(per)), "e2e": float(e2e), "steps": steps, "warmup": warmup,
            "ms_depth_cg": float(np.mean(cg)), "cg_iters": int(cg_k), "phases_ms": phases, "stack": stack, "h2d_bytes": float(h2d)}


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local):
    import numpy as np
    import torch
    from srmeetsps_cuda_b200 import Context
    from srmeetsps_cuda_b200.synth import synth_scene_torch

    from srmeetsps_cuda_b200.dist import make_strip_context, strip_bounds
    h, w, sf, n, seed = WORKLOADS[args.workload]
    torch.cuda.set_device(local)
    strips = world > 1 and args.parallelism == "strips"
    parity = strip_parity_check(rank, world, local, args.albedo) if strips else None
    if strips:
        # ONE scene, strip-partitioned along the image columns (BASELINE config 4 at 2/4/8 GPUs): ghost lines and
        # CG scalars travel inside the library's kernels over NVLink peer memory (csrc/srps_comm.cuh)
        j0, j1 = strip_bounds(w, world)[rank]
        sc = synth_scene_torch(h, w, sf, n, seed, device=f"cuda:{local}", j0=j0, j1=j1)
        t_ctx = time.perf_counter()
        ctx = make_strip_context(sc["mask"], n, sf, sc["K"], rank, world, local, albedo_mode=args.albedo)
    else:
        # N independent replicas, one scene per GPU (BASELINE config 5 pattern, no communication)
        sc = synth_scene_torch(h, w, sf, n, seed + rank, device=f"cuda:{local}")
        t_ctx = time.perf_counter()
        ctx = Context(sc["mask"], n, sf, sc["K"], device=local, albedo_mode=args.albedo)
    npix = ctx.npix                  # pixels this rank owns
    t_ctx = (time.perf_counter() - t_ctx) * 1e3
    ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
    torch.cuda.empty_cache()

    # ---- device-resident timing: W warm-up + K timed outer iterations
    for _ in range(args.warmup):
        ctx.outer_iteration()
    # per-phase split and pass counts: a few separately synchronised iterations (srps_outer_iteration returns after each)
    phases = {"ms_lighting": [], "ms_albedo": [], "ms_depth": [], "ms_normals": [], "ms_depth_cg": []}
    cg_iters, per_call, zskips = [], [], []
    for _ in range(min(3, max(1, args.steps))):
        _, k = ctx.outer_iteration()
        t = ctx.timings()
        per_call.append(t["ms_total"])
        for key in phases:
            phases[key].append(t[key])
        cg_iters.append(k)
        zskips.append(t["cg_zskip"])
    # the timed region: EXACTLY K outer iterations through srps_run(fixed_iters=K) -- the loop the library itself runs,
    # iterations queued back to back -- between two CUDA events on the context's stream, barrier + synchronise on both sides
    sampler = ClockSampler(local)
    barrier(world)
    sampler.start()
    l0 = ctx.timings()["launches"]
    ctx.timer_start()
    energies = ctx.run(fixed_iters=args.steps)
    ms_region = ctx.timer_stop()
    t_last = ctx.timings()                # per-phase device times of the LAST iteration of the timed region (events inside it)
    launches = t_last["launches"] - l0
    barrier(world)
    clocks = sampler.stop()
    assert len(energies) == args.steps
    ms_step = max_over_ranks(ms_region / args.steps, world)
    ms_per_call = max_over_ranks(float(np.mean(per_call)), world)

    # ---- end to end through the C ABI from pinned host memory: upload + K iterations + download
    out = {k: torch.empty(s, dtype=torch.float32, pin_memory=True).numpy() for k, s in
           (("z", (npix,)), ("rho", (3, npix)), ("N", (4, npix)), ("s", (n, 3, 4)))}
    barrier(world)
    ctx.timer_start()
    ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
    e2e_energies = ctx.run(fixed_iters=args.steps)
    for k in out:
        ctx.download(k, out=out[k])
    e2e_ms_total = ctx.timer_stop()
    e2e_ms = max_over_ranks(e2e_ms_total / args.steps, world)
    h2d = (sc["I"].nbytes + sc["z"].nbytes + sc["z0s"].nbytes) / args.steps
    stack_gb = sc["I"].nbytes / 1e9
    d2h = sum(v.nbytes for v in out.values()) / args.steps + 16

    # ---- roofline of the dominant kernel of the CG driver in use, timed alone
    prof = ctx.profile_kernels(reps=30)
    peak, peak_src = measured_peaks()
    fused = prof["cg_driver"] in ("fused", "persistent_fused")
    passes = float(np.mean(cg_iters))
    ms_cg = float(np.mean(phases["ms_depth_cg"]))
    ms_cg_timed = float(t_last["ms_depth_cg"]) if t_last["ms_depth_cg"] > 0 else ms_cg      # the CG launch of the last timed iteration
    zskip = float(np.mean(zskips))
    if prof["cg_driver"] == "persistent_fused":
        # the whole solve is ONE cooperative launch: its duration is the device time between the events around it, taken
        # from the last iteration of the timed region (the mean over the three separately synchronised iterations before
        # it is printed next to it); its algorithmic bytes are those of the passes it ran: 44 B per pixel, 36 in the passes
        # that leave z to the next one (srps_timings.cg_zskip)
        kernel = ("cg_persistent_fused_kernel<sf> (whole CG solve in one launch; per pass: r -= alpha y; p <- r + beta p; "
                  "y <- (KtK + GtMG) p; r.r, p.y, r.y, y.y; one grid barrier; z += two steps in every other pass)")
        alg_bytes, ms_kernel, tkey = (44.0 * passes - 8.0 * zskip) * npix, ms_cg_timed, "cg_persistent_fused"
    elif fused:    # one kernel per pass: reads r, y, p, z, w0..2 ; writes r, p, y, z  
        kernel = "cg_fused_kernel<sf> (CG pass: r -= alpha y; z += alpha p; p <- r + beta p; y <- (KtK + GtMG) p; r.r, p.y, r.y, y.y)"
        alg_bytes, ms_kernel, tkey = 44.0 * npix, prof["cg_fused"], "cg_fused"
    else:          # operator kernel of the two-kernel form: reads r, p, w0..2 ; writes p, y
        kernel = "stencil_strip_kernel<MODE_ITER, sf> (CG operator: p <- r + beta p; y <- (KtK + GtMG) p; p.y)"
        alg_bytes, ms_kernel, tkey = 28.0 * npix, prof["cg_stencil"], "cg_operator"
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel, from the committed ncu --set full capture
    # of this workload on ONE GPU (profiles/traffic.json, written by profiles/summarize.py).  A strip of the scene is
    # another launch geometry: no capture, no number (ncu cannot wrap a multi-rank run).
    traffic = None
    if world == 1:
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get(args.workload, {}).get(tkey)
        except Exception:
            pass
    pass_bytes = ((44.0 * passes - 8.0 * zskip) / passes if prof["cg_driver"] == "persistent_fused" else 44.0) if fused else 52.0
    roofline = {"bound": "hbm", "kernel": kernel,
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "profiles/traffic.json (ncu --set full capture of this kernel, 1 GPU)" if traffic else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_kernel, "passes_per_launch": passes if tkey == "cg_persistent_fused" else 1,
                "passes_without_z": zskip if tkey == "cg_persistent_fused" else 0,
                "ms_per_launch_separately_synchronised": ms_cg if tkey == "cg_persistent_fused" else None,
                "cg_driver": prof["cg_driver"],
                "other_kernels": {
                    "stencil_strip_kernel (two-kernel form)": {"ms": prof["cg_stencil"], "GBps": 28.0 * npix / (prof["cg_stencil"] * 1e-3) / 1e9},
                    "cg_update_kernel (two-kernel form)": {"ms": prof["cg_update"], "GBps": 24.0 * npix / (prof["cg_update"] * 1e-3) / 1e9},
                    "cg_fused_kernel": {"ms": prof["cg_fused"], "GBps": 44.0 * npix / (max(prof["cg_fused"], 1e-9) * 1e-3) / 1e9},
                    "lighting_pass": {"ms": prof["lighting_pass"], "GBps": (4.0 * n * 3 + 24) * npix / (prof["lighting_pass"] * 1e-3) / 1e9},
                    "project_pass": {"ms": prof["project_pass"], "GBps": (4.0 * n * 3 + 72) * npix / (prof["project_pass"] * 1e-3) / 1e9}},
                "cg_loop_GBps_survey_64B": 64.0 * npix * passes / (ms_cg * 1e-3) / 1e9,
                "cg_loop_GBps_actual": pass_bytes * npix * passes / (ms_cg * 1e-3) / 1e9,
                "cg_loop_bytes_per_pixel_pass": pass_bytes}
    ctx.close()

    # ---- the same-size leg of the reference ratio and the other BASELINE configs (N = 1, headline workload only)
    same_size = extra = None
    if world == 1 and args.workload == "4k" and not args.no_extras:
        del sc, out
        torch.cuda.empty_cache()
        same_size = time_workload("1k", local, args.albedo, steps=10, warmup=3)
        same_size["note"] = ("the reference CUDA build runs this size itself (bench.py --impl reference prints its un-scaled "
                             "measurement under the same key): a same-configuration pair for the speed-up")
        extra = {"config3_1080p": time_workload("1080p", local, args.albedo, steps=10, warmup=3)}
        try:            # the headline workload with the 8-bit stack (N3): the fp32 line above stays the headline
            extra["config4_u8_stack"] = time_workload("4k", local, args.albedo, steps=10, warmup=3, stack="u8")
        except Exception as ex:
            extra["config4_u8_stack"] = {"error": repr(ex)}
        try:
            extra["config5_batch_1gpu"] = batch_scenes(local, 1, 0, 4, args.albedo)
        except Exception as ex:                                     # never lose the headline line to an extra
            extra["config5_batch_1gpu"] = {"error": repr(ex)}
    if rank != 0:
        return
    cb = None
    if world == 1 and not args.no_cpu:
        cb = cpu_baseline(args.cpu_sample, args.cpu_sample, sf, n, seed, npix)
    line = {
        "metric": METRIC if args.workload == "4k" else f"ms per outer iteration ({args.workload})",
        "value": ms_step if (strips or world == 1) else ms_step / world, "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False,
        "scaling": "strong" if (strips or world == 1) else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{h}x{w} HR, sf={sf}, {n} images, full mask (BASELINE config {'4' if args.workload == '4k' else '3'})",
                   "albedo": args.albedo, "depth_cg": "reference schedule: un-preconditioned, 101 passes",
                   "parallelism": "1 GPU" if world == 1 else (
                       f"{world} strips of the same scene along the image columns, ghost lines pulled from the neighbours + one 4-value all-reduce per CG pass, "
                       f"exchanged in-kernel over NVLink peer memory" if strips else
                       f"{world} independent replicas, one scene per GPU (value = ms per scene-iteration)"),
                   "l2": (f"per GPU: image stack {stack_gb:.2f} GB, CG working set {pass_bytes * npix / 1e6:.0f} MB per pass; "
                          + ("both exceed the 126 MB L2: no flush needed" if pass_bytes * npix > 126e6 else
                             "the stack exceeds the 126 MB L2, the CG vectors fit it (algorithmic GB/s of the CG may exceed the HBM peak)"))},
        "e2e": {"value": e2e_ms if (strips or world == 1) else e2e_ms / world, "unit": "ms",
                "h2d_bytes_per_step": h2d * (world if strips else 1), "d2h_bytes_per_step": d2h * (world if strips else 1),
                "what": f"srps_upload_state (pinned host) + {args.steps} outer iterations + download of z, rho, N, s; total {e2e_ms_total:.1f} ms / {args.steps}"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cb,
        "strip_parity": parity,
        "same_size": same_size,
        "extra": extra,
        "phases_ms": {k: float(np.mean(v)) for k, v in phases.items()},
        "ms_per_separately_synchronised_iteration": ms_per_call,
        "cg_iters_per_s": float(np.mean(cg_iters)) / (float(np.mean(phases["ms_depth_cg"])) * 1e-3),
        "cg_iters": int(np.mean(cg_iters)),
        "energy_last": float(energies[-1]),
        "one_shot_ms": {"ctx_create": t_ctx,
                        "note": "mask analysis + allocation (the reference's one-shot init, SRPS.cu:151-203) is outside the per-iteration metric"},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10, help="outer iterations timed (default 10 = the reference's MAX_ITERATIONS, SRPS.cu:86)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k", choices=list(WORKLOADS))
    ap.add_argument("--albedo", default="closed_form", choices=["closed_form", "reference_cg"])
    ap.add_argument("--cpu-sample", type=int, default=2048, help="edge of the square sample the CPU port is timed on")
    ap.add_argument("--ref-sample", type=int, default=1024, help="edge of the square sample the reference CUDA build is timed on")
    ap.add_argument("--ref-cpu", action="store_true", help="reference arm: use the CPU port instead of the reference CUDA build")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the same-size / config 3 / config 5 legs of the N=1 headline run")
    ap.add_argument("--batch-scenes", type=int, default=0,
                    help="> 0: BASELINE config 5 instead of the headline metric -- this many 1080p scenes per GPU, scenes/s")
    ap.add_argument("--parallelism", default="strips", choices=["strips", "replicas"],
                    help="N > 1: strip-partition ONE scene (default, strong scaling) or run N independent scenes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        # single-GPU program: under torchrun rank 0 alone runs it, the other ranks exit 0 without work
        rank = int(os.environ.get("RANK", "0"))
        if rank == 0:
            run_reference(args, 0, 1)
        return
    rank, world, local = dist_setup(args.gpus)
    if args.batch_scenes > 0:
        run_batch(args, rank, world, local)
    else:
        run_ours(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

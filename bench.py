#!/usr/bin/env python
"""bench.py -- throughput of the B200 rendering core on BASELINE.json's configurations.

    python bench.py --gpus N --steps K --warmup W [--config C1|C2|C3|C4|C5]      (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W [--config ...]

Default workload (BASELINE.json configs[4], "C5"): Cornell box `default_scene`, path_tracing_iteration_t(depth 5,
both_mis), 3840x2160.  The whole job is 16384 spp; ONE STEP renders a slice of `--spp-per-step` sample indices of
every pixel (throughput does not depend on which slice).  With N GPUs each rank renders its own sample indices of
every pixel (weak scaling: per-GPU work per step is fixed) into a rank-local partial film; the partial films are
summed by one NCCL reduce at the end of the timed region and clamped on rank 0, which is how the job ends (the
reference clamps after the spp-sum, ky.cpp:3726).

The line printed by rank 0 follows the driver's contract: `value` is Msamples/s with everything resident in HBM,
`e2e` is the same metric for the job as a user runs it, host buffers and copies inside the timed region (N = 1: the
host-buffer C ABI call integrator_t::render() makes; N > 1: render on every rank -> one NCCL reduce -> clamp -> ONE
device-to-host copy on rank 0).  `configs` carries short measurements of the OTHER BASELINE configurations made after
the timed region (C1-C4 at their full sizes, sample-split over the same N GPUs), and -- N > 1 -- `multi_gpu_parity` is
the outcome of a film comparison made before timing (the run exits non-zero if it fails).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SM_COUNT, LANES_PER_SM = 148, 128
STAGES = ["raygen", "intersect", "shade", "light_sample", "shadow", "-", "accumulate", "pixel"]
# cost of one shape test that misses (FP32 operations, FMA off), by kyd_shape_kind
MISS_FLOP = {0: 16, 1: 68, 2: 51, 3: 23}


def flop_per_ray(scene):
    """brute-force closest-hit walk that misses everything: sum of the per-shape miss costs over the surface list"""
    return sum(MISS_FLOP[scene.desc.shapes[scene.desc.surfaces[i].shape].kind] for i in range(scene.desc.surface_count))


# ---- workloads: BASELINE.json configs -----------------------------------------------------------------------------------
def workloads(ky):
    """name -> dict(title, job_spp, panels=[(scene_id, scene_flags, w, h, integrator, depth, direct_sample)], reference=...)"""
    PT, DL = ky.INT_PT_ITERATION, ky.INT_DIRECT_LIGHTING
    lights4 = [ky.CB_LIGHT_POINT, ky.CB_LIGHT_DIRECTION, ky.CB_LIGHT_AREA, ky.CB_LIGHT_ENVIRONMENT]
    return {
        "C1": dict(title="C1 smallpt scene (9 spheres) 1024x768, path_tracing_iteration depth 5 both_mis, job 64 spp", job_spp=64,
                   panels=[(ky.SCENE_SMALLPT, 0, 1024, 768, PT, 5, ky.DS_BOTH_MIS)]),
        "C2": dict(title="C2 cornell default_scene 1024x768, direct_lighting, panels bsdf / light / both_mis, job 256 spp", job_spp=256,
                   panels=[(ky.SCENE_CORNELL, ky.CB_DEFAULT, 1024, 768, DL, 0, ds) for ds in (ky.DS_BSDF, ky.DS_LIGHT, ky.DS_BOTH_MIS)]),
        "C3": dict(title="C3 veach MIS scene 1280x720, path_tracing_iteration depth 5 both_mis (render_mis_scene's headline panel), job 1024 spp",
                   job_spp=1024, panels=[(ky.SCENE_VEACH, 0, 1280, 720, PT, 5, ky.DS_BOTH_MIS)]),
        "C4": dict(title="C4 render_multiple_scene: 4 light variants x {bsdf, light, both_mis}, depth 8, 12 panels of 480x360 (1920x1080), job 1024 spp",
                   job_spp=1024, panels=[(ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | l, 480, 360, PT, 8, ds)
                                         for l in lights4 for ds in (ky.DS_BSDF, ky.DS_LIGHT, ky.DS_BOTH_MIS)]),
        "C5": dict(title="C5 cornell default_scene 3840x2160, path_tracing_iteration depth 5 both_mis, job 16384 spp", job_spp=16384,
                   panels=[(ky.SCENE_CORNELL, ky.CB_DEFAULT, 3840, 2160, PT, 5, ky.DS_BOTH_MIS)]),
    }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


def source_hash():
    """identifies the kernel sources a profile was captured from"""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "ky_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_summary():
    """Per-kernel ncu figures of the committed capture (profiles/r*_ncu_summary.json, written by scripts/ncu_summary.py json):
    DRAM traffic and issue-slot utilisation cannot be measured inside a timed run, so the line quotes the capture -- and says
    whether the capture is of this build.  A capture that names kernels this build does not contain is refused."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_summary.json")))
    if not files:
        return None
    path = files[-1]
    d = json.load(open(path))
    lib = open(os.path.join(ROOT, "ky_b200", "lib", "libkyd.so"), "rb").read()
    missing = [k for k in d.get("kernel_symbols", []) if k.encode() not in lib]
    if missing:
        raise SystemExit(f"bench.py: {os.path.relpath(path, ROOT)} describes kernels {missing} that libkyd.so does not contain; "
                         "re-capture it (scripts/gpu_prof.sh) or remove it")
    d["file"] = os.path.relpath(path, ROOT)
    d["same_sources"] = d.get("source_hash") == source_hash()
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- reference arm: the reference's own CPU implementation ------------------------------------------------------------
def cpu_reference(config="C5", bounded_seconds=15.0, threads=0):
    """The reference's own CPU implementation (oracle/_ref verbatim build: reference + compile-only patches, glibc libm,
    per-row mt19937_64 sampler, OpenMP over rows) on a bounded sample of the workload's first panel... of every panel for
    multi-panel configurations the LAST panel (both_mis) stands for the job."""
    import kyref
    import ky_b200 as ky
    kind = "verbatim"
    if not kyref.available(kind):
        return None
    wl = workloads(ky)[config]
    sid, sflags, w, h, integ, depth, ds = wl["panels"][-1]
    scene_of = {ky.SCENE_CORNELL: kyref.CORNELL, ky.SCENE_VEACH: kyref.VEACH, ky.SCENE_SMALLPT: kyref.SMALLPT}
    cores = threads or os.cpu_count() or 1
    common = dict(integrator=integ, max_depth=depth, direct_sample=ds, scene_flags=sflags, sampler=kyref.RANDOM_SAMPLER, threads=cores, kind=kind)
    # calibrate on a thumbnail, then size the real sample: full resolution, as many spp as fit the budget
    tw, th = max(16, w // 8), max(16, h // 8)
    _, sec, _ = kyref.render(scene_of[sid], tw, th, 2, **common)
    rate = tw * th * 2 / max(sec, 1e-6)
    spp = int(max(1, min(wl["job_spp"], bounded_seconds * rate / (w * h))))
    _, sec, rays = kyref.render(scene_of[sid], w, h, spp, **common)
    samples = w * h * spp
    return {"value": samples / sec / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
            "sample": f"{w}x{h} @ {spp} spp of the {config} job ({samples / 1e6:.1f} Msamples, {sec:.1f} s wall)",
            "mrays_per_s": rays / sec / 1e6, "seconds": sec}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ky_b200 as ky
    t = []
    base = None
    for i in range(args.warmup + args.steps):
        r = cpu_reference(args.config, bounded_seconds=max(2.0, 40.0 / max(1, args.warmup + args.steps)))
        if r is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref/libky_ref_verbatim.so was not built (needs /root/reference at build time)"})
            return
        if i >= args.warmup:
            t.append(r)
        base = r
    value = sum(r["value"] for r in t) / len(t)
    ms = 1e3 * sum(r["seconds"] for r in t) / len(t)
    line = {"impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workloads(ky)[args.config]["title"] + "; reference CPU path, each step a bounded spp slice: " + base["sample"]},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "mrays_per_s": sum(r["mrays_per_s"] for r in t) / len(t),
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["cpu_baseline"]["value"] = value
    emit(line)


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when
    NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the real stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


# ---- our arm --------------------------------------------------------------------------------------------------------------
class Job:
    """One BASELINE configuration on this rank: its panels (scene + request + a device film each) and counters."""

    def __init__(self, ky, torch, dev, name, wl, mode_flags, rank, world):
        self.ky, self.torch, self.dev, self.name, self.wl = ky, torch, dev, name, wl
        self.mode_flags, self.rank, self.world = mode_flags, rank, world
        self.job_spp = wl["job_spp"]
        self.panels = []
        floats = 0
        for (sid, sflags, w, h, integ, depth, ds) in wl["panels"]:
            scene = ky.Scene(sid, w, h, sflags)
            self.panels.append(dict(scene=scene, w=w, h=h, integ=integ, depth=depth, ds=ds, offset=floats, flop=flop_per_ray(scene)))
            floats += w * h * 3
        self.film = torch.zeros(floats, dtype=torch.float32, device="cuda")   # every panel's film, one reduce for all
        self.pixels = sum(p["w"] * p["h"] for p in self.panels)
        self.reset_counters()

    def reset_counters(self):
        self.launches = self.rays = self.traced = self.irays = 0
        self.flops_ref = self.flops_traced = self.flops_intersect = 0.0
        self.stage_ms = [0.0] * 8
        self.shade = [0, 0]

    def film_ptr(self, p):
        return self.film.data_ptr() + 4 * p["offset"]

    def render_slice(self, begin, end, collect=False, flags=None):
        """sample indices [begin, end) of every pixel of every panel, accumulated into the panels' device films"""
        ky = self.ky
        for p in self.panels:
            if len(self.panels) > 1 or getattr(self.dev, "_scene", None) is not p["scene"]:
                self.dev.upload(p["scene"])
            d = ky.render_desc(p["w"], p["h"], self.job_spp, integrator=p["integ"], max_depth=p["depth"], direct_sample=p["ds"],
                               sample_begin=begin, sample_end=end, flags=(ky.FLAG_ACCUMULATE | self.mode_flags) if flags is None else flags)
            self.dev.render_device(d, self.film_ptr(p), None)   # the context's own (blocking) stream; returns when done
            if collect:
                st = self.dev.stats()
                self.launches += st.kernel_launches
                self.rays += st.rays
                self.traced += st.rays_traced
                self.irays += st.intersect_rays
                self.flops_ref += st.rays * p["flop"]
                self.flops_traced += st.rays_traced * p["flop"]
                self.flops_intersect += st.intersect_rays * p["flop"]
                for j in range(8):
                    self.stage_ms[j] += st.stage_ms[j]
                self.shade[0] += st.shade_vertices
                self.shade[1] += st.shade_light_lines


def multi_gpu_parity(ky, torch, dist, dev, rank, world):
    """Before anything is timed: a small fixed job through the multi-GPU path (ky_b200.distributed.render_job: sample split,
    NCCL reduce, clamp on the root) against the single-rank film, and the rank-ordered sum of the ranks' partial films
    (gathered to the root, added in rank order) against the same partial films rendered and added on the root alone."""
    from ky_b200.distributed import render_job, sample_range
    out = {"status": "ok", "cases": []}
    for sid, name, w, h, spp in ((ky.SCENE_CORNELL, "cornell", 160, 96, 24), (ky.SCENE_VEACH, "veach", 128, 72, 16)):
        scene = ky.Scene(sid, w, h)
        dev.upload(scene)
        desc = ky.render_desc(w, h, spp)
        film = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
        render_job(dev, scene, desc, film)                     # NCCL reduce + clamp
        # the same partial films, gathered and added in rank order
        b, e = sample_range(spp, world, rank)
        part = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
        if e > b:
            dev.render_device(ky.render_desc(w, h, spp, sample_begin=b, sample_end=e, flags=0), part.data_ptr(), None)
        parts = [torch.empty_like(part) for _ in range(world)] if rank == 0 else None
        dist.gather(part, parts, dst=0)
        if rank == 0:
            ordered = parts[0].clone()
            for r in range(1, world):
                if sample_range(spp, world, r)[1] > sample_range(spp, world, r)[0]:
                    ordered += parts[r]
            ordered.clamp_(0.0, 1.0)
            # rank 0 alone: one-shot film, and every rank's share rendered here and added in rank order
            single = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
            dev.render_device(ky.render_desc(w, h, spp), single.data_ptr(), None)
            alone = None
            for r in range(world):
                rb, re_ = sample_range(spp, world, r)
                if re_ == rb:
                    continue
                tmp = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
                dev.render_device(ky.render_desc(w, h, spp, sample_begin=rb, sample_end=re_, flags=0), tmp.data_ptr(), None)
                alone = tmp if alone is None else alone + tmp
            alone.clamp_(0.0, 1.0)
            torch.cuda.synchronize()
            tol = 2e-6 * torch.clamp(single.abs(), min=1.0)
            err_nccl = float(((film - single).abs() / torch.clamp(single.abs(), min=1.0)).max())
            ok_nccl = bool(((film - single).abs() <= tol).all())
            ok_order = bool(torch.equal(ordered.view(torch.int32), alone.view(torch.int32)))
            ok_order_tol = bool(((ordered - single).abs() <= tol).all())
            out["cases"].append({"scene": name, "size": f"{w}x{h}@{spp}", "nccl_reduce_max_rel_err_vs_single_gpu": err_nccl,
                                 "nccl_reduce_within_2e-6": ok_nccl, "rank_order_sum_bit_exact": ok_order})
            if not (ok_nccl and ok_order and ok_order_tol):
                out["status"] = "FAILED"
    flag = torch.tensor([1 if out["status"] == "ok" else 0], device="cuda")
    dist.broadcast(flag, src=0)
    out["status"] = "ok" if int(flag.item()) == 1 else "FAILED"
    return out


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C5", choices=["C1", "C2", "C3", "C4", "C5"], help="BASELINE.json configuration the line is measured on")
    ap.add_argument("--spp-per-step", type=int, default=0, help="sample indices per pixel and step (0: 32 for C5, else sized to ~256 Mi samples per step)")
    ap.add_argument("--mode", default="wavefront", choices=["wavefront", "wavefront-split", "pixel"], help="kernel organisation (kyd flags)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short measurements of the other configurations")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--wave-paths", type=int, default=0, help="paths per wavefront (0 = library default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import ky_b200 as ky

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # CUDA events around every kernel of the wavefront, on the stream it is launched on: the live per-stage
    # durations the roofline of the dominant kernel is computed from
    os.environ["KYD_STAGE_TIMING"] = "1"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    dev = ky.Device(local)
    dev.set_wave_paths(args.wave_paths)
    mode_flags = {"wavefront": 0, "wavefront-split": ky.FLAG_SPLIT_LIGHT_SAMPLE, "pixel": ky.FLAG_FUSED}[args.mode]
    WL = workloads(ky)
    peaks = measured_peaks()
    peak_lane_ops = SM_COUNT * LANES_PER_SM * peaks["sm_max_mhz"] * 1e6

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if world > 1:
        parity = multi_gpu_parity(ky, torch, dist, dev, rank, world)
        if parity["status"] != "ok":
            if rank == 0:
                print("bench.py: multi-GPU film parity FAILED: " + json.dumps(parity), file=sys.stderr)
            dist.destroy_process_group()
            sys.exit(3)

    job = Job(ky, torch, dev, args.config, WL[args.config], mode_flags, rank, world)
    JOB_SPP = job.job_spp
    S = args.spp_per_step or (32 if args.config == "C5" else max(1, min(JOB_SPP // world, (256 << 20) // job.pixels)))
    share = max(S, JOB_SPP // world)
    base = rank * (JOB_SPP // world)
    if base + share > JOB_SPP:       # tiny jobs on many GPUs: every rank still renders S valid sample indices
        base = max(0, JOB_SPP - share)
    S = min(S, JOB_SPP)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step(k, collect=False):
        b = base + (k * S) % max(1, share - S + 1)
        job.render_slice(b, b + S, collect)

    for k in range(args.warmup):
        step(k)
    if world > 1:
        warm = job.film.clone()
        dist.reduce(warm, dst=0)
        del warm
    barrier()
    job.film.zero_()
    job.reset_counters()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations, outside the per-step events
        ev[k][0].record()
        step(args.warmup + k, collect=True)
        ev[k][1].record()
    # the job's last act: one reduce of the partial films over NVLink, clamp on the root
    stream = torch.cuda.current_stream().cuda_stream
    ev[-1][0].record()
    if world > 1:
        dist.reduce(job.film, dst=0)
    if rank == 0:
        dev.clamp_device(job.film.data_ptr(), job.film.numel(), stream)
        job.launches += 1
    ev[-1][1].record()
    barrier()
    clock_info = clocks.stop() if rank == 0 else None

    step_ms = [a.elapsed_time(b) for a, b in ev[:-1]]
    reduce_ms = ev[-1][0].elapsed_time(ev[-1][1])
    total_ms = torch.tensor([sum(step_ms) + reduce_ms, sum(step_ms)], dtype=torch.float64, device="cuda")
    counts = torch.tensor([job.rays, job.traced, job.launches, job.flops_ref, job.flops_traced], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_ms, kernels_ms = total_ms.tolist()
    rays_all, traced_all, launches_all, flops_ref_all, flops_traced_all = counts.tolist()
    samples = job.pixels * S * args.steps * world
    value = samples / (total_ms * 1e-3) / 1e6

    # ---- e2e: the job as a user runs it, host buffers, copies inside the timed region --------------------------------
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    host_film = np.empty(job.film.numel(), np.float32)
    pinned = torch.empty(job.film.numel(), dtype=torch.float32, pin_memory=True) if world > 1 else None
    scene_bytes = 0
    for p in job.panels:
        sd = p["scene"].desc
        scene_bytes += (ky.C.sizeof(ky.SceneDesc) + sd.shape_count * ky.C.sizeof(ky.Shape) + sd.material_count * ky.C.sizeof(ky.Material)
                        + sd.light_count * ky.C.sizeof(ky.Light) + sd.surface_count * ky.C.sizeof(ky.Surface) + ky.C.sizeof(ky.RenderDesc))

    def e2e_step(k):
        b = base + (k * S) % max(1, share - S + 1)
        if world == 1:
            # the host-buffer C ABI call integrator_t::render makes: scene + request in, clamped film out
            for p in job.panels:
                d = ky.render_desc(p["w"], p["h"], JOB_SPP, integrator=p["integ"], max_depth=p["depth"], direct_sample=p["ds"],
                                   sample_begin=b, sample_end=b + S, flags=ky.FLAG_CLAMP | mode_flags)
                dev.upload(p["scene"])
                dev.render(d, host_film[p["offset"]:p["offset"] + p["w"] * p["h"] * 3])
        else:
            # the multi-GPU job: every rank uploads the scene and renders its sample indices into a device-resident partial
            # film, ONE NCCL reduce, clamp on the root, ONE device-to-host copy of the finished film on the root
            for p in job.panels:
                dev.upload(p["scene"])
                d = ky.render_desc(p["w"], p["h"], JOB_SPP, integrator=p["integ"], max_depth=p["depth"], direct_sample=p["ds"],
                                   sample_begin=b, sample_end=b + S, flags=mode_flags)
                dev.render_device(d, job.film_ptr(p), None)
            dist.reduce(job.film, dst=0)
            if rank == 0:
                dev.clamp_device(job.film.data_ptr(), job.film.numel(), stream)
                pinned.copy_(job.film, non_blocking=True)
                torch.cuda.synchronize()
                host_film[:] = pinned.numpy()

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(1 + k)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = job.pixels * S * e2e_steps * world / e2e_s.item() / 1e6

    # ---- the other BASELINE configurations, briefly (after the timed region): whole jobs, sample-split over the ranks ----
    configs = {}
    if not args.no_configs:
        from ky_b200.distributed import sample_range
        for name in ("C1", "C2", "C3", "C4"):
            if name == args.config:
                continue
            cj = Job(ky, torch, dev, name, WL[name], mode_flags, rank, world)
            b, e = sample_range(cj.job_spp, world, rank)
            cj.render_slice(b, min(e, b + 1))            # warm-up: buffers, clocks
            cj.film.zero_()
            cj.reset_counters()
            barrier()
            t_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            t_ev[0].record()
            cj.render_slice(b, e, collect=True)
            if world > 1:
                dist.reduce(cj.film, dst=0)
            if rank == 0:
                dev.clamp_device(cj.film.data_ptr(), cj.film.numel(), stream)
            t_ev[1].record()
            barrier()
            ms = torch.tensor([t_ev[0].elapsed_time(t_ev[1])], dtype=torch.float64, device="cuda")
            c = torch.tensor([cj.rays, cj.traced, cj.flops_ref, cj.flops_traced, cj.irays, cj.flops_intersect], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.all_reduce(c, op=dist.ReduceOp.SUM)
            ms = ms.item()
            c = c.tolist()
            n_samples = cj.pixels * cj.job_spp
            configs[name] = {
                "workload": WL[name]["title"], "n_gpus": world, "samples": n_samples, "ms": ms,
                "msamples_per_s": n_samples / ms / 1e3,
                "mrays_per_s_reference_equivalent": c[0] / ms / 1e3, "mrays_traced_per_s": c[1] / ms / 1e3,
                "rays_per_sample": c[0] / n_samples,
                "stage_ms_rank0": {n: cj.stage_ms[j] for j, n in enumerate(STAGES) if cj.stage_ms[j] > 0},
                "roofline_fp32_issue": {"frac_traced": c[3] / (ms * 1e-3) / world / peak_lane_ops,
                                        "frac_reference_equivalent": c[2] / (ms * 1e-3) / world / peak_lane_ops,
                                        "intersect_kernel_frac": (cj.flops_intersect / (cj.stage_ms[1] * 1e-3) / peak_lane_ops) if cj.stage_ms[1] > 0 else None},
                "vs_north_star_625_msamples_per_gpu": n_samples / ms / 1e3 / world / 625.0,
            }
            if name == "C1" and world == 1:
                # the reference's OWN CUDA kernel for this scene (smallpt2pbrt/smallpt_kernel.cu: FP64, recursive, no light
                # sampling, one thread per pixel; built for sm_100a by oracle/ref/build_ref.sh) beside this library's FP64
                # validation mode of the same algorithm -- a different algorithm from the FP32 both_mis path above, reported
                # as the reference's GPU datum
                ref_gpu = {"workload": "smallpt_kernel.cu Device::Render(1024, 768, 64 spp), FP64, wall clock incl. managed film allocation; child process"}
                try:
                    script = os.path.join(ROOT, "scripts", "ref_cuda_smallpt.py")
                    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsmallpt_kernel_cuda_ref.so")):
                        for stack in (0, 8192):   # as written (default stack x 3), then with a larger base stack limit
                            r = subprocess.run([sys.executable, script, "1024", "768", "64", str(stack)], capture_output=True, text=True, timeout=300)
                            key = "as_written" if stack == 0 else "with_base_stack_8192"
                            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
                            if r.returncode == 0 and lines:
                                ref_gpu[key] = json.loads(lines[-1])
                            else:
                                ref_gpu[key] = {"failed": f"exit code {r.returncode}", "message": (r.stderr or r.stdout).strip()[-200:]}
                        dev.render_smallpt_f64(256, 192, 4)
                        t0 = time.perf_counter()
                        dev.render_smallpt_f64(1024, 768, 64)
                        ours_s = time.perf_counter() - t0
                        ref_gpu["kyd_render_smallpt_f64_msamples_per_s"] = 1024 * 768 * 64 / ours_s / 1e6
                        ref_gpu["kyd_render_smallpt_f64_kernel_msamples_per_s"] = 1024 * 768 * 64 / dev.stats().device_ms / 1e3
                    else:
                        ref_gpu["unavailable"] = "oracle/_ref/libsmallpt_kernel_cuda_ref.so not built"
                except Exception as e:  # noqa: BLE001  (a measurement beside the contract: never fatal)
                    ref_gpu["unavailable"] = repr(e)
                configs[name]["reference_gpu"] = ref_gpu
            del cj

    if rank == 0:
        sm_mhz = (clock_info or {}).get("sm_mhz") or peaks["sm_max_mhz"]
        ncu = ncu_summary()
        stage_ms = job.stage_ms
        # FP32-issue roofline of the ray-query work, per GPU: brute-force miss cost of the scene's surface list
        # x scene queries.  `frac` counts the queries the device TRAVERSED; the reference-equivalent figure (queries the
        # reference issues for the same samples, 29 % of which the device never needs to trace) is reported beside it.
        frac_traced = flops_traced_all / (kernels_ms * 1e-3) / world / peak_lane_ops
        frac_ref = flops_ref_all / (kernels_ms * 1e-3) / world / peak_lane_ops
        intersect_frac = (job.flops_intersect / (stage_ms[1] * 1e-3) / peak_lane_ops) if stage_ms[1] > 0 else None
        # dominant kernel = the stage with the largest live duration (rank 0's events)
        dom = max(range(8), key=lambda j: stage_ms[j])
        if dom == 2 and job.shade[0] > 0:
            # shade: gathers a 64-byte path record per vertex and rewrites it (+4 B queue entry in, +4..8 B out); multi-light
            # scenes also write 64-byte light-sampling lines
            # (shade[1]: light-sampling lines written for the shadow stage, 64 B each; with k_nee the counter holds (vertex, light) pairs)
            alg_bytes = 136 * job.shade[0] + 64 * job.shade[1]
            kshade = (ncu or {}).get("kernels", {}).get("k_shade<lambert>", {})
            roofline = {"bound": "hbm", "kernel": "k_shade<lobe> (4 launches per bounce)", "achieved": alg_bytes / (stage_ms[2] * 1e-3) / 1e9,
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": alg_bytes / (stage_ms[2] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        "traffic": kshade.get("dram_bytes"),
                        "algorithmic_bytes_per_step": alg_bytes / args.steps, "ms_per_step": stage_ms[2] / args.steps,
                        "note": f"peak = {peaks['source']} copy bandwidth; gather/scatter of 64-byte records through lobe-sorted queues "
                                "(tools/membench.cu: 3.7-3.9 TB/s for that pattern); the kernel also traces the vertex' light queries and -- fused "
                                "configuration -- the path's next ray, which adds instructions but no bytes, so it is issue-bound, not HBM-bound: "
                                "see roofline_fp32_issue and ncu.kernels; traffic = dram bytes of ONE k_shade<lambert> launch of the committed "
                                "capture (ncu.file), not of a whole step"}
        else:
            roofline = {"bound": "fp32_issue", "kernel": STAGES[dom], "achieved": flops_traced_all / (kernels_ms * 1e-3) / world / 1e12,
                        "peak": peak_lane_ops / 1e12, "unit": "TFLOP/s (FP32 lane-ops/s, FMA off)", "frac": frac_traced, "traffic": None}
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WL[args.config]["title"] + f"; step = {S} spp slice per GPU (sample-index split across GPUs, one NCCL film reduce + clamp at the end)",
                       "mode": args.mode, "wave_paths": args.wave_paths, "l2": "flushed between timed steps (256 MiB write); each step also renders new sample indices",
                       "parallelism": f"spp-split x{world}"},
            "mrays_per_s": rays_all / (total_ms * 1e-3) / 1e6,
            "mrays_traced_per_s": traced_all / (total_ms * 1e-3) / 1e6,
            "rays_per_sample": rays_all / samples,
            "reduce_ms": reduce_ms,
            "stage_ms_per_step": {n: stage_ms[j] / args.steps for j, n in enumerate(STAGES) if stage_ms[j] > 0},
            "gpu_launches": int(launches_all),
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": scene_bytes * world, "d2h_bytes_per_step": job.film.numel() * 4,
                    "steps": e2e_steps,
                    "path": "kyd_render: scene upload, render, clamp, film download (host buffer)" if world == 1 else
                            "per rank: scene upload + kyd_render_device; then ONE NCCL reduce, clamp on rank 0, ONE film download on rank 0"},
            "roofline": roofline,
            "roofline_fp32_issue": {"achieved": flops_traced_all / (kernels_ms * 1e-3) / world / 1e12, "peak": peak_lane_ops / 1e12,
                                    "unit": "TFLOP/s (FP32 lane-ops/s, FMA off)", "frac": frac_traced,
                                    "frac_reference_equivalent_rays": frac_ref,
                                    "intersect_kernel": {"frac": intersect_frac, "rays": job.irays, "ms": stage_ms[1],
                                                         "note": "k_intersect alone: its own closest-hit queries x miss cost / its own event time (rank 0)"},
                                    "note": f"whole step: {job.panels[-1]['flop']} flop per scene query (miss cost of a brute-force walk; any-hit queries leave early, "
                                            f"so this is an upper bound of the work a query needs) x queries the device traversed, per GPU; peak = {SM_COUNT} SMs x "
                                            f"{LANES_PER_SM} lanes x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} max SM clock; median under load {sm_mhz}); "
                                            "tensor cores unused"},
            "ncu": ncu,
        }
        if configs:
            line["configs"] = configs
        if parity is not None:
            line["multi_gpu_parity"] = parity["status"]
            line["multi_gpu_parity_detail"] = parity["cases"]
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(args.config)
            if cb:
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "mrays_per_s")}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

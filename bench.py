#!/usr/bin/env python
"""bench.py -- throughput of the B200 rendering core on BASELINE.json's multi-GPU configuration.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4], "C5"): Cornell box `default_scene`, path_tracing_iteration_t(depth 5,
both_mis), 3840x2160.  The whole job is 16384 spp; ONE STEP renders a slice of `--spp-per-step` sample
indices of every pixel (throughput does not depend on which slice).  With N GPUs each rank renders its own
sample indices of every pixel (weak scaling: per-GPU work per step is fixed) into a rank-local partial
film; the partial films are summed by one NCCL reduce at the end of the timed region and clamped on rank 0,
which is how the job ends (the reference clamps after the spp-sum, ky.cpp:3726).

The line printed by rank 0 follows the driver's contract; `value` is Msamples/s with everything resident in
HBM, `e2e` is the same metric through the host-buffer C ABI call integrator_t::render() makes.
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

WIDTH, HEIGHT, JOB_SPP, DEPTH = 3840, 2160, 16384, 5
# brute-force miss cost per ray query, Cornell default scene = 10 rectangles x 68 + 2 spheres x 16
FLOP_PER_RAY = 10 * 68 + 2 * 16
SM_COUNT, LANES_PER_SM = 148, 128
STAGES = ["raygen", "intersect", "shade", "light_sample", "shadow", "-", "accumulate", "pixel"]
# dram__bytes_read.sum + dram__bytes_write.sum of one k_shade<Lambert> launch (ncu --set full, profiles/r01_final_kernels.txt)
NCU_SHADE_TRAFFIC = 768990208  # bytes, k_shade<Lambert> of bounce 1 of a 16.6 M-path wave
# smsp__issue_active.avg.pct_of_peak_sustained_active of the same capture (profiles/r01_final_kernels.txt)
NCU_ISSUE_ACTIVE = {"k_intersect": 0.81, "k_shade<Lambert>": 0.59, "k_shade<Phong>": 0.40, "source": "profiles/r01_final_kernels.txt"}


def JOB_WAVES(spp_per_step, capacity=1 << 24):
    """waves one step is cut into by the default wave size"""
    per_wave = max(1, capacity // (WIDTH * HEIGHT))
    return -(-spp_per_step // per_wave)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


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


def cpu_reference(bounded_seconds=15.0, threads=0):
    """The reference's own CPU implementation (oracle/_ref verbatim build: reference + compile-only patches,
    glibc libm, per-row mt19937_64 sampler, OpenMP over rows) on a bounded sample of the workload."""
    import kyref
    kind = "verbatim"
    if not kyref.available(kind):
        return None
    cores = threads or os.cpu_count() or 1
    common = dict(integrator=kyref.PT_ITERATION, max_depth=DEPTH, direct_sample=kyref.BOTH_MIS, scene_flags=kyref.DEFAULT_SCENE,
                  sampler=kyref.RANDOM_SAMPLER, threads=cores, kind=kind)
    # calibrate on a thumbnail, then size the real sample: full resolution, as many spp as fit the budget
    _, sec, _ = kyref.render(kyref.CORNELL, WIDTH // 8, HEIGHT // 8, 2, **common)
    rate = (WIDTH // 8) * (HEIGHT // 8) * 2 / max(sec, 1e-6)
    spp = int(max(1, min(64, bounded_seconds * rate / (WIDTH * HEIGHT))))
    _, sec, rays = kyref.render(kyref.CORNELL, WIDTH, HEIGHT, spp, **common)
    samples = WIDTH * HEIGHT * spp
    return {"value": samples / sec / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
            "sample": f"{WIDTH}x{HEIGHT} @ {spp} spp of the C5 job ({samples / 1e6:.1f} Msamples, {sec:.1f} s wall)",
            "mrays_per_s": rays / sec / 1e6, "seconds": sec}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t = []
    base = None
    for i in range(args.warmup + args.steps):
        r = cpu_reference(bounded_seconds=max(2.0, 40.0 / max(1, args.warmup + args.steps)))
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
            "config": {"workload": "C5 cornell default_scene 3840x2160, path_tracing_iteration depth 5 both_mis; reference CPU path, "
                                   "each step a bounded spp slice: " + base["sample"]},
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


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=32)
    ap.add_argument("--mode", default="wavefront", choices=["wavefront", "wavefront-split", "pixel"], help="kernel organisation (kyd flags)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--wave-paths", type=int, default=0, help="paths per wavefront (0 = library default)")
    ap.add_argument("--direct-sample", default="both_mis", choices=["idle", "bsdf", "light", "bsdf_mis", "light_mis", "both_mis"],
                    help="direct_sample_enum_t of the workload (the headline workload is both_mis; others are diagnostics)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import ky_b200 as ky
    DS = {"idle": ky.DS_IDLE, "bsdf": ky.DS_BSDF, "light": ky.DS_LIGHT, "bsdf_mis": ky.DS_BSDF_MIS, "light_mis": ky.DS_LIGHT_MIS,
          "both_mis": ky.DS_BOTH_MIS}[args.direct_sample]

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
    scene = ky.Scene(ky.SCENE_CORNELL, WIDTH, HEIGHT, ky.CB_DEFAULT)
    dev.upload(scene)
    dev.set_wave_paths(args.wave_paths)
    S = args.spp_per_step
    share = JOB_SPP // world
    base = rank * share
    mode_flags = {"wavefront": 0, "wavefront-split": ky.FLAG_SPLIT_LIGHT_SAMPLE, "pixel": ky.FLAG_FUSED}[args.mode]

    film = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def step(k):
        b = base + (k * S) % max(S, share - S + 1)
        d = ky.render_desc(WIDTH, HEIGHT, JOB_SPP, integrator=ky.INT_PT_ITERATION, max_depth=DEPTH, direct_sample=DS,
                           sample_begin=b, sample_end=b + S, flags=ky.FLAG_ACCUMULATE | mode_flags)
        dev.render_device(d, film.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step(k)
    if world > 1:
        warm = film.clone()
        dist.reduce(warm, dst=0)
        del warm
    barrier()
    film.zero_()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    launches = rays = traced = 0
    stage_ms = [0.0] * 8
    shade = [0, 0]
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations, outside the per-step events
        ev[k][0].record()
        step(args.warmup + k)
        ev[k][1].record()
        st = dev.stats()  # synchronises; counters of this step
        launches += st.kernel_launches
        rays += st.rays
        traced += st.rays_traced
        for j in range(8):
            stage_ms[j] += st.stage_ms[j]
        shade[0] += st.shade_vertices
        shade[1] += st.shade_light_lines
    # the job's last act: one reduce of the partial films over NVLink, clamp on the root
    ev[-1][0].record()
    if world > 1:
        dist.reduce(film, dst=0)
    if rank == 0:
        dev.clamp_device(film.data_ptr(), film.numel(), stream)
        launches += 1
    ev[-1][1].record()
    barrier()
    clock_info = clocks.stop() if rank == 0 else None

    step_ms = [a.elapsed_time(b) for a, b in ev[:-1]]
    reduce_ms = ev[-1][0].elapsed_time(ev[-1][1])
    total_ms = torch.tensor([sum(step_ms) + reduce_ms, sum(step_ms)], dtype=torch.float64, device="cuda")
    counts = torch.tensor([rays, traced, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_ms, kernels_ms = total_ms.tolist()
    rays_all, traced_all, launches_all = counts.tolist()
    samples = WIDTH * HEIGHT * S * args.steps * world
    value = samples / (total_ms * 1e-3) / 1e6

    # ---- e2e: the host-buffer C ABI call (what integrator_t::render makes), copies inside the timed region
    host_film = np.empty((HEIGHT, WIDTH, 3), np.float32)
    e2e_steps = max(1, min(args.e2e_steps, args.steps))

    def e2e_step(k):
        b = base + (k * S) % max(S, share - S + 1)
        d = ky.render_desc(WIDTH, HEIGHT, JOB_SPP, integrator=ky.INT_PT_ITERATION, max_depth=DEPTH, direct_sample=DS,
                           sample_begin=b, sample_end=b + S, flags=ky.FLAG_CLAMP | mode_flags)
        dev.upload(scene)            # host->device: the flattened scene + request
        dev.render(d, host_film)     # device->host: the film
    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(1 + k)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = WIDTH * HEIGHT * S * e2e_steps * world / e2e_s.item() / 1e6
    scene_bytes = (ky.C.sizeof(ky.SceneDesc) + scene.desc.shape_count * ky.C.sizeof(ky.Shape) + scene.desc.material_count * ky.C.sizeof(ky.Material)
                   + scene.desc.light_count * ky.C.sizeof(ky.Light) + scene.desc.surface_count * ky.C.sizeof(ky.Surface) + ky.C.sizeof(ky.RenderDesc))

    if rank == 0:
        peaks = measured_peaks()
        sm_mhz = (clock_info or {}).get("sm_mhz") or peaks["sm_max_mhz"]
        # FP32-issue roofline of the ray-query work; per GPU
        peak_lane_ops = SM_COUNT * LANES_PER_SM * peaks["sm_max_mhz"] * 1e6
        achieved_flops = rays_all * FLOP_PER_RAY / (kernels_ms * 1e-3) / world
        # dominant kernel = the stage with the largest live duration (rank 0's events)
        dom = max(range(8), key=lambda j: stage_ms[j])
        if dom == 2 and shade[0] > 0:
            # shade: gathers a 64-byte path record per vertex and rewrites it (+4 B queue entry in, +4..8 B out); multi-light
            # scenes also write 64-byte light-sampling lines.  Single-light scenes (this workload) trace
            # the shadow queries inside shade, which makes the kernel issue-bound (59 % issue-active, 12 % of DRAM peak under
            # ncu): the HBM figure is reported because SURVEY.md 8(d) asks for both rooflines; roofline_fp32_issue is the binding one
            alg_bytes = 136 * shade[0] + 64 * shade[1]
            launches_dom = 4 * (DEPTH + 1) * (JOB_WAVES(S)) * args.steps
            roofline = {"bound": "hbm", "kernel": "k_shade<lobe> (4 launches per bounce)", "achieved": alg_bytes / (stage_ms[2] * 1e-3) / 1e9,
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": alg_bytes / (stage_ms[2] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        "traffic": NCU_SHADE_TRAFFIC,
                        "algorithmic_bytes_per_step": alg_bytes / args.steps, "ms_per_step": stage_ms[2] / args.steps,
                        "note": f"peak = {peaks['source']} copy bandwidth; gather/scatter of 64-byte records through lobe-sorted queues "
                                "(tools/membench.cu: 3.7-3.9 TB/s for that pattern, profiles/r01_membench.txt); with the shadow queries traced "
                                "inside shade the kernel is issue-bound (59 % issue-active under ncu, profiles/r01_final_kernels.txt): see roofline_fp32_issue"}
        else:
            roofline = {"bound": "fp32_issue", "kernel": STAGES[dom], "achieved": achieved_flops / 1e12, "peak": peak_lane_ops / 1e12,
                        "unit": "TFLOP/s (FP32 lane-ops/s, FMA off)", "frac": achieved_flops / peak_lane_ops, "traffic": None}
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C5 cornell default_scene {WIDTH}x{HEIGHT}, path_tracing_iteration depth {DEPTH} {args.direct_sample}, job {JOB_SPP} spp; "
                                   f"step = {S} spp slice per GPU (sample-index split across GPUs, one NCCL film reduce + clamp at the end)",
                       "mode": args.mode, "wave_paths": args.wave_paths, "l2": "flushed between timed steps (256 MiB write); each step also renders new sample indices",
                       "parallelism": f"spp-split x{world}"},
            "mrays_per_s": rays_all / (total_ms * 1e-3) / 1e6,
            "mrays_traced_per_s": traced_all / (total_ms * 1e-3) / 1e6,
            "rays_per_sample": rays_all / samples,
            "reduce_ms": reduce_ms,
            "stage_ms_per_step": {n: stage_ms[j] / args.steps for j, n in enumerate(STAGES) if stage_ms[j] > 0},
            "gpu_launches": int(launches_all),
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": scene_bytes, "d2h_bytes_per_step": WIDTH * HEIGHT * 3 * 4,
                    "steps": e2e_steps},
            "roofline": roofline,
            "roofline_fp32_issue": {"achieved": achieved_flops / 1e12, "peak": peak_lane_ops / 1e12, "unit": "TFLOP/s (FP32 lane-ops/s, FMA off)",
                                    "frac": achieved_flops / peak_lane_ops,
                                    "note": f"whole step: {FLOP_PER_RAY} flop/ray x reference-equivalent rays, per GPU; peak = {SM_COUNT} SMs x {LANES_PER_SM} lanes x "
                                            f"{peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} max SM clock; median under load {sm_mhz}); tensor cores unused",
                                    # issue-slot utilisation of the kernels themselves, from the ncu capture of this build
                                    "issue_active_ncu": NCU_ISSUE_ACTIVE},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference()
            if cb:
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "mrays_per_s")}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

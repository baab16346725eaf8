#!/usr/bin/env python
"""Benchmark of the Monte Carlo flux hot path (BASELINE.json metric: rays/s on
the synthetic 1M-disk trench, two particles: diffuse neutral + coned-cosine
ion).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm
    python bench.py --config c5 ...                          # 4M-disk hole array (HBM-bound)
    python bench.py --scaling strong ...                     # --rays is the whole job's count

One step = one pass of the hot path over the config's batch.  `--rays` is the
number of rays per particle of one step: per GPU with `--scaling weak` (the
default: BASELINE.json's "1e9 rays" of the 1M-disk trench for each of the two
particles on every GPU), for the whole job with `--scaling strong` (the same 1e9
rays split over the ranks by `distributed.shard_bounds`).  Ray indices of
consecutive steps and of different ranks are disjoint slices of one job.
Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 12345 + 1
ION = dict(kind=2, sticking=0.5, power=100.0, cone=float(np.deg2rad(85.0)))
NEUTRAL = dict(kind=0, sticking=0.1, power=1.0, cone=0.0)
HOLE_PARTICLE = dict(kind=0, sticking=0.2, power=100.0, cone=0.0)


def workload(name, slices):
    """The synthetic scene and trace set-up of a BASELINE.json config."""
    from viennaray_b200 import scenes
    if name == "c4":
        points, normals, gd = scenes.trench(num_slices=slices)
        text = "C4 synthetic trench, %s disks (gridDelta 1, periodic), diffuse neutral " \
               "(sticking 0.1, cosine) + coned-cosine ion (sticking 0.5, power 100)" % \
               format(len(points), ",")
        if slices != 999:
            text += " [%d slices]" % slices
        return dict(name=name, points=points, normals=normals, gd=gd, bc=[1, 1, 1],
                    particles=[NEUTRAL, ION], text=text, rays=1e9, smooth=False)
    points, normals, gd = scenes.hole_array()
    text = "C5 synthetic hole array, %s disks (gridDelta 1, 10 x 10 holes of aspect ratio 6, " \
           "reflective), diffuse particle (sticking 0.2) from a power-cosine source n = 100, " \
           "normalizeFlux + smoothFlux on the device" % format(len(points), ",")
    return dict(name=name, points=points, normals=normals, gd=gd, bc=[0, 0, 0],
                particles=[HOLE_PARTICLE], text=text, rays=4e8, smooth=True)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop = index, [], threading.Event()

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_rate(wl, target_seconds):
    """rays/s of the reference's own TraceKernel (oracle/_ref: unmodified
    reference headers + substitute intersector) -- or of the plain-C oracle
    when oracle/_ref was not built -- over the config's particles, on all host
    cores.  Two runs of different size; the rate is the slope, which removes the
    scene build that the reference's timer includes (rtcJoinCommitScene inside
    the timed region, rayTraceKernel.hpp:84-91).

    The thread count is set explicitly: a launcher such as torch.distributed.run
    exports OMP_NUM_THREADS=1 to its workers, which would silently turn "all
    host cores" into one."""
    from oracle import pyoracle as po
    want = os.cpu_count() or 1
    points, normals, gd, parts = wl["points"], wl["normals"], wl["gd"], wl["particles"]
    if po.have_ref():
        kind = "reference"
        L = po.ref_lib()
        L.ref_set_threads(want)
        cores = int(L.ref_max_threads())
        assert cores == want, "reference arm runs on %d threads, asked for %d" % (cores, want)

        def run(nrays):
            tot = 0.0
            for p in parts:
                _, _, sec = po.ref_trace_disk(3, points, normals, gd, wl["bc"], po.POS_Z,
                                              p["kind"], p["sticking"], p["power"], p["cone"],
                                              rays_fixed=nrays, seed=12345, threads=want)
                tot += sec
            return tot
    else:
        kind = "port"
        from viennaray_b200 import host
        cores = int(po.oracle_set_threads(want))
        assert cores == want, "oracle port runs on %d threads, asked for %d" % (cores, want)
        sc = po.OracleScene(3)
        r = host.disk_radius(gd, 3)
        sc.set_disks(points, normals, r)
        sc.setup(po.POS_Z, wl["bc"], r)

        def run(nrays):
            t = time.perf_counter()
            for k, p in enumerate(parts):
                sc.trace(po.Particle(p["kind"], p["sticking"], p["power"], p["cone"]),
                         sc.config(nrays, SEED, stream=k))
            return time.perf_counter() - t
    npart = len(parts)
    n1 = 200_000
    t1 = run(n1)
    t2 = run(2 * n1)
    rate = npart * n1 / max(t2 - t1, 1e-9)  # rays/s over all particles
    n3 = int(min(max(rate * target_seconds / npart, 4 * n1), 5e7))
    t3 = run(n3)
    rate = npart * (n3 - n1) / max(t3 - t1, 1e-9)
    sample = "%d + %d rays per particle (slope of two runs), %d host threads" % (n1, n3, cores)
    return rate, kind, cores, sample, n3


def traffic_capture(config_name):
    """DRAM bytes per traversal of the traverse kernel from the newest committed ncu capture
    (profiles/r*_traffic.json, stamped with the commit and command it was taken at)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if name.endswith("traffic.json"):
            with open(os.path.join(pdir, name)) as f:
                t = json.load(f)
            if t.get("config", "c4") == config_name:
                best = t
    return best


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else a library
    prints on fd 1 while the bench runs (NCCL's version banner, for one) was sent to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=["c4", "c5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--rays", type=float, default=None,
                    help="rays per particle per step: per GPU (weak) or of the whole job (strong); "
                         "default 1e9 (c4) / 4e8 (c5)")
    ap.add_argument("--warmup-rays", type=float, default=2e7)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slices", type=int, default=999, help="trench length (999 = the 1M config)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    wl = workload(args.config, args.slices)
    points, normals, gd = wl["points"], wl["normals"], wl["gd"]
    n = len(points)
    npart = len(wl["particles"])
    rays_arg = int(args.rays if args.rays is not None else wl["rays"])
    strong = args.scaling == "strong"
    step_rays = rays_arg if strong else rays_arg * world  # rays per particle of one step, all ranks
    config = {"workload": wl["text"], "disks": n, "particles": npart,
              "seed": SEED, "parallelism": "ray-sharded x%d, scene replicated" % world,
              "l2": "L2 flushed between steps (256 MiB write)"}
    if strong:
        config["rays_per_particle_per_step_whole_job"] = rays_arg
    else:
        config["rays_per_particle_per_gpu_per_step"] = rays_arg

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        per_step = []
        kind = cores = sample = None
        for k in range(args.warmup + args.steps):
            if k < args.warmup and k > 0:
                continue  # one warm-up pass is enough for a CPU loop
            t = time.perf_counter()
            rate, kind, cores, sample, _ = cpu_reference_rate(
                wl, max(args.cpu_seconds / max(args.steps, 1), 3.0))
            if k >= args.warmup:
                per_step.append((rate, time.perf_counter() - t))
        value = float(np.mean([r for r, _ in per_step]))
        line = {"impl": "reference", "metric": "rays/s", "value": value, "unit": "rays/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * float(np.mean([t for _, t in per_step])),
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": kind,
                                 "sample": sample + "; reference TraceKernel + substitute "
                                 "intersector (Embree absent)" if kind == "reference" else sample},
                "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------ CUDA arm
    import torch
    import torch.distributed as dist
    from viennaray_b200 import capi, host
    from viennaray_b200 import distributed as vdist

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    D = 3
    r = host.disk_radius(gd, D)
    xyzr = np.concatenate([points, np.full((n, 1), r, np.float32)], 1)
    glo, ghi = host.geometry_bbox(points, D)
    lo, hi = host.adjust_bbox(glo, ghi, host.POS_Z, r, D)
    _, first, second, _, _ = host.trace_settings(host.POS_Z)
    src = host.source_desc(lo, hi, host.POS_Z)
    parts = [capi.ParticleDesc(p["kind"], p["sticking"], p["power"], p["cone"])
             for p in wl["particles"]]
    # normalizeFlux(SOURCE): flux *= sourceArea / numRays / diskArea (rayTraceDisk.hpp:121-138);
    # whole-disk areas here (the clipping of boundary disks is host set-up, not the hot path)
    areas = np.full(n, np.float32(np.pi) * r * r, np.float32)
    source_area = float((hi[first] - lo[first]) * (hi[second] - lo[second]))

    def upload_and_commit(c):
        # what TraceDisk::setGeometry + apply() hand over (rayTraceDisk.hpp:63-70,19-57): points,
        # normals, radius; the neighbour lists (PointNeighborhood) and the BVH are built on the
        # device
        c.set_disks(xyzr, normals)
        c.build_neighbors_device(D, points, np.float32(2) * r)
        c.set_boundary(lo, hi, first, second, wl["bc"][first], wl["bc"][second], D)
        c.commit()

    ctx = capi.Context(local_rank)
    t0 = time.perf_counter()
    upload_and_commit(ctx)
    t_first_commit = time.perf_counter() - t0
    bvh = ctx.bvh_stats()
    total_steps = args.steps * 2 + args.warmup + 8
    total_rays_job = step_rays * total_steps
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(stream):
        flush.zero_()  # also loads torch's fill kernel outside the timed region

    def shard(step, count=None):
        """this rank's ray-index slice of step `step` (count: a shorter warm-up slice)"""
        b, e = vdist.shard_bounds(step_rays, rank, world)
        if count is not None:
            e = min(e, b + count)
        base = step * step_rays
        return host.config(total_rays_job, SEED, base + b, base + e)

    my_rays = vdist.shard_bounds(step_rays, rank, world)
    my_rays = my_rays[1] - my_rays[0]

    def all_reduce_flux(c=None):
        # the path's one exchange step: sum of the fixed-point flux words (original primitive
        # order) and counters
        if world == 1:
            return
        ptr, words = (c or ctx).flux_device()
        t = vdist.as_int64_tensor(ptr, words, torch.device("cuda", local_rank))
        with torch.cuda.stream(stream):
            vdist.all_reduce_words(t)

    # per-ray work of this workload (same kernels; counters ride in registers)
    os.environ["VR_COUNT_WORK"] = "1"
    cctx = capi.Context(local_rank)
    upload_and_commit(cctx)
    os.environ.pop("VR_COUNT_WORK")
    count_rays = 4_000_000
    cctx.trace_device(src, parts, host.config(count_rays, SEED), sync=True)
    work = cctx.work_counters()
    _, cinfo = cctx.flux_download()
    # the level an L2-resident scene's fetches come from: measured read bandwidth
    # of a 48 MB buffer, outside the timed region
    l2_gbps, l2_bytes = cctx.l2_read_bandwidth(48 << 20, 20)
    cctx.close()
    node_b, prim_b = bvh["node_bytes"], 32
    per_ray = {k: v / float(npart * count_rays) for k, v in work.items()}
    trav_bytes_per_ray = per_ray["node_visits"] * node_b + per_ray["prim_tests"] * prim_b
    bytes_per_ray = trav_bytes_per_ray + per_ray["nb_tests"] * (4 + 32) + per_ray["flux_adds"] * 8
    # what the kernels fetch from: disk records, nodes, 8-wide neighbour rows (the CSR behind
    # the rows is read for the few disks with more than eight neighbours only)
    scene_bytes = n * 32 + bvh["nodes"] * node_b + n * 32

    # warm-up
    for k in range(args.warmup):
        ctx.trace_device(src, parts, shard(k, int(args.warmup_rays)))
        all_reduce_flux()
    ctx.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    launches = 0
    ev0.record(stream)
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        ctx.trace_device(src, parts, shard(args.warmup + k))
        all_reduce_flux()
        launches += ctx.last_launch_count()[0]
    ev1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.stop.set()
    sampler.join()
    elapsed_ms = ev0.elapsed_time(ev1)
    _ = ctx.flux_download()
    if world > 1:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    rays_total = float(npart) * step_rays * args.steps
    value = rays_total / (elapsed_ms * 1e-3)

    # the same steps once more with CUDA events around every kernel launch (on the
    # launching stream), for the duration of the dominant kernel
    phase = None
    if rank == 0:
        ctx.phase_timing(True)
        for k in range(args.steps):
            with torch.cuda.stream(stream):
                flush.zero_()
            ctx.trace_device(src, parts, shard(args.warmup + args.steps + k))
        phase = ctx.phase_ms()
        ctx.phase_timing(False)
    if world > 1:
        dist.barrier()

    # ---- end to end through the host C ABI, host buffers in, host flux out: points + normals
    # up, neighbour lists and BVH built on the device, trace, (all-reduce,) flux down; for C5
    # normalizeFlux + smoothFlux on the device before the download.
    def e2e_step(c, step):
        t = time.perf_counter()
        upload_and_commit(c)
        c.trace_device(src, parts, shard(step))
        all_reduce_flux(c)
        if wl["smooth"]:
            out = c.flux_postprocess(0, areas, source_area / step_rays, True)
        else:
            out, _ = c.flux_download()
        return time.perf_counter() - t, out

    base_step = args.warmup + 2 * args.steps
    # cold: a fresh context (ray pools allocated, Morton cell shape searched: three BVH builds)
    fresh = capi.Context(local_rank)
    t_cold, flux = e2e_step(fresh, base_step)
    fresh.close()
    # warm: the time-stepping case (a context that has traced this scene before: pools
    # exist, the Morton cell shape of the last search is reused, one BVH build)
    t_warm = [e2e_step(ctx, base_step + 1 + k)[0] for k in range(max(1, min(args.steps, 3)))]

    def over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e2e_warm_s = over_ranks(float(np.mean(t_warm)))
    e2e_cold_s = over_ranks(t_cold)
    e2e_value = float(npart) * step_rays / e2e_warm_s
    h2d = xyzr.nbytes + normals.nbytes + points.nbytes + (areas.nbytes if wl["smooth"] else 0)
    d2h = flux.nbytes + (0 if wl["smooth"] else npart * 72)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, hbm_src = peaks()
    # roofline of the dominant kernel (traverseKernel): algorithmic node + primitive
    # bytes it fetched during the instrumented steps / its summed launch durations
    trav_ms = phase["traverse_ms"]
    trav_launches = max(phase["traverse_launches"], 1)
    trav_bytes = trav_bytes_per_ray * float(npart) * my_rays * args.steps
    achieved = trav_bytes / (trav_ms * 1e-3) / 1e9
    traces_per_ray = sum(i.totalRaysTraced for i in cinfo) / float(npart * count_rays)
    traversals_per_launch = (traces_per_ray - per_ray["sky_finished"]) * float(npart) * my_rays \
        * args.steps / trav_launches
    # The level the traversal's bytes come from: a scene that fits the L2 cache is served from
    # L2 (ncu: DRAM at a few per cent of peak), a larger one from HBM.
    l2_resident = scene_bytes < l2_bytes
    bound = "l2" if l2_resident else "hbm"
    peak = l2_gbps if l2_resident else hbm_peak
    peak_src = ("measured in this run: vr_debug_l2_read_bandwidth, a 48 MB buffer streamed 20 "
                "times with 16-byte ld.global.cg loads") if l2_resident else hbm_src
    traffic = traffic_capture(args.config)
    step_gbps = value / world * bytes_per_ray / 1e9
    clocks = sampler.summary()
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    lsu_peak = sms * (clocks["sm_mhz"] or 1965.0) * 1e6
    fetches = (per_ray["node_visits"] + per_ray["prim_tests"]) * float(npart) * my_rays * args.steps
    lsu = {"unit": "L1 data-stage wavefronts/s", "peak": lsu_peak,
           "peak_source": "%d SMs x %.0f MHz (sampled under load), one wavefront per cycle per SM"
                          % (sms, clocks["sm_mhz"] or 1965.0),
           "achieved": fetches / (trav_ms * 1e-3), "frac": fetches / (trav_ms * 1e-3) / lsu_peak}
    if traffic and traffic.get("traverse_lsu_wavefronts_per_traversal"):
        w = traffic["traverse_lsu_wavefronts_per_traversal"]
        rate = traversals_per_launch * trav_launches / (trav_ms * 1e-3)
        lsu["measured_wavefronts_per_traversal"] = w
        lsu["measured_frac"] = w * rate / lsu_peak
    line = {
        "metric": "rays/s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "seconds_per_step": e2e_warm_s,
                "cold_value": float(npart) * step_rays / e2e_cold_s, "cold_seconds": e2e_cold_s,
                "what": "per step, host buffers in and out: vr_scene_set_disks (points, normals) + "
                        "vr_scene_build_neighbors + vr_scene_commit (device BVH) + vr_trace_device"
                        + (" + NCCL all-reduce" if world > 1 else "")
                        + (" + vr_flux_postprocess (normalise, smooth) with the flux read back"
                           if wl["smooth"] else " + vr_flux_download")
                        + ". `value`: a context that has committed this scene before (time "
                        "stepping: ray pools exist, the Morton cell shape of its last search is "
                        "reused, one BVH build); `cold_value`: the first step of a fresh context "
                        "(pool allocation, three BVH builds)"},
        "gpu_launches": launches,
        "lanes": ("the particles of a trace (or the two halves of a lone particle's ray range) run "
                  "two at a time on two streams (VR_LANES=2); the per-kernel durations under "
                  "`roofline` come from a second pass over the same steps with one lane and CUDA "
                  "events around every launch"),
        "roofline": {
            "kernel": "traverseKernel<0,0,0,0>", "bound": bound, "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak if peak else None,
            "peak_source": peak_src,
            "hbm": {"peak": hbm_peak, "peak_source": hbm_src, "frac": achieved / hbm_peak},
            "l2": {"size_bytes": l2_bytes, "read_gbps": l2_gbps,
                   "frac_of_l2_read": achieved / l2_gbps if l2_gbps else None},
            # The unit that saturates first: every lane's 32-byte node / disk
            # fetch is one wavefront of the L1 data stage, which passes one per cycle per SM.
            # achieved: counted node visits + primitive tests per second of traverse-kernel
            # time; measured: ncu l1tex__data_pipe_lsu_wavefronts of the captured launch per
            # traversal (pool and stack accesses included) x this run's traversal rate
            "lsu": lsu,
            "scene_bytes": scene_bytes,
            # DRAM bytes of an average launch: ncu's dram__bytes_{read,write}.sum per traversed
            # slot of one captured launch (profiles/r*_traffic.json) x traversals per launch
            "traffic": (traffic["traverse_dram_bytes_per_traversal"] * traversals_per_launch
                        if traffic else None),
            "traffic_source": ("%s (commit %s)" % (traffic["source"], traffic.get("commit", "?"))
                               if traffic else None),
            "algorithmic_bytes_per_launch": trav_bytes / trav_launches,
            "avg_launch_ms": trav_ms / trav_launches, "launches": trav_launches,
            "kernel_share_of_step": trav_ms / max(trav_ms + phase["shade_ms"] + phase["other_ms"],
                                                  1e-9),
            "shade_ms_per_step": phase["shade_ms"] / args.steps,
            "traverse_ms_per_step": trav_ms / args.steps,
            "other_ms_per_step": phase["other_ms"] / args.steps,
            "step": {"bytes_per_ray": bytes_per_ray, "achieved": step_gbps,
                     "frac": step_gbps / peak if peak else None,
                     "frac_of_hbm": step_gbps / hbm_peak},
            "per_ray": per_ray,
            "note": "algorithmic bytes = counted node visits x %d B + primitive tests x 32 B for "
                    "the traverse kernel (+ neighbour tests x 36 B + flux adds x 8 B for the whole "
                    "step), all four counted by the same kernels on this workload. `bound` / "
                    "`peak`: the memory level the scene is served from -- L2 when it fits the L2 "
                    "cache (measured read bandwidth of this run), HBM otherwise "
                    "(MEASURED_PEAKS.json); `hbm.frac` is the same rate over the HBM peak" % node_b},
        "clocks": clocks,
        "bvh": bvh, "first_commit_s": t_first_commit,
        "walk": {"traces_per_ray": [i.totalRaysTraced / count_rays for i in cinfo],
                 "geo_hits_per_ray": [i.geometryHits / count_rays for i in cinfo]},
    }
    if not args.no_cpu_baseline:
        rate, kind, cores, sample, _ = cpu_reference_rate(wl, args.cpu_seconds)
        line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": cores, "kind": kind,
                                "sample": sample}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

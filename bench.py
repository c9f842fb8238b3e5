#!/usr/bin/env python
"""bench.py -- closest-hit Mrays/s on BASELINE config 3 (sponza.obj 1920x1080 incoherent diffuse-bounce path
tracing; sponza.obj is absent from the reference tree, so the declared stand-in of SURVEY 8d is used: the
reference's makeBunny20Scene geometry, 1 389 021 triangles, with its own camera).

    python bench.py --gpus N --steps K --warmup W            # this engine, one rank per GPU (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

A step is one pass of the hot path over one batch of synthetic rays: SPP jittered samples of the frame --
device-side Camera::eyeRay -> closest-hit -> device-side Ray::diffuse at every hit -> closest-hit.
`value` = (primary + live bounce rays of all ranks) / max-over-ranks device time, inputs resident in HBM; the timed region keeps
          three steps in flight (consecutive steps go round-robin over three streams, each with its own ray / hit buffers, all K
          steps finished inside the timed region), and the same steps are replayed one launch after the other right afterwards
          for the per-kernel durations and the roofline (--sequential times that schedule instead; MIRO_BENCH_INFLIGHT=0 the
          round-2 schedule of two half-batches per step on two streams).
`e2e`   = the same metric through the reference-facing call with HOST buffers: Scene::raytraceImage ->
          mirogpu_render (diffuse-bounce mode, no shadow rays so the ray work equals a step's), framebuffer
          gathered over NCCL/NVLink for N > 1 (two frames in flight) and copied to pinned host memory, every step.
Image rows are interleaved across ranks (row % N == rank); the BVH is replicated; fixed frame => strong scaling.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # objio (fixture -> .obj) and, for the CPU legs only, miro_driver

METRIC = "closest-hit Mrays/s (sponza, primary+diffuse bounce)"
SCENE = "bunny20"
WIDTH, HEIGHT = 1920, 1080
SPP = 16   # samples per step: sized so that one rank's share of a step is still milliseconds of device work at N = 8
SEED = 168
# SURVEY 8d yardstick: bytes/ray = 32*V + 36*T + 48 with V, T the scalar reference's node entries / triangle
# tests per ray.  Measured live with the oracle on a subsample of the step's own rays; when the oracle is not run
# (--no-cpu, or the oracle library missing) the figures it measured on this very workload are used (profiles/
# r02g_bench_full.json: the bench's rays hit mostly the floor triangle; SURVEY 8d's 18.48 / 22.43 node entries are for a
# 512 x 512 view and would overstate the roofline fraction 1.35x).
FALLBACK_VT = {"primary": (13.489, 2.515), "bounce": (16.234, 3.508)}


NTRIS = 1389021   # makeBunny20Scene: 20 x 69 451 bunny triangles + the floor triangle (checked against the built scene)


def workload(ntris=NTRIS):
    return (f"{SCENE}: the reference's makeBunny20Scene geometry ({ntris} triangles, bit-identical to the script's: tests/test_scene_scripts.py) and camera, "
            f"declared stand-in for BASELINE config 3's sponza.obj (absent from the reference tree); {WIDTH}x{HEIGHT}, {SPP} jittered samples per step, "
            f"one Ray::diffuse bounce ray per hit")


def shared_config():
    """The `config` both arms print, key for key: one workload.  What differs per arm (how much of a step a CPU step samples,
    kernel and layout choices, measured figures) goes under `detail`."""
    return {"workload": workload(), "scene": SCENE, "triangles": NTRIS, "width": WIDTH, "height": HEIGHT, "samples_per_step": SPP,
            "rays": "jittered Camera::eyeRay -> closest hit -> one Ray::diffuse per hit -> closest hit", "seed": SEED,
            "l2": "inputs larger than L2: each step writes and re-reads its ray / hit buffers (3.2 GB per step at N = 1, 96 B per ray) between "
                  "kernels, far beyond the 126 MB L2; no explicit flush"}


def load_scenes():
    """scenes.py by path: the CPU arms must not import the package (which maps libmirogpu.so)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_miro_scenes", os.path.join(ROOT, "cse168-raytracer_b200", "scenes.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL's version banner, the reference's progress
    lines), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def bytes_per_ray(V, T):
    return 32.0 * V + 36.0 * T + 48.0


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.path = tempfile.mktemp(prefix="miro_clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def build_host_scene(pkg, scenes, layout, builder=None):
    import objio
    H = pkg.HostScene(layout) if builder is None else pkg.HostScene(layout, builder=builder)
    scenes.realise(H, SCENE, objio.obj_path)
    H.precalc()
    return H


# ------------------------------------------------------------------------------------------------------------------
def _quiet():
    """Context: file descriptor 1 -> /dev/null (the reference prints progress to stdout)."""
    class Q:
        def __enter__(self):
            self.devnull = os.open(os.devnull, os.O_WRONLY); self.saved = os.dup(1); os.dup2(self.devnull, 1)
        def __exit__(self, *a):
            os.dup2(self.saved, 1); os.close(self.devnull); os.close(self.saved)
    return Q()


def measure_vt(rays_p, rays_b, threads):
    """V (node entries) and T (triangle tests) per ray of the SCALAR REFERENCE on these rays: the oracle restatement's counters
    (bit-exact with the reference's -DSTATS build, tests/test_oracle_vs_reference.py)."""
    import miro_driver as md
    import objio
    scenes = load_scenes()
    O = md.oracle()
    with _quiet():
        scenes.realise(O, SCENE, objio.obj_path)
        O.precalc()
        vt = {}
        for name, r in (("primary", rays_p), ("bounce", rays_b)):
            O.stats_reset_rays()
            O.trace(r, threads)
            st = O.stats()
            vt[name] = (st["ray_box"] / max(1, r.shape[0]), st["ray_tri"] / max(1, r.shape[0]))
        st = O.stats()
    return vt, (st["nodes"], st["leaves"])


def cpu_leg(rays_p, rays_b, kind_pref="reference"):
    """Times the reference's CPU path (oracle/_ref; else the oracle port) on the given rays with all host threads and KEEPS its
    answers (t, prim id) for the parity block."""
    import miro_driver as md
    import objio
    scenes = load_scenes()
    cores = os.cpu_count() or 1
    out = {}
    kind = "reference" if (kind_pref == "reference" and os.path.exists(md.REF_SO)) else "port"
    D = md.reference("scalar") if kind == "reference" else md.oracle()
    rays = np.concatenate([rays_p, rays_b])
    with _quiet():
        scenes.realise(D, SCENE, objio.obj_path)
        build_s = D.precalc()
        D.trace_time(rays[:100000], cores)   # warm the caches / thread pool
        secs, t, ids = D.trace_time_hits(rays, cores)
    out["hits"] = (t, ids)
    out["cpu"] = dict(value=rays.shape[0] / secs / 1e6, unit="Mrays/s", cores=cores, kind=kind,
                      sample=f"{rays_p.shape[0]} primary + {rays_b.shape[0]} live bounce rays = one whole step ({SPP} samples of {WIDTH}x{HEIGHT}, the GPU's own rays), "
                             f"Scene::trace over OpenMP dynamic chunks of 4096 on {cores} threads, (t, object) kept per ray; reference BVH build {build_s:.1f} s not timed",
                      seconds=secs)
    # also the reference's shipped configuration (-msse4.1: 4-wide box tests with _mm_rcp_ps, triangle packets; not
    # parity grade) as a second throughput baseline on the same rays
    if kind == "reference" and os.path.exists(md.REF_SSE_SO):
        E = md.reference("sse")
        with _quiet():
            scenes.realise(E, SCENE, objio.obj_path)
            E.precalc()
            E.trace_time(rays[:100000], cores)
            secs_sse, _ = E.trace_time(rays, cores)
        out["cpu"]["sse_build"] = dict(value=rays.shape[0] / secs_sse / 1e6, unit="Mrays/s", seconds=secs_sse,
                                       note="the reference compiled with -msse4.1 (Makedefs:14-15), same rays, same threads")
    return out


def parity_block(gpu_hits, ref_t, ref_id):
    """GPU closest hits against the reference's Scene::trace on the same rays (north_star: ids exact except <= 1e-5 of rays in the
    documented classes, t within 1e-5 relative)."""
    n = int(ref_id.shape[0])
    gid = gpu_hits[:, 1].copy().view(np.int32)          # MIROGPU_MISS = 0xFFFFFFFF reads as -1, the reference's miss id
    gt = gpu_hits[:, 0]
    mism = gid != ref_id
    both = (~mism) & (ref_id >= 0)
    rel = np.abs(gt[both].astype(np.float64) - ref_t[both]) / np.maximum(np.abs(ref_t[both].astype(np.float64)), 1e-30)
    bit_same = gt.view(np.uint32) == ref_t.view(np.uint32)
    mi = np.flatnonzero(mism)
    ties = int(np.count_nonzero(gt[mi] == ref_t[mi]))
    gpu_closer = int(np.count_nonzero(gt[mi] < ref_t[mi]))      # hits the reference's own tree culls (its boxes lack the slop band)
    gpu_farther = int(np.count_nonzero(gt[mi] > ref_t[mi]))     # would be a hit the GPU lost: must be 0
    nontie = int(mi.shape[0]) - ties
    return {"rays": n, "against": "the reference's Scene::trace (oracle/_ref, scalar build) on the identical rays of the last timed step",
            "id_mismatches": int(mi.shape[0]), "id_mismatch_frac": float(mi.shape[0]) / max(1, n),
            "mismatch_classes": {"equal_t_tie": ties, "gpu_closer_reference_culled": gpu_closer, "gpu_farther": gpu_farther},
            "equal_t_tie_frac": ties / max(1, n), "id_mismatch_frac_excluding_equal_t_ties": nontie / max(1, n),
            "max_rel_t": float(rel.max()) if rel.size else 0.0, "t_bit_identical_frac": float(np.count_nonzero(bit_same)) / max(1, n),
            "gate": "gpu_farther == 0 (no hit lost); max_rel_t <= 1e-5; ids differing at a different t (reference false culls) <= 1e-5 of rays; "
                    "equal-t ties (two edge-sharing triangles accept the ray in the reference's epsilon band at bit-identical t; the reference "
                    "keeps whichever its own tree visits first, this engine the smaller id -- DESIGN.md section 4) <= 5e-5 of rays",
            "north_star_1e-5_all_classes": bool(mi.shape[0] <= 1e-5 * n)}


def parity_ok(p):
    c = p["mismatch_classes"]
    return (c["gpu_farther"] == 0 and p["max_rel_t"] <= 1e-5 and p["id_mismatch_frac_excluding_equal_t_ties"] <= 1e-5 and p["equal_t_tie_frac"] <= 5e-5)


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores (rank 0 only).  A step of this arm is ONE
    of the workload's 16 samples of the frame (2 073 600 jittered camera rays + a Ray::diffuse bounce ray per hit), so that
    --steps 20 --warmup 5 ends within minutes; the rate is per ray, so it compares directly."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ctypes
    import miro_driver as md
    import objio
    scenes = load_scenes()
    cores = os.cpu_count() or 1
    kind = "reference" if os.path.exists(md.REF_SO) else "port"
    D = md.reference("scalar") if kind == "reference" else md.oracle()
    O = md.oracle()
    threads = D.host_threads()
    with _quiet():
        scenes.realise(D, SCENE, objio.obj_path)
        build_s = D.precalc()
        scenes.realise(O, SCENE, objio.obj_path)   # camera + Ray::diffuse restatement for ray generation (no BVH build)
        rng = np.random.default_rng(SEED)
        times, nrays = [], []
        D.trace_time(np.zeros((100000, 8), np.float32), threads)   # thread pool up
        for it in range(args.warmup + args.steps):
            jit = rng.random((WIDTH * HEIGHT, 2), dtype=np.float32)
            rays = np.zeros((WIDTH * HEIGHT, 8), np.float32)
            O.lib.orc_eye_rays_jitter(WIDTH, HEIGHT, md._fp(jit), md._fp(rays))
            s_primary, _ = D.trace_time(rays, threads)     # timed: Scene::trace over all host threads, no output traffic
            t, ids, P, N = D.trace(rays, threads)          # untimed: the hits, to generate the bounce rays from
            u = rng.random((rays.shape[0], 2), dtype=np.float32)
            br = np.zeros_like(rays)
            O.lib.orc_diffuse_rays(md._fp(P), md._fp(N), md._fp(ids), md._fp(u), ctypes.c_long(rays.shape[0]), md._fp(br))
            br = np.ascontiguousarray(br[ids >= 0])
            s_bounce, _ = D.trace_time(br, threads)
            if it >= args.warmup:
                times.append(s_primary + s_bounce); nrays.append(rays.shape[0] + br.shape[0])
    total_t, total_r = float(np.sum(times)), int(np.sum(nrays))
    value = total_r / total_t / 1e6
    sample = (f"{total_r // max(1, args.steps)} rays per step = one of the workload's {SPP} samples of the {WIDTH}x{HEIGHT} frame + its bounce rays, "
              f"Scene::trace per ray, OpenMP dynamic chunks of 4096 on {threads} host threads ({cores} logical CPUs)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(),
        "detail": {"sample": sample, "rays_per_step": total_r // max(1, args.steps), "threads": threads, "bvh_build_s": build_s,
                   "library": "oracle/_ref/libmiro_ref.so (the unmodified reference, scalar build)" if kind == "reference" else "oracle/libmiro_oracle.so (port)"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback (use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("cse168-raytracer_b200")
    scenes = importlib.import_module("cse168-raytracer_b200.scenes")
    layout = {"bvh2": pkg.LAYOUT_BVH2, "cwbvh8": pkg.LAYOUT_CWBVH8, "bvh4": pkg.LAYOUT_BVH4, "qbvh4": pkg.LAYOUT_QBVH4}[args.layout]
    builder = {"sah": None, "lbvh": pkg.BUILDER_LBVH_DEVICE, "ploc": pkg.BUILDER_PLOC_DEVICE}[args.builder]
    H = build_host_scene(pkg, scenes, layout, builder)       # BVH replicated on every GPU
    S = H.scene()
    S.set_kernel_variant(args.variant)
    cam = H.camera()
    info = S.info
    rows = (0, HEIGHT, world, rank)
    nrows = len(range(rank, HEIGHT, world))
    if args.emulate_shard > 1 and world == 1:
        rows = (0, HEIGHT, args.emulate_shard, 0)
        nrows = len(range(0, HEIGHT, args.emulate_shard))
    npix = nrows * WIDTH
    n = npix * SPP
    dev = torch.device("cuda", local)
    d_rays = torch.empty((n, 8), dtype=torch.float32, device=dev)
    d_hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
    d_b = torch.empty((n, 8), dtype=torch.float32, device=dev)
    d_h2 = torch.empty((n, 4), dtype=torch.float32, device=dev)
    d_live = torch.zeros(1, dtype=torch.int64, device=dev)
    index_base = (rank * 0x01000000) & 0xFFFFFFFF

    def step_sequential(it, ev=None):
        if ev: ev[4].record()
        S.generate_primary(cam, WIDTH, HEIGHT, d_rays, rows=rows, jitter=1, seed=SEED, sample=it * SPP, samples=SPP)
        if ev: ev[0].record()
        S.intersect_device(d_rays, d_hits, mode=pkg.CLOSEST_HIT | pkg.HINT_COHERENT)
        if ev: ev[1].record()
        S.generate_bounce(d_rays, d_hits, d_b, seed=SEED, sample=it, index_base=index_base, d_live_count=d_live)
        if ev: ev[2].record()
        S.intersect_device(d_b, d_h2)
        if ev: ev[3].record()

    # The timed step: the same work as step_sequential -- same rays, same buffers, same results -- issued as two half-batches
    # (samples 0-7 and 8-15 of the step) on two streams.  Each half is its own dependent chain (eye rays -> trace -> bounce rays
    # -> trace); the persistent trace kernels fill the GPU, so the other stream's kernel gets its CTAs exactly while the first
    # one's last warps finish their longest rays: the ragged end of every launch (~0.1 ms, a quarter of a rank's launch at
    # N = 8) is covered by the other half's work instead of idling the chip.
    chunks = int(os.environ.get("MIRO_BENCH_CHUNKS", "2"))
    nstreams = int(os.environ.get("MIRO_BENCH_STREAMS", "2"))
    half_spp = SPP // chunks
    nh = npix * half_spp
    sides = [torch.cuda.Stream(dev) for _ in range(nstreams - 1)]

    def half_step(it, h):
        lo, hi = h * nh, (h + 1) * nh
        S.generate_primary(cam, WIDTH, HEIGHT, d_rays[lo:hi], rows=rows, jitter=1, seed=SEED, sample=it * SPP + h * half_spp, samples=half_spp)
        S.intersect_device(d_rays[lo:hi], d_hits[lo:hi], mode=pkg.CLOSEST_HIT | pkg.HINT_COHERENT)
        S.generate_bounce(d_rays[lo:hi], d_hits[lo:hi], d_b[lo:hi], seed=SEED, sample=it, index_base=(index_base + lo) & 0xFFFFFFFF, d_live_count=d_live)
        S.intersect_device(d_b[lo:hi], d_h2[lo:hi])

    def step(it):
        for c in range(chunks):
            k = c % nstreams
            if k == 0:
                half_step(it, c)
            else:
                with torch.cuda.stream(sides[k - 1]):
                    half_step(it, c)

    # Default schedule (MIRO_BENCH_INFLIGHT = k, default 3; 0 or 1 = the half-batch schedule above): whole steps go round-robin over
    # k streams, each with its own ray / hit buffers, so consecutive steps cover each other's ragged ends with launches twice the
    # size of the half-batches' (a persistent launch pays a fixed ramp and tail whatever its size).  Measured on one B200 for one
    # rank's share of an N-rank run (--emulate-shard N; profiles/r02v_inflight.txt): N = 8: 7.44 Grays/s with half-batches, 8.16 with
    # two steps in flight, 8.35 with three; N = 4: 8.13 -> 8.62 (two); N = 1: 8.97 -> 9.11 -> 9.13.
    inflight = int(os.environ.get("MIRO_BENCH_INFLIGHT", "3"))
    if inflight > 1 and not args.sequential:
        bufs = [(d_rays, d_hits, d_b, d_h2)] + [tuple(torch.empty_like(t) for t in (d_rays, d_hits, d_b, d_h2)) for _ in range(inflight - 1)]
        while len(sides) < inflight - 1:
            sides.append(torch.cuda.Stream(dev))

        def whole_step(it, B):
            r, h, b, h2 = B
            S.generate_primary(cam, WIDTH, HEIGHT, r, rows=rows, jitter=1, seed=SEED, sample=it * SPP, samples=SPP)
            S.intersect_device(r, h, mode=pkg.CLOSEST_HIT | pkg.HINT_COHERENT)
            S.generate_bounce(r, h, b, seed=SEED, sample=it, index_base=index_base, d_live_count=d_live)
            S.intersect_device(b, h2)

        def step(it):
            k = it % inflight
            if k == 0:
                whole_step(it, bufs[0])
            else:
                with torch.cuda.stream(sides[k - 1]):
                    whole_step(it, bufs[k])

    def join_streams():
        for side in sides:
            done = torch.cuda.Event()
            done.record(side)
            torch.cuda.current_stream().wait_event(done)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.sequential:
        step = step_sequential
    for it in range(args.warmup):
        step(it)
    barrier()
    d_live.zero_()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record()
    for side in sides:
        side.wait_event(t_begin)
    for it in range(args.steps):
        step(args.warmup + it)
    join_streams()
    t_end.record()
    barrier()
    ms_local = t_begin.elapsed_time(t_end)
    live_local = int(d_live.item())
    # Per-kernel durations (and with them roofline.achieved): the same steps once more, one launch after the other on one stream,
    # CUDA events around each kernel -- under two streams an event interval would include the other stream's kernel.
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    t_sb, t_se = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(min(3, args.warmup)):
        step_sequential(it)
    barrier()
    t_sb.record()
    for it in range(args.steps):
        step_sequential(args.warmup + it, evs[it])
    t_se.record()
    barrier()
    ms_sequential = t_sb.elapsed_time(t_se)
    clocks = None
    prim_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    bounce_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
    genb_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    genp_ms = float(np.mean([e[4].elapsed_time(e[0]) for e in evs]))

    # ---- e2e: Scene::raytraceImage-shaped call, host framebuffer out, every step --------------------------------
    e2e_steps = args.steps
    p = S.render_params(WIDTH, HEIGHT, spp=SPP, jitter=1, max_depth=10, mode=pkg.RENDER_DIFFUSE_BOUNCE, seed=SEED, tonemap=0,
                        rows=rows, shadows=0)
    host_fb = torch.empty((HEIGHT, WIDTH, 3), dtype=torch.uint8).pin_memory()   # the reference's Image: 3 bytes per pixel
    e2e_rays = 0
    e2e_call = "mirogpu_render_rgb8 (Scene::raytraceImage -> 8-bit Image), diffuse-bounce mode, pinned host framebuffer; rays = primary + LIVE bounce rays"
    # development: MIRO_BENCH_FORCE_PIPE=1 takes the multi-rank e2e path (frames in flight, NCCL exchange) on a single rank
    piped = world > 1 or os.environ.get("MIRO_BENCH_FORCE_PIPE") == "1"
    if piped and world == 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local))
    if piped:
        e2e_call = ("mirogpu_render_device (Scene::raytraceImage of this rank's rows, diffuse-bounce mode, float radiance in HBM; rays = primary + LIVE bounce rays)"
                    " per rank -> mirogpu_frame_max_device, all_reduce(max) of the tone-map constant, mirogpu_tonemap_rows_rgb8_device, NCCL all_gather of the 8-bit row shards, rank 0 copies the frame to pinned host memory; "
                     "frames in flight (the exchange of frame i overlaps the render of frame i + 1; the host queues the timed frames without waiting on any -- "
                     "their live-ray counts, a function of the seed alone, are read from an untimed pass over the same seeds -- and all frames are delivered inside the timed region)")
    if not piped:
        def e2e_step(it):
            p.seed = SEED + it
            S.render_rgb8(cam, p, out=host_fb.numpy())
            return S.last_call_stats()[0]
    else:
        sharding = importlib.import_module("cse168-raytracer_b200.sharding")
        # frames in flight (sharding.FramePipeline): rank r renders frame i + 1 while frame i's exchange -- all_reduce(max) of the
        # tone map's frame-wide constant (Scene.cpp:157-202), 8-bit rows, NCCL all_gather, rank 0's copy to the host -- runs on a
        # side stream; every frame is still delivered to host memory inside the timed region (drained before the clock stops)
        # MIRO_BENCH_E2E_HANDLES = k (default 3): k handles of the scene per rank (0.25 GB each), k consecutive frames rendering
        # concurrently on k streams -- a handle renders one frame at a time
        n_handles = max(1, int(os.environ.get("MIRO_BENCH_E2E_HANDLES", "3")))
        if n_handles > 1:
            # with several frames in flight the library's own split of a frame into two half-batches on two streams only makes the
            # launches smaller (one rank's share of an N = 8 run on one GPU: 1.229 -> 1.177 ms per frame without it; read once, at
            # the process's first render, which is the one below)
            os.environ.setdefault("MIROGPU_RENDER_STREAMS", "1")
        replicas = [scenes.handle_replica(pkg, H, SCENE, layout) for _ in range(n_handles - 1)]
        for Sx in replicas:
            Sx.set_kernel_variant(args.variant)
        pipe = sharding.FramePipeline(S, HEIGHT, WIDTH, world, rank, dev, replicas=replicas)
        e2e_call += f"; {n_handles} handles of the scene per rank, {n_handles} consecutive frames rendering concurrently, each as whole 16-sample batches"

        # Untimed pass over the seeds of the timed frames: rays traced per frame (primary + LIVE bounce rays -- the device counter
        # mirogpu_last_call_stats reads is valid once the frame's render has finished, so reading it costs a host wait per frame;
        # the counts depend on the seed alone).  The timed loop then queues its frames without waiting on any of them.
        rays_of_seed = {}

        def e2e_step(it):
            p.seed = SEED + it
            if it in rays_of_seed:
                pipe.submit(cam, p)
                return rays_of_seed[it]
            rays_of_seed[it] = pipe.rays_traced(pipe.submit(cam, p))
            return rays_of_seed[it]
    for it in range(min(3, args.warmup)):
        e2e_step(it)
    if piped:
        for it in range(e2e_steps):
            e2e_step(100 + it)
        pipe.drain()
    barrier()
    t0 = time.perf_counter()
    for it in range(e2e_steps):
        e2e_rays += e2e_step(100 + it)
    if piped:
        pipe.drain()
    barrier()
    e2e_s_local = time.perf_counter() - t0
    exchange_ms = pipe.exchange_ms(last=e2e_steps) if piped else 0.0
    if rank == 0:
        clocks = sampler.stop()

    # ---- aggregate over ranks: max time, summed rays ------------------------------------------------------------
    stats = torch.tensor([ms_local, e2e_s_local, prim_ms, bounce_ms, genb_ms, genp_ms, ms_sequential, exchange_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(n * args.steps), float(live_local), float(e2e_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_total, e2e_s, prim_ms, bounce_ms, genb_ms, genp_ms, ms_sequential, exchange_ms = [float(x) for x in stats.tolist()]
    prim_total, live_total, e2e_total = [float(x) for x in sums.tolist()]
    rays_total = prim_total + live_total
    value = rays_total / (ms_total * 1e-3) / 1e6
    e2e_value = e2e_total / e2e_s / 1e6

    line = None
    rc = 0
    if rank == 0:
        peak, peak_src = peaks()
        cpu, vt, ref_nodes, parity = None, dict(FALLBACK_VT), None, None
        vt_src = "oracle not run: V, T as measured on this workload in profiles/r02g_bench_full.json"
        threads = os.cpu_count() or 1
        # the rays of the last timed step, exactly as the GPU generated them (all SPP samples of this rank's rows)
        rp = d_rays.cpu().numpy()
        rb_all = d_b.cpu().numpy()
        live = rb_all[:, 7] >= rb_all[:, 3]
        rb = np.ascontiguousarray(rb_all[live])
        del rb_all
        if not args.no_cpu:
            try:
                # V, T per ray: the scalar reference's counters on every 16th of this rank's rays -- at every N, so that the
                # roofline fractions of a scaling run are comparable
                vt, ref_nodes = measure_vt(rp[::16], rb[::16], threads)
                vt_src = "scalar reference counters (oracle restatement, bit-exact with -DSTATS) on every 16th of rank 0's rays of the last timed step"
            except Exception as exc:
                vt_src = "oracle failed (%r): V, T as measured on this workload in profiles/r02g_bench_full.json" % (exc,)
        if world == 1 and not args.no_cpu:
            try:
                leg = cpu_leg(rp, rb)
                cpu = leg["cpu"]
                ref_t, ref_id = leg["hits"]
                gh = np.concatenate([d_hits.cpu().numpy(), d_h2.cpu().numpy()[live]])
                parity = parity_block(gh, ref_t, ref_id)
                if not parity_ok(parity):
                    rc = 3
                    sys.stderr.write("bench.py: PARITY GATE FAILED: %s\n" % json.dumps(parity))
            except Exception as exc:   # the CPU leg is a reported baseline, never a reason to lose the GPU number
                cpu = {"error": repr(exc)}
        bpr_p, bpr_b = bytes_per_ray(*vt["primary"]), bytes_per_ray(*vt["bounce"])
        live_per_launch = live_total / max(1, args.steps) / world
        achieved = live_per_launch * bpr_b / (bounce_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
                js = json.load(f)
                traffic, traffic_src = js.get("bounce_trace_dram_bytes_per_launch"), js.get("source")
        except Exception:
            pass
        kernel = ("k_trace_hybrid" if (args.layout in ("bvh2", "bvh4", "qbvh4") and args.variant in (-1, 2)) else
                  "k_trace_simple" if args.variant == 1 else "k_trace_persistent")
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": shared_config(),
            "detail": {
                "rays_per_step": rays_total / args.steps, "primary_rays_per_step": prim_total / args.steps, "bounce_rays_per_step": live_total / args.steps,
                "layout": args.layout, "builder": args.builder, "kernel_variant": args.variant, "nodes": info.num_nodes, "node_mb": info.node_bytes / 1e6,
                "triangle_mb": info.triangle_bytes / 1e6, "build_s": info.build_seconds + info.flatten_seconds,
                "sharding": f"image rows interleaved over {world} rank(s), BVH replicated",
                "schedule": ("one launch after the other on one stream (--sequential)" if args.sequential else
                             f"{inflight} steps in flight: consecutive steps round-robin over {inflight} streams, each with its own ray / hit buffers, "
                             "so each persistent launch's ragged end overlaps the neighbouring steps' kernels; all steps end inside the timed region"
                             if inflight > 1 else
                             "two half-batches (8 samples each) per step on two streams: each launch's ragged end overlaps the other half's kernels"),
                "ms_per_step_sequential": ms_sequential / args.steps,
                "per_gpu_ms": {"primary_trace": prim_ms, "gen_bounce": genb_ms, "bounce_trace": bounce_ms, "gen_primary": genp_ms,
                               "e2e_frame_exchange_elapsed": exchange_ms,   # N > 1: elapsed device time of a frame's exchange on the side stream, first to last operation -- it runs underneath the following frames' persistent trace kernels and waits for SM slots between them, so this is a latency (what the slots of the frame pipeline have to cover), not a cost added to the frame
                               "note": "max over ranks of each rank's mean over a sequential replay of the timed steps right after the timed region "
                                       "(same rays, whole-step launches one after the other on one stream, CUDA events around each kernel on that stream); "
                                       "ms_per_step_sequential is that replay's step time"},
                "primary_mrays_s": (prim_total / args.steps / world) / (prim_ms * 1e-3) / 1e6 * world,
                "bounce_mrays_s": (live_total / args.steps / world) / (bounce_ms * 1e-3) / 1e6 * world,
                "bytes_per_ray": {"primary": bpr_p, "bounce": bpr_b, "V_T": vt, "source": vt_src},
                "reference_bvh_nodes": ref_nodes,
                "e2e_rays_per_step": e2e_total / max(1, e2e_steps), "e2e_ms_per_step": 1e3 * e2e_s / max(1, e2e_steps),
            },
            "parity": parity,
            "roofline": {"bound": "hbm", "kernel": kernel + " (bounce rays, closest hit)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "note": "algorithmic bytes = live bounce rays per launch x (32V + 36T + 48) B; the scene is mostly L2-resident, so DRAM traffic is far below this"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 40 + 17 * 4, "d2h_bytes_per_step": WIDTH * HEIGHT * 3,
                    "call": e2e_call},
            "gpu_launches": int((4 if (args.sequential or inflight > 1) else 4 * chunks) * args.steps * world),
            "clocks": clocks,
        }
        if world == 1 and not args.no_extras:
            line["extra"] = run_extras()
        emit(line)
    if world > 1:
        dist.barrier()
    if dist.is_initialized():
        dist.destroy_process_group()
    if rc:
        raise SystemExit(rc)
    return line


def run_extras():
    """The other BASELINE configs and entry points, measured outside the timed region by the tools/ scripts (one subprocess each,
    one JSON object each): frame times of configs 1 / 2 / 4 / 5 with the fraction of pixels within 2/255 of the REAL reference's
    image, photon gather rate on config 5 against its yardstick, BVH::build seconds per builder, the host-buffer batch API, the
    photon pass on the device, and the same step as the timed one on a deep-traversal scene (56 node entries per camera ray, like
    sponza's published 54.8) with its own V, T and roofline fraction."""
    out = {}
    # the reference renders configs 1, 2 and 5 next to the GPU inside this run; config 4's reference frame needs 140 s of its own
    # preCalc (both photon passes) + 13.5 s, so it is only run with MIRO_REF_ALL=1 (recorded: profiles/r02p_configs.json)
    env = dict(os.environ, MIRO_REF_CONFIG5="1")
    for key, script, tmo in (("configs", "bench_configs.py", 420), ("photon_gather", "bench_gather.py", 120), ("bvh_build", "bench_build.py", 120),
                             ("host_batch_api", "bench_host_batch.py", 120), ("photon_pass", "bench_photon_pass.py", 120),
                             ("deep_traversal", "bench_deep.py", 180)):
        path = os.path.join(ROOT, "tools", script)
        if not os.path.exists(path):
            continue
        try:
            r = subprocess.run([sys.executable, path], capture_output=True, text=True, timeout=tmo, env=env)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            out[key] = json.loads(lines[-1]) if (r.returncode == 0 and lines) else {"error": (r.stderr or r.stdout)[-400:]}
        except Exception as exc:
            out[key] = {"error": repr(exc)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layout", default=os.environ.get("MIROGPU_LAYOUT", "qbvh4"), choices=["bvh2", "cwbvh8", "bvh4", "qbvh4"])
    ap.add_argument("--variant", type=int, default=int(os.environ.get("MIROGPU_VARIANT", "-1")))
    ap.add_argument("--builder", default="sah", choices=["sah", "lbvh", "ploc"], help="BVH::build: host binned SAH (default) or a device builder")
    ap.add_argument("--sequential", action="store_true", help="time the step as four whole-batch launches on one stream (round-1 schedule)")
    ap.add_argument("--emulate-shard", type=int, default=0, help="development: trace rank 0's rows of an N-rank run on one GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity legs (faster iteration)")
    ap.add_argument("--no-extras", action="store_true", help="skip the untimed extra measurements (other configs, gather, builders)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun when called as plain `python bench.py --gpus N`
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd, stdout=_REAL_STDOUT))
        run_ours(args)


if __name__ == "__main__":
    main()

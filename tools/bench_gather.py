"""Photon-map kNN gather throughput on BASELINE config 5 (cornell box + drops, 200 000 + 200 000 photons traced on the device,
k = 500, r_max = 1e10 -- PhotonMap usage of Scene.cpp:286-299): queries = the primary hit points of the 512 x 512 frame.
Prints one JSON object (queries/s per map, algorithmic GB/s by SURVEY 8d's yardstick 28 * visited + 48 bytes per query,
visited counted by the oracle's instrumented locate_photons on a subsample, and the CPU figure of the oracle on the host cores)."""
import ctypes, importlib, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import miro_driver as md, objio
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
H = pkg.HostScene()
scenes.realise(H, "cornell_drops", objio.obj_path)
H.set_photon_counts(200000, 200000)
t0 = time.perf_counter(); H.precalc(); t_pre = time.perf_counter() - t0
S = H.scene(); cam = H.camera()
W = Hh = 512
n = W * Hh
dev = torch.device("cuda", 0)
d_rays = torch.empty((n, 8), dtype=torch.float32, device=dev); d_hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
S.generate_primary(cam, W, Hh, d_rays)
S.intersect_device(d_rays, d_hits, mode=pkg.CLOSEST_HIT | pkg.HINT_COHERENT)
d_P = torch.empty((n, 3), dtype=torch.float32, device=dev); d_N = torch.empty((n, 3), dtype=torch.float32, device=dev)
S.resolve_hits_device(d_hits, d_P, d_N)
hit = (d_hits.view(torch.int32)[:, 1] != -1)
P = d_P[hit].contiguous(); N = d_N[hit].contiguous()
nq = P.shape[0]
irr = torch.empty_like(P)
out = {"scene": "cornell_drops 512x512, k=500, r_max=1e10", "queries": int(nq), "photon_pass_and_build_s": t_pre, "maps": {}}
O = md.oracle()
O.new_scene()   # creates the oracle's two scene-owned maps (which = 0, 1)
EXACT = os.environ.get("MIRO_GATHER_EXACT") == "1"
out["search"] = "reference verbatim, one query per thread (exact)" if EXACT else "one query per warp (default)"
for which, name in ((0, "global"), (1, "caustic")):
    S.photon_set_exact(which, EXACT)
    for _ in range(2):
        S.photon_gather_device(which, P, N, irr, 1e10, 500)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        S.photon_gather_device(which, P, N, irr, 1e10, 500)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ph = H.pm_dump(which)
    O.lib.orc_pm_reset(which, ctypes.c_int(len(ph)))
    O.lib.orc_pm_load(which, md._fp(ph.view(np.uint8)), ctypes.c_int(len(ph) - 1))
    sub = slice(0, nq, 64)
    Ps, Ns = P[sub].cpu().numpy(), N[sub].cpu().numpy()
    O.lib.orc_pm_visited.restype = ctypes.c_longlong
    visited = O.lib.orc_pm_visited(which, md._fp(Ps), md._fp(Ns), ctypes.c_long(Ps.shape[0]), ctypes.c_float(1e10), 500) / Ps.shape[0]
    t0 = time.perf_counter(); ref = O.pm_irradiance(which, Ps, Ns, 1e10, 500); cpu_s = time.perf_counter() - t0
    got = irr[sub].cpu().numpy()
    same = np.array_equal(ref.view(np.uint32), got.view(np.uint32))
    relerr = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)))
    bpq = 28.0 * visited + 48.0
    out["maps"][name] = {"stored": int(len(ph) - 1), "ms": ms, "mqueries_s": nq / ms / 1e3, "visited_per_query": visited, "bytes_per_query": bpq,
                         "algorithmic_gb_s": nq * bpq / (ms * 1e-3) / 1e9, "bit_identical_to_oracle_on_subsample": bool(same), "max_rel_err_vs_oracle_on_subsample": relerr,
                         "cpu_oracle_mqueries_s": Ps.shape[0] / cpu_s / 1e6, "cpu_threads": os.cpu_count()}
print(json.dumps(out))

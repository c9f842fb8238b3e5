"""The photon pass on the device on BASELINE config 5 (cornell box + drops, 200 000 + 200 000 photons):
mirogpu_photon_pass (emit -> store -> scale -> balance -> gather records, nothing crosses PCIe) against the round-1 route
(mirogpu_photon_trace records to the host, Photon_map::store there, the oracle's balance() as the reference's host balance,
mirogpu_photon_upload), and mirogpu_photon_balance alone against the oracle's balance() on the same store-order array.
Prints one JSON object."""
import importlib, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import miro_driver as md, objio
from photon_helpers import consume
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
H = pkg.HostScene()
scenes.realise(H, "cornell_drops", objio.obj_path)
H.set_photon_counts(0, 0)
H.precalc()
S = H.scene()
O = md.oracle(); O.new_scene()
out = {"scene": "cornell_drops (config 5), target 200000 photons per map", "maps": {}}
for which, name in ((0, "global"), (1, "caustic")):
    seed = 168 + which
    S.photon_pass(which, which, seed, 200000)          # warm-up (allocations, module load)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        emissions, stored = S.photon_pass(which, which, seed, 200000)
        ts.append(time.perf_counter() - t0)
    # round-1 route: records over PCIe, host store, host balance (the oracle's = the reference's), upload
    t0 = time.perf_counter()
    counts, records = S.photon_trace(0, which, seed, 0, int(emissions * 1.02) + 1000)
    t_trace = time.perf_counter() - t0
    rec, expect = consume(counts, records, 200000)
    O.lib.orc_pm_reset(which, stored)
    t0 = time.perf_counter(); O.pm_store(which, rec[:, 0:3], rec[:, 3:6], rec[:, 6:9]); O.pm_scale(which, 1.0 / emissions); t_store = time.perf_counter() - t0
    store_order = O.pm_dump(which).copy()
    t0 = time.perf_counter(); O.pm_balance(which); t_bal = time.perf_counter() - t0
    ref = O.pm_dump(which)
    t0 = time.perf_counter(); S.photon_upload(which, ref); t_up = time.perf_counter() - t0
    lo = np.minimum(np.float32(1e8), rec[:, 3:6].min(axis=0)); hi = np.maximum(np.float32(-1e8), rec[:, 3:6].max(axis=0))
    pkg.photon_balance(store_order, lo, hi)
    tb = []
    for _ in range(3):
        t0 = time.perf_counter(); got = pkg.photon_balance(store_order, lo, hi); tb.append(time.perf_counter() - t0)
    same = all(np.array_equal(ref[f][1:], got[f][1:]) for f in ("pos", "power", "theta", "phi"))
    out["maps"][name] = {"emissions": emissions, "stored": stored, "device_pass_s": min(ts), "device_pass_runs_s": ts,
                         "host_route_s": {"trace_and_records_to_host": t_trace, "store_scale": t_store, "balance_reference_single_thread": t_bal, "upload": t_up,
                                          "total": t_trace + t_store + t_bal + t_up},
                         "balance_abi_host_array_s": min(tb), "balance_equals_reference_heap": bool(same),
                         "stop_rule_equals_sequential_replay": bool(emissions == expect)}
    O.lib.orc_pm_reset(which, 1)
print(json.dumps(out))

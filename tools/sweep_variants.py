"""Kernel-variant sweep on the bench workload (bunny20, QBVH4): for each environment-knob setting a fresh process builds the scene,
generates SPP jittered samples of the 1920x1080 frame and their bounce rays, and times the two closest-hit launches (CUDA events,
median of 5 after 2 warm-ups).  A hash of the hits proves every variant returns the same answers.  One JSON line per setting.
    python tools/sweep_variants.py 'MIROGPU_SHORT=8' 'MIROGPU_SHORT=8 MIROGPU_MINB=10' ...
"""
import hashlib, importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPP = int(os.environ.get("SWEEP_SPP", "16"))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np, torch, objio
    pkg = importlib.import_module("cse168-raytracer_b200")
    scenes = importlib.import_module("cse168-raytracer_b200.scenes")
    saved = os.dup(1); os.dup2(2, 1)
    H = pkg.HostScene(); scenes.realise(H, os.environ.get("SWEEP_SCENE", "bunny20"), objio.obj_path); H.precalc()
    S = H.scene(); cam = H.camera()
    W, Hh = 1920, 1080
    n = W * Hh * SPP
    d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda"); d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    d_b = torch.empty((n, 8), dtype=torch.float32, device="cuda"); d_h2 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    d_live = torch.zeros(1, dtype=torch.int64, device="cuda")
    S.generate_primary(cam, W, Hh, d_rays, jitter=1, seed=168, sample=0, samples=SPP)
    S.intersect_device(d_rays, d_hits, mode=pkg.CLOSEST_HIT | pkg.HINT_COHERENT)
    S.generate_bounce(d_rays, d_hits, d_b, seed=168, sample=0, d_live_count=d_live)
    torch.cuda.synchronize()
    live = int(d_live.item())
    def timed(fn):
        ts = []
        for it in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts[2:]))
    tp = timed(lambda: S.intersect_device(d_rays, d_hits, mode=pkg.CLOSEST_HIT | pkg.HINT_COHERENT))
    tb = timed(lambda: S.intersect_device(d_b, d_h2))
    hsh = hashlib.sha1(d_hits.cpu().numpy().tobytes() + d_h2.cpu().numpy().tobytes()).hexdigest()[:16]
    os.dup2(saved, 1)
    print(json.dumps({"primary_ms": tp, "bounce_ms": tb, "primary_grays_s": n / tp / 1e6, "bounce_grays_s": live / tb / 1e6,
                      "step_grays_s": (n + live) / (tp + tb + 0.74) / 1e6, "hits_sha1": hsh}))
    sys.exit(0)
for setting in sys.argv[1:] or [""]:
    env = dict(os.environ, MIROGPU_STRICT="1")
    for kv in setting.split():
        k, v = kv.split("="); env[k] = v
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    out = json.loads(line[-1]) if line else {"error": (r.stderr or "")[-300:]}
    out["setting"] = setting or "(default)"
    print(json.dumps(out), flush=True)

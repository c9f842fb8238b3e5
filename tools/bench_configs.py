"""Frame times of every BASELINE config through the reference-facing call (Scene::raytraceImage -> mirogpu_render_rgb8,
host framebuffer out) next to the reference's own raytraceImage timer (Scene.cpp:206) on the box's host cores
(SURVEY 8d asks for the latter on configs 1-2; configs 4 and 5 are added where the reference finishes in reasonable time).
Prints one JSON object.  GPU: median of 5 frames after 2 warm-up frames; reference: one frame (OpenMP, all host threads)."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import miro_driver as md, objio
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
if len(sys.argv) > 1 and sys.argv[1] == "--reference-frame":
    import ctypes
    name, w, h, path = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    R = md.reference("scalar")
    scenes.realise(R, name, objio.obj_path)
    R._f("srand")(168)
    t0 = time.perf_counter(); R.precalc(); pre = time.perf_counter() - t0
    buf = np.zeros((h, w, 3), np.uint8)
    fs = float(R._f("render")(w, h, buf.ctypes.data_as(ctypes.c_void_p)))
    np.savez(path, img=buf, precalc_s=pre, frame_s=fs)
    sys.exit(0)

import torch  # noqa: F401,E402
pkg = importlib.import_module("cse168-raytracer_b200")
CASES = [  # name, scene, (w, h), spp, photon maps, run the reference?
    ("config 1: cornell_box 512x512 primary + shadow", "cornell", (512, 512), 1, 0, True),
    ("config 2: bunny + teapot 1024x1024 primary + shadow", "bunny_teapot", (1024, 1024), 1, 0, True),
    ("config 4: flower 2048x1365, 4 spp, refractive tree, procedural textures (assignment3.cpp:93-105)", "flower_a3", (2048, 1365), 4, 0, os.environ.get("MIRO_REF_ALL") == "1"),   # the reference's preCalc alone takes 140 s here
    ("config 5: cornell + drops 512x512, photon maps (200k + 200k), gather k=500", "cornell_drops", (512, 512), 1, 1, os.environ.get("MIRO_REF_ALL") == "1" or os.environ.get("MIRO_REF_CONFIG5") == "1"),
]
out = {"threads": os.cpu_count(), "cases": []}
saved = os.dup(1); os.dup2(2, 1)   # the reference prints progress to stdout
try:
    for label, name, (w, h), spp, pm, run_ref in CASES:
        H = pkg.HostScene()
        sc = scenes.realise(H, name, objio.obj_path)
        if pm:
            H.set_photon_counts(200000, 200000)
        t0 = time.perf_counter(); H.precalc(); pre = time.perf_counter() - t0
        H.set_render(spp=spp, jitter=1 if spp > 1 else 0, mode=pkg.RENDER_WHITTED, shadows=1, seed=168, use_photon_maps=pm)
        ts, own = [], []
        img = np.zeros((h, w, 3), np.uint8)
        for it in range(7):
            t0 = time.perf_counter(); H.render(w, h, out=img); ts.append(time.perf_counter() - t0); own.append(H.last_render_seconds)
        rays = H.scene().last_call_stats()[0]
        gpu_s = float(np.median(ts[2:]))     # wall clock around Camera::click (Image::clear + raytraceImage + the copy out to numpy)
        case = {"config": label, "triangles": H.num_objects(), "gpu_precalc_s": pre, "gpu_frame_ms": 1e3 * gpu_s,
                "gpu_frame_ms_own_timer": 1e3 * float(np.median(own[2:])),   # Scene::raytraceImage's own timer, what the reference prints
                "rays_per_frame": int(rays), "gpu_mrays_s": rays / gpu_s / 1e6}
        if run_ref and pm == 0 and any(l["kind"] == 1 for l in sc["lights"]):
            # the reference traces both photon maps whenever the scene has a DirectionalAreaLight (Scene.cpp:76-82) and adds
            # their irradiance at every diffuse hit (Scene.cpp:286-299): give the device the same job for the comparison
            H.set_photon_counts(200000, 200000)
            t0 = time.perf_counter(); H.precalc(); case["gpu_precalc_with_photon_maps_s"] = time.perf_counter() - t0
            pm = 1
        if run_ref:
            if spp != 1 or pm:   # the reference's default build takes one pixel-centre sample (Scene.cpp:140): compare like with like
                H.set_render(spp=1, jitter=0, mode=pkg.RENDER_WHITTED, shadows=1, seed=168, use_photon_maps=pm)
                ts = []
                for it in range(3):
                    t0 = time.perf_counter(); H.render(w, h, out=img); ts.append(time.perf_counter() - t0)
                gpu_s = float(np.median(ts[1:]))
                case["gpu_frame_ms_like_reference"] = 1e3 * gpu_s   # 1 spp, photon maps as the reference has them
            # a fresh process per frame: Camera::eyeRay keeps its basis in function statics (Camera.cpp:106-125), so a second
            # camera in the same process would be rendered with the first one's basis
            import subprocess, tempfile
            tmp = tempfile.mktemp(suffix=".npz")
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--reference-frame", name, str(w), str(h), tmp], stdout=2)
            z = np.load(tmp); os.unlink(tmp)
            buf = z["img"]; case["reference_precalc_s"] = float(z["precalc_s"]); case["reference_frame_s"] = float(z["frame_s"])
            case["speedup_frame"] = case["reference_frame_s"] / gpu_s
            a, b = img.astype(int), buf.astype(int)
            case["pixels_within_2_of_255"] = float((np.abs(a - b).max(axis=2) <= 2).mean())
        out["cases"].append(case)
finally:
    os.dup2(saved, 1)
print(json.dumps(out))

#!/bin/bash
# GPU-box pass for the device photon pass: parity tests, sanitizer, timings, launch list.
set -u
T=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_photon_build.py tests/test_gpu_photon_trace.py -x -q > gpurun_out/${T}_pb_tests.log 2>&1; echo "pytest rc=$? $(tail -3 gpurun_out/${T}_pb_tests.log)"
cat > /tmp/pb_small.py <<'P'
import importlib, sys, os
import numpy as np
sys.path.insert(0, "tests")
pkg = importlib.import_module("cse168-raytracer_b200")
from test_photon_build_model import _photons, balance_model
for n, kind in ((3000, "walls"), (1500, "grid"), (700, "random")):
    pos, d, pw = _photons(n, kind, 5)
    ph = np.zeros(n + 1, pkg.PHOTON_DTYPE); ph["pos"][1:] = pos; ph["power"][1:] = pw
    lo = np.minimum(np.float32(1e8), pos.min(axis=0)); hi = np.maximum(np.float32(-1e8), pos.max(axis=0))
    got = pkg.photon_balance(ph, lo, hi)
    p1 = np.zeros((n + 1, 3), np.float32); p1[1:] = pos
    heap, plane = balance_model(p1, lo, hi)
    assert np.array_equal(got["pos"][1:], p1[heap[1:]]), (n, kind)
print("balance ok")
P
timeout 600 compute-sanitizer --tool memcheck python /tmp/pb_small.py > gpurun_out/${T}_pb_memcheck.log 2>&1; echo "memcheck rc=$? $(tail -2 gpurun_out/${T}_pb_memcheck.log)"
timeout 600 compute-sanitizer --tool racecheck python /tmp/pb_small.py > gpurun_out/${T}_pb_racecheck.log 2>&1; echo "racecheck rc=$? $(tail -2 gpurun_out/${T}_pb_racecheck.log)"
timeout 600 python tools/bench_photon_pass.py > gpurun_out/${T}_photon_pass.json 2> gpurun_out/${T}_photon_pass.err; echo "bench_photon_pass rc=$?"; cat gpurun_out/${T}_photon_pass.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_photon_pass_launches.csv python tools/bench_photon_pass.py > /dev/null 2>&1; echo "ncu rc=$?"
python bench.py --no-cpu --no-extras --steps 10 > gpurun_out/${T}_bench_quick.json 2> gpurun_out/${T}_bench_quick.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${T}_bench_quick.json

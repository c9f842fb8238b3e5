#!/bin/bash
# Photon gather variants on config 5 (tools/bench_gather.py): chunk size, neighbour seeding on/off.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_render.py -x -q -m gpu -k "photon or config5 or config4" > gpurun_out/pytest_photon.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_photon.log
run() { tag=$1; shift; env "$@" python tools/bench_gather.py > gpurun_out/gather_$tag.json 2> gpurun_out/gather_$tag.err; echo "$tag rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/gather_$tag.json"))
print("$tag", {k:(round(v["ms"],2), round(v["mqueries_s"],2), v["max_rel_err_vs_oracle_on_subsample"]) for k,v in d["maps"].items()})
PY
}
run default A=1
run chunk4 MIROGPU_GATHER_CHUNK=4
run chunk16 MIROGPU_GATHER_CHUNK=16

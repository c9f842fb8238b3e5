"""Generates tests/golden/ref_photon_stats.json by running the REAL reference (compiled in place into oracle/_ref by
oracle/Makefile) on BASELINE config 5: Scene::preCalc traces PhotonsPerLightSource = CausticPhotonsPerLightSource =
200000 photons (Scene.h:67-68) from the DirectionalAreaLight and balances both maps.  Run in the build container, where
/root/reference exists:   python tests/golden/make_photon_stats.py
The reference draws from rand(), so only distribution statistics are kept; the photon tests compare against them with
statistical tolerances (stored count and mean power within 1 %)."""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import miro_driver as md   # noqa: E402
import objio               # noqa: E402


def stats_of(ph, base_power):
    """ph: structured photon array without entry 0.  base_power = color*wattage*PI r^2 (/10 for the caustic pass)."""
    pw, pos = ph["power"].astype(np.float64), ph["pos"].astype(np.float64)
    white = (ph["power"][:, 0] == ph["power"][:, 1]) & (ph["power"][:, 1] == ph["power"][:, 2])
    vals, cnt = np.unique(ph["power"][white][:, 0], return_counts=True)
    scaled_unit = float(vals[cnt.argmax()])            # power of a photon that only met white surfaces = base / emissions
    edges = [np.linspace(0, 5.5, 5), np.linspace(0, 5.5, 5), np.linspace(-5.5, 0, 5)]
    hist, _ = np.histogramdd(pos, bins=edges)
    return dict(stored=int(len(ph)), emissions=float(base_power / scaled_unit), mean_pos=pos.mean(0).tolist(), std_pos=pos.std(0).tolist(),
                mean_power_times_emissions=(pw.mean(0) * (base_power / scaled_unit)).tolist(), floor_fraction=float((pos[:, 1] < 0.01).mean()),
                hist4=(hist / len(ph)).ravel().tolist())


if __name__ == "__main__":
    scenes = importlib.import_module("cse168-raytracer_b200.scenes")
    R = md.reference("scalar")
    scenes.realise(R, "cornell_drops", objio.obj_path)
    R._f("srand")(168)
    R.precalc()
    out = {"scene": "cornell_drops", "source": "reference Scene::preCalc (tracePhotons + traceCausticPhotons), srand(168), oracle/_ref/libmiro_ref.so"}
    for which, name, base in ((0, "global", 160 * np.pi), (1, "caustic", 160 * np.pi / 10)):
        out[name] = stats_of(R.pm_dump(which)[1:], base)
        print(name, {k: v for k, v in out[name].items() if k != "hist4"})
    json.dump(out, open(os.path.join(HERE, "ref_photon_stats.json"), "w"), indent=1)

"""Shared by the photon-tracing tests: the reference's sequential stop rule and the statistics compared with the
golden figures of the real reference (tests/golden/ref_photon_stats.json, made by tests/golden/make_photon_stats.py)."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_photon_stats.json")
BASE_POWER = {0: 160 * np.pi, 1: 160 * np.pi / 10}   # color * wattage * PI r^2 (/ 10 caustic): Scene.cpp:379-385, 431-434


def golden():
    with open(GOLD) as f:
        return json.load(f)


def consume(counts, records, target):
    """Scene.cpp:370-396 executed sequentially: emission i happens only while fewer than `target` photons are stored.
    Returns (records of the stored photons (n, 9), emissions consumed)."""
    csum = np.cumsum(counts.astype(np.int64))
    before = csum - counts
    used = before < target                       # the test `photonsAdded < PhotonsPerLightSource` made before emission i
    emissions = int(used.sum())
    assert emissions < len(counts), "not enough emissions traced to reach the target"
    keep = used[:, None] & (np.arange(5)[None, :] < counts[:, None])
    return records[keep], emissions


def stats(rec, emissions):
    """Same figures as make_photon_stats.stats_of, from unscaled {power, pos, dir} records."""
    pos = rec[:, 3:6].astype(np.float64)
    edges = [np.linspace(0, 5.5, 5), np.linspace(0, 5.5, 5), np.linspace(-5.5, 0, 5)]
    hist, _ = np.histogramdd(pos, bins=edges)
    return dict(stored=len(rec), emissions=float(emissions), mean_pos=pos.mean(0), std_pos=pos.std(0),
                mean_power_times_emissions=rec[:, 0:3].astype(np.float64).mean(0), floor_fraction=float((pos[:, 1] < 0.01).mean()),
                hist4=(hist / len(rec)).ravel())


def assert_statistical_parity(mine, ref, target):
    """photon tracing is statistical -- stored count and mean power within 1 % (2 % on the noisier figures)."""
    assert target <= mine["stored"] <= target + 5 and target <= ref["stored"] <= target + 30
    assert abs(mine["emissions"] / ref["emissions"] - 1) < 0.02, (mine["emissions"], ref["emissions"])
    assert np.allclose(mine["mean_pos"], ref["mean_pos"], atol=0.03), (mine["mean_pos"], ref["mean_pos"])
    assert np.allclose(mine["std_pos"], ref["std_pos"], atol=0.03), (mine["std_pos"], ref["std_pos"])
    assert np.allclose(mine["mean_power_times_emissions"], ref["mean_power_times_emissions"], rtol=0.02), \
        (mine["mean_power_times_emissions"], ref["mean_power_times_emissions"])
    assert abs(mine["floor_fraction"] - ref["floor_fraction"]) < 0.01
    assert np.abs(np.asarray(mine["hist4"]) - np.asarray(ref["hist4"])).max() < 0.01

"""The oracle against the committed golden vectors (tests/golden/ref_hits.npz, produced by the real
reference through oracle/_ref with tests/golden/make_golden_hits.py).  Runs anywhere, GPU box included."""
import os

import numpy as np
import pytest

import objio
from conftest import bits

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_hits.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("name", ["testobj", "cornell", "teapot", "bunny_teapot"])
def test_oracle_matches_golden(oracle, scenes, gold, name):
    scenes.realise(oracle, name, objio.obj_path)
    oracle.precalc()
    st = oracle.stats()
    assert [st["nodes"], st["leaves"]] == list(gold[f"{name}__nodes"])
    for kind in ("primary", "bounce"):
        rays = gold[f"{name}__{kind}_rays"]
        oracle.stats_reset_rays()
        t, ids, P, N = oracle.trace(rays, 1)
        assert np.array_equal(ids, gold[f"{name}__{kind}_id"])
        assert np.array_equal(bits(t), bits(gold[f"{name}__{kind}_t"]))
        assert np.array_equal(bits(P), bits(gold[f"{name}__{kind}_P"]))
        assert np.array_equal(bits(N), bits(gold[f"{name}__{kind}_N"]))
        st = oracle.stats()
        assert [st["ray_box"], st["ray_tri"]] == list(gold[f"{name}__{kind}_counters"])


def test_primary_rays_match_golden(oracle, scenes, gold):
    for name, (w, h) in {"cornell": (64, 64), "teapot": (96, 96)}.items():
        scenes.realise(oracle, name, objio.obj_path)
        assert np.array_equal(bits(oracle.eye_rays(w, h)), bits(gold[f"{name}__primary_rays"]))

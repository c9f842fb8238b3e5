"""GPU tier: photon emission and tracing on the device (mirogpu_photon_trace <- Scene::tracePhoton, Scene.cpp:526-641)
and Scene::tracePhotons / traceCausticPhotons of the host layer (Scene.cpp:351-472), on BASELINE config 5.

  * emission by emission against the oracle, which draws the same counter-based uniforms: the walks must agree except
    where CUDA's sinf/cosf/asinf/acosf round differently from glibc's (the same last-ulp effect as for bounce rays) and
    the photon then lands on another triangle -- <= 0.3 % of emissions;
  * the whole pass against the REAL reference's maps, statistically (stored count, mean power within 1 %);
  * the host layer's stop rule ("emit while fewer than the target are stored", in emission order) exactly, and
    independence of how the emissions are batched.
"""
import numpy as np
import pytest

import objio
from photon_helpers import assert_statistical_parity, consume, golden, stats

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def drops(pkg, scenes, oracle):
    H = pkg.HostScene(pkg.LAYOUT_BVH2)
    for d in (oracle, H):
        scenes.realise(d, "cornell_drops", objio.obj_path)
        d.precalc()
    return H, H.scene(), oracle


@pytest.mark.parametrize("caustic", [0, 1])
def test_device_walks_match_the_oracle_emission_by_emission(drops, caustic):
    H, S, O = drops
    n = 200000
    gc, gr = S.photon_trace(0, caustic, 168 + caustic, 0, n)
    oc, orr = O.trace_photons(0, caustic, 168 + caustic, 0, n)
    assert gc.max() <= 5 and gc.sum() > (1000 if caustic else 100000)
    same = gc == oc
    assert same.mean() > 0.997, same.mean()
    assert abs(int(gc.sum()) - int(oc.sum())) <= 0.003 * oc.sum()
    keep = same[:, None] & (np.arange(5)[None, :] < gc[:, None])
    a, b = gr[keep], orr[keep]
    close = np.isclose(a, b, rtol=2e-4, atol=2e-4).all(axis=1)
    assert close.mean() > 0.995, close.mean()
    # a record is {power, pos, incoming dir}: directions are unit vectors (refracted ones are not normalised by the
    # reference, Ray.h:233, so allow their slack), positions lie in the box
    assert (np.abs(np.linalg.norm(a[:, 6:9], axis=1) - 1) < 0.2).all()
    assert (a[:, 3:6].min(0) > [-0.01, -0.01, -5.51]).all() and (a[:, 3:6].max(0) < [5.51, 5.51, 0.01]).all()


def test_layouts_and_call_splits_give_the_same_walks(pkg, scenes, drops):
    H, S, O = drops
    c0, r0 = S.photon_trace(0, 0, 168, 5000, 30000)
    c1, r1 = S.photon_trace(0, 0, 168, 15000, 4000)
    assert np.array_equal(c0[10000:14000], c1) and np.array_equal(r0[10000:14000], r1)
    H8 = pkg.HostScene(pkg.LAYOUT_CWBVH8)
    scenes.realise(H8, "cornell_drops", objio.obj_path)
    H8.precalc()
    c8, r8 = H8.scene().photon_trace(0, 0, 168, 5000, 30000)
    assert np.array_equal(c0, c8) and np.array_equal(r0, r8)          # closest hits are layout-independent, so are the walks
    scenes.realise(H, "cornell_drops", objio.obj_path); H.precalc()    # the host layer holds one global scene: restore it


@pytest.mark.parametrize("which,name", [(0, "global"), (1, "caustic")])
def test_scene_trace_photons_against_the_reference(pkg, scenes, drops, which, name):
    H, _, O = drops
    S = H.scene()      # the handle of the CURRENT BVH (the test above rebuilt the global scene: the fixture's wrapper is stale)
    target = 200000                                                     # Scene.h:67-68
    H.set_photon_counts(target if which == 0 else 0, target if which == 1 else 0)
    emissions = H.trace_photons(which)
    ph = H.pm_dump(which)[1:]
    assert target <= len(ph) <= target + 5
    # exact stop rule: replay it on the device's own per-emission counts
    n = int(emissions * 1.02) + 1000
    counts, records = S.photon_trace(0, which, 168 + which, 0, n)
    rec, expect = consume(counts, records, target)
    assert emissions == expect and len(ph) == len(rec)
    # Photon_map::balance only permutes: same multiset of positions, powers scaled by 1 / emissions (Scene.cpp:400)
    assert np.array_equal(np.sort(ph["pos"][:, 0]), np.sort(rec[:, 3]))
    assert np.allclose(np.sort(ph["power"][:, 0]), np.sort(rec[:, 0] * np.float32(1.0 / emissions)), rtol=1e-6)
    assert_statistical_parity(stats(rec, emissions), golden()[name], target)
    # and the map is live on the device: the gather sees it
    q = np.array([[2.5, 0.0, -2.5], [1.0, 0.0, -1.0]], np.float32); qn = np.array([[0, 1, 0], [0, 1, 0]], np.float32)
    S.photon_set_exact(which, True)
    irr = S.photon_gather(which, q, qn, 1e10, 100)
    assert (irr > 0).all()
    O.lib.orc_pm_reset(which, len(ph))
    O.pm_store(which, rec[:, 0:3], rec[:, 3:6], rec[:, 6:9]); O.pm_scale(which, 1.0 / emissions); O.pm_balance(which)
    assert np.array_equal(O.pm_irradiance(which, q, qn, 1e10, 100).view(np.uint32), irr.view(np.uint32))
    O.lib.orc_pm_reset(which, 1)

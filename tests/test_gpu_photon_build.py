"""GPU tier: the photon pass behind the walks on the device (csrc/photon_build.cu).

  * mirogpu_photon_balance <- Photon_map::balance (PhotonMap.cpp:314-466): the heap array must equal the oracle's (whose
    balance() is pinned to the real reference in test_oracle_vs_reference.py) bit for bit -- positions, powers, direction
    bytes and split planes -- also where keys tie (photons on axis-aligned walls, quantised coordinates, one shared
    coordinate) and where the input is already sorted: the quickselect's rounds are reproduced, not just its medians;
  * mirogpu_photon_pass <- Scene::tracePhotons / traceCausticPhotons (Scene.cpp:351-472): emissions consumed, photons
    stored (stop rule in emission order), Photon_map::store's direction bytes, scale_photon_power and the balanced array
    against the oracle fed with the device's own per-emission records;
  * the host layer's Photon_map::balance (host-filled maps) goes through the same device code.
"""
import numpy as np
import pytest

import objio
from photon_helpers import consume
from test_photon_build_model import _photons

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _same_map(a, b, n):
    for f in ("pos", "power"):
        assert np.array_equal(a[f][1:].view(np.uint32), b[f][1:].view(np.uint32)), f
    for f in ("theta", "phi"):
        assert np.array_equal(a[f][1:], b[f][1:]), f
    # plane is set on the nodes balance_segment visits (every node with a child); the gather reads it below n / 2 - 1
    inner = np.arange(1, n + 1) * 2 <= n
    assert np.array_equal(a["plane"][1:][inner], b["plane"][1:][inner])


@pytest.mark.parametrize("n,kind", [(1, "random"), (2, "random"), (3, "grid"), (100, "random"), (513, "walls"), (1025, "grid"), (4097, "random"),
                                    (4097, "walls"), (70001, "random"), (70001, "walls"), (50000, "equal"), (30000, "sorted"),
                                    (200000, "random"), (200003, "walls"), (150000, "grid")])
def test_device_balance_equals_the_reference_heap(pkg, oracle, n, kind):
    pos, d, pw = _photons(n, kind, 7000 + n)
    oracle.new_scene(); w = oracle.pm_new(n)
    oracle.pm_store(w, pw, pos, d); oracle.pm_scale(w, 0.25)
    store_order = oracle.pm_dump(w).copy()
    oracle.pm_balance(w)
    ref = oracle.pm_dump(w)
    lo = np.minimum(np.float32(1e8), pos.min(axis=0)); hi = np.maximum(np.float32(-1e8), pos.max(axis=0))
    got = pkg.photon_balance(store_order, lo, hi)
    _same_map(ref, got, n)


@pytest.mark.parametrize("n", [4097, 70001])
def test_host_layer_balance_goes_through_the_device(pkg, oracle, n):
    pos, d, pw = _photons(n, "walls", 9)
    H = pkg.HostScene(); H.new_scene()
    oracle.new_scene(); w = oracle.pm_new(n)
    oracle.pm_store(w, pw, pos, d); oracle.pm_scale(w, 0.25); oracle.pm_balance(w)
    H.pm_store(0, pw, pos, d); H.pm_scale(0, 0.25); H.pm_balance(0)
    _same_map(oracle.pm_dump(w), H.pm_dump(0), n)


@pytest.fixture(scope="module")
def drops(pkg, scenes, oracle):
    H = pkg.HostScene(pkg.LAYOUT_QBVH4)
    for d in (oracle, H):
        scenes.realise(d, "cornell_drops", objio.obj_path)
        d.precalc()
    return H, H.scene(), oracle


@pytest.mark.parametrize("which,target", [(0, 200000), (1, 200000), (0, 5000), (1, 3), (0, 1)])
def test_device_pass_equals_store_scale_balance_of_its_own_records(drops, which, target):
    H, S, O = drops
    seed = 168 + which
    emissions, stored = S.photon_pass(which, which, seed, target)
    got = S.photon_download(which)
    assert stored == len(got) - 1 and target <= stored <= target + 4
    # the same emissions through the per-emission ABI, consumed sequentially on the host (Scene.cpp:370-396)
    counts, records = S.photon_trace(0, which, seed, 0, int(emissions * 1.02) + 1000)
    rec, expect = consume(counts, records, target)
    assert emissions == expect and stored == len(rec)
    O.lib.orc_pm_reset(which, stored)
    O.pm_store(which, rec[:, 0:3], rec[:, 3:6], rec[:, 6:9]); O.pm_scale(which, 1.0 / emissions); O.pm_balance(which)
    _same_map(O.pm_dump(which), got, stored)
    # the map is live: exact-mode gather equals the oracle's on its own balanced array
    q = np.array([[2.5, 0.0, -2.5], [1.0, 0.0, -1.0], [0.0, 2.0, -3.0]], np.float32); qn = np.array([[0, 1, 0], [0, 1, 0], [1, 0, 0]], np.float32)
    S.photon_set_exact(which, True)
    k = min(100, max(1, stored))
    assert np.array_equal(S.photon_gather(which, q, qn, 1e10, k).view(np.uint32), O.pm_irradiance(which, q, qn, 1e10, k).view(np.uint32))
    S.photon_set_exact(which, False)
    O.lib.orc_pm_reset(which, 1)


def test_device_built_maps_are_replicated_over_a_device_list(pkg, scenes):
    """mirogpu_photon_pass builds the map on the first device of a multi-device handle and clones it to the others with peer
    copies: a photon-mapped frame rendered over two devices equals the one-device frame."""
    if pkg.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    frames = []
    for ndev in (1, 2):
        H = pkg.HostScene(pkg.LAYOUT_QBVH4)
        scenes.realise(H, "cornell_drops", objio.obj_path)
        H.set_device_count(ndev)
        H.set_photon_counts(20000, 5000)
        H.precalc()
        S = H.scene()
        assert S.devices() == list(range(ndev))
        p = S.render_params(96, 96, mode=pkg.RENDER_WHITTED, shadows=1, use_photon_maps=1, seed=5)
        frames.append(S.render(H.camera(), p))
    assert frames[0].std() > 0
    assert np.allclose(frames[0], frames[1], rtol=2e-5, atol=1e-7)

"""CPU tier: the oracle's restatement of Scene::tracePhoton / tracePhotons (counter-based uniforms in place of rand())
against distribution statistics of the REAL reference's photon maps on BASELINE config 5 (golden fixture made by
tests/golden/make_photon_stats.py from oracle/_ref).  This pins the oracle for the photon-tracing path; the GPU tier
(test_gpu_photon_trace.py) then compares the device walks with the oracle emission by emission."""
import numpy as np
import pytest

import objio
from photon_helpers import BASE_POWER, assert_statistical_parity, consume, golden, stats


@pytest.fixture(scope="module")
def drops_oracle(oracle, scenes):
    scenes.realise(oracle, "cornell_drops", objio.obj_path)
    oracle.precalc()
    return oracle


@pytest.mark.parametrize("which,name,emit", [(0, "global", 300000), (1, "caustic", 1600000)])
def test_oracle_photon_pass_matches_reference_statistics(drops_oracle, which, name, emit):
    target = 200000                                           # Scene.h:67-68
    counts, records = drops_oracle.trace_photons(0, which, 168 + which, 0, emit)
    rec, emissions = consume(counts, records.reshape(emit, 5, 9), target)
    assert_statistical_parity(stats(rec, emissions), golden()[name], target)
    # every photon leaves the light with the full power of the pass, and white surfaces keep it unchanged:
    # the most common stored power is exactly that value
    vals, cnt = np.unique(rec[:, 0], return_counts=True)
    assert np.isclose(vals[cnt.argmax()], BASE_POWER[which], rtol=1e-6)


def test_walks_are_a_pure_function_of_the_emission_index(drops_oracle):
    a_c, a_r = drops_oracle.trace_photons(0, 0, 168, 1000, 5000)
    b_c, b_r = drops_oracle.trace_photons(0, 0, 168, 3000, 1000)
    assert np.array_equal(a_c[2000:3000], b_c) and np.array_equal(a_r[2000:3000], b_r)
    c_c, _ = drops_oracle.trace_photons(0, 0, 169, 1000, 5000)
    assert not np.array_equal(a_c, c_c)
    # caustic photons are only stored after a specular surface: never at the first hit, and the first stored one
    # of an emission sits below the drops' top (y <= 2.3) or on the walls
    cc, cr = drops_oracle.trace_photons(0, 1, 169, 0, 20000)
    assert cc.sum() > 1000 and cc.max() <= 5

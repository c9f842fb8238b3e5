"""CPU tier: the algorithm behind the device photon balance (csrc/photon_build.cu), replayed in numpy against the oracle.

Photon_map::balance (reference PhotonMap.cpp:314-466) selects each node's median with Jensen's median_split, a Hoare
quickselect; where keys are equal (photons on an axis-aligned wall) the resulting heap array depends on the order in
which its two scan pointers meet the elements.  The device evaluates every partition round from a closed form instead
of walking the pointers:

    G = positions in [left, right) with key not < pivot, ascending;  S = positions with key not > pivot, descending;
    K = #{k : g_k < s_k};  the round exchanges g_k <-> s_k for k <= K, then the pivot with i = min(g_{K+1}, s_K)
    (s_0 = right).

This file states that closed form in numpy (round for round what hoare_round<> does with ballots and prefix counts),
runs the whole balance with it and checks the heap array against the oracle's restatement of balance() -- which
tests/test_oracle_vs_reference.py pins to the real reference -- on inputs with and without ties.  The GPU tier
(tests/test_gpu_photon_build.py) then checks the kernels themselves the same way.
"""
import numpy as np
import pytest


def hoare_round(key, perm, left, right):
    """One partition round of median_split on perm[left..right] (inclusive); returns where the pivot lands."""
    v = key[perm[right]]
    q = np.arange(left, right)
    k = key[perm[left:right]]
    G = q[~(k < v)]                    # where the left pointer can stop
    S = q[~(k > v)][::-1]              # where the right pointer can stop, in its scan order
    m = min(len(G), len(S))
    K = int(np.count_nonzero(G[:m] < S[:m]))     # monotone predicate: the first K pairs are exchanged
    assert np.all(G[:K] < S[:K]) and (K == m or G[K] >= S[K])
    a, b = G[:K], S[:K]
    perm[a], perm[b] = perm[b].copy(), perm[a].copy()
    gi = G[K] if K < len(G) else np.iinfo(np.int64).max
    si = S[K - 1] if K >= 1 else right
    i = int(min(gi, si))
    perm[i], perm[right] = perm[right], perm[i]
    return i


def median_of(start, end):             # PhotonMap.cpp:416-425
    count = end - start + 1
    median = 1
    while 4 * median <= count:
        median += median
    if 3 * median <= count:
        median += median
        median += start - 1
    else:
        median = end - median + 1
    return median


def balance_model(pos, lo, hi):
    """pos: (n + 1, 3) float32 with entry 0 unused; returns heap[1..n] = store index and plane[1..n] (-1 where unset)."""
    n = pos.shape[0] - 1
    heap = np.zeros(n + 1, np.int64); plane = -np.ones(n + 1, np.int64)
    if n == 1:
        heap[1] = 1
    if n <= 1:
        return heap, plane
    perm = np.arange(n + 1)
    stack = [(1, 1, n, np.array(lo, np.float32), np.array(hi, np.float32))]
    while stack:
        index, start, end, blo, bhi = stack.pop()
        median = median_of(start, end)
        ext = bhi - blo                # float32 subtraction, as the reference's
        axis = 0 if (ext[0] > ext[1] and ext[0] > ext[2]) else (1 if ext[1] > ext[2] else 2)
        key = pos[:, axis]
        left, right = start, end
        while right > left:
            i = hoare_round(key, perm, left, right)
            if i >= median:
                right = i - 1
            if i <= median:
                left = i + 1
        heap[index] = perm[median]; plane[index] = axis
        split = key[perm[median]]
        if median > start:
            if start < median - 1:
                h2 = bhi.copy(); h2[axis] = split
                stack.append((2 * index, start, median - 1, blo, h2))
            else:
                heap[2 * index] = perm[start]
        if median < end:
            if median + 1 < end:
                l2 = blo.copy(); l2[axis] = split
                stack.append((2 * index + 1, median + 1, end, l2, bhi))
            else:
                heap[2 * index + 1] = perm[end]
    return heap, plane


def _photons(n, kind, seed):
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3), dtype=np.float32) * 3
    if kind == "grid":                 # few distinct values per axis: ties everywhere
        pos = np.floor(pos * 3).astype(np.float32) / 3
    elif kind == "walls":              # photons on the faces of a box, like a Cornell box's
        face = rng.integers(0, 6, n)
        for f in range(6):
            pos[face == f, f % 3] = 0.0 if f < 3 else 3.0
    elif kind == "equal":              # one coordinate shared by all
        pos[:, 1] = 1.5
    elif kind == "sorted":
        pos = pos[np.argsort(pos[:, 0])]
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    pw = rng.random((n, 3), dtype=np.float32)
    return pos, d, pw


@pytest.mark.parametrize("n,kind", [(1, "random"), (2, "random"), (3, "grid"), (7, "walls"), (100, "random"), (257, "grid"), (1000, "walls"),
                                    (1500, "equal"), (777, "sorted"), (4097, "walls"), (3000, "grid")])
def test_closed_form_rounds_reproduce_balance(oracle, n, kind):
    pos, d, pw = _photons(n, kind, 1000 + n)
    oracle.new_scene(); w = oracle.pm_new(n)
    oracle.pm_store(w, pw, pos, d); oracle.pm_balance(w)
    ref = oracle.pm_dump(w)
    p1 = np.zeros((n + 1, 3), np.float32); p1[1:] = pos
    lo = np.minimum(np.float32(1e8), pos.min(axis=0)); hi = np.maximum(np.float32(-1e8), pos.max(axis=0))
    heap, plane = balance_model(p1, lo, hi)
    assert sorted(heap[1:]) == list(range(1, n + 1))
    assert np.array_equal(ref["pos"][1:].view(np.uint32), p1[heap[1:]].view(np.uint32))
    assert np.array_equal(ref["power"][1:].view(np.uint32), pw[heap[1:] - 1].view(np.uint32))
    inner = plane[1:] >= 0
    assert np.array_equal(ref["plane"][1:][inner], plane[1:][inner])


def test_round_against_pointer_walk():
    """The closed form against a literal two-pointer walk of one partition round, on short arrays with many ties."""
    rng = np.random.default_rng(5)
    for trial in range(3000):
        m = int(rng.integers(2, 24))
        key = rng.integers(0, 4, m).astype(np.float32)
        left, right = 0, m - 1
        p = list(range(m))
        v = key[p[right]]; i = left - 1; j = right
        while True:
            i += 1
            while key[p[i]] < v:
                i += 1
            j -= 1
            while key[p[j]] > v and j > left:
                j -= 1
            if i >= j:
                break
            p[i], p[j] = p[j], p[i]
        p[i], p[right] = p[right], p[i]
        perm = np.arange(m)
        i2 = hoare_round(key, perm, left, right)
        assert i2 == i and list(perm) == p, (trial, key)

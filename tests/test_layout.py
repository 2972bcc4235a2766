"""CPU tests of the device layout cfmm_finalize builds for ProductTwoCoin pools
(csrc/pool_layout.hpp), through the device-free hook cfmm_debug_product_layout:
b-bucketing, per-bucket padding to whole 96-pool chunks, a-order inside buckets,
hub detection and degree orientation."""
import ctypes as C

import numpy as np
import pytest


def layout(cr, n, Ai, orient=-1, variant=0):
    lib = cr.load_library()
    Ai = np.ascontiguousarray(Ai, dtype=np.int64).reshape(-1, 2)
    m = len(Ai)
    info = np.zeros(6, dtype=np.int64)
    ip = C.POINTER(C.c_int64)
    rc = lib.cfmm_debug_product_layout(n, m, Ai.ctypes.data_as(ip), orient, variant, 0, None, None, None,
                                       info.ctypes.data_as(ip))
    assert rc == 0
    mp, tile = int(info[0]), int(info[4])
    order = np.zeros(mp, dtype=np.int64)
    tb = np.zeros(max(mp // tile, 1) if tile else 1, dtype=np.int32)
    sw = np.zeros(max(m, 1), dtype=np.uint8)
    rc = lib.cfmm_debug_product_layout(n, m, Ai.ctypes.data_as(ip), orient, variant, mp,
                                       order.ctypes.data_as(ip), tb.ctypes.data_as(C.POINTER(C.c_int32)),
                                       sw.ctypes.data_as(C.POINTER(C.c_uint8)), info.ctypes.data_as(ip))
    assert rc == 0
    return dict(m_padded=mp, nb=int(info[1]), bucketed=bool(info[2]), skewed=bool(info[3]), tile=tile,
                variant=int(info[5]), order=order, tile_bucket=tb[:mp // tile] if tile else tb[:0], swapped=sw[:m])


def check_invariants(Ai, n, lay):
    Ai = np.asarray(Ai).reshape(-1, 2) - 1
    m = len(Ai)
    order, sw = lay["order"], lay["swapped"].astype(bool)
    real = order[order >= 0]
    assert sorted(real.tolist()) == list(range(m))  # every pool exactly once
    oa = np.where(sw, Ai[:, 1], Ai[:, 0])
    ob = np.where(sw, Ai[:, 0], Ai[:, 1])
    if not lay["bucketed"]:
        assert lay["m_padded"] == m and np.all(np.diff(oa[order]) >= 0)  # a-sorted, stable
        return
    tile, nb = lay["tile"], lay["nb"]
    assert lay["m_padded"] % tile == 0 and len(lay["tile_bucket"]) == lay["m_padded"] // tile
    assert np.all(np.diff(lay["tile_bucket"]) >= 0)  # buckets in order
    for t, bk in enumerate(lay["tile_bucket"]):
        seg = order[t * tile:(t + 1) * tile]
        r = seg[seg >= 0]
        assert np.all(ob[r] // nb == bk)  # a tile never straddles buckets
        assert np.all(seg[len(r):] == -1)  # padding trails inside the tile
    for bk in np.unique(lay["tile_bucket"]):
        tiles = np.nonzero(lay["tile_bucket"] == bk)[0]
        seg = order[tiles[0] * tile:(tiles[-1] + 1) * tile]
        r = seg[seg >= 0]
        assert len(r) > 0 and np.all(seg[:len(r)] >= 0)  # padding only at the end of the bucket
        assert np.all(np.diff(oa[r]) >= 0)  # a non-decreasing inside the bucket
        # stable: equal a keep insertion order
        same = np.diff(oa[r]) == 0
        assert np.all(np.diff(r)[same] > 0)
    assert lay["m_padded"] - m < len(np.unique(lay["tile_bucket"])) * tile  # < one tile of padding per bucket


@pytest.mark.parametrize("variant", [0, -1])
@pytest.mark.parametrize("m,n", [(1, 2), (7, 3), (5000, 7), (40_000, 3001), (60_000, 20_011)])
def test_uniform_graph_layout(cr, m, n, variant):
    from cfmmrouter_b200 import synth
    _, _, Ai = synth.product_pools(m, n, seed=m + n)
    lay = layout(cr, n, Ai, variant=variant)
    assert not lay["skewed"] and not lay["swapped"].any() and lay["variant"] == variant
    check_invariants(Ai, n, lay)
    if variant == -1:
        assert not lay["bucketed"]
    else:
        assert lay["bucketed"] and lay["tile"] == 96
        B = -(-n // 1600)
        assert lay["nb"] == -(-n // B) and lay["nb"] <= 1600


@pytest.mark.parametrize("m,n,variant", [(300_000, 12_000, 0), (300_000, 1_500, 0), (200_000, 3_200, 0),
                                         (300_000, 12_000, -1)])
def test_large_sets_take_the_threaded_paths(cr, m, n, variant):
    """Above 65 536 pools the layout runs its OpenMP forms (threaded counting sort on the bucket
    keys, buckets sorted concurrently when there are >= 4 of them, the threaded sort inside a
    bucket otherwise): the order must be exactly the stable (bucket(b), a, insertion index) order."""
    from cfmmrouter_b200 import synth
    _, _, Ai = synth.product_pools(m, n, seed=m + n)
    lay = layout(cr, n, Ai, variant=variant)
    check_invariants(Ai, n, lay)
    a, b = Ai[:, 0] - 1, Ai[:, 1] - 1
    real = lay["order"][lay["order"] >= 0]
    if variant == -1:
        assert not lay["bucketed"]
        assert np.array_equal(real, np.argsort(a, kind="stable"))
    else:
        assert lay["bucketed"]
        assert np.array_equal(real, np.lexsort((np.arange(m), a, b // lay["nb"])))


def test_sparse_buckets_fall_back_to_a_sorted(cr):
    from cfmmrouter_b200 import synth
    _, _, Ai = synth.product_pools(3000, 200_000, seed=1)  # 63+ buckets, ~50 pools each
    lay = layout(cr, 200_000, Ai)
    assert not lay["bucketed"] and lay["m_padded"] == 3000
    check_invariants(Ai, 200_000, lay)


@pytest.mark.parametrize("orient", [-1, 1, 0])
def test_skewed_graph_orientation(cr, orient):
    from cfmmrouter_b200 import synth
    n, m = 4000, 80_000
    _, _, Ai = synth.product_pools_skewed(m, n, alpha=1.0, seed=5)
    lay = layout(cr, n, Ai, orient=orient)
    check_invariants(Ai, n, lay)
    deg = np.bincount((Ai - 1).ravel(), minlength=n)
    a, b = Ai[:, 0] - 1, Ai[:, 1] - 1
    if orient == 0:
        assert not lay["swapped"].any() and lay["variant"] == 0
    else:
        assert lay["skewed"] and lay["variant"] == 0  # default shape, SKEW instantiation
        assert np.array_equal(lay["swapped"].astype(bool), deg[b] > deg[a])  # higher-degree token first
        hub = int(np.argmax(deg))
        first = np.where(lay["swapped"].astype(bool), b, a)
        assert np.all(first[(a == hub) | (b == hub)] == hub)  # the hub is always on the run side


def test_uniform_graph_orientation_forced(cr):
    from cfmmrouter_b200 import synth
    _, _, Ai = synth.product_pools(30_000, 500, seed=3)
    lay = layout(cr, 500, Ai, orient=1)
    assert not lay["skewed"] and lay["swapped"].any() and lay["variant"] == 0
    check_invariants(Ai, 500, lay)


def test_bad_arguments(cr):
    lib = cr.load_library()
    info = np.zeros(6, dtype=np.int64)
    ip = C.POINTER(C.c_int64)
    bad = np.array([[1, 9]], dtype=np.int64)
    assert lib.cfmm_debug_product_layout(5, 1, bad.ctypes.data_as(ip), -1, 0, 0, None, None, None,
                                         info.ctypes.data_as(ip)) == -1
    assert lib.cfmm_debug_product_layout(5, 1, bad.ctypes.data_as(ip), -1, 99, 0, None, None, None,
                                         info.ctypes.data_as(ip)) == -1

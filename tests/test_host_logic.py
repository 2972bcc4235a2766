"""Host-side logic that needs no GPU: objectives (test/objectives.jl), pool
descriptors, sharding, and the Router/route! driver wired to an ORACLE-backed
stand-in for DevicePools (tests may use the oracle; the product never does)."""
import os

import numpy as np
import pytest


# ---- test/objectives.jl ---------------------------------------------------

def test_linear_nonnegative(cr):
    with pytest.raises(ValueError):
        cr.LinearNonnegative([1.0, -1.0])  # objectives.jl test :4
    assert cr.LinearNonnegative([1, 2]).c.dtype == np.float64  # :5-6
    obj = cr.LinearNonnegative(np.ones(2))
    assert obj.f(np.array([1.0, 2.0])) == 0.0 and obj.f(np.array([0.5, 2.0])) == np.inf  # :9-10
    g = np.ones(2)
    obj.grad(g, np.array([1.0, 2.0]))
    assert not g.any()
    obj.grad(g, np.array([0.5, 2.0]))
    assert np.all(np.isinf(g))  # :12-16
    assert np.array_equal(obj.lower_limit(), np.ones(2) + 1e-8) and np.all(np.isinf(obj.upper_limit()))


def test_basket_liquidation_and_swap(cr):
    with pytest.raises(ValueError):
        cr.BasketLiquidation(3, [0.0, 1.0])  # :20
    obj = cr.BasketLiquidation(1, [0, 1])
    assert obj.f(np.array([2.0, 3.0])) == 3.0  # :24
    assert obj.f(np.array([0.5, 3.0])) == np.inf  # :25
    g = np.zeros(2)
    obj.grad(g, np.array([2.0, 3.0]))
    assert np.array_equal(g, [0.0, 1.0])  # :27-29
    sw = cr.Swap(1, 2, 1.0, 2)
    assert sw.i == 1 and np.array_equal(sw.delta_in, [0.0, 1.0])  # :35-44
    eps = np.sqrt(np.finfo(float).eps)
    assert np.array_equal(obj.lower_limit(), [1 + eps, eps])


def test_pool_descriptors(cr):
    p = cr.ProductTwoCoin([1, 1], 1, [1, 2])
    assert p.R.dtype == np.float64 and len(p) == 2  # ints cast to Float64, cfmms.jl:80-84
    with pytest.raises(ValueError):
        cr.ProductTwoCoin([1, 1], .9, [1])  # ArgumentError, test/cfmms.jl:90
    with pytest.raises(ValueError):
        cr.GeometricMeanTwoCoin([1, 1, 1], [.5, .5], 1, [1, 2])
    u = cr.UniV3(15.0, [30.0, 20, 10, 5], [1.0, 2.0, 1.5, 0.0], 0.997, [1, 2])
    assert u.current_tick == 2  # searchsortedlast rev=true


def test_shard_range(cr):
    for m in (0, 1, 7, 100, 10_000_001):
        for world in (1, 2, 3, 8):
            parts = [cr.shard_range(m, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == m
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


# ---- Router / route! driver over an oracle-backed stand-in ------------------

class OraclePools:
    """Test double with DevicePools' interface, computing with the CPU oracle."""

    def __init__(self, n_tokens, device=0):
        import oracle_lib
        self.o = oracle_lib.load()
        self.n_tokens = n_tokens
        self.parts = []
        self.peer_attached = False
        self._trades = None

    def add_product(self, R, gamma, Ai):
        self.parts.append(("p", np.array(R, float), np.array(gamma, float), np.array(Ai, np.int64)))

    def add_geomean(self, R, gamma, Ai, w):
        self.parts.append(("g", np.array(R, float), np.array(gamma, float), np.array(Ai, np.int64), np.array(w, float)))

    def add_univ3(self, cp, gamma, Ai, off, lt, lq):
        self.parts.append(("u", np.array(cp, float), np.array(gamma, float), np.array(Ai, np.int64),
                           np.array(off, np.int64), np.array(lt, float), np.array(lq, float)))

    def finalize(self):
        pass

    def sweep(self, v, materialize=False):
        Ds, Ls, As = [], [], []
        for part in self.parts:
            if part[0] == "p":
                D, L = self.o.sweep_product(part[1], part[2], part[3], v)
                A = part[3]
            elif part[0] == "g":
                D, L = self.o.sweep_geomean(part[1], part[2], part[3], part[4], v)
                A = part[3]
            else:
                D, L = self.o.sweep_univ3(part[1], part[2], part[3], part[4], part[5], part[6], v)
                A = part[3]
            Ds.append(D), Ls.append(L), As.append(A)
        if not Ds:
            return np.zeros(self.n_tokens), 0.0
        D, L, A = np.concatenate(Ds), np.concatenate(Ls), np.concatenate(As)
        acc, G = self.o.fold(A, D, L, v, self.n_tokens)
        self._trades = (D, L)
        return G, acc

    def trades(self):
        return self._trades

    def update_reserves(self, t, first, R):
        kinds = [p for p in self.parts if p[0] == ("p" if t == 0 else "g")]
        kinds[0][1][first:first + len(R)] = R

    def close(self):
        pass


TOL = 1e-4


def check_primal(cr, r, arb=True):
    flows = np.zeros_like(r.v)
    for D, L, c in zip(r.Δs, r.Λs, r.cfmms):
        assert np.all(D >= -TOL) and np.all(L >= -TOL)
        assert c.phi(c.R + c.gamma * D - L) >= c.phi() - np.sqrt(np.finfo(float).eps)
        flows[c.Ai - 1] += L - D
    assert np.array_equal(flows, cr.netflows(r))
    if arb:
        assert np.all(flows >= -TOL)
    else:
        assert np.sum(flows >= -TOL) == 1


def test_route_driver_readme(cr):
    pools = [cr.ProductTwoCoin([1e6, 1e6], 1, [1, 2]), cr.ProductTwoCoin([1e3, 2e3], 1, [1, 2])]
    r = cr.Router(cr.LinearNonnegative(np.ones(2)), pools, 2, _pools_factory=OraclePools)
    cr.route(r)
    psi = cr.netflows(r)
    # (the absolute ϕ slack of test/arb.jl:11 is only meaningful for the small
    # pools that test uses; with R = 1e6 one ulp of ϕ is already 1e-4)
    assert np.all(r.Δs >= -TOL) and np.all(r.Λs >= -TOL) and np.all(psi >= -TOL)
    assert abs(psi[1] - 171.40) < 0.05 and abs(psi[0]) < 1e-3
    assert np.all(r.v >= r.objective.lower_limit() - TOL)
    # the cache trick (router.jl:74-77, 92-95): one sweep per (fn, g!) pair
    assert r.last_result.nfev < 50


def test_route_driver_mixed_types_and_swap(cr):
    rng = np.random.default_rng(1234)
    pools = []
    for k in range(60):
        Ai = rng.choice(np.arange(1, 11), size=2, replace=False)
        if k % 3 == 2:
            w1 = rng.uniform(0.2, 0.8)
            pools.append(cr.GeometricMeanTwoCoin(1000 * rng.random(2) + 1, [w1, 1 - w1], 1.0, Ai))
        else:
            pools.append(cr.ProductTwoCoin(1000 * rng.random(2) + 1, 1.0, Ai))
    r = cr.Router(cr.LinearNonnegative(rng.random(10) + 1e-2), pools, 10, _pools_factory=OraclePools)
    cr.route(r)
    check_primal(cr, r)
    delta_in = np.concatenate([[0.0], 100 * rng.random(9)])
    prod = [c for c in pools if isinstance(c, cr.ProductTwoCoin)]
    r = cr.Router(cr.BasketLiquidation(1, delta_in), prod, 10, _pools_factory=OraclePools)
    cr.route(r)
    check_primal(cr, r, arb=False)


def test_router_ignores_pools_appended_later(cr):
    eq, sm = cr.ProductTwoCoin([100, 100], 1, [1, 2]), cr.ProductTwoCoin([1, 2], 1, [1, 2])
    r = cr.Router(cr.LinearNonnegative(np.ones(2)), [eq, sm], 2, _pools_factory=OraclePools)
    r.cfmms += [eq, sm]  # test/arb.jl:51
    cr.route(r)
    assert r.Δs.shape == (2, 2)
    check_primal(cr, r)


def test_unknown_pool_type_rejected(cr):
    class Curve(cr.CFMM):
        Ai = np.array([1, 2])
    with pytest.raises(TypeError):
        cr.Router(cr.LinearNonnegative(np.ones(2)), [Curve()], 2, _pools_factory=OraclePools)


def test_flat_pool_file_roundtrip(cr, tmp_path):
    """The flat ingest format (cfmm_pool_file_write / _info): header, sizes, rejection of
    foreign or truncated files.  (Adding a file to a context needs a GPU: tests/test_gpu_parity.py.)"""
    from cfmmrouter_b200 import synth
    R, g, Ai = synth.product_pools(1000, 50, seed=1)
    f = tmp_path / "p.cfmm"
    cr.write_pool_file(f, 50, R, g, Ai)
    assert cr.pool_file_info(f) == (0, 50, 1000)
    assert os.path.getsize(f) == 64 + 1000 * 40
    raw = np.fromfile(f, dtype=np.uint8)
    assert bytes(raw[:8]) == b"CFMMPOOL"
    body = raw[64:].view(np.float64)
    assert np.array_equal(body[:2000].reshape(-1, 2), R) and np.array_equal(body[2000:3000], g)
    assert np.array_equal(raw[64 + 24000:].view(np.int64).reshape(-1, 2), Ai)
    Rg, gg, Ag, wg = synth.geomean_pools(10, 50, seed=2)
    f2 = tmp_path / "g.cfmm"
    cr.write_pool_file(f2, 50, Rg, gg, Ag, wg)
    assert cr.pool_file_info(f2) == (1, 50, 10) and os.path.getsize(f2) == 64 + 10 * 56
    bad = tmp_path / "bad.cfmm"
    bad.write_bytes(raw[:-8].tobytes())          # truncated
    with pytest.raises(OSError):
        cr.pool_file_info(bad)
    bad.write_bytes(b"NOTAPOOL" + raw[8:].tobytes())
    with pytest.raises(OSError):
        cr.pool_file_info(bad)

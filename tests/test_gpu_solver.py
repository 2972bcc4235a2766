"""route! with the outer iteration on the device (cfmm_solve, csrc/solver.cuh; SURVEY §8f
rank 2).  The reference pins nothing at the optimizer boundary beyond feasibility of the
resulting trades (test/arb.jl:3-28, test/swap.jl:2-46): those predicates are restated for
optimizer="device", and the optimal dual value / net flows are compared with the host path
(scipy L-BFGS-B driving one device sweep per evaluation), which is a different algorithm
converging to the same minimiser of the convex dual."""
import numpy as np
import pytest

from test_gpu_parity import TOL, check_dual_feasibility, check_primal_feasibility

pytestmark = pytest.mark.gpu


def dual_value(cr, r):
    """g(ν) at r.v through one host-driven sweep."""
    psi, acc = r._pools.sweep(r.v)
    return r.objective.f(r.v) + acc


def test_device_route_readme_quickstart(cr):
    pools = [cr.ProductTwoCoin([1e6, 1e6], 1, [1, 2]), cr.ProductTwoCoin([1e3, 2e3], 1, [1, 2])]
    r = cr.Router(cr.LinearNonnegative(np.ones(2)), pools, 2)
    cr.route(r, optimizer="device")
    psi = cr.netflows(r)
    # (with R = 1e6 the dual is so flat that its stopping test, relative decrease <= 10 eps, admits
    # |Ψ_1| up to ~1e-3 -- the host path's README test uses the same 1e-3 -- hence not TOL here)
    assert np.all(r.Δs >= -TOL) and np.all(r.Λs >= -TOL) and np.all(psi >= -1e-3), (psi, r.last_result)
    check_dual_feasibility(r)
    # SURVEY App. B (README.md:27-39): Ψ ≈ [0, 171.40].  The stopping rule is the reference's (relative
    # decrease of the dual <= factr·eps), which on this very flat dual leaves |Ψ_1| of a few 1e-3
    assert abs(psi[1] - 171.40) < 0.05 and abs(psi[0]) < 1e-2, (psi, r.last_result)
    assert r.last_result["status"] in (0, 1) and r.last_result["fun_evals"] >= r.last_result["iterations"]


def _random_market(cr, seed, m=100, n=10, fee=1.0):
    rng = np.random.default_rng(seed)
    pools = []
    for _ in range(m):
        Ai = rng.choice(np.arange(1, n + 1), size=2, replace=False)
        pools.append(cr.ProductTwoCoin(1000 * rng.random(2), fee, Ai))
    return rng, pools


def test_device_route_arbitrage_markets(cr):
    # test/arb.jl:42-58
    eq, sm = cr.ProductTwoCoin([100, 100], 1, [1, 2]), cr.ProductTwoCoin([1, 2], 1, [1, 2])
    r = cr.Router(cr.LinearNonnegative(np.ones(2)), [eq, sm], 2)
    cr.route(r, optimizer="device")
    check_primal_feasibility(cr, r)
    check_dual_feasibility(r)
    # test/arb.jl:60-85, and agreement with the host path
    for seed, fee in ((1234, 1.0), (7, 0.997)):
        rng, pools = _random_market(cr, seed, fee=fee)
        c = rng.random(10) + 1e-3
        rd = cr.Router(cr.LinearNonnegative(c), pools, 10)
        cr.route(rd, optimizer="device")
        check_primal_feasibility(cr, rd)
        check_dual_feasibility(rd)
        rh = cr.Router(cr.LinearNonnegative(c), pools, 10)
        cr.route(rh)
        gd, gh = dual_value(cr, rd), dual_value(cr, rh)
        assert abs(gd - gh) <= 1e-6 * max(1.0, abs(gh)), (gd, gh)
        # the arbitrage profit cᵀΨ is the primal optimum: both paths reach it
        pd, ph = float(c @ cr.netflows(rd)), float(c @ cr.netflows(rh))
        assert abs(pd - ph) <= 1e-4 * max(1.0, abs(ph)), (pd, ph)


def test_device_route_swap_markets(cr):
    # test/swap.jl:2-46
    eq, sm = cr.ProductTwoCoin([100, 100], 1, [1, 2]), cr.ProductTwoCoin([1, 2], 1, [1, 2])
    r = cr.Router(cr.BasketLiquidation(1, [5.0, 0.0]), [eq, sm], 2)
    cr.route(r, optimizer="device")
    check_primal_feasibility(cr, r)
    check_dual_feasibility(r)
    rng, pools = _random_market(cr, 1234)
    delta_in = np.concatenate([[0.0], 100 * rng.random(9)])
    rd = cr.Router(cr.BasketLiquidation(1, delta_in), pools, 10)
    cr.route(rd, optimizer="device")
    check_primal_feasibility(cr, rd, arb=False)
    check_dual_feasibility(rd)
    rh = cr.Router(cr.BasketLiquidation(1, delta_in), pools, 10)
    cr.route(rh)
    gd, gh = dual_value(cr, rd), dual_value(cr, rh)
    assert abs(gd - gh) <= 1e-6 * max(1.0, abs(gh)), (gd, gh)
    # the amount of token 1 received is the primal optimum
    od, oh = cr.netflows(rd)[0], cr.netflows(rh)[0]
    assert abs(od - oh) <= 1e-4 * max(1.0, abs(oh)), (od, oh)


def test_device_solver_larger_market_and_options(cr, synth):
    """config-2-shaped market (ProductTwoCoin with fees) through the raw solve() call: bounds are
    respected, the projected gradient meets pgtol or the decrease test fires, maxfun is honoured."""
    n = 300
    R, g, Ai = synth.product_pools(20_000, n, seed=3)
    c = np.random.default_rng(0).random(n) + 0.5
    p = cr.DevicePools(n)
    p.add_product(R, g, Ai)
    p.finalize()
    lower = c + 1e-8
    x, info = p.solve(lower)
    assert info["status"] in (0, 1) and np.all(x >= lower)
    psi, acc = p.sweep(x)
    # KKT of the box-constrained dual: Ψ_i ≈ 0 where ν_i is free, Ψ_i >= 0 where it sits on the bound
    free = x > lower * (1 + 1e-12)
    scale = np.max(np.abs(psi)) + 1.0
    assert np.all(np.abs(psi[free]) <= max(1e-5, info["pg_norm"]) * 10 + 1e-9 * scale)
    assert np.all(psi[~free] >= -1e-5 * 10 - 1e-9 * scale)
    x2, info2 = p.solve(lower, maxfun=3)
    assert info2["status"] == 3 and info2["fun_evals"] <= 4
    D, L = p.trades()   # the trades at the returned ν are materialised
    assert D.shape == (20_000, 2) and np.all(D >= 0) and np.all(L >= 0)
    p.close()

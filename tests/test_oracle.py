"""Pins the CPU oracle (oracle/cfmm_oracle.c) -- runs without a GPU.

1. the reference's own known-answer tests for this path (test/cfmms.jl:74-86),
2. the reference's optimality predicates (test/cfmms.jl:3-22 two-coin, :25-56
   UniV3) on seeded random pools and on its UniV3 scenarios (:117-201),
3. 50-digit mpmath goldens (tests/golden/closed_forms.json).
"""
import json
import os

import numpy as np
import pytest

SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "closed_forms.json")


# ---------------------------------------------------------------------------
# predicates restated from the reference's tests
# ---------------------------------------------------------------------------

def phi_product(R):
    return R[0] * R[1]


def grad_phi_product(R):
    return np.array([R[1], R[0]])


def phi_geomean(R, w):
    return R[0] ** w[0] * R[1] ** w[1]


def grad_phi_geomean(R, w):
    return np.array([w[0] * (R[1] / R[0]) ** w[1], w[1] * (R[0] / R[1]) ** w[0]])


def optimality_conditions_met(c, D, L, R, gamma, phi, grad_phi):
    """test_optimality_conditions_met, test/cfmms.jl:3-22."""
    Rp = R + gamma * D - L
    pfeas = np.all(D >= 0) and np.all(L >= 0)
    pR, pRp = phi(R), phi(Rp)
    g = grad_phi(Rp)
    # ≈ is isapprox with rtol = sqrt(eps)
    cfmm_sat = abs(pR - pRp) <= SQRT_EPS * max(abs(pR), abs(pRp)) and pRp >= pR - SQRT_EPS
    opt = max(gamma * g[i] / c[i] for i in range(2)) <= min(g[i] / c[i] for i in range(2)) + SQRT_EPS
    return pfeas and cfmm_sat and opt


def univ3_conditions_met(o, c, D, L, cp, lt, lq, gamma):
    """test_optimality_conditions_met(c, Δ, Λ, cfmm::UniV3), test/cfmms.jl:25-56,
    with ForwardDiff replaced by a central difference of forward_trade."""
    p_opt = c[0] / c[1]
    q = cp
    if gamma * q <= p_opt <= q / gamma:
        return D[0] == 0 and D[1] == 0
    ft = lambda d: o.univ3_forward_trade(cp, lt, lq, gamma, d)
    if p_opt > lt[0]:
        lam = ft(D)
        return np.isclose(lam, L[0], rtol=SQRT_EPS) and L[1] == 0
    if p_opt < lt[-1] and lq[-1] == 0:
        lam = ft(D)
        return np.isclose(lam, L[1], rtol=SQRT_EPS) and L[0] == 0
    j = 0 if q > p_opt else 1
    h = 1e-6 * max(D[j], 1e-3)
    dp, dm = D.copy(), D.copy()
    dp[j] += h
    dm[j] -= h
    impact = (ft(dp) - ft(dm)) / (2 * h)
    target = p_opt if j == 0 else 1 / p_opt
    return abs(impact - target) <= 1e-6


# ---------------------------------------------------------------------------
# 1. reference KATs
# ---------------------------------------------------------------------------

def test_product_kats(oracle):
    # test/cfmms.jl:74-76, 78-80: no arb in the fee-less case
    for v in ([1.0, 1.0], [2.0, 2.0]):
        D, L = oracle.product_arb([1, 1], 1, v)
        assert not D.any() and not L.any()
    # test/cfmms.jl:82-86: easy arb
    D, L = oracle.product_arb([1, 1], 1, [2.0, 1.0])
    assert abs(D[0]) < 1e-15 and np.isclose(D[1], np.sqrt(2) - 1, rtol=SQRT_EPS)
    assert np.isclose(L[0], 1 - np.sqrt(0.5), rtol=SQRT_EPS) and abs(L[1]) < 1e-15
    # SURVEY Appendix B values
    assert D[1] == 0.41421356237309515 and L[0] == 0.2928932188134524


def test_geomean_equal_weights_is_product(oracle):
    rng = np.random.default_rng(5)
    for _ in range(50):
        R = 1000 * rng.random(2) + 1
        v = rng.random(2) + 0.1
        Dp, Lp = oracle.product_arb(R, 0.997, v)
        Dg, Lg = oracle.geomean_arb(R, [0.5, 0.5], 0.997, v)
        np.testing.assert_allclose(Dg, Dp, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(Lg, Lp, rtol=1e-12, atol=1e-12)


def test_univ3_current_tick(oracle):
    lt = [30.0, 20.0, 10.0, 5.0]
    assert oracle.univ3_current_tick(lt, 15.0) == 2
    assert oracle.univ3_current_tick(lt, 30.0) == 1
    assert oracle.univ3_current_tick(lt, 31.0) == 0
    assert oracle.univ3_current_tick(lt, 5.0) == 4
    assert oracle.univ3_current_tick(lt, 1.0) == 4


def test_univ3_tick_tuples(oracle):
    # tick tuples (k, α, β, R1, R2) for cp=15
    cp, lt, lq = 15.0, [30.0, 20.0, 10.0, 5.0], [1.0, 2.0, 1.5, 0.0]
    exp = [(1, .182574, 4.472136, .041033, 0), (2, .316228, 4.472136, .048921, 1.005090),
           (1.5, .387298, 2.738613, 0, 1.134371), (0, 0, 0, 0, 0)]
    for i, e in enumerate(exp):
        np.testing.assert_allclose(oracle.univ3_tick(cp, lt, lq, i + 1), e, atol=2e-6)


# ---------------------------------------------------------------------------
# 2. the reference's predicates
# ---------------------------------------------------------------------------

def test_product_optimality_random(oracle):
    # test/cfmms.jl:64-68, 92-96 (3x3x3 there; 6x6x6 here, own seed)
    rng = np.random.default_rng(1234)
    gammas = rng.random(6) * 0.98 + 0.01
    Rs = [rng.random(2) * 10 + 1e-3 for _ in range(6)]
    nus = [rng.random(2) + 1e-3 for _ in range(6)]
    for R in Rs:
        for g in gammas:
            for nu in nus:
                D, L = oracle.product_arb(R, g, nu)
                assert optimality_conditions_met(nu, D, L, R, g, phi_product, grad_phi_product)


def test_geomean_optimality_random(oracle):
    # test/cfmms.jl:100-107
    rng = np.random.default_rng(4321)
    gammas = rng.random(4) * 0.98 + 0.01
    Rs = [rng.random(2) * 10 + 1e-3 for _ in range(4)]
    nus = [rng.random(2) + 1e-3 for _ in range(4)]
    ws = [np.array([w1, 1 - w1]) for w1 in rng.uniform(0.02, 0.98, size=4)]
    for R in Rs:
        for g in gammas:
            for nu in nus:
                for w in ws:
                    D, L = oracle.geomean_arb(R, w, g, nu)
                    assert optimality_conditions_met(
                        nu, D, L, R, g, lambda r: phi_geomean(r, w), lambda r: grad_phi_geomean(r, w))


@pytest.mark.parametrize("gamma", [1.0, 0.997])
def test_univ3_reference_scenarios(oracle, gamma):
    # test/cfmms.jl:117-201
    cp, lt, lq = 15.0, np.array([30.0, 20, 10, 5]), np.array([1.0, 2.0, 1.5, 0.0])
    first = [15.0, 1.0] if gamma == 1.0 else [15.0 * (1 + gamma) / 2, 1.0]
    for v in (first, [16.0, 1.0], [14.0, 1.0], [25.0, 1.0], [7.5, 1.0], [4.0, 1.0], [35.0, 1.0]):
        v = np.array(v)
        D, L = oracle.univ3_arb(cp, lt, lq, gamma, v)
        assert univ3_conditions_met(oracle, v, D, L, cp, lt, lq, gamma), (gamma, v, D, L)


def test_univ3_drained_equals_tick_reserves(oracle):
    # SURVEY App. B sanity: drained Λ is the sum of the tick reserves
    cp, lt, lq = 15.0, [30.0, 20, 10, 5], [1.0, 2.0, 1.5, 0.0]
    _, L = oracle.univ3_arb(cp, lt, lq, 1.0, [4.0, 1.0])
    t2, t3 = oracle.univ3_tick(cp, lt, lq, 2), oracle.univ3_tick(cp, lt, lq, 3)
    assert np.isclose(L[1], t2[4] + t3[4], rtol=1e-14)


# ---------------------------------------------------------------------------
# 3. mpmath goldens
# ---------------------------------------------------------------------------

@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def _close(got, want_str, scale, ulps):
    want = np.array([float(s) for s in want_str])
    tol = ulps * np.finfo(np.float64).eps * scale
    return np.all(np.abs(got - want) <= tol), got, want


def test_golden_product(oracle, golden):
    for case in golden["product"]:
        D, L = oracle.product_arb(case["R"], case["gamma"], case["v"])
        scale = max(case["R"]) / min(case["gamma"], 1.0)
        # a handful of correctly rounded ops: within a few ulp of the operand scale
        for got, want in ((D, case["Delta"]), (L, case["Lambda"])):
            ok, g, w = _close(got, want, scale, 8)
            assert ok, (case, g, w)


def test_golden_geomean(oracle, golden):
    for case in golden["geomean"]:
        D, L = oracle.geomean_arb(case["R"], case["w"], case["gamma"], case["v"])
        scale = max(case["R"]) / min(case["gamma"], 1.0)
        for got, want in ((D, case["Delta"]), (L, case["Lambda"])):
            ok, g, w = _close(got, want, scale, 64)
            assert ok, (case, g, w)


def test_golden_univ3(oracle, golden):
    for case in golden["univ3"]:
        D, L = oracle.univ3_arb(case["cp"], case["lower_ticks"], case["liquidity"], case["gamma"], case["v"])
        wantD = np.array([float(s) for s in case["Delta"]])
        wantL = np.array([float(s) for s in case["Lambda"]])
        scale = max(1.0, np.max(wantD), np.max(wantL))
        np.testing.assert_allclose(D, wantD, rtol=0, atol=1e-12 * scale, err_msg=str(case))
        np.testing.assert_allclose(L, wantL, rtol=0, atol=1e-12 * scale, err_msg=str(case))


# ---------------------------------------------------------------------------
# router-level restatement (router.jl:38-42, 79-83, 98-100)
# ---------------------------------------------------------------------------

def test_sweep_and_fold_match_per_pool(oracle):
    import cfmmrouter_b200  # noqa: F401  (path set-up)
    from cfmmrouter_b200 import synth
    n = 50
    R, g, Ai = synth.product_pools(500, n)
    v = synth.dual_prices(n, "wide")
    D, L = oracle.sweep_product(R, g, Ai, v, threads=1)
    D4, L4 = oracle.sweep_product(R, g, Ai, v, threads=4)
    assert np.array_equal(D, D4) and np.array_equal(L, L4)
    acc, G = oracle.fold(Ai, D, L, v, n)
    a2, G2 = 0.0, np.zeros(n)
    for i in range(len(g)):
        d, l = oracle.product_arb(R[i], g[i], v[Ai[i] - 1])
        assert np.array_equal(d, D[i]) and np.array_equal(l, L[i])
        vi = v[Ai[i] - 1]
        a2 += (l[0] * vi[0] + l[1] * vi[1]) - (d[0] * vi[0] + d[1] * vi[1])
        G2[Ai[i] - 1] += l - d
    assert acc == a2 and np.array_equal(G, G2)
    accx, Gx, absG = oracle.fold_compensated(Ai, D, L, v, n)
    assert abs(float(accx) - acc) <= 1e-12 * np.sum(absG * v)
    assert np.all(np.abs(Gx.astype(np.float64) - G) <= 1e-13 * absG + 1e-300)
    # the timing flavours compute the same thing
    f = oracle.faithful(n)
    f.add_product(R, g, Ai)
    accf, Gf = f.sweep(v, threads=2)
    f.close()
    assert accf == acc and np.array_equal(Gf, G)
    accs, Gs = oracle.soa_sweep_product(R, g, Ai, v, n, threads=3)
    assert np.isclose(accs, acc, rtol=1e-12) and np.allclose(Gs, G, rtol=1e-12, atol=1e-9)


# ---------------------------------------------------------------------------
# property tests (hypothesis): the reference's predicate on arbitrary pools
# ---------------------------------------------------------------------------

try:
    from hypothesis import given, settings, strategies as st

    pos = st.floats(min_value=1e-3, max_value=1e6, allow_nan=False, allow_infinity=False)
    fee = st.floats(min_value=0.5, max_value=1.0)

    @settings(max_examples=300, deadline=None, derandomize=True)
    @given(R1=pos, R2=pos, gamma=fee, v1=pos, v2=pos)
    def test_product_optimality_hypothesis(R1, R2, gamma, v1, v2):
        import oracle_lib
        o = oracle_lib.load()
        R, nu = np.array([R1, R2]), np.array([v1, v2])
        D, L = o.product_arb(R, gamma, nu)
        # feasibility + at most one side trades + KKT (test/cfmms.jl:3-22; the ≈ / sqrt(eps)
        # slacks of that predicate are absolute, so scale ϕ to O(1) first)
        assert np.all(D >= 0) and np.all(L >= 0)
        assert not (D[0] > 0 and D[1] > 0) and not (L[0] > 0 and L[1] > 0)
        Rp = R + gamma * D - L
        assert abs(Rp[0] * Rp[1] - R1 * R2) <= 1e-9 * R1 * R2
        g = np.array([Rp[1], Rp[0]])
        assert max(gamma * g[i] / nu[i] for i in range(2)) <= min(g[i] / nu[i] for i in range(2)) * (1 + 1e-9)

    @settings(max_examples=200, deadline=None, derandomize=True)
    @given(R1=pos, R2=pos, gamma=fee, v1=pos, v2=pos, w1=st.floats(min_value=0.05, max_value=0.95))
    def test_geomean_optimality_hypothesis(R1, R2, gamma, v1, v2, w1):
        import oracle_lib
        o = oracle_lib.load()
        R, nu, w = np.array([R1, R2]), np.array([v1, v2]), np.array([w1, 1 - w1])
        D, L = o.geomean_arb(R, w, gamma, nu)
        assert np.all(D >= 0) and np.all(L >= 0)
        Rp = R + gamma * D - L
        # (when a trade nearly drains a reserve, R − Λ cancels and the closed form's own
        # rounding is amplified: the tolerances are relative to what is left of the reserve)
        amp = float(np.max(R / np.maximum(Rp, 1e-300)))
        phi0, phi1 = phi_geomean(R, w), phi_geomean(Rp, w)
        assert abs(phi1 - phi0) <= 1e-12 * amp * phi0 + 1e-9 * phi0
        g = grad_phi_geomean(Rp, w)
        assert max(gamma * g[i] / nu[i] for i in range(2)) <= \
            min(g[i] / nu[i] for i in range(2)) * (1 + 1e-12 * amp + 1e-9)
except ImportError:  # hypothesis is optional
    pass

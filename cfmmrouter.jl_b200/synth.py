"""Synthetic pool sets for the parity tests and bench.py.

Template: the reference's own benchmark generator (benchmark/scaling.jl:13-38)
and its random-market tests (test/arb.jl:60-78):  R = 1000*rand(2),
γ = rand((0.997, 1.0)), Ai = sample(1:n, 2, replace=false),
LinearNonnegative(rand(n)).  Julia's RNG stream is not reproducible outside
Julia, so inputs are drawn from numpy's PCG64 with a fixed seed instead.
"""
from __future__ import annotations

import numpy as np


def token_pairs(rng, m: int, n_tokens: int):
    """m pairs of distinct 1-based token ids, uniform (scaling.jl:22)."""
    a = rng.integers(1, n_tokens + 1, size=m, dtype=np.int64)
    b = rng.integers(1, n_tokens, size=m, dtype=np.int64)
    b = b + (b >= a)  # skip a: uniform over the other n-1 tokens
    return np.stack([a, b], axis=1)


def product_pools(m: int, n_tokens: int, seed: int = 1234):
    """(R [m,2], gamma [m], Ai [m,2] 1-based)."""
    rng = np.random.default_rng(seed)
    R = 1000.0 * rng.random((m, 2))
    R = np.maximum(R, 1e-3)  # rand() can return exactly 0; a pool has reserves
    gamma = rng.choice(np.array([0.997, 1.0]), size=m)
    return R, gamma, token_pairs(rng, m, n_tokens)


def product_pools_skewed(m: int, n_tokens: int, alpha: float = 1.0, seed: int = 1234):
    """Like product_pools, but token popularity follows a Zipf law (rank^-alpha):
    a few hub tokens (the WETH / USDC of a real DEX graph) sit in a large share
    of the pools, on either side.  The reference's own generators are uniform;
    this one exists to exercise hub contention."""
    rng = np.random.default_rng(seed)
    R = np.maximum(1000.0 * rng.random((m, 2)), 1e-3)
    gamma = rng.choice(np.array([0.997, 1.0]), size=m)
    w = 1.0 / np.arange(1, n_tokens + 1) ** alpha
    w /= w.sum()
    a = rng.choice(n_tokens, size=m, p=w) + 1
    b = rng.choice(n_tokens, size=m, p=w) + 1
    clash = a == b
    b[clash] = (a[clash] % n_tokens) + 1
    return R, gamma, np.stack([a, b], axis=1).astype(np.int64)


def geomean_pools(m: int, n_tokens: int, seed: int = 4321):
    """(R, gamma, Ai, w) with w1 ~ U(0.05, 0.95), w2 = 1 - w1 (test/cfmms.jl:101)."""
    rng = np.random.default_rng(seed)
    R = np.maximum(1000.0 * rng.random((m, 2)), 1e-3)
    gamma = rng.choice(np.array([0.997, 1.0]), size=m)
    w1 = rng.uniform(0.05, 0.95, size=m)
    w = np.stack([w1, 1.0 - w1], axis=1)
    return R, gamma, token_pairs(rng, m, n_tokens), w


def univ3_pools(m: int, n_tokens: int, seed: int = 777, ragged: bool = False):
    """Pools shaped like examples/Univ3.jl:11-15 (4 ticks: cp·[2, 4/3, 2/3, 1/3],
    liquidity [1, 2, 1.5, 0]·s); with ragged=True the tick count varies 1..16.
    Returns (current_price, gamma, Ai, tick_off, lower_ticks, liquidity)."""
    rng = np.random.default_rng(seed)
    cp = np.exp(rng.uniform(np.log(0.1), np.log(10.0), size=m))
    s = rng.uniform(1.0, 1000.0, size=m)
    gamma = np.full(m, 0.997)
    Ai = token_pairs(rng, m, n_tokens)
    if not ragged:
        lower = cp[:, None] * np.array([2.0, 4.0 / 3.0, 2.0 / 3.0, 1.0 / 3.0])[None, :]
        liq = s[:, None] * np.array([1.0, 2.0, 1.5, 0.0])[None, :]
        off = np.arange(m + 1, dtype=np.int64) * 4
        return cp, gamma, Ai, off, lower.reshape(-1), liq.reshape(-1)
    T = rng.integers(1, 17, size=m)
    off = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
    lower = np.empty(off[-1])
    liq = np.empty(off[-1])
    for i in range(m):
        t = int(T[i])
        # strictly decreasing ladder whose first rung is above the current price
        rungs = cp[i] * 2.0 * np.cumprod(np.concatenate([[1.0], rng.uniform(0.5, 0.9, size=t - 1)]))
        lower[off[i]:off[i + 1]] = rungs
        lq = s[i] * rng.uniform(0.0, 2.0, size=t)
        lq[rng.random(t) < 0.15] = 0.0  # some empty ticks (skipped, not terminal)
        liq[off[i]:off[i + 1]] = lq
    return cp, gamma, Ai, off, lower, liq


def objective_prices(n_tokens: int, seed: int = 99):
    """c = rand(n) for LinearNonnegative (scaling.jl:31), kept away from 0."""
    rng = np.random.default_rng(seed)
    return rng.uniform(0.05, 1.0, size=n_tokens)


def dual_prices(n_tokens: int, kind: str = "near", seed: int = 7):
    """ν for timing/parity: 'ones' = the benchmark's start (scaling.jl:16);
    'near' = c·(1 + 0.05·U(0,1)), an in-bounds, near-optimal point;
    'wide' = LogU(0.2, 5), every pool far from its no-trade band."""
    rng = np.random.default_rng(seed)
    if kind == "ones":
        return np.ones(n_tokens)
    if kind == "near":
        return objective_prices(n_tokens) * (1.0 + 0.05 * rng.random(n_tokens))
    if kind == "wide":
        return np.exp(rng.uniform(np.log(0.2), np.log(5.0), size=n_tokens))
    raise ValueError(kind)

"""CFMM pool types -- host-side descriptors mirroring src/cfmms.jl.  They carry
the same fields the reference structs do (R, γ, Ai, w / current_price,
lower_ticks, liquidity); the arithmetic lives in the CUDA kernels."""
from __future__ import annotations

import numpy as np

__all__ = ["CFMM", "ProductTwoCoin", "GeometricMeanTwoCoin", "UniV3"]


def _two_coin_check_cast(R, gamma, idx):
    """two_coin_check_cast (src/cfmms.jl:76-90)."""
    R = np.asarray(R)
    idx = np.asarray(idx)
    if R.shape != (2,):
        raise ValueError("length of R must be 2 for *TwoCoin constructors")
    if idx.shape != (2,):
        raise ValueError("length of idx must be 2 for *TwoCoin constructors")
    if np.any(idx < 0):
        raise ValueError("idx must be convertible to UInt")  # convert.(UInt, idx), :87
    return R.astype(np.float64), float(gamma), idx.astype(np.int64)


class CFMM:
    """abstract type CFMM{T} (src/cfmms.jl:5)"""

    def __len__(self):  # Base.length(c::CFMM), cfmms.jl:19
        return len(self.Ai)


class ProductTwoCoin(CFMM):
    """ProductTwoCoin(R, γ, idx): φ(R) = R1·R2 (src/cfmms.jl:101-111).
    `idx` is 1-based, as in the reference."""

    def __init__(self, R, gamma, idx):
        self.R, self.gamma, self.Ai = _two_coin_check_cast(R, gamma, idx)

    def phi(self, R=None):  # ϕ, cfmms.jl:113-116
        R = self.R if R is None else R
        return R[0] * R[1]

    def grad_phi(self, R=None):  # ∇ϕ!, cfmms.jl:117-122
        R = self.R if R is None else R
        return np.array([R[1], R[0]])


class GeometricMeanTwoCoin(CFMM):
    """GeometricMeanTwoCoin(R, w, γ, idx): φ(R) = R1^w1·R2^w2 (src/cfmms.jl:152-165)."""

    def __init__(self, R, w, gamma, idx):
        self.R, self.gamma, self.Ai = _two_coin_check_cast(R, gamma, idx)
        w = np.asarray(w, dtype=np.float64)
        if w.shape != (2,):
            raise ValueError("length of w must be 2")
        self.w = w

    def phi(self, R=None):  # cfmms.jl:167-171
        R = self.R if R is None else R
        return R[0] ** self.w[0] * R[1] ** self.w[1]

    def grad_phi(self, R=None):  # cfmms.jl:172-178
        R = self.R if R is None else R
        w = self.w
        return np.array([w[0] * (R[1] / R[0]) ** w[1], w[1] * (R[0] / R[1]) ** w[0]])


class UniV3(CFMM):
    """UniV3(current_price, lower_ticks, liquidity, γ, Ai) (src/cfmms.jl:226-245).
    lower_ticks in decreasing order; current_tick (1-based) is derived exactly
    as the reference ctor does (searchsortedlast rev=true, cfmms.jl:235)."""

    def __init__(self, current_price, lower_ticks, liquidity, gamma, Ai):
        self.current_price = float(current_price)
        self.lower_ticks = np.asarray(lower_ticks, dtype=np.float64)
        self.liquidity = np.asarray(liquidity, dtype=np.float64)
        if self.lower_ticks.shape != self.liquidity.shape or self.lower_ticks.ndim != 1:
            raise ValueError("lower_ticks and liquidity must be vectors of equal length")
        self.gamma = float(gamma)
        self.Ai = np.asarray(Ai, dtype=np.int64)
        if self.Ai.shape != (2,):
            raise ValueError("length of Ai must be 2")
        self.current_tick = int(np.sum(self.lower_ticks >= self.current_price))

"""Objectives of the routing problem -- host side, unchanged in spirit from the
reference (src/objectives.jl): the conjugate f(ν), its gradient and the box
bounds on ν.  O(n_tokens) numpy work per L-BFGS-B evaluation; stays on the host
by design (BASELINE north_star)."""
from __future__ import annotations

import numpy as np

__all__ = ["Objective", "LinearNonnegative", "BasketLiquidation", "Swap"]


class Objective:
    """abstract type Objective (src/objectives.jl:3)"""

    def f(self, v):
        raise NotImplementedError

    def grad(self, g, v):
        raise NotImplementedError

    def lower_limit(self):
        raise NotImplementedError

    def upper_limit(self):
        raise NotImplementedError

    def linear_term(self):
        """lin such that f(ν) = linᵀν on the box [lower_limit, upper_limit] (None: the objective is
        not of that form and the device solver cannot be used)."""
        return None


class LinearNonnegative(Objective):
    """U(Ψ) = cᵀΨ − I(Ψ ≥ 0)  (src/objectives.jl:51-79)."""

    def __init__(self, c):
        c = np.array(c, dtype=np.float64)
        if c.ndim != 1 or not np.all(c > 0):
            # ArgumentError("all elements must be strictly positive"), objectives.jl:54
            raise ValueError("all elements must be strictly positive")
        self.c = c

    def f(self, v):  # objectives.jl:62-67
        return 0.0 if np.all(self.c <= v) else np.inf

    def grad(self, g, v):  # objectives.jl:69-76
        g[:] = 0.0 if np.all(self.c <= v) else np.inf

    def lower_limit(self):  # objectives.jl:78
        return self.c + 1e-8

    def linear_term(self):  # f = 0 on the box
        return np.zeros_like(self.c)

    def upper_limit(self):  # objectives.jl:79
        return np.full_like(self.c, np.inf)


class BasketLiquidation(Objective):
    """Ψ_i − I(Ψ_{-i} + Δin_{-i} = 0, Ψ_i ≥ 0)  (src/objectives.jl:92-129).
    `i` is 1-based, as in the reference."""

    def __init__(self, i, delta_in):
        delta_in = np.array(delta_in, dtype=np.float64)
        if not (0 < i <= len(delta_in)):
            raise ValueError("Invalid index i")  # objectives.jl:97
        self.i = int(i)
        self.delta_in = delta_in

    def f(self, v):  # objectives.jl:106-111
        if v[self.i - 1] >= 1.0:
            s = 0.0
            for j in range(len(v)):
                s += 0.0 if j == self.i - 1 else self.delta_in[j] * v[j]
            return s
        return np.inf

    def grad(self, g, v):  # objectives.jl:113-121
        if v[self.i - 1] >= 1.0:
            g[:] = self.delta_in
            g[self.i - 1] = 0.0
        else:
            g[:] = np.inf

    def linear_term(self):  # f = Σ_{j≠i} Δin_j ν_j on the box (ν_i >= 1 there)
        lin = self.delta_in.copy()
        lin[self.i - 1] = 0.0
        return lin

    def lower_limit(self):  # objectives.jl:123-128
        eps = np.sqrt(np.finfo(np.float64).eps)
        ret = np.full(len(self.delta_in), eps)
        ret[self.i - 1] = 1.0 + eps
        return ret

    def upper_limit(self):  # objectives.jl:129
        return np.full(len(self.delta_in), np.inf)


def Swap(i, j, delta, n):
    """Swap(i, j, δ, n): one-hot BasketLiquidation (src/objectives.jl:142-146)."""
    delta_in = np.zeros(n)
    delta_in[j - 1] = delta
    return BasketLiquidation(i, delta_in)

"""ctypes wrapper of oracle/liboracle.so (the CPU oracle; test infrastructure).
Builds the library with oracle/Makefile if it is missing."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def _i64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a if shape is None else a.reshape(shape)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.oracle_product_arb.argtypes = [_dp, C.c_double, _dp, _dp, _dp]
        L.oracle_geomean_arb.argtypes = [_dp, _dp, C.c_double, _dp, _dp, _dp]
        L.oracle_univ3_current_tick.restype = C.c_int64
        L.oracle_univ3_current_tick.argtypes = [_dp, C.c_int64, C.c_double]
        L.oracle_univ3_arb.argtypes = [C.c_double, C.c_int64, _dp, _dp, C.c_int64, C.c_double, _dp, _dp, _dp]
        L.oracle_univ3_forward_trade.restype = C.c_double
        L.oracle_univ3_forward_trade.argtypes = [C.c_double, C.c_int64, _dp, _dp, C.c_int64, C.c_double, _dp]
        L.oracle_univ3_tick.argtypes = [C.c_double, C.c_int64, _dp, _dp, C.c_int64, C.c_int64, _dp]
        L.oracle_sweep_product.argtypes = [C.c_int64, _dp, _dp, _ip, _dp, _dp, _dp, C.c_int]
        L.oracle_sweep_geomean.argtypes = [C.c_int64, _dp, _dp, _ip, _dp, _dp, _dp, _dp, C.c_int]
        L.oracle_sweep_univ3.argtypes = [C.c_int64, _dp, _dp, _ip, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int]
        L.oracle_fold.argtypes = [C.c_int64, _ip, _dp, _dp, _dp, _dp, _dp]
        L.oracle_fold_compensated.argtypes = [C.c_int64, _ip, _dp, _dp, _dp, C.c_void_p, C.c_void_p, _dp]
        L.oracle_faithful_create.restype = C.c_void_p
        L.oracle_faithful_create.argtypes = [C.c_int64]
        L.oracle_faithful_add_product.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _ip]
        L.oracle_faithful_add_geomean.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _ip, _dp]
        L.oracle_faithful_sweep.restype = C.c_double
        L.oracle_faithful_sweep.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.oracle_faithful_destroy.argtypes = [C.c_void_p]
        L.oracle_soa_sweep_product.restype = C.c_double
        L.oracle_soa_sweep_product.argtypes = [C.c_int64, _dp, _dp, _ip, _dp, C.c_int64, _dp, C.c_int]
        L.oracle_max_threads.restype = C.c_int

    # ---- per pool ----
    def product_arb(self, R, gamma, v):
        R, v = _f64(R), _f64(v)
        D, L = np.zeros(2), np.zeros(2)
        self.lib.oracle_product_arb(_d(R), float(gamma), _d(v), _d(D), _d(L))
        return D, L

    def geomean_arb(self, R, w, gamma, v):
        R, w, v = _f64(R), _f64(w), _f64(v)
        D, L = np.zeros(2), np.zeros(2)
        self.lib.oracle_geomean_arb(_d(R), _d(w), float(gamma), _d(v), _d(D), _d(L))
        return D, L

    def univ3_current_tick(self, lower_ticks, cp):
        lt = _f64(lower_ticks)
        return int(self.lib.oracle_univ3_current_tick(_d(lt), len(lt), float(cp)))

    def univ3_arb(self, cp, lower_ticks, liquidity, gamma, v):
        lt, lq, v = _f64(lower_ticks), _f64(liquidity), _f64(v)
        ct = self.univ3_current_tick(lt, cp)
        D, L = np.zeros(2), np.zeros(2)
        self.lib.oracle_univ3_arb(float(cp), ct, _d(lt), _d(lq), len(lt), float(gamma), _d(v), _d(D), _d(L))
        return D, L

    def univ3_forward_trade(self, cp, lower_ticks, liquidity, gamma, Delta):
        lt, lq, Dl = _f64(lower_ticks), _f64(liquidity), _f64(Delta)
        ct = self.univ3_current_tick(lt, cp)
        return float(self.lib.oracle_univ3_forward_trade(float(cp), ct, _d(lt), _d(lq), len(lt), float(gamma), _d(Dl)))

    def univ3_tick(self, cp, lower_ticks, liquidity, idx):
        lt, lq = _f64(lower_ticks), _f64(liquidity)
        ct = self.univ3_current_tick(lt, cp)
        out = np.zeros(5)
        self.lib.oracle_univ3_tick(float(cp), ct, _d(lt), _d(lq), len(lt), int(idx), _d(out))
        return out

    # ---- sweeps (router.jl:38-42) ----
    def sweep_product(self, R, gamma, Ai, v, threads=1):
        R, gamma, Ai, v = _f64(R, (-1, 2)), _f64(gamma), _i64(Ai, (-1, 2)), _f64(v)
        m = len(gamma)
        D, L = np.zeros((m, 2)), np.zeros((m, 2))
        self.lib.oracle_sweep_product(m, _d(R), _d(gamma), _i(Ai), _d(v), _d(D), _d(L), threads)
        return D, L

    def sweep_geomean(self, R, gamma, Ai, w, v, threads=1):
        R, gamma, Ai, w, v = _f64(R, (-1, 2)), _f64(gamma), _i64(Ai, (-1, 2)), _f64(w, (-1, 2)), _f64(v)
        m = len(gamma)
        D, L = np.zeros((m, 2)), np.zeros((m, 2))
        self.lib.oracle_sweep_geomean(m, _d(R), _d(gamma), _i(Ai), _d(w), _d(v), _d(D), _d(L), threads)
        return D, L

    def sweep_univ3(self, cp, gamma, Ai, tick_off, lower, liq, v, threads=1):
        cp, gamma, Ai = _f64(cp), _f64(gamma), _i64(Ai, (-1, 2))
        off, lower, liq, v = _i64(tick_off), _f64(lower), _f64(liq), _f64(v)
        m = len(gamma)
        D, L = np.zeros((m, 2)), np.zeros((m, 2))
        self.lib.oracle_sweep_univ3(m, _d(cp), _d(gamma), _i(Ai), _i(off), _d(lower), _d(liq), _d(v), _d(D), _d(L), threads)
        return D, L

    # ---- folds (router.jl:79-83, 98-100) ----
    def fold(self, Ai, D, L, v, n_tokens, acc=0.0, G=None):
        Ai, D, L, v = _i64(Ai, (-1, 2)), _f64(D, (-1, 2)), _f64(L, (-1, 2)), _f64(v)
        if G is None:
            G = np.zeros(n_tokens)
        a = C.c_double(acc)
        self.lib.oracle_fold(len(Ai), _i(Ai), _d(D), _d(L), _d(v), C.byref(a), _d(G))
        return float(a.value), G

    def fold_compensated(self, Ai, D, L, v, n_tokens):
        """(acc, Psi, abs_Psi): extended-precision pool-order sums."""
        Ai, D, L, v = _i64(Ai, (-1, 2)), _f64(D, (-1, 2)), _f64(L, (-1, 2)), _f64(v)
        G = np.zeros(n_tokens, dtype=np.longdouble)
        acc = np.zeros(1, dtype=np.longdouble)
        absG = np.zeros(n_tokens)
        self.lib.oracle_fold_compensated(len(Ai), _i(Ai), _d(D), _d(L), _d(v),
                                         acc.ctypes.data, G.ctypes.data, _d(absG))
        return acc[0], G, absG

    # ---- timing baselines ----
    def max_threads(self):
        return int(self.lib.oracle_max_threads())

    def faithful(self, n_tokens):
        return Faithful(self, n_tokens)

    def soa_sweep_product(self, R, gamma, Ai, v, n_tokens, threads):
        R, gamma, Ai, v = _f64(R, (-1, 2)), _f64(gamma), _i64(Ai, (-1, 2)), _f64(v)
        G = np.zeros(n_tokens)
        acc = self.lib.oracle_soa_sweep_product(len(gamma), _d(R), _d(gamma), _i(Ai), _d(v), n_tokens, _d(G), threads)
        return float(acc), G


class Faithful:
    def __init__(self, o: Oracle, n_tokens):
        self.o, self.n = o, n_tokens
        self.h = o.lib.oracle_faithful_create(n_tokens)

    def add_product(self, R, gamma, Ai):
        R, gamma, Ai = _f64(R, (-1, 2)), _f64(gamma), _i64(Ai, (-1, 2))
        self.o.lib.oracle_faithful_add_product(self.h, len(gamma), _d(R), _d(gamma), _i(Ai))

    def add_geomean(self, R, gamma, Ai, w):
        R, gamma, Ai, w = _f64(R, (-1, 2)), _f64(gamma), _i64(Ai, (-1, 2)), _f64(w, (-1, 2))
        self.o.lib.oracle_faithful_add_geomean(self.h, len(gamma), _d(R), _d(gamma), _i(Ai), _d(w))

    def sweep(self, v, threads=1):
        v = _f64(v)
        G = np.zeros(self.n)
        acc = self.o.lib.oracle_faithful_sweep(self.h, _d(v), _d(G), threads)
        return float(acc), G

    def close(self):
        if self.h:
            self.o.lib.oracle_faithful_destroy(self.h)
            self.h = None


_oracle = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def load() -> Oracle:
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _oracle = Oracle(C.CDLL(ORACLE_SO))
    return _oracle

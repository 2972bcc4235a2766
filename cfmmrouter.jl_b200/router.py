"""Router / route! / find_arb! / netflows -- the host-side mirror of
src/router.jl, with the inner loop (the find_arb! sweep over every pool and the
two fold loops of the L-BFGS-B callback) replaced by libcfmm_b200.so.

What stays on the host, as in BASELINE's north_star: the Objective
(objectives.py) and the L-BFGS-B outer iteration (scipy's L-BFGS-B, the same
Nocedal/Zhu/Byrd algorithm LBFGSB.jl wraps; src/router.jl:60, 105).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .cfmms import CFMM, GeometricMeanTwoCoin, ProductTwoCoin, UniV3
from .objectives import Objective

__all__ = ["DevicePools", "write_pool_file", "pool_file_info", "Router", "route", "find_arb", "netflows", "netflows_",
           "update_reserves", "shard_range"]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def shard_range(m: int, world: int, rank: int):
    """Contiguous, balanced split of m pools over `world` ranks."""
    lo = (m * rank) // world
    hi = (m * (rank + 1)) // world
    return lo, hi


def write_pool_file(path, n_tokens: int, R, gamma, Ai, w=None):
    """Write ProductTwoCoin (w is None) or GeometricMeanTwoCoin pools as a flat pool file."""
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 2)
    gamma = np.ascontiguousarray(gamma, dtype=np.float64).reshape(-1)
    Ai = np.ascontiguousarray(Ai, dtype=np.int64).reshape(-1, 2)
    wp = None
    if w is not None:
        w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1, 2)
        wp = _dp(w)
    rc = _lib.load().cfmm_pool_file_write(str(path).encode(), _lib.POOL_GEOMEAN if w is not None else _lib.POOL_PRODUCT,
                                          int(n_tokens), len(gamma), _dp(R), _dp(gamma), _ip(Ai), wp)
    if rc != _lib.CFMM_OK:
        raise OSError(f"cannot write pool file {path}")


def pool_file_info(path):
    """(pool type, n_tokens, m) of a flat pool file."""
    t, n, m = C.c_int(), C.c_int64(), C.c_int64()
    rc = _lib.load().cfmm_pool_file_info(str(path).encode(), C.byref(t), C.byref(n), C.byref(m))
    if rc != _lib.CFMM_OK:
        raise OSError(f"{path} is not a CFMM pool file")
    return t.value, n.value, m.value


class DevicePools:
    """One GPU's shard of the pool set: a thin object wrapper over cfmm_ctx."""

    def __init__(self, n_tokens: int, device: int = 0):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        self._ptr_args = {}
        rc = self._lib.cfmm_create(C.byref(self._ctx), int(device), int(n_tokens))
        if rc != _lib.CFMM_OK:
            msg = self._lib.cfmm_last_error(None)
            raise _lib.CFMMError(rc, msg.decode() if msg else "")
        self.n_tokens = int(n_tokens)
        self.device = int(device)
        self.peer_attached = False
        self._psi = np.zeros(self.n_tokens)
        self._acc = C.c_double(0.0)
        # pinned staging for sweep(): ν in, [Ψ; acc] out (contiguous) -- the buffers the library
        # replays its {H2D, sweep, D2H} graph on (cfmm_b200.h, option "sweep_graphs")
        nbytes = 8 * (2 * self.n_tokens + 1)
        self._pin = self._lib.cfmm_host_alloc(nbytes)
        if not self._pin:
            self._lib.cfmm_destroy(self._ctx)
            raise MemoryError("cfmm_host_alloc failed")
        buf = (C.c_double * (2 * self.n_tokens + 1)).from_address(self._pin)
        arr = np.frombuffer(buf, dtype=np.float64)
        self._pin_nu, self._pin_out = arr[:self.n_tokens], arr[self.n_tokens:]

    # -- ingest -------------------------------------------------------------
    def _chk(self, rc):
        _lib.check(self._lib, self._ctx, rc)

    def add_product(self, R, gamma, Ai):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 2)
        gamma = np.ascontiguousarray(gamma, dtype=np.float64).reshape(-1)
        Ai = np.ascontiguousarray(Ai, dtype=np.int64).reshape(-1, 2)
        if not (len(R) == len(gamma) == len(Ai)):
            raise ValueError("R, gamma, Ai must describe the same number of pools")
        self._chk(self._lib.cfmm_add_product(self._ctx, len(gamma), _dp(R), _dp(gamma), _ip(Ai)))

    def add_geomean(self, R, gamma, Ai, w):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 2)
        w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1, 2)
        gamma = np.ascontiguousarray(gamma, dtype=np.float64).reshape(-1)
        Ai = np.ascontiguousarray(Ai, dtype=np.int64).reshape(-1, 2)
        if not (len(R) == len(gamma) == len(Ai) == len(w)):
            raise ValueError("R, gamma, Ai, w must describe the same number of pools")
        self._chk(self._lib.cfmm_add_geomean(self._ctx, len(gamma), _dp(R), _dp(gamma), _ip(Ai), _dp(w)))

    def add_univ3(self, current_price, gamma, Ai, tick_off, lower_ticks, liquidity):
        cp = np.ascontiguousarray(current_price, dtype=np.float64).reshape(-1)
        gamma = np.ascontiguousarray(gamma, dtype=np.float64).reshape(-1)
        Ai = np.ascontiguousarray(Ai, dtype=np.int64).reshape(-1, 2)
        off = np.ascontiguousarray(tick_off, dtype=np.int64).reshape(-1)
        lt = np.ascontiguousarray(lower_ticks, dtype=np.float64).reshape(-1)
        lq = np.ascontiguousarray(liquidity, dtype=np.float64).reshape(-1)
        if not (len(cp) == len(gamma) == len(Ai) == len(off) - 1) or len(lt) != len(lq) \
                or (len(off) and off[-1] != len(lt)):
            raise ValueError("inconsistent UniV3 CSR arrays")
        self._chk(self._lib.cfmm_add_univ3(self._ctx, len(cp), _dp(cp), _dp(gamma), _ip(Ai),
                                           _ip(off), _dp(lt), _dp(lq)))

    def add_file(self, path: str):
        """Add the pools of a flat pool file (write_pool_file / cfmm_pool_file_write)."""
        self._chk(self._lib.cfmm_add_pool_file(self._ctx, str(path).encode()))

    def finalize(self):
        self._chk(self._lib.cfmm_finalize(self._ctx))

    @property
    def num_pools(self) -> int:
        return int(self._lib.cfmm_num_pools(self._ctx))

    def set_option(self, key: str, value: int):
        self._chk(self._lib.cfmm_set_option(self._ctx, key.encode(), int(value)))

    # -- the hot path ---------------------------------------------------------
    def sweep(self, v, materialize: bool = False):
        """find_arb!(r, v) + both folds; returns (psi [n_tokens], acc)."""
        v = np.ascontiguousarray(v, dtype=np.float64)
        if v.shape != (self.n_tokens,):
            raise ValueError(f"v must have length {self.n_tokens}")
        np.copyto(self._pin_nu, v)
        base = self._pin
        n8 = 8 * self.n_tokens
        dp = C.POINTER(C.c_double)
        self._chk(self._lib.cfmm_sweep(self._ctx, C.cast(base, dp), C.cast(base + n8, dp),
                                       C.cast(base + 2 * n8, dp), 1 if materialize else 0))
        return self._pin_out[:self.n_tokens].copy(), float(self._pin_out[self.n_tokens])

    def sweep_into(self, v_ptr: int, psi_ptr: int, acc_ptr: int, materialize: bool = False):
        """Raw-pointer form of sweep (host pointers, e.g. pinned buffers)."""
        key = (v_ptr, psi_ptr, acc_ptr)
        args = self._ptr_args.get(key)
        if args is None:  # ctypes casts cost microseconds each: keep them for buffers that come back every sweep
            if len(self._ptr_args) > 64:
                self._ptr_args.clear()
            args = self._ptr_args[key] = tuple(C.cast(q, C.POINTER(C.c_double)) for q in key)
        rc = self._lib.cfmm_sweep(self._ctx, args[0], args[1], args[2], 1 if materialize else 0)
        if rc != 0:
            self._chk(rc)

    def sweep_device(self, d_v_ptr: int, d_psi_acc_ptr: int, materialize: bool = False,
                     stream_ptr: int = 0):
        self._chk(self._lib.cfmm_sweep_device(self._ctx, d_v_ptr, d_psi_acc_ptr,
                                              1 if materialize else 0, stream_ptr or None))

    def sweep_device_view(self, d_v_ptr: int, materialize: bool = False, stream_ptr: int = 0) -> int:
        """Zero-copy device sweep: returns the device address of [psi; acc]
        (valid until the next sweep on this context)."""
        out = C.c_void_p()
        self._chk(self._lib.cfmm_sweep_device_view(self._ctx, d_v_ptr, 1 if materialize else 0,
                                                   stream_ptr or None, C.byref(out)))
        return int(out.value)

    def last_sweep_ms(self) -> float:
        ms = C.c_float(0.0)
        self._chk(self._lib.cfmm_last_sweep_ms(self._ctx, C.byref(ms)))
        return float(ms.value)

    @property
    def launch_count(self) -> int:
        return int(self._lib.cfmm_launch_count(self._ctx))

    def profile_read(self, pool_type: int):
        """(total_ms, launches) of the event-timed kernels of one pool type."""
        ms, cnt = C.c_double(0.0), C.c_int64(0)
        self._chk(self._lib.cfmm_profile_read(self._ctx, int(pool_type), C.byref(ms), C.byref(cnt)))
        return float(ms.value), int(cnt.value)

    def profile_times(self, pool_type: int):
        """Per-launch durations (ms, launch order) of the event-timed kernels of one pool type."""
        cnt = C.c_int64(0)
        self._chk(self._lib.cfmm_profile_read_times(self._ctx, int(pool_type), None, 0, C.byref(cnt)))
        out = np.zeros(int(cnt.value), dtype=np.float32)
        if len(out):
            self._chk(self._lib.cfmm_profile_read_times(self._ctx, int(pool_type),
                                                        out.ctypes.data_as(C.POINTER(C.c_float)),
                                                        len(out), C.byref(cnt)))
        return out

    def profile_reset(self):
        self._chk(self._lib.cfmm_profile_reset(self._ctx))

    def trades(self):
        m = self.num_pools
        D = np.zeros((m, 2))
        L = np.zeros((m, 2))
        self._chk(self._lib.cfmm_get_trades(self._ctx, _dp(D), _dp(L)))
        return D, L

    def solve(self, lower, lin=None, upper=None, v0=None, pgtol=1e-5, factr=1e1, maxfun=15_000, maxiter=15_000):
        """cfmm_solve: minimise linᵀν + Σ arb_i(ν) over the box on the device.  Returns (ν, info dict);
        the trades at ν are materialised (trades())."""
        n = self.n_tokens
        lower = np.ascontiguousarray(lower, dtype=np.float64)
        arrs = [lower]
        def opt(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != (n,):
                raise ValueError(f"vector arguments must have length {n}")
            arrs.append(a)
            return _dp(a)
        if lower.shape != (n,):
            raise ValueError(f"lower must have length {n}")
        opts = _lib.SolveOpts(int(maxiter), int(maxfun), float(pgtol), float(factr))
        info = _lib.SolveInfo()
        out = np.empty(n)
        self._chk(self._lib.cfmm_solve(self._ctx, opt(lin), _dp(lower), opt(upper), opt(v0), C.byref(opts),
                                       _dp(out), C.byref(info)))
        return out, {k: getattr(info, k) for k, _ in _lib.SolveInfo._fields_}

    def apply_trades(self):
        """R <- R + γΔ − Λ on the device, from the last materialising sweep."""
        self._chk(self._lib.cfmm_apply_trades(self._ctx))

    def update_reserves(self, pool_type: int, first: int, R):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 2)
        self._chk(self._lib.cfmm_update_reserves(self._ctx, int(pool_type), int(first), len(R), _dp(R)))

    # -- multi-GPU --------------------------------------------------------------
    def detach_group(self):
        self._chk(self._lib.cfmm_comm_detach(self._ctx))
        self.peer_attached = False

    def comm_check(self):
        """After sweep_device* calls: wait for them and raise CFMMError (CFMM_ERR_COMM) if an
        exchange gave up on a peer (cfmm_comm_check)."""
        self._chk(self._lib.cfmm_comm_check(self._ctx))

    def attach_group(self, group=None, barrier=True):
        """Join the NVLink peer exchange of a torch.distributed group (one
        process per GPU): export this rank's handle, all-gather the handles
        (torch.distributed is only the transport for 128 bytes per rank),
        attach.  After this, sweep() returns the sum over all ranks."""
        import torch
        import torch.distributed as dist

        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        if world == 1:
            return
        buf = (C.c_ubyte * _lib.COMM_HANDLE_BYTES)()
        self._chk(self._lib.cfmm_comm_export(self._ctx, buf))
        dev = torch.device("cuda", self.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        allh = bytes(torch.cat(gathered).cpu().numpy().tobytes())
        self._chk(self._lib.cfmm_comm_attach(self._ctx, world, rank, allh))
        self.peer_attached = True
        if barrier:
            dist.barrier(group)

    def close(self):
        if self._ctx:
            self._lib.cfmm_destroy(self._ctx)
            self._ctx = C.c_void_p()
            self._pin_nu = self._pin_out = None
            if self._pin:
                self._lib.cfmm_host_free(self._pin)
                self._pin = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _pack(cfmms):
    """Vector{CFMM} (AoS of Python objects) -> per-type SoA + the map from the
    library's global insertion order back to list positions."""
    idx = {0: [], 1: [], 2: []}
    for i, c in enumerate(cfmms):
        if isinstance(c, ProductTwoCoin):
            idx[0].append(i)
        elif isinstance(c, GeometricMeanTwoCoin):
            idx[1].append(i)
        elif isinstance(c, UniV3):
            idx[2].append(i)
        else:
            # the reference would hit a MethodError in find_arb! for these
            raise TypeError(f"no find_arb! method for {type(c).__name__}")
    return idx


class Router:
    """Router(objective, cfmms, n_tokens) (src/router.jl:4-36).

    Fields as in the reference: objective, cfmms, Δs, Λs, v.  Δs / Λs are
    (m, 2) arrays; Δs[i] / Λs[i] are the trade vectors of cfmms[i].  As in the
    reference (router.jl:39, loop bound length(r.Δs)), pools appended to
    r.cfmms after construction are ignored.

    Multi-GPU: with `group` (a torch.distributed process group, one process per
    GPU) every rank holds the full Python pool list but uploads only its
    contiguous shard; Ψ/acc are summed across ranks after each sweep and every
    rank runs the same L-BFGS-B iteration on the identical reduced vector.
    """

    def __init__(self, objective: Objective, cfmms=None, n_tokens: int = None, *,
                 device: int = 0, group=None, exchange: str = "peer", _pools_factory=None):
        if n_tokens is None and isinstance(cfmms, (int, np.integer)):
            cfmms, n_tokens = None, int(cfmms)  # Router(objective, n_tokens): empty pool list, router.jl:36
        if n_tokens is None:
            raise TypeError("n_tokens is required")
        self.objective = objective
        self.cfmms = list(cfmms) if cfmms is not None else []
        m = len(self.cfmms)
        self.Δs = np.zeros((m, 2))  # zerotrade(c), router.jl:23-26
        self.Λs = np.zeros((m, 2))
        self.v = np.zeros(int(n_tokens))
        self._group = group
        self._world, self._rank = 1, 0
        if group is not None:
            import torch.distributed as dist
            self._world, self._rank = dist.get_world_size(group), dist.get_rank(group)
        self._lo, self._hi = shard_range(m, self._world, self._rank)
        self._exchange = exchange
        factory = _pools_factory or DevicePools
        self._pools = factory(int(n_tokens), device)
        self._upload()
        if self._world > 1 and exchange == "peer":
            self._attach_with_agreement(group)
        self._psi = np.zeros(int(n_tokens))
        self._acc = 0.0

    def _attach_with_agreement(self, group):
        """Join the NVLink peer exchange, or -- if ANY rank cannot (no P2P / IPC) -- fall back to
        torch.distributed on every rank: a rank that raised alone would leave the others blocked in
        their first exchange."""
        import torch
        import torch.distributed as dist
        ok = 1
        try:
            self._pools.attach_group(group, barrier=False)
        except _lib.CFMMError:
            ok = 0
        dev = torch.device("cuda", self._pools.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            if ok:
                self._pools.detach_group()
            self._exchange = "dist"
        dist.barrier(group)

    # ascii aliases
    @property
    def Ds(self):
        return self.Δs

    @property
    def Ls(self):
        return self.Λs

    def _upload(self):
        shard = self.cfmms[self._lo:self._hi]
        idx = _pack(shard)
        p = self._pools
        order = []  # library global order -> shard-local list position
        if idx[0]:
            cs = [shard[i] for i in idx[0]]
            p.add_product(np.array([c.R for c in cs]), np.array([c.gamma for c in cs]),
                          np.array([c.Ai for c in cs]))
            order += idx[0]
        if idx[1]:
            cs = [shard[i] for i in idx[1]]
            p.add_geomean(np.array([c.R for c in cs]), np.array([c.gamma for c in cs]),
                          np.array([c.Ai for c in cs]), np.array([c.w for c in cs]))
            order += idx[1]
        if idx[2]:
            cs = [shard[i] for i in idx[2]]
            off = np.concatenate([[0], np.cumsum([len(c.lower_ticks) for c in cs])]).astype(np.int64)
            p.add_univ3(np.array([c.current_price for c in cs]), np.array([c.gamma for c in cs]),
                        np.array([c.Ai for c in cs]), off,
                        np.concatenate([c.lower_ticks for c in cs]),
                        np.concatenate([c.liquidity for c in cs]))
            order += idx[2]
        p.finalize()
        self._order = np.asarray(order, dtype=np.int64)
        self._type_lists = idx

    # one find_arb!(r, v) + folds; caches Ψ and acc like the reference caches Δs/Λs
    def _sweep(self, v, materialize=False):
        psi, acc = self._pools.sweep(v, materialize)
        if self._world > 1 and not self._pools.peer_attached:
            import torch
            import torch.distributed as dist
            buf = torch.from_numpy(np.append(psi, acc))
            if dist.get_backend(self._group) == "nccl":  # NCCL reduces device tensors only
                buf = buf.to(torch.device("cuda", self._pools.device))
            dist.all_reduce(buf, group=self._group)
            out = buf.cpu().numpy()
            psi, acc = out[:-1].copy(), float(out[-1])
        self._psi, self._acc = psi, acc
        if materialize:
            self._fetch_trades()

    def _fetch_trades(self):
        """r.Δs / r.Λs <- the trades of the last materialising sweep, in list order."""
        if True:
            D, L = self._pools.trades()
            Dl = np.zeros_like(D)
            Ll = np.zeros_like(L)
            Dl[self._order] = D
            Ll[self._order] = L
            if self._world > 1:
                import torch
                import torch.distributed as dist
                # shards are contiguous in list order: gather them back
                parts = [None] * self._world
                dist.all_gather_object(parts, (Dl, Ll), group=self._group)
                Dl = np.concatenate([q[0] for q in parts])
                Ll = np.concatenate([q[1] for q in parts])
            n = len(self.Δs)
            self.Δs[:] = Dl[:n]
            self.Λs[:] = Ll[:n]

    def sync_reserves(self):
        """Push cfmm.R of every Product/GeoMean pool to the device (the reference
        reads cfmm.R live on each sweep; call this after mutating reserves)."""
        shard = self.cfmms[self._lo:self._hi]
        for t in (0, 1):
            ids = self._type_lists[t]
            if ids:
                self._pools.update_reserves(t, 0, np.array([shard[i].R for i in ids]))


def find_arb(*args):
    """find_arb!(r::Router, v) (src/router.jl:38-42) or
    find_arb!(Δ, Λ, cfmm, v) (src/cfmms.jl:130, 185, 339).  Both run on the GPU."""
    if len(args) == 2 and isinstance(args[0], Router):
        r, v = args
        r._sweep(np.asarray(v, dtype=np.float64), materialize=True)
        return None
    if len(args) == 4:
        D, L, cfmm, v = args
        if not isinstance(cfmm, CFMM):
            raise TypeError("find_arb!(Δ, Λ, cfmm, v): cfmm must be a CFMM")
        v = np.asarray(v, dtype=np.float64)
        if v.shape != (2,):
            raise ValueError("v must have length 2 (the prices of the pool's two tokens)")
        one = type(cfmm).__new__(type(cfmm))
        one.__dict__.update(cfmm.__dict__)
        one.Ai = np.array([1, 2], dtype=np.int64)
        r = Router(_NoObjective(), [one], 2)
        r._sweep(v, materialize=True)
        D[:] = r.Δs[0]
        L[:] = r.Λs[0]
        r._pools.close()
        return None
    raise TypeError("find_arb(r, v) or find_arb(Δ, Λ, cfmm, v)")


class _NoObjective(Objective):
    pass


def route(r: Router, v=None, verbose=False, m=5, factr=1e1, pgtol=1e-5,
          maxfun=15_000, maxiter=15_000, optimizer="host"):
    """route!(r; v, verbose, m, factr, pgtol, maxfun, maxiter) (src/router.jl:58-108).
    Overwrites r.Δs, r.Λs and r.v.

    optimizer="host" (default): L-BFGS-B on the host (scipy), one device sweep per function /
    gradient evaluation, as in the reference.  optimizer="device": the whole outer iteration runs
    on the GPU (cfmm_solve: projected L-BFGS, m = 5; objectives of the form linᵀν on a box, which
    both reference objectives are) and only scalars cross PCIe per evaluation; single GPU."""
    if optimizer == "device":
        lin = r.objective.linear_term()
        if lin is None:
            raise TypeError("the device solver needs an objective with linear_term()")
        if r._world > 1:
            raise NotImplementedError("optimizer='device' drives one GPU")
        n = len(r.v)
        x, info = r._pools.solve(r.objective.lower_limit(), lin=lin, upper=r.objective.upper_limit(),
                                 v0=None if v is None else np.asarray(v, dtype=np.float64),
                                 pgtol=pgtol, factr=factr, maxfun=maxfun, maxiter=maxiter)
        r.v[:] = x
        r._fetch_trades()
        r.last_result = info
        return None
    from scipy.optimize import minimize

    n = len(r.v)
    if v is None:
        r.v[:] = np.ones(n) / n  # router.jl:61
    else:
        r.v[:] = v
    lower = np.asarray(r.objective.lower_limit(), dtype=np.float64)
    upper = np.asarray(r.objective.upper_limit(), dtype=np.float64)
    bounds = [(lo if np.isfinite(lo) else None, up if np.isfinite(up) else None)
              for lo, up in zip(lower, upper)]

    def fn(x):  # router.jl:73-86
        if not np.array_equal(x, r.v):
            r._sweep(x)
            r.v[:] = x
        return r.objective.f(x) + r._acc

    def g(x):  # router.jl:89-102
        G = np.zeros(n)
        if not np.array_equal(x, r.v):
            r._sweep(x)
            r.v[:] = x
        r.objective.grad(G, x)
        G += r._psi  # G[c.Ai] .+= Λ .- Δ over all pools
        return G

    r._sweep(r.v)  # find_arb!(r, r.v), router.jl:104
    res = minimize(fn, r.v.copy(), jac=g, method="L-BFGS-B", bounds=bounds,
                   options=dict(maxcor=m, ftol=factr * np.finfo(np.float64).eps, gtol=pgtol,
                                maxfun=maxfun, maxiter=maxiter,
                                **({"iprint": 1} if verbose else {})))
    r.v[:] = res.x  # router.jl:106
    r._sweep(r.v, materialize=True)  # find_arb!(r, v), router.jl:107
    r.last_result = res
    return None


def netflows_(psi, r: Router):
    """netflows!(ψ, r) (src/router.jl:111-119): serial pool-order sum of the
    stored trades on the host (the reference's tests compare it for exact
    equality with another pool-order sum, test/arb.jl:16)."""
    psi[:] = 0.0
    for D, L, c in zip(r.Δs, r.Λs, r.cfmms):
        psi[c.Ai - 1] += L - D
    return None


def netflows(r: Router):
    psi = np.zeros_like(r.v)
    netflows_(psi, r)
    return psi


def update_reserves(r: Router):
    """A working update_reserves!(r) (the reference's, src/router.jl:127-132,
    calls a per-CFMM method that is defined nowhere): R ← R + γΔ − Λ for the
    two-coin pools, on the host objects and on the device."""
    for D, L, c in zip(r.Δs, r.Λs, r.cfmms):
        if isinstance(c, (ProductTwoCoin, GeometricMeanTwoCoin)):
            c.R = c.R + c.gamma * D - L  # same operation order as the device kernel: bit-identical
    if any(isinstance(c, UniV3) for c in r.cfmms):
        import warnings
        warnings.warn("update_reserves: UniV3 pools keep their state (current_price / ticks are not advanced: "
                      "the reference defines no reserve update for them)", stacklevel=2)
        r.sync_reserves()  # mixed set: push the two-coin reserves from the host objects
    else:
        r._pools.apply_trades()  # on the device, from the materialised trades: no upload
    return None

"""Multi-GPU path on real hardware (skipped with fewer than 2 GPUs): pools
sharded over ranks, [Ψ; acc] summed over NVLink peer memory inside
cfmm_sweep; every rank must hold the bitwise-identical global result, equal to
the unsharded oracle within summation-order noise."""
import os
import socket
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _workload():
    from cfmmrouter_b200 import synth
    n = 4_000
    R, g, Ai = synth.product_pools(120_001, n, seed=3)
    Rg, gg, Ag, wg = synth.geomean_pools(30_000, n, seed=4)
    v = synth.dual_prices(n, "wide")
    return n, (R, g, Ai), (Rg, gg, Ag, wg), v


def _reference(oracle):
    n, (R, g, Ai), (Rg, gg, Ag, wg), v = _workload()
    D1, L1 = oracle.sweep_product(R, g, Ai, v, threads=8)
    D2, L2 = oracle.sweep_geomean(Rg, gg, Ag, wg, v, threads=8)
    A = np.concatenate([Ai, Ag])
    D, L = np.concatenate([D1, D2]), np.concatenate([L1, L2])
    accx, Gx, absG = oracle.fold_compensated(A, D, L, v, n)
    slack = np.zeros(n)
    w = 32 * np.finfo(float).eps * (R[:, 0] + R[:, 1]) / g
    np.add.at(slack, Ai[:, 0] - 1, w)
    np.add.at(slack, Ai[:, 1] - 1, w)
    tol = 1e-9 * absG + slack + 1e-300  # geomean trades: CUDA pow vs glibc pow
    return Gx.astype(np.float64), float(accx), tol, float(np.sum(tol * v))


def _shard_product_only(cr, rank, world):
    """Product-only shard: the sweep is ONE kernel (TMA sweep + fused exchange)."""
    from cfmmrouter_b200 import synth
    n = 9_000
    R, g, Ai = synth.product_pools(400_003, n, seed=13)
    v = synth.dual_prices(n, "near")
    lo, hi = cr.shard_range(len(g), world, rank)
    p = cr.DevicePools(n, device=rank)
    p.add_product(R[lo:hi], g[lo:hi], Ai[lo:hi])
    p.finalize()
    return p, v, (R, g, Ai, n)


def _shard(cr, rank, world):
    n, (R, g, Ai), (Rg, gg, Ag, wg), v = _workload()
    lo, hi = cr.shard_range(len(g), world, rank)
    lo2, hi2 = cr.shard_range(len(gg), world, rank)
    p = cr.DevicePools(n, device=rank)
    p.add_product(R[lo:hi], g[lo:hi], Ai[lo:hi])
    p.add_geomean(Rg[lo2:hi2], gg[lo2:hi2], Ag[lo2:hi2], wg[lo2:hi2])
    p.finalize()
    return p, v


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_contexts_one_process_peer_exchange(cr, oracle):
    """Single-process multi-GPU (what a Julia host would do): raw-pointer peer
    mapping, one host thread per context because the exchange kernels wait for
    each other."""
    import ctypes as C
    from cfmmrouter_b200 import _lib
    world = 2
    pools = [_shard(cr, r, world) for r in range(world)]
    # Strictly one device after the other, before any exchange: per-kernel state
    # (the > 48 KB shared-memory opt-in, occupancy) is per device, so the second
    # context must set it up again rather than inherit a process-wide cache.
    local = [p.sweep(v) for p, v in pools]
    assert all(np.all(np.isfinite(psi)) and np.isfinite(acc) for psi, acc in local)
    handles = b""
    for p, _ in pools:
        buf = (C.c_ubyte * _lib.COMM_HANDLE_BYTES)()
        p._chk(p._lib.cfmm_comm_export(p._ctx, buf))
        handles += bytes(buf)
    for r, (p, _) in enumerate(pools):
        p._chk(p._lib.cfmm_comm_attach(p._ctx, world, r, handles))
    ref, accref, tol, atol = _reference(oracle)
    for sweep in range(11):  # several epochs: slot double-buffering; all three protocols
        if sweep == 3:  # from the direct push to the LL two-shot (reduce-scatter + gather) protocol
            for p, _ in pools:
                p.set_option("exchange_two_shot", 1)
        if sweep == 6:  # LL one-shot
            for p, _ in pools:
                p.set_option("exchange_protocol", 1)
        if sweep == 8:  # and back to the direct push: its slots must have been left empty
            for p, _ in pools:
                p.set_option("exchange_protocol", 3)
        out = [None] * world

        def run(r):
            out[r] = pools[r][0].sweep(pools[r][1])

        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [t.start() for t in th]
        [t.join(timeout=120) for t in th]
        assert all(o is not None for o in out), "exchange dead-locked"
        assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]  # identical bits
        assert np.all(np.abs(out[0][0] - ref) <= tol)
        assert abs(out[0][1] - accref) <= atol
    for p, _ in pools:
        p.close()


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import cfmmrouter_b200 as cr
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        p, v = _shard(cr, rank, world)
        p.attach_group(dist.group.WORLD)  # cudaIpc handles over torch.distributed
        res = [p.sweep(v) for _ in range(3)]  # direct push (default)
        for proto in (2, 1, 3):  # LL two-shot, LL one-shot, direct again
            p.set_option("exchange_protocol", proto)
            res += [p.sweep(v) for _ in range(2)]
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), psi=np.stack([r[0] for r in res]),
                 acc=np.array([r[1] for r in res]))
        # product-only pool set: fused compute+collective kernel (and, for contrast, the unfused path)
        q, vq, _ = _shard_product_only(cr, rank, world)
        q.attach_group(dist.group.WORLD)
        l0 = q.launch_count
        fres = [q.sweep(vq) for _ in range(3)]
        fused_launches = q.launch_count - l0
        q.set_option("fused_exchange", 0)
        fres += [q.sweep(vq) for _ in range(2)]
        q.set_option("fused_exchange", 1)
        for proto in (2, 1, 3):
            q.set_option("exchange_protocol", proto)
            fres += [q.sweep(vq) for _ in range(2)]
        q.set_option("coop_launch", 1)  # the same fused kernel under cudaLaunchCooperativeKernel
        fres += [q.sweep(vq) for _ in range(2)]
        q.set_option("coop_launch", 0)
        np.savez(os.path.join(out_dir, f"fused{rank}.npz"), psi=np.stack([r[0] for r in fres]),
                 acc=np.array([r[1] for r in fres]), launches=fused_launches)
        q.close()
        # the Router-level path: every rank ends with the same route! result
        rng = np.random.default_rng(7)
        pools = [cr.ProductTwoCoin(1000 * rng.random(2) + 1, 0.997, rng.choice(np.arange(1, 9), 2, replace=False))
                 for _ in range(200)]
        r = cr.Router(cr.LinearNonnegative(rng.random(8) + 0.05), pools, 8, device=rank, group=dist.group.WORLD)
        cr.route(r)
        np.savez(os.path.join(out_dir, f"route{rank}.npz"), v=r.v, D=r.Δs, L=r.Λs)
        dist.barrier()
        p.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_one_process_per_gpu_ipc_exchange(cr, oracle, tmp_path):
    import torch.multiprocessing as mp
    world = min(_ngpu(), 8)  # every visible GPU: 2, 4 or 8 ranks
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref, accref, tol, atol = _reference(oracle)
    outs = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    for o in outs[1:]:  # bitwise-identical reduced vectors on every rank
        assert np.array_equal(outs[0]["psi"], o["psi"]) and np.array_equal(outs[0]["acc"], o["acc"])
    for k in range(outs[0]["psi"].shape[0]):
        assert np.all(np.abs(outs[0]["psi"][k] - ref) <= tol)
        assert abs(outs[0]["acc"][k] - accref) <= atol
    # fused path: one launch per sweep, same bits on every rank, right answer
    from cfmmrouter_b200 import synth
    fo = [np.load(tmp_path / f"fused{k}.npz") for k in range(world)]
    assert int(fo[0]["launches"]) == 3  # 3 sweeps -> 3 kernels: no separate exchange launch
    for o in fo[1:]:
        assert np.array_equal(fo[0]["psi"], o["psi"]) and np.array_equal(fo[0]["acc"], o["acc"])
    n = 9_000
    R, g, Ai = synth.product_pools(400_003, n, seed=13)
    vq = synth.dual_prices(n, "near")
    Do, Lo = oracle.sweep_product(R, g, Ai, vq, threads=8)
    accx, Gx, absG = oracle.fold_compensated(Ai, Do, Lo, vq, n)
    slack = np.zeros(n)
    w = 32 * np.finfo(float).eps * (R[:, 0] + R[:, 1]) / g
    np.add.at(slack, Ai[:, 0] - 1, w)
    np.add.at(slack, Ai[:, 1] - 1, w)
    for k in range(fo[0]["psi"].shape[0]):
        assert np.all(np.abs(fo[0]["psi"][k] - Gx.astype(np.float64)) <= 1e-12 * absG + slack + 1e-300)
        assert abs(fo[0]["acc"][k] - float(accx)) <= 1e-12 * float(np.sum(absG * vq)) + float(np.sum(slack * vq))
    routes = [np.load(tmp_path / f"route{k}.npz") for k in range(world)]
    for r in routes[1:]:
        assert np.array_equal(routes[0]["v"], r["v"]) and np.array_equal(routes[0]["D"], r["D"]) \
            and np.array_equal(routes[0]["L"], r["L"])
    ra = routes[0]
    assert ra["D"].shape == (200, 2) and np.all(ra["D"] >= 0) and np.any(ra["D"] > 0)

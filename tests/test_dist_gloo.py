"""The N>1 host path on CPU: world_size-2 `gloo` process group.  Each rank holds
a contiguous shard of the pools (an oracle-backed stand-in computes the shard's
sweep), Ψ/acc are all-reduced, the trades are gathered back in list order, and
every rank must end with the same ν, Ψ and trades as a single-process run."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(cr, group=None):
    from test_host_logic import OraclePools
    rng = np.random.default_rng(2024)
    pools = []
    for k in range(101):  # odd count: uneven shards
        Ai = rng.choice(np.arange(1, 13), size=2, replace=False)
        if k % 4 == 3:
            w1 = rng.uniform(0.2, 0.8)
            pools.append(cr.GeometricMeanTwoCoin(1000 * rng.random(2) + 1, [w1, 1 - w1], 0.997, Ai))
        else:
            pools.append(cr.ProductTwoCoin(1000 * rng.random(2) + 1, 0.997, Ai))
    obj = cr.LinearNonnegative(rng.random(12) + 0.05)
    return cr.Router(obj, pools, 12, group=group, exchange="dist", _pools_factory=OraclePools)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import cfmmrouter_b200 as cr
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r = _build(cr, group=dist.group.WORLD)
        lo, hi = cr.shard_range(101, world, rank)
        assert (r._lo, r._hi) == (lo, hi) and r._pools is not None
        cr.route(r)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), v=r.v, D=r.Δs, L=r.Λs, psi=r._psi, acc=r._acc)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_matches_single_process(cr, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = _build(cr)
    cr.route(r)
    outs = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    # both ranks ran the same L-BFGS-B iteration on the same reduced vector
    assert np.array_equal(outs[0]["v"], outs[1]["v"])
    assert np.array_equal(outs[0]["D"], outs[1]["D"]) and np.array_equal(outs[0]["L"], outs[1]["L"])
    # and agree with the unsharded run (sum order differs: tolerance, not bits)
    np.testing.assert_allclose(outs[0]["v"], r.v, rtol=1e-6, atol=1e-9)
    scale = np.max(np.abs(r.Δs)) + np.max(np.abs(r.Λs))
    np.testing.assert_allclose(outs[0]["D"], r.Δs, rtol=0, atol=1e-5 * scale)
    np.testing.assert_allclose(outs[0]["L"], r.Λs, rtol=0, atol=1e-5 * scale)
    assert outs[0]["D"].shape == (101, 2)

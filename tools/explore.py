#!/usr/bin/env python
"""Measurement helper (not product code): times the sweep kernels under
different options to attribute time (atomics / segmented reduce / acc / math
mode / occupancy) and reports achieved algorithmic GB/s.  Run on the GPU box."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmmrouter_b200 as cr
from cfmmrouter_b200 import synth


def time_sweeps(pools, d_nu, d_psi, iters, flush=None):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        pools.sweep_device(d_nu.data_ptr(), d_psi.data_ptr(), False, st)
    torch.cuda.synchronize()
    pools.set_option("profile", iters * 3)
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        pools.sweep_device(d_nu.data_ptr(), d_psi.data_ptr(), False, st)
    torch.cuda.synchronize()
    out = {}
    for t in (0, 1, 2):
        ms, cnt = pools.profile_read(t)
        if cnt:
            out[t] = ms / cnt * 1e3  # us
    pools.set_option("profile", 0)
    return out


def main():
    dev = torch.device("cuda", 0)
    res = []
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    PRE = ("tma_variant", "orient_by_degree")

    def run(tag, m, n, kind, nu_kind="near", iters=30, use_flush=False, **opts):
        p = cr.DevicePools(n)
        for k in PRE:
            if k in opts:
                p.set_option(k, opts[k])
        if kind == "product":
            p.add_product(*synth.product_pools(m, n))
            bpp = 32
        elif kind == "geomean":
            p.add_geomean(*synth.geomean_pools(m, n))
            bpp = 48
        else:
            args = synth.univ3_pools(m, n)
            p.add_univ3(*args)
            bpp = 32 + 16 * 4
        p.finalize()
        for k, v in opts.items():
            if k not in PRE:
                p.set_option(k, v)
        if kind == "univ3":
            rng = np.random.default_rng(3)
            nu = np.exp(rng.uniform(np.log(0.5), np.log(2.0), size=n))
        else:
            nu = synth.dual_prices(n, nu_kind)
        d_nu = torch.from_numpy(nu).to(dev)
        d_psi = torch.zeros(n + 1, dtype=torch.float64, device=dev)
        us = time_sweeps(p, d_nu, d_psi, iters, flush if use_flush else None)
        t = list(us.values())[0]
        row = {"tag": tag, "m": m, "n": n, "kind": kind, "nu": nu_kind, "opts": opts, "flush": use_flush,
               "us": round(t, 2), "Gpools_s": round(m / t / 1e3, 2), "GBs": round(m * bpp / t / 1e3, 1)}
        print(json.dumps(row), flush=True)
        res.append(row)
        p.close()

    if len(sys.argv) > 1 and sys.argv[1] == "skew":
        for alpha in (1.0, 1.2):
            for orient in (1, 2):  # 1 = default sequential SKEW shape, 2 = interleaved SKEW shape
                p = cr.DevicePools(50_000)
                p.set_option("skew_interleaved", 1 if orient == 2 else 0)
                p.add_product(*synth.product_pools_skewed(10_000_000, 50_000, alpha=alpha))
                p.finalize()
                nu = synth.dual_prices(50_000, "near")
                d_nu = torch.from_numpy(nu).to(dev)
                d_psi = torch.zeros(50_001, dtype=torch.float64, device=dev)
                us = time_sweeps(p, d_nu, d_psi, 10)
                print(json.dumps({"tag": f"c5 zipf alpha={alpha} orient={orient}", "us": round(list(us.values())[0], 2)}), flush=True)
                p.close()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "other":
        run("geomean 500k warm", 500_000, 10_000, "geomean", "near", iters=50)
        run("geomean 5M econ", 5_000_000, 10_000, "geomean", "near", iters=10)
        run("geomean 5M reference-order", 5_000_000, 10_000, "geomean", "near", iters=10, gradient_math=0)
        run("geomean 5M exact", 5_000_000, 10_000, "geomean", "near", iters=5, exact=1)
        run("univ3 500k warm", 500_000, 5_000, "univ3", iters=50)
        run("univ3 500k flushed", 500_000, 5_000, "univ3", iters=20, use_flush=True)
        run("univ3 5M", 5_000_000, 50_000, "univ3", iters=10)
        return
    M, N = 10_000_000, 50_000
    if len(sys.argv) > 1 and sys.argv[1] == "seq":
        run("c5 tma (default)", M, N, "product", "near")
        for var in (17, 20, 21, 22):
            run(f"c5 tma_variant={var} (sequential)", M, N, "product", "near", tma_variant=var)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        run("c5 tma", M, N, "product", "near")
        run("c5 tma orient=0", M, N, "product", "near", orient_by_degree=0)
        run("c5 tma", M, N, "product", "near")
        run("c5 tma orient=0", M, N, "product", "near", orient_by_degree=0)
        return
    for nu in ("near", "wide", "ones"):
        run("c5 tma", M, N, "product", nu)


    return
    run("c5 tma warp-aggregated a-RED", M, N, "product", "near", a_red_per_thread=0)
    run("c5 tma reference-order math", M, N, "product", "near", gradient_math=0)
    run("c5 tma exact(generic in tma kernel)", M, N, "product", "near", exact=1)
    run("c5 gen1 (a-sorted)", M, N, "product", "near", tma_variant=-1)
    run("c5 gen1 (bucketed layout)", M, N, "product", "near", use_tma=0)
    run("c2 tma L2-warm", 100_000, 1_000, "product", "near", iters=200)
    run("1M tma L2-warm", 1_000_000, 10_000, "product", "near", iters=100)
    return
    run("c5 exact", M, N, "product", "near", exact=1)
    for skip, name in ((1, "no b-RED"), (2, "no a-segRED"), (4, "no acc"), (3, "no RED at all"), (7, "math+loads only")):
        run("c5 " + name, M, N, "product", "wide", debug_skip=skip)
    for bps in (1, 2, 3, 4):
        run(f"c5 blocks_per_sm={bps}", M, N, "product", "wide", blocks_per_sm=bps)
    run("c2 L2-warm", 100_000, 1_000, "product", "near", iters=200)
    run("c2 flushed", 100_000, 1_000, "product", "near", iters=30, use_flush=True)
    run("1M L2-warm", 1_000_000, 10_000, "product", "near", iters=100)
    run("1M flushed", 1_000_000, 10_000, "product", "near", iters=30, use_flush=True)
    run("geomean 500k warm", 500_000, 10_000, "geomean", "near", iters=50)
    run("geomean 5M", 5_000_000, 10_000, "geomean", "near", iters=10)
    run("geomean 5M exact", 5_000_000, 10_000, "geomean", "near", iters=5, exact=1)
    run("univ3 500k warm", 500_000, 5_000, "univ3", iters=50)
    run("univ3 500k flushed", 500_000, 5_000, "univ3", iters=20, use_flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/explore.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

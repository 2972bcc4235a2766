#!/usr/bin/env python
"""First GPU call of the next round (not product code): the two experiments staged
after this round's GPU budget was spent -- `tma_variant` 23 (bulk-reduction slice
flush), 24 (warp-merged Ψ[a] runs), 25 (both) and option `geomean_log2` -- are checked against the validated default path
on the same inputs, then timed beside it.  Prints one JSON object per check; any
"ok": false means the experiment stays off.

    python tools/staged_check.py            # on a B200 box
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmmrouter_b200 as cr
from cfmmrouter_b200 import synth

EPS = np.finfo(np.float64).eps
STAGED = (23, 24, 25)


def bracketed_us(pools, d_nu, iters=200):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        pools.sweep_device_view(d_nu.data_ptr(), False, st)
    torch.cuda.synchronize()
    pools.set_option("profile", iters * 2)
    for _ in range(iters):
        pools.sweep_device_view(d_nu.data_ptr(), False, st)
    torch.cuda.synchronize()
    out = None
    for t in (0, 1):
        times = pools.profile_times(t)
        if len(times):
            out = float(np.median(times.astype(np.float64))) * 1e3
    pools.set_option("profile", 0)
    return out


def product_case(m, n, nu_kind, seed):
    R, g, Ai = synth.product_pools(m, n, seed=seed)
    v = synth.dual_prices(n, nu_kind)
    res, us = {}, {}
    for variant in (0,) + STAGED:
        p = cr.DevicePools(n)
        p.set_option("tma_variant", variant)
        p.add_product(R, g, Ai)
        p.finalize()
        psi = [p.sweep(v) for _ in range(3)]          # several sweeps: ping-pong accumulators, re-zeroing
        p.set_option("gradient_math", 0)              # reference operation order: only summation noise differs
        psi_ref_order = p.sweep(v)
        p.set_option("gradient_math", 1)
        res[variant] = (psi, psi_ref_order)
        us[variant] = bracketed_us(p, torch.from_numpy(v).cuda())
        p.close()
    # per-component bound: summation-order noise of Σ(|Λ|+|Δ|)_j plus the economized-math slack
    absG = np.zeros(n)
    w = (R[:, 0] + R[:, 1]) / g
    np.add.at(absG, Ai[:, 0] - 1, w)
    np.add.at(absG, Ai[:, 1] - 1, w)
    tol = 64 * EPS * absG + 1e-300
    out = []
    for var in STAGED:
        ok, worst = True, 0.0
        for k in range(3):
            d = np.abs(res[var][0][k][0] - res[0][0][k][0])
            worst = max(worst, float(np.max(d / tol)))
            ok &= bool(np.all(d <= tol)) and abs(res[var][0][k][1] - res[0][0][k][1]) <= float(np.sum(tol * v))
        d = np.abs(res[var][1][0] - res[0][1][0])
        ok &= bool(np.all(d <= tol))
        out.append({"check": f"tma_variant_{var}_vs_0", "m": m, "n": n, "nu": nu_kind, "ok": ok,
                    "worst_err_over_tol": worst, "median_us": {"variant0": us[0], f"variant{var}": us[var]}})
    return out


def geomean_case(m, n, seed):
    R, g, Ai, w = synth.geomean_pools(m, n, seed=seed)
    v = synth.dual_prices(n, "wide")
    p = cr.DevicePools(n)
    p.add_geomean(R, g, Ai, w)
    p.finalize()
    d_nu = torch.from_numpy(v).cuda()
    psi0, acc0 = p.sweep(v)
    us0 = bracketed_us(p, d_nu, 50)
    p.set_option("geomean_log2", 1)
    psi1, acc1 = p.sweep(v)
    us1 = bracketed_us(p, d_nu, 50)
    p.close()
    absG = np.zeros(n)
    s = (R[:, 0] + R[:, 1]) / g
    np.add.at(absG, Ai[:, 0] - 1, s)
    np.add.at(absG, Ai[:, 1] - 1, s)
    tol = 1e-12 * absG + 1e-300                       # the bound the parity tests use for this type
    d = np.abs(psi1 - psi0)
    return {"check": "geomean_log2_vs_pow", "m": m, "n": n, "ok": bool(np.all(d <= tol)),
            "worst_err_over_tol": float(np.max(d / tol)), "median_us": {"pow": us0, "log2": us1}}


def main():
    assert torch.cuda.is_available()
    out = []
    for m, n, kind, seed in ((5_000, 7, "wide", 1), (200_003, 3_001, "near", 2), (300_000, 20_011, "wide", 3),
                             (1_000_000, 49_999, "near", 4), (10_000_000, 50_000, "near", 1234)):
        for r in product_case(m, n, kind, seed):
            print(json.dumps(r), flush=True)
            out.append(r)
    for m, n, seed in ((50_000, 500, 5), (5_000_000, 10_000, 4321)):
        r = geomean_case(m, n, seed)
        print(json.dumps(r), flush=True)
        out.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/staged_check.json", "w") as f:
        json.dump(out, f, indent=1)
    sys.exit(0 if all(r["ok"] for r in out) else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""End-to-end cost of cfmm_sweep with pinned host buffers (measurement tool): wall time per call for
option sets, next to the raw pieces (H2D / D2H of n doubles, an empty stream round trip)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cfmmrouter_b200 as cr  # noqa: E402
from cfmmrouter_b200 import synth  # noqa: E402


def main():
    m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
    R, g, Ai = synth.product_pools(m, n, seed=1234)
    v = synth.dual_prices(n, "near")
    h_nu = torch.from_numpy(v).pin_memory()
    h_out = torch.zeros(n + 1, dtype=torch.float64).pin_memory()
    d = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    out = {}
    with torch.cuda.stream(st):
        for name, fn in (("h2d_n_doubles", lambda: d[:n].copy_(h_nu, non_blocking=True)),
                         ("d2h_n1_doubles", lambda: h_out.copy_(d, non_blocking=True)),
                         ("empty_sync", lambda: None)):
            for _ in range(20):
                fn(); st.synchronize()
            t0 = time.perf_counter()
            for _ in range(500):
                fn(); st.synchronize()
            out[name + "_us"] = (time.perf_counter() - t0) / 500 * 1e6
    for opts in ({"sweep_graphs": 1}, {"sweep_graphs": 0}, {"sweep_graphs": 1, "balance": 0}, {"sweep_graphs": 0, "balance": 0}):
        p = cr.DevicePools(n)
        p.add_product(R, g, Ai)
        p.finalize()
        for k, val in opts.items():
            p.set_option(k, val)
        args = (h_nu.data_ptr(), h_out.data_ptr(), h_out.data_ptr() + 8 * n)
        for _ in range(200):
            p.sweep_into(*args)
        t0 = time.perf_counter()
        for _ in range(1000):
            p.sweep_into(*args)
        dt = (time.perf_counter() - t0) / 1000 * 1e6
        out["sweep_into " + json.dumps(opts)] = dt
        p.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""How long after an idle period do event-bracketed sweeps reach their steady duration?
(measurement tool, not product code)   python tools/ramp_probe.py [--idle-ms 50]"""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=10_000_000)
    ap.add_argument("--n", type=int, default=50_000)
    ap.add_argument("--idle-ms", type=float, default=50.0)
    ap.add_argument("--launches", type=int, default=600)
    a = ap.parse_args()
    R, g, Ai = synth.product_pools(a.m, a.n, seed=1234)
    p = cr.DevicePools(a.n, device=0)
    p.add_product(R, g, Ai)
    p.finalize()
    p.set_option("sweep_events", 0)
    d_nu = torch.from_numpy(synth.dual_prices(a.n, "near")).cuda()
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for label, idle in (("after_idle", a.idle_ms), ("after_short_sync", 0.0)):
        for _ in range(300):
            p.sweep_device_view(d_nu.data_ptr(), False, st)
        torch.cuda.synchronize()
        time.sleep(idle * 1e-3)
        p.set_option("profile", a.launches)
        for _ in range(a.launches):
            p.sweep_device_view(d_nu.data_ptr(), False, st)
        torch.cuda.synchronize()
        t = np.asarray(p.profile_times(0)) * 1e3
        p.set_option("profile", 0)
        edges = [0, 5, 10, 20, 50, 100, 200, 400, a.launches]
        out[label] = {f"{lo}-{hi}": float(np.median(t[lo:hi])) for lo, hi in zip(edges[:-1], edges[1:]) if hi <= len(t)}
    print(json.dumps(out, indent=1))
    p.close()


if __name__ == "__main__":
    main()

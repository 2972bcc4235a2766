#!/usr/bin/env python
"""Per-CTA phase timeline of the TMA gradient sweep (measurement tool, option "trace"):
%globaltimer stamps at CTA entry / slice ready / own range done / all chunks done / flushed /
exit (and the chunks every CTA ended up processing: work stealing), relative to
the first CTA's entry, plus the CUDA-event duration of the same launch.

    python tools/trace_phases.py [--m 10000000 --n 50000] [--json out.json]"""
import argparse
import ctypes as C
import json
import os
import sys

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
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--opt", default="")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    R, g, Ai = synth.product_pools(a.m, a.n, seed=1234)
    v = synth.dual_prices(a.n, "near")
    p = cr.DevicePools(a.n)
    p.add_product(R, g, Ai)
    p.finalize()
    for kv in a.opt.split(","):
        if kv:
            p.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    d_nu = torch.from_numpy(v).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(10):
        p.sweep_device_view(d_nu.data_ptr(), False, st)
    torch.cuda.synchronize()
    p.set_option("trace", 1)
    rows = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.sweep_device_view(d_nu.data_ptr(), False, st)
        e1.record()
        torch.cuda.synchronize()
        grid = C.c_int64()
        p._lib.cfmm_debug_read_trace(p._ctx, None, 0, C.byref(grid))
        buf = np.zeros(grid.value * 8, dtype=np.uint64)
        rc = p._lib.cfmm_debug_read_trace(p._ctx, buf.ctypes.data_as(C.POINTER(C.c_uint64)), grid.value, C.byref(grid))
        assert rc == 0
        t = buf.reshape(-1, 8).astype(np.int64)
        t0 = t[:, 0].min()
        rel = (t[:, :6] - t0) / 1e3  # us
        chunks = t[:, 7] & 0xffffffff
        rows.append({"event_us": e0.elapsed_time(e1) * 1e3,
                     "entry_max": rel[:, 0].max(), "entry_med": np.median(rel[:, 0]),
                     "slice_ready_med": np.median(rel[:, 1] - rel[:, 0]), "slice_ready_max": (rel[:, 1] - rel[:, 0]).max(),
                     "loop_done_min": rel[:, 3].min(), "loop_done_med": np.median(rel[:, 3]), "loop_done_max": rel[:, 3].max(),
                     "flush_med": np.median(rel[:, 4] - rel[:, 3]), "exit_max": rel[:, 5].max(),
                     "loop_len_min": (rel[:, 3] - rel[:, 1]).min(), "loop_len_med": np.median(rel[:, 3] - rel[:, 1]),
                     "loop_len_max": (rel[:, 3] - rel[:, 1]).max(), "grid": int(grid.value),
                     "own_done_min": rel[:, 2].min(), "own_done_med": np.median(rel[:, 2]), "own_done_max": rel[:, 2].max(),
                     "chunks_min": int(chunks.min()), "chunks_max": int(chunks.max())})
    med = {k: float(np.median([r[k] for r in rows])) for k in rows[0]}
    print(json.dumps(med, indent=1))
    # per-SM spread of the last rep: which SMs finish last
    sm = t[:, 7] >> 32
    done = rel[:, 3]
    order = np.argsort(done)
    print("slowest CTAs (sm, loop_done us):", [(int(sm[i]), round(float(done[i]), 1)) for i in order[-6:]])
    print("fastest CTAs (sm, loop_done us):", [(int(sm[i]), round(float(done[i]), 1)) for i in order[:6]])
    if a.json:
        with open(a.json, "w") as f:
            json.dump(med, f, indent=1)
    p.close()


if __name__ == "__main__":
    main()

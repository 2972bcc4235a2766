#!/usr/bin/env python
"""bench.py -- the find_arb! dual-gradient sweep on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

A "step" is ONE dual-gradient sweep: find_arb! over every pool at the current
ν plus the Ψ / acc folds (src/router.jl:38-42, 79-83, 98-100) -- exactly what
one L-BFGS-B function/gradient evaluation of route! costs on the pool side.

Workload (config.workload): BASELINE.json configs[4] -- 10M ProductTwoCoin
pools, 50k tokens -- per GPU.  It is the configuration the metric's target is
quoted on (">= 10M find_arb! evaluations per sweep"), it fits one GPU, and at
320 MB it is larger than the 126 MB L2, so every timed sweep streams from HBM
without an explicit flush.  For N > 1 each rank owns its own 10M-pool shard
(weak scaling) and the only exchange is the sum of [Ψ; acc] over NVLink peer
memory after each sweep; `--scaling strong` splits the same 10M pools instead.

Prints ONE JSON line (rank 0).  `value` = pools evaluated per second with ν and
Ψ resident in HBM (CUDA events, max over ranks); `e2e` = the same through the
public C-ABI call cfmm_sweep() with pinned HOST buffers (H2D ν and D2H Ψ inside
the timed region).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "find_arb_pools_per_sec_per_dual_gradient_sweep"
UNIT = "pools/s"

WORKLOADS = {
    # name: (pools per GPU, n_tokens, kind)
    "config5_10M_product_50k_tokens": (10_000_000, 50_000, "product"),
    "config2_100k_product_1k_tokens": (100_000, 1_000, "product"),
    "config3_1M_mixed_10k_tokens": (1_000_000, 10_000, "mixed"),
    "config4_500k_univ3_5k_tokens": (500_000, 5_000, "univ3"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default="config5_10M_product_50k_tokens", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer")
    ap.add_argument("--two-shot", type=int, default=-1, help="peer exchange protocol: -1 auto (two-shot for N>2), 0, 1")
    ap.add_argument("--nu", choices=["near", "wide", "ones"], default="near")
    ap.add_argument("--exact", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps (capped)")
    return ap.parse_args()


def make_shard(workload, rank, world, scaling):
    from cfmmrouter_b200 import synth
    m, n, kind = WORKLOADS[workload]
    if scaling == "strong":
        lo, hi = (m * rank) // world, (m * (rank + 1)) // world
    else:
        lo, hi = 0, m
    seed = 1234 + (rank if scaling == "weak" else 0)
    out = {"n": n, "kind": kind, "m_local": hi - lo}
    if kind == "product":
        R, g, Ai = synth.product_pools(m, n, seed=seed)
        out["product"] = (R[lo:hi], g[lo:hi], Ai[lo:hi])
        out["bytes"] = (hi - lo) * 32
    elif kind == "mixed":
        h = m // 2
        R, g, Ai = synth.product_pools(h, n, seed=seed)
        Rg, gg, Ag, wg = synth.geomean_pools(h, n, seed=seed + 1)
        l2, h2 = lo // 2, hi // 2
        out["product"] = (R[l2:h2], g[l2:h2], Ai[l2:h2])
        out["geomean"] = (Rg[l2:h2], gg[l2:h2], Ag[l2:h2], wg[l2:h2])
        out["m_local"] = 2 * (h2 - l2)
        out["bytes"] = (h2 - l2) * (32 + 48)
    else:
        cp, g, Ai, off, lt, lq = synth.univ3_pools(m, n, seed=seed)
        sl = slice(lo, hi)
        out["univ3"] = (cp[sl], g[sl], Ai[sl], off[lo:hi + 1] - off[lo], lt[off[lo]:off[hi]], lq[off[lo]:off[hi]])
        out["bytes"] = (hi - lo) * 32 + (off[hi] - off[lo]) * 16
    return out


# ---------------------------------------------------------------------------
# clocks: sampled DURING the timed region (NVML, falls back to nvidia-smi)
# ---------------------------------------------------------------------------

class ClockSampler:
    def __init__(self, device_index):
        self.idx = device_index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self._nvml = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    _BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
             0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def _loop(self):
        n = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self._BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def _sample_once(self):
        n = self._nvml
        try:
            self.samples.append(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
            try:
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
            except Exception:
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            for bit, name in self._BITS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        if self._nvml:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._sample_once()  # at least one sample taken while the last steps are in flight
            self._stop.set()
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle "faithful" port)
# ---------------------------------------------------------------------------

def cpu_faithful_rate(workload, sample_pools, sweeps, threads):
    """pools/s of the faithful-layout CPU restatement (oracle/) on a sample of
    the workload: threaded sweep + serial acc / scatter folds, like
    src/router.jl:38-42, 79-83, 98-100."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from cfmmrouter_b200 import synth
    o = oracle_lib.load()
    m, n, kind = WORKLOADS[workload]
    sample_pools = int(min(sample_pools, m))
    f = o.faithful(n)
    if kind in ("product", "univ3"):  # (univ3 has no faithful flavour: product stands in)
        R, g, Ai = synth.product_pools(sample_pools, n)
        f.add_product(R, g, Ai)
    else:
        R, g, Ai = synth.product_pools(sample_pools // 2, n)
        Rg, gg, Ag, wg = synth.geomean_pools(sample_pools // 2, n)
        f.add_product(R, g, Ai)
        f.add_geomean(Rg, gg, Ag, wg)
    v = synth.dual_prices(n, "near")
    f.sweep(v, threads)  # warm
    t0 = time.perf_counter()
    for _ in range(sweeps):
        f.sweep(v, threads)
    dt = time.perf_counter() - t0
    f.close()
    return sample_pools * sweeps / dt, dt / sweeps, sample_pools


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    o = oracle_lib.load()
    threads = o.max_threads()
    m, n, kind = WORKLOADS[args.workload]
    # bound the whole run to ~2 minutes of CPU time: pools per step from a probe
    # all host threads, unless half of them (one per physical core) is faster on this box
    probe_rate, _, _ = cpu_faithful_rate(args.workload, 200_000, 2, threads)
    if threads >= 4:
        half_rate, _, _ = cpu_faithful_rate(args.workload, 200_000, 2, threads // 2)
        if half_rate > probe_rate:
            threads, probe_rate = threads // 2, half_rate
    budget = 120.0 / max(1, args.steps + args.warmup)
    sample = int(max(10_000, min(m, 2_000_000, probe_rate * budget)))
    from cfmmrouter_b200 import synth
    f = o.faithful(n)
    R, g, Ai = synth.product_pools(sample, n)
    f.add_product(R, g, Ai)
    v = synth.dual_prices(n, "near")
    for _ in range(args.warmup):
        f.sweep(v, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        f.sweep(v, threads)
    dt = time.perf_counter() - t0
    f.close()
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "pools_per_step_sample": sample, "n_tokens": n,
                   "note": "reference = CFMMRouter.jl's CPU algorithm; Julia is not installed, so this is "
                           "the oracle's faithful-layout C restatement (threaded sweep, serial folds)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} ProductTwoCoin pools of the workload per step, "
                                   f"{args.steps} steps, {threads} OpenMP threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    import cfmmrouter_b200 as cr

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    shard = make_shard(args.workload, rank, world, args.scaling)
    n = shard["n"]
    pools = cr.DevicePools(n, device=local_rank)
    if "product" in shard:
        pools.add_product(*shard["product"])
    if "geomean" in shard:
        pools.add_geomean(*shard["geomean"])
    if "univ3" in shard:
        pools.add_univ3(*shard["univ3"])
    pools.finalize()
    pools.set_option("exact", args.exact)
    pools.set_option("sweep_events", 0)
    m_local = shard["m_local"]
    alg_bytes = float(shard["bytes"])
    del shard

    exchange = "none"
    if world > 1:
        exchange = args.exchange
        if exchange == "peer":
            try:
                pools.attach_group(dist.group.WORLD)
            except cr.CFMMError as e:
                if rank == 0:
                    print(f"[bench] peer exchange unavailable ({e}); using NCCL", file=sys.stderr)
                exchange = "nccl"
            if exchange == "peer" and args.two_shot >= 0:
                pools.set_option("exchange_two_shot", args.two_shot)
            flag = torch.tensor([1 if exchange == "peer" else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if flag.item() == 0 and exchange == "peer":
                exchange = "nccl"

    from cfmmrouter_b200 import synth
    nu_host = synth.dual_prices(n, args.nu)
    d_nu = torch.from_numpy(nu_host).to(dev)
    d_psi = torch.zeros(n + 1, dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    sptr = stream.cuda_stream

    def step():
        if exchange == "nccl":  # NCCL needs the partial in a torch tensor
            pools.sweep_device(d_nu.data_ptr(), d_psi.data_ptr(), False, sptr)
            dist.all_reduce(d_psi)
        else:  # zero-copy: [Ψ; acc] stays in the context's device buffer
            pools.sweep_device_view(d_nu.data_ptr(), False, sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler_ref = [None]

    def timed_region(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        if sampler_ref[0] is not None:
            sampler_ref[0]._sample_once()  # GPU still busy with the queued steps
        barrier()
        return e0.elapsed_time(e1)

    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            step()
        barrier()
        # ---- timed region 1: K steps, nothing but the sweeps on the stream -> `value`
        l0 = pools.launch_count
        sampler = ClockSampler(local_rank)
        sampler_ref[0] = sampler if sampler._nvml else None
        sampler.start()
        ms_total = timed_region(args.steps)
        sampler.stop()
        sampler_ref[0] = None
        launches = pools.launch_count - l0
        # ---- timed region 2: the same K steps with a CUDA-event pair around every
        # kernel launch (on the launching stream) -> per-kernel durations for `roofline`
        n_kernels = 2 if args.workload.startswith("config3") else 1
        pools.set_option("profile", args.steps * (n_kernels + (1 if exchange == "peer" else 0)))
        timed_region(args.steps)
        prof = {t: pools.profile_read(t) for t in (0, 1, 2, 3)}
        prof_times = {t: pools.profile_times(t) for t in (0, 1, 2)}
        pools.set_option("profile", 0)

        # ---- e2e: public C-ABI call with pinned host buffers, copies inside ----
        e2e_steps = args.e2e_steps or min(args.steps, 2000)
        h_nu = torch.from_numpy(nu_host).pin_memory()
        h_out = torch.zeros(n + 1, dtype=torch.float64).pin_memory()  # [psi ; acc] contiguous
        h_psi, h_acc = h_out[:n], h_out[n:]
        for _ in range(3):
            pools.sweep_into(h_nu.data_ptr(), h_psi.data_ptr(), h_acc.data_ptr())
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pools.sweep_into(h_nu.data_ptr(), h_psi.data_ptr(), h_acc.data_ptr())
            if exchange == "nccl":  # host-visible result must be the global sum
                t = torch.cat([h_psi, h_acc]).to(dev)
                dist.all_reduce(t)
                t.cpu()
        barrier()
        e2e_s = time.perf_counter() - t0

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = t[0].item(), t[1].item()
    total_pools = m_local * world if args.scaling == "weak" else WORKLOADS[args.workload][0]
    if args.scaling == "weak" and world > 1:
        tp = torch.tensor([m_local], dtype=torch.float64, device=dev)
        dist.all_reduce(tp)
        total_pools = int(tp.item())
    value = total_pools * args.steps / (ms_total * 1e-3)
    e2e_value = total_pools * e2e_steps / e2e_s

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kind = WORKLOADS[args.workload][2]
        roofline = roofline_object(prof, prof_times, ms_total, args.steps, launches, kind, m_local,
                                   alg_bytes, peak, peak_src, read_traffic())
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib
            threads = oracle_lib.load().max_threads()
            rate, per_sweep, sample = cpu_faithful_rate(args.workload, 2_000_000, 3, threads)
            rate1, _, sample1 = cpu_faithful_rate(args.workload, 500_000, 2, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{sample} pools of the workload x 3 sweeps, faithful-layout C restatement of "
                             f"router.jl:38-42,79-83,98-100 (Julia unavailable), {threads} OpenMP threads; "
                             f"1 thread (Julia's default): {rate1:.3g} pools/s on {sample1} pools"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "pools_per_gpu": m_local, "pools_total": total_pools,
                       "n_tokens": n, "nu": args.nu, "exact_mode": args.exact, "exchange": exchange,
                       "l2": "inputs larger than L2 (320 MB/GPU > 126 MB)" if alg_bytes > 126e6
                             else "L2-WARM: working set fits in L2, no flush between steps"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n,
                    "d2h_bytes_per_step": 8 * (n + 1), "steps": e2e_steps,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "api": "cfmm_sweep (C ABI), pinned host buffers"},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        emit(line)
    pools.close()
    if world > 1:
        dist.destroy_process_group()


def roofline_object(prof, prof_times, ms_total, steps, launches, kind, m_local, alg_bytes, peak, peak_src,
                    traffic):
    """The `roofline` object of the JSON line, from the event-bracketed launches of timed
    region 2.  prof[t] = (total_ms, launches) and prof_times[t] = per-launch ms for pool type
    t (3 = peer exchange); ms_total / launches belong to timed region 1 (no events between
    launches)."""
    # dominant kernel = the one with the most event-timed device time
    dom = max((0, 1, 2), key=lambda t: prof[t][0])
    dom_ms, dom_cnt = prof[dom]
    dom_name = {0: "product_sweep_tma (ProductTwoCoin gradient sweep)", 1: "sweep_kernel<GeomeanPools>",
                2: "sweep_kernel<Univ3Pools>"}[dom]
    if kind == "mixed":
        dom_bytes = (m_local // 2) * (32 if dom == 0 else 48)
    else:
        dom_bytes = alg_bytes
    achieved = dom_bytes / (dom_ms / max(dom_cnt, 1) * 1e-3) / 1e9 if dom_cnt else None
    roofline = {
        "bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_us": 1e3 * dom_ms / max(dom_cnt, 1),
        "launches_timed": dom_cnt, "traffic": traffic,
    }
    if len(prof_times[dom]):
        # spread of the individual event-bracketed launches: `avg_launch_us` is their mean
        # (what `achieved` uses); a bracket also holds the stream's event/launch front-end
        # latency, which differs between hosts, so the quantiles are reported beside it
        us = 1e3 * np.sort(np.asarray(prof_times[dom], dtype=np.float64))
        roofline["launch_us"] = {"mean": float(us.mean()), "min": float(us[0]),
                                 "p05": float(us[int(0.05 * (len(us) - 1))]),
                                 "median": float(us[(len(us) - 1) // 2]),
                                 "p95": float(us[int(0.95 * (len(us) - 1))]), "max": float(us[-1])}
    if dom_cnt and launches == steps and ms_total > 0:
        # one launch per step and nothing else on the stream: timed region 1 is the same
        # kernel back to back.  Consecutive launches of the persistent kernel overlap their
        # ramp and tail, so this is shorter than a bracketed (serialised) launch; reported
        # beside `frac`, not instead of it.
        step_us = 1e3 * ms_total / steps
        b2b = dom_bytes / (step_us * 1e-6) / 1e9
        roofline["back_to_back"] = {"launch_us": step_us, "achieved": b2b, "frac": b2b / peak}
    if prof[3][1]:
        roofline["exchange_avg_us"] = 1e3 * prof[3][0] / prof[3][1]
    return roofline



def read_traffic():
    """dram bytes (read+write) per launch of the dominant kernel from the
    committed ncu --set full capture, if present (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line, on the real stdout (libraries such as NCCL print
    banners to fd 1; everything else this process writes goes to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — seed-to-multi-MUM throughput of the B200 path (metric of BASELINE.json).

One "step" = one pass of the whole hot path (seed extraction -> radix sort -> bucket policy ->
extension + de-dup -> canonical match CSR) over one batch of synthetic genomes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--scale S]
  python bench.py --impl reference ...    # the CPU restatement (oracle/) timed on the host cores

Keys of the JSON line: see the task contract; `value` = inputs resident in HBM, `e2e` = through the
C ABI with pinned host buffers (H2D of the ASCII genomes + D2H of the match CSR inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"  # keeps NCCL's version banner out of stdout: the bench prints ONE JSON line there

CONFIG_NAMES = {
    1: "C1 mauveAligner 2 x 5 Mbp, weight-15 spaced seed, MODE_UNIQUE",
    2: "C2 progressiveMauve seed matching, 8 x 5 Mbp with rearrangements, default (coding) seed w15, MODE_UNIQUE",
    3: "C3 uniqueMerCount, 1 x 100 Mbp, weight-19 seed, MODE_UNIQUE_COUNT",
    4: "C4 repeatoire self-match, 1 x 200 Mbp repeat-rich, w15, rmin 2 rmax 500, MODE_SEED_ENUM",
    5: "C5 progressiveMauve seeds, 64 x 5 Mbp (320 Mbp), coding seed w15, MODE_UNIQUE",
}


def config_params(mb, config):
    if config == 1:
        return mb.get_seed(15, 0), mb.MODE_UNIQUE, {}
    if config in (2, 5):
        return mb.get_seed(15, mb.CODING_SEED), mb.MODE_UNIQUE, {}
    if config == 3:
        return mb.get_seed(19, 0), mb.MODE_UNIQUE_COUNT, {}
    return mb.get_seed(15, 0), mb.MODE_SEED_ENUM, dict(min_multi=2, max_multi=500)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons of the GPUs in use, sampled during the timed region through NVML in-process
    (pynvml; a `nvidia-smi` subprocess per sample stalls the driver for milliseconds and showed up as multi-ms
    outliers in the per-step times).  Falls back to the nvidia-smi query line of B200_PROFILING.md without pynvml."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, indices, period=0.05):
        super().__init__(daemon=True)
        self.indices = [indices] if isinstance(indices, int) else list(indices)
        self.period, self.samples, self.stop_flag = period, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.indices]
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        N = self.nvml
        bits = (("hw_slowdown", getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        for h in self.handles:
            sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.samples.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for _, b in bits])

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
        for line in out.splitlines():
            self.samples.append([x.strip() for x in line.split(",")])

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(self.period if self.nvml else 0.25)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi", "gpus": self.indices}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (oracle/, kind 'port': libMems, which holds
    the reference's own implementation, is not in /root/reference, so nothing can be compiled into oracle/_ref)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import mauvealigner_b200 as mb
    config = args.config
    scale = args.ref_scale
    seqs = mb.synth_genomes(config, scale)
    pattern, mode, kw = config_params(mb, config)
    bp = sum(len(s) for s in seqs)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        O.find(seqs, pattern, mode, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.find(seqs, pattern, mode, **kw)
    dt = time.perf_counter() - t0
    val = bp * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": "seed-to-multi-MUM input throughput", "value": val, "unit": "Gbp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": CONFIG_NAMES[config], "sample": f"same generator at scale 1/{scale} ({bp} bp)", "bp_per_step": bp},
        "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": 1, "kind": "port",
                         "sample": f"config C{config} at 1/{scale} length ({bp} bp), oracle/liboracle.so single thread"},
        "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=0, help="BASELINE.json config 1..5 (default: 2 at N=1, 5 at N>1)")
    ap.add_argument("--scale", type=int, default=1, help="divide genome lengths (debug only; 1 = the named size)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-scale", type=int, default=16, help="length divisor of the CPU sample per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.config == 0:
        args.config = 2 if args.gpus == 1 else 5
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import mauvealigner_b200 as mb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 or world > 1:
        from mauvealigner_b200 import dist as mbdist
        return mbdist.bench_main(args, CONFIG_NAMES, config_params, peaks, ClockSampler)

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    W = max(3, args.warmup)
    config = args.config
    pattern, mode, kw = config_params(mb, config)
    seqs = mb.synth_genomes(config, args.scale)
    bp = sum(len(s) for s in seqs)
    ctx = mb.Context(local)
    stream = torch.cuda.Stream(dev)  # explicit: the library launches on it and the timing events are recorded on it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_seed(pattern)

    # ---------------- value: packed genomes resident in HBM -> match CSR resident in HBM
    dev_ascii = [torch.from_numpy(s).to(dev) for s in seqs]
    ctx.clear_sequences()
    for t in dev_ascii:
        ctx.add_sequence_device(t.data_ptr(), t.numel())
    torch.cuda.synchronize()
    for _ in range(W):
        ctx.find_device(mode, **kw)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {}
    radix_ms, radix_launches, launches = 0.0, 0, 0
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        ctx.find_device(mode, **kw)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    # per-stage / per-kernel device times come from the library's own CUDA events of the last timed steps
    res = ctx.fetch()
    st = ctx.stats()
    for k in ("ms_extract", "ms_sort", "ms_bucket", "ms_dedup", "ms_output", "ms_total_device"):
        stage[k] = round(st[k], 4)
    radix_ms, radix_launches, launches = st["ms_radix_kernels"], st["radix_launches"], st["kernel_launches"]
    ms_per_step = ms_total / args.steps
    value = bp / (ms_per_step * 1e-3) / 1e9

    # ---------------- e2e: pinned host ASCII -> C ABI -> host CSR
    pinned = [torch.from_numpy(s).pin_memory() for s in seqs]

    def e2e_step():
        ctx.clear_sequences()
        for t in pinned:
            ctx.add_sequence_ptr(t.data_ptr(), t.numel())
        return ctx.find(mode, copy=False, **kw)

    for _ in range(2):
        r = e2e_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = e2e_step()
    e1.record(stream)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = max(e0.elapsed_time(e1) / args.steps, wall_ms)
    st2 = ctx.stats()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e_val = bp / (e2e_ms * 1e-3) / 1e9

    # ---------------- roofline of the dominant kernel (one radix pass = read + write of every record)
    peak, peak_src = peaks()
    R = st["record_bytes"]
    n_seeds = st["n_seeds"]
    alg_bytes_per_launch = 2.0 * R * n_seeds
    roofline = None
    traffic = None  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (same workload only)
    try:
        if config == 2 and args.scale == 1:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_c2.json")))["k_onesweep_traffic_bytes_per_launch"]
    except Exception:
        traffic = None
    if radix_launches:
        avg_ms = radix_ms / radix_launches
        achieved = alg_bytes_per_launch / (avg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_onesweep (one LSD radix pass over all seed records)", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": avg_ms, "launches_per_step": radix_launches,
                    "share_of_step": radix_ms / st["ms_total_device"] if st["ms_total_device"] else None}
    # whole-path figure of SURVEY.md §8d: B_alg = 0.25 + R (3 + 2P) bytes per input base
    P = (2 * mb.seed_weight(pattern) + 7) // 8
    b_alg = 0.25 + R * (3 + 2 * P)
    path = {"b_alg_bytes_per_bp": b_alg, "achieved": b_alg * bp / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s"}
    path["frac"] = path["achieved"] / peak

    cpu = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        cs = mb.synth_genomes(config, max(args.scale, 1) * 4)
        cbp = sum(len(s) for s in cs)
        t0 = time.perf_counter()
        O.find(cs, pattern, mode, **kw)
        dt = time.perf_counter() - t0
        cpu = {"value": cbp / dt / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "port", "seconds": dt,
               "sample": f"config C{config} at 1/{max(args.scale, 1) * 4} length ({cbp} bp), oracle/liboracle.so, single thread "
                         "(the reference is single-threaded; libMems itself is not buildable here)"}

    line = {
        "metric": "seed-to-multi-MUM input throughput", "value": value, "unit": "Gbp/s", "n_gpus": 1, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": CONFIG_NAMES[config], "bp_per_step": bp, "n_genomes": len(seqs), "seed_pattern": mb.seeds.pattern_text(pattern),
                   "scale": args.scale, "l2": "inputs larger than L2 (records: %d MB per buffer)" % (n_seeds * R // 2 ** 20),
                   "n_matches": res["n_matches"], "n_candidates": st["n_candidates"], "n_extended": st["n_extended"]},
        "roofline": roofline, "path_roofline": path, "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": "Gbp/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": st2["h2d_bytes"],
                "d2h_bytes_per_step": st2["d2h_bytes"]},
        "gpu_launches": int(launches) * args.steps, "stages_ms": stage, "clocks": sampler.summary(),
        "dedup": {"batches": st["dedup_batches"], "iters": st["dedup_iters"]},
    }
    print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

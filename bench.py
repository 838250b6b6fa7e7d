#!/usr/bin/env python
"""bench.py — seed-to-multi-MUM throughput of the B200 path (metric of BASELINE.json).

One "step" = one pass of the whole hot path (seed extraction -> radix sort -> bucket policy ->
extension + de-dup -> canonical match CSR) over one batch of synthetic genomes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--no-extra]
  python bench.py --impl reference ...    # the CPU restatement (oracle/) timed on the host cores

Every N runs the SAME workload — C5 of BASELINE.json (64 x 5 Mbp, the configuration the metric's
"1/2/4/8 B200" is quoted on; it fits one GPU) — so the N = 1 line is the strong-scaling denominator.
C1..C4 at their full sizes ride along at N = 1 under `extra` (value, ms, path_roofline, digest_ok).

Keys of the JSON line: see the task contract; `value` = inputs resident in HBM, `e2e` = through the
C ABI with pinned host buffers (H2D of the ASCII genomes + D2H of the match CSR inside the timed region);
`parity.digest_ok` = SHA-256 of the complete result equals the oracle's committed full-size digest
(tests/golden/fullsize_digests.json).
"""
from __future__ import annotations

import argparse
import importlib.util
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

METRIC = "seed-to-multi-MUM input throughput"
HEADLINE_CONFIG = 5
CPU_SAMPLE_SCALE = 32  # the CPU legs (cpu_baseline and --impl reference) run the same configuration at 1/32 of every genome length
MODE_UNIQUE, MODE_SEED_ENUM, MODE_UNIQUE_COUNT = 0, 1, 2
CONFIG_NAMES = {
    1: "C1 mauveAligner 2 x 5 Mbp, weight-15 spaced seed, MODE_UNIQUE",
    2: "C2 progressiveMauve seed matching, 8 x 5 Mbp with rearrangements, default (coding) seed w15, MODE_UNIQUE",
    3: "C3 uniqueMerCount, 1 x 100 Mbp, weight-19 seed, MODE_UNIQUE_COUNT",
    4: "C4 repeatoire self-match, 1 x 200 Mbp repeat-rich, w15, rmin 2 rmax 500, MODE_SEED_ENUM",
    5: "C5 progressiveMauve seeds, 64 x 5 Mbp (320 Mbp), coding seed w15, MODE_UNIQUE",
}
CONFIG_BP = {1: 9999842, 2: 39999840, 3: 100000000, 4: 200000000, 5: 319999084}  # full-size input bases (tests/golden/fullsize_digests.json)
CONFIG_NSEQ = {1: 2, 2: 8, 3: 1, 4: 1, 5: 64}


def _load(name, *path):
    """Import a pure-Python helper by file path (the reference arm must not import the product package)."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, *path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


seeds = _load("mb_seeds", "mauvealigner_b200", "seeds.py")         # seed tables: pure Python
synth = _load("mb_synth_tools", "tools", "synth", "__init__.py")     # libmbsynth.so: host-only generator, not the product library
golden = _load("mb_golden", "tests", "golden", "make_fullsize_digests.py")


def config_params(_mb, config):
    if config == 1:
        return seeds.get_seed(15, 0), MODE_UNIQUE, {}
    if config in (2, 5):
        return seeds.get_seed(15, seeds.CODING_SEED), MODE_UNIQUE, {}
    if config == 3:
        return seeds.get_seed(19, 0), MODE_UNIQUE_COUNT, {}
    return seeds.get_seed(15, 0), MODE_SEED_ENUM, dict(min_multi=2, max_multi=500)


def config_dict(config, scale=1):
    """`config` of the JSON line: names the workload only, identical in the repo arm and in the reference arm."""
    pattern, _, _ = config_params(None, config)
    return {"workload": CONFIG_NAMES[config], "bp_per_step": CONFIG_BP[config] if scale == 1 else None, "n_genomes": CONFIG_NSEQ[config],
            "seed_pattern": seeds.pattern_text(pattern), "scale": scale, "l2": "inputs larger than L2 (every record buffer exceeds the 126 MB L2)"}


def expected_digest(config):
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize_digests.json")))[str(config)]
    except Exception:
        return None


def parity_of(res, config, mode, scale=1):
    """Bit-exact check inside the bench line: SHA-256 of the complete result against the oracle's committed digest."""
    out = {"n_matches": int(res["n_matches"]), "n_comps": int(res["n_comps"])}
    want = expected_digest(config) if scale == 1 else None
    if want is None:
        out["digest_ok"] = None
        return out
    got = golden.digest(res, mode == MODE_UNIQUE_COUNT)
    out.update(digest_ok=bool(got == want["sha256"]), sha256=got, oracle="tests/golden/fullsize_digests.json")
    return out


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


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_sample_text(config, scale, bp):
    return (f"{CONFIG_NAMES[config].split(' ')[0]} through the same generator at 1/{scale} of every genome length ({bp} bp per step), "
            "oracle/liboracle.so (CPU restatement of the reference path; libMems itself is not buildable here), one thread as the reference")


def cpu_leg(config, scale, warmup, steps):
    """The oracle (kind 'port') on the host cores over a bounded sample of `config`; returns (Gbp/s, seconds per step, bp)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    pattern, mode, kw = config_params(None, config)
    seqs = synth.synth_genomes(config, scale)
    bp = sum(len(s) for s in seqs)
    for _ in range(warmup):
        O.find(seqs, pattern, mode, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.find(seqs, pattern, mode, **kw)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return bp / dt / 1e9, dt, bp


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (oracle/, kind 'port': libMems, which holds the
    reference's own implementation, is not in /root/reference, so nothing can be compiled into oracle/_ref).  Same
    configuration as the repo arm, every step a bounded sample of it (CPU_SAMPLE_SCALE); loads liboracle.so and
    libmbsynth.so only — never the product library."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    config, scale = args.config, args.ref_scale * args.scale
    val, dt, bp = cpu_leg(config, scale, args.warmup, args.steps)
    sample = cpu_sample_text(config, scale, bp)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_dict(config, args.scale),
        "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": 1, "kind": "port", "sample": sample, "bp_per_step": bp},
        "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(config, scale):
    val, dt, bp = cpu_leg(config, CPU_SAMPLE_SCALE * scale, 1, 2)
    return {"value": val, "unit": "Gbp/s", "cores": 1, "kind": "port", "seconds": 2 * dt, "sample": cpu_sample_text(config, CPU_SAMPLE_SCALE * scale, bp),
            "bp_per_step": bp}


# ------------------------------------------------------------------------------------------ one GPU
def time_config(mb, torch, ctx, stream, dev, config, scale, W, K, want_e2e):
    """Device-timed passes (+ the e2e leg) of one configuration on one GPU; returns a dict of everything measured."""
    pattern, mode, kw = config_params(None, config)
    seqs = synth.synth_genomes(config, scale)
    bp = sum(len(s) for s in seqs)
    ctx.set_seed(pattern)
    dev_ascii = [torch.from_numpy(s).to(dev) for s in seqs]
    ctx.clear_sequences()
    for t in dev_ascii:
        ctx.add_sequence_device(t.data_ptr(), t.numel())
    torch.cuda.synchronize()
    del dev_ascii
    for _ in range(W):
        ctx.find_device(mode, **kw)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(K):
        ctx.find_device(mode, **kw)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms_per_step = ev0.elapsed_time(ev1) / K
    res = ctx.fetch(copy=False)
    st = ctx.stats()
    out = dict(config=config, bp=bp, n_genomes=len(seqs), pattern=pattern, mode=mode, ms_per_step=ms_per_step, value=bp / (ms_per_step * 1e-3) / 1e9,
               stats=st, parity=parity_of(res, config, mode, scale))
    out["stages_ms"] = {k: round(st[k], 4) for k in ("ms_extract", "ms_sort", "ms_bucket", "ms_dedup", "ms_output", "ms_total_device")}
    if want_e2e:
        pinned = [torch.from_numpy(s).pin_memory() for s in seqs]

        def e2e_step():
            ctx.clear_sequences()
            for t in pinned:
                ctx.add_sequence_ptr(t.data_ptr(), t.numel())
            return ctx.find(mode, copy=False, compact=True, **kw)  # the compact result form (5 B / component over PCIe)

        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(K):
            r = e2e_step()
        e1.record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3 / K
        e2e_ms = max(e0.elapsed_time(e1) / K, wall_ms)
        st2 = ctx.stats()
        out["e2e"] = {"value": bp / (e2e_ms * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": st2["h2d_bytes"],
                      "d2h_bytes_per_step": st2["d2h_bytes"], "digest_ok": parity_of(r, config, mode, scale)["digest_ok"]}
    return out


def rooflines(m, peak, peak_src, traffic=None):
    """roofline of the dominant kernel (one radix pass = one read + one write of every record) and the whole-path figure of
    SURVEY.md §8d: B_alg = 0.25 + R (3 + 2 P) bytes per input base with P = ceil(2w / 8) — the floor is defined on 8-bit
    digits whatever digit width the sort really uses, so fewer, wider passes show up as a higher fraction."""
    st = m["stats"]
    R, n_seeds = st["record_bytes"], st["n_seeds"]
    roofline = None
    if st["radix_launches"]:
        avg_ms = st["ms_radix_kernels"] / st["radix_launches"]
        alg = 2.0 * R * n_seeds
        ach = alg / (avg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_onesweep (one LSD radix pass over all seed records)", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "avg_launch_ms": avg_ms,
                    "launches_per_step": st["radix_launches"],
                    "share_of_step": st["ms_radix_kernels"] / st["ms_total_device"] if st["ms_total_device"] else None}
    P = (2 * seeds.seed_weight(m["pattern"]) + 7) // 8
    b_alg = 0.25 + R * (3 + 2 * P)
    path = {"b_alg_bytes_per_bp": b_alg, "achieved": b_alg * m["bp"] / (m["ms_per_step"] * 1e-3) / 1e9, "unit": "GB/s"}
    path["frac"] = path["achieved"] / peak
    return roofline, path


def committed_traffic(config):
    """DRAM bytes per launch of k_onesweep from the committed `ncu --set full` capture of the same workload (or None)."""
    for name in (f"r02_ncu_full_c{config}.json", f"r01_ncu_full_c{config}.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))["k_onesweep_traffic_bytes_per_launch"]
        except Exception:
            continue
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=HEADLINE_CONFIG, help="BASELINE.json config 1..5 (default: 5, at every N)")
    ap.add_argument("--scale", type=int, default=1, help="divide genome lengths (debug only; 1 = the named size)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-scale", type=int, default=CPU_SAMPLE_SCALE, help="length divisor of the CPU sample per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1..C4 entries of `extra` (N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import mauvealigner_b200 as mb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 or world > 1:
        from mauvealigner_b200 import dist as mbdist
        return mbdist.bench_main(args, sys.modules[__name__])

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    W, K = max(3, args.warmup), args.steps
    ctx = mb.Context(local)
    stream = torch.cuda.Stream(dev)  # explicit: the library launches on it and the timing events are recorded on it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    sampler = ClockSampler(local)
    sampler.start()
    m = time_config(mb, torch, ctx, stream, dev, args.config, args.scale, W, K, True)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    peak, peak_src = peaks()
    roofline, path = rooflines(m, peak, peak_src, committed_traffic(args.config) if args.scale == 1 else None)
    st = m["stats"]

    extra = {}
    if not args.no_extra and args.scale == 1:
        for c in (1, 2, 3, 4):
            if c == args.config:
                continue
            x = time_config(mb, torch, ctx, stream, dev, c, 1, 3, 5, False)
            xr, xp = rooflines(x, peak, peak_src)
            extra[f"C{c}"] = {"workload": CONFIG_NAMES[c], "value": x["value"], "unit": "Gbp/s", "ms_per_step": x["ms_per_step"], "bp_per_step": x["bp"],
                              "path_roofline": xp, "k_onesweep_frac": xr["frac"] if xr else None, "digest_ok": x["parity"]["digest_ok"],
                              "n_matches": x["parity"]["n_matches"], "stages_ms": x["stages_ms"], "steps": 5, "warmup": 3}
    ctx.close()

    cpu = None if args.no_cpu_baseline else cpu_baseline(args.config, args.scale)
    line = {
        "metric": METRIC, "value": m["value"], "unit": "Gbp/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": m["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_dict(args.config, args.scale),
        "parity": m["parity"], "roofline": roofline, "path_roofline": path, "cpu_baseline": cpu, "e2e": m["e2e"],
        "gpu_launches": int(st["kernel_launches"]) * K, "stages_ms": m["stages_ms"], "clocks": sampler.summary(),
        "counts": {"n_seeds": st["n_seeds"], "n_candidates": st["n_candidates"], "n_extended": st["n_extended"], "dedup_rounds": st["dedup_iters"]},
        "extra": extra,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

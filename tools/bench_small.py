"""Many small problems (recursive anchoring, src/mauveAligner.cpp:94,698): searches per second for
problems of 2 sequences x LEN bases — one mb_find per problem on one context, a pool of contexts on concurrent streams,
and mb_find_batch (all problems in ONE pass of the pipeline, called through the C ABI with host buffers).
usage: python tools/bench_small.py [LEN ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mauvealigner_b200 as mb  # noqa: E402
from mauvealigner_b200 import _lib as L  # noqa: E402
from toygen import family  # noqa: E402


def batch_rate(ctx, problems, pattern, reps):
    """mb_find_batch straight through ctypes (the per-problem Python wrapping of Context.find_batch is not the library)"""
    nprob, nseq = len(problems), len(problems[0])
    ptrs, lens = (C.c_void_p * (nprob * nseq))(), (C.c_uint64 * (nprob * nseq))()
    for i, p in enumerate(problems):
        for g, a in enumerate(p):
            ptrs[i * nseq + g] = a.ctypes.data
            lens[i * nseq + g] = a.size
    ctx.set_seed(pattern)
    prm = L.MbParams(L.MODE_UNIQUE, 0, 2, 1000, 0)
    out = C.POINTER(L.MbBatchResult)()
    lib = mb.lib()
    assert lib.mb_find_batch(ctx._h, C.byref(prm), nprob, nseq, ptrs, lens, C.byref(out)) == 0
    t0 = time.perf_counter()
    for _ in range(reps):
        assert lib.mb_find_batch(ctx._h, C.byref(prm), nprob, nseq, ptrs, lens, C.byref(out)) == 0
    dt = (time.perf_counter() - t0) / reps
    return nprob / dt, int(out.contents.n_matches), ctx.stats()["ms_total_device"]


def main():
    lens = [int(x) for x in sys.argv[1:]] or [500, 5000, 50000]
    rng = np.random.default_rng(1)
    pattern = mb.get_seed(11, 0)
    for n in lens:
        base = [family(rng, n, 2, sub=0.05, indel=0.005, inv=0) for _ in range(64)]
        base = [[np.frombuffer(s.encode(), dtype=np.uint8) for s in p] for p in base]
        problems = base * 8  # 512 searches
        line = f"2 x {n} bp:"
        for streams in (1, 16):
            pool = mb.ContextPool(streams)
            pool.find_many(problems[:2 * streams], pattern)  # warm-up: workspaces allocated
            t0 = time.perf_counter()
            res = pool.find_many(problems, pattern)
            dt = time.perf_counter() - t0
            pool.close()
            line += f"  {streams:2d} context(s) {len(problems) / dt:8.0f}/s"
        want = sum(r["n_matches"] for r in res)
        ctx = mb.Context(0)
        for nb in (512, 16384, 131072):
            if nb * n * 2 > 600_000_000:
                continue
            probs = base * (nb // 64)
            rate, nm, ms_dev = batch_rate(ctx, probs, pattern, 3)
            assert nb != 512 or nm == want, (nm, want)
            line += f"  batch of {nb}: {rate:9.0f}/s ({ms_dev:.2f} ms on the device)"
        ctx.close()
        print(line + f"  [{want} matches per 512]", flush=True)


if __name__ == "__main__":
    main()

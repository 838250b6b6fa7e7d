"""mb_find_multi on N GPUs of this box, one process (the library's own host threads + NVLink peer access, no NCCL):
bit-exact against the single-GPU path, and timed.   usage: python tools/multi_check.py N [CONFIG SCALE STEPS]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mauvealigner_b200 as mb  # noqa: E402


def main():
    n = int(sys.argv[1])
    config = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    scale = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    seqs = mb.synth_genomes(config, scale)
    bp = sum(len(s) for s in seqs)
    pattern = mb.get_seed(15, 0) if config == 1 else mb.get_seed(15, mb.CODING_SEED)
    ctxs = [mb.Context(r) for r in range(n)]
    for c in ctxs:
        for s in seqs:
            c.add_sequence(s)
        c.set_seed(pattern)
    got = mb.find_multi(ctxs)
    want = ctxs[0].find(mb.MODE_UNIQUE)
    ok = got["n_matches"] == want["n_matches"] and all(np.array_equal(np.asarray(got[k], dtype=np.int64), np.asarray(want[k], dtype=np.int64))
                                                       for k in ("length", "comp_off", "comp_seq", "comp_start"))
    import ctypes as C
    from mauvealigner_b200 import _lib as L
    p = ctxs[0]._params(L.MODE_UNIQUE, 2, 1000, False, 0)
    arr = (C.c_void_p * n)(*[c._h for c in ctxs])
    for _ in range(2):
        L.lib().mb_find_multi(arr, n, C.byref(p))
    t0 = time.perf_counter()
    for _ in range(steps):
        rc = L.lib().mb_find_multi(arr, n, C.byref(p))
        assert rc == 0, rc
    dt = (time.perf_counter() - t0) / steps
    print(f"MULTI_CHECK gpus={n} config=C{config}/{scale} matches={got['n_matches']} {'OK bit-exact vs single-GPU path' if ok else 'MISMATCH'} "
          f"{1e3 * dt:.2f} ms per search (wall, results left on the devices) = {bp / dt / 1e9:.2f} Gbp/s", flush=True)
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    main()

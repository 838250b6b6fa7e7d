"""Golden digests of the oracle's results at the FULL BASELINE.json sizes (C1..C5), so that the CUDA path can be pinned
bit for bit at sizes where running the oracle inside the GPU tests would take minutes (C5: 320 Mbp).

    python tests/golden/make_fullsize_digests.py [configs...]   ->  tests/golden/fullsize_digests.json

The digest is SHA-256 over n_matches, n_comps (uint64 LE) and the raw little-endian bytes of length (uint32),
comp_off (uint64), comp_seq (uint32), comp_start (int64), then — in MODE_UNIQUE_COUNT only — unique_mers (uint64).  Inputs: the deterministic synthetic
generator (mb_synth_create, host only), same seeds and parameters as bench.py's config_params."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def digest(res, with_unique_mers=False):
    h = hashlib.sha256()
    h.update(np.asarray([int(res["n_matches"]), int(res["n_comps"])], dtype="<u8").tobytes())
    if int(res["n_matches"]):  # (an empty result has no arrays worth hashing: comp_off would be a lone 0)
        for key, dt in (("length", "<u4"), ("comp_off", "<u8"), ("comp_seq", "<u4"), ("comp_start", "<i8")):
            h.update(np.ascontiguousarray(np.asarray(res[key]).astype(dt, copy=False)).tobytes())
    if with_unique_mers:
        h.update(np.asarray([int(res["unique_mers"])], dtype="<u8").tobytes())
    return h.hexdigest()


def config_params(mb, config):
    if config == 1:
        return mb.get_seed(15, 0), mb.MODE_UNIQUE, {}
    if config in (2, 5):
        return mb.get_seed(15, mb.CODING_SEED), mb.MODE_UNIQUE, {}
    if config == 3:
        return mb.get_seed(19, 0), mb.MODE_UNIQUE_COUNT, {}
    return mb.get_seed(15, 0), mb.MODE_SEED_ENUM, dict(min_multi=2, max_multi=500)


def main():
    import mauvealigner_b200 as mb
    import oracle_lib as O
    path = os.path.join(HERE, "fullsize_digests.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for config in [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4, 5]:
        pattern, mode, kw = config_params(mb, config)
        seqs = mb.synth_genomes(config, 1)
        t0 = time.time()
        res = O.find(seqs, pattern, mode, **kw)
        out[str(config)] = dict(sha256=digest(res, mode == mb.MODE_UNIQUE_COUNT), n_matches=int(res["n_matches"]), n_comps=int(res["n_comps"]),
                                unique_mers=int(res.get("unique_mers", 0)), bp=int(sum(len(s) for s in seqs)), pattern=int(pattern), mode=int(mode),
                                params=kw, oracle_seconds=round(time.time() - t0, 1))
        print(config, out[str(config)], flush=True)
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

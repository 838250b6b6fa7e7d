"""SHA-256 of every synthetic genome of BASELINE.json configs C1..C5 (full size) -> tests/golden/genome_sha256.json.
BASELINE.md §5 cites this file; tests/test_golden.py::test_generator_reproduces_genome_digests re-checks it.
usage: python tools/genome_hashes.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.synth import synth_genomes  # noqa: E402


def config_digests(config, scale=1):
    seqs = synth_genomes(config, scale)
    per = [hashlib.sha256(s.tobytes()).hexdigest() for s in seqs]
    return dict(n_genomes=len(seqs), lengths=[int(len(s)) for s in seqs], bp=int(sum(len(s) for s in seqs)), sha256=per,
                sha256_of_all=hashlib.sha256("".join(per).encode()).hexdigest())


if __name__ == "__main__":
    out = {f"C{c}": config_digests(c) for c in (1, 2, 3, 4, 5)}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "genome_sha256.json"), "w"), indent=1)
    for k, v in out.items():
        print(k, v["n_genomes"], v["bp"], v["sha256_of_all"])

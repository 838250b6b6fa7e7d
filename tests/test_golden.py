"""Committed golden vectors (tests/golden/seed_match_golden.json, made by tests/golden/make_golden.py).
CPU: the oracle must still reproduce them.  GPU: the CUDA path (through the C ABI) must reproduce them bit for bit."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seed_match_golden.json")))
KEYS = ("length", "comp_off", "comp_seq", "comp_start")


def _check(res, case):
    assert int(res["n_matches"]) == case["n_matches"], case["name"]
    for k in KEYS:
        assert [int(x) for x in res[k]] == case[k], (case["name"], k)
    assert int(res["unique_mers"]) == case["unique_mers"], case["name"]


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_reproduces_golden(case):
    res = O.find(case["seqs"], case["pattern"], case["mode"], **case["params"])
    _check(res, case)
    assert [int(x) for x in res["unique_mers_per_seq"]] == case["unique_mers_per_seq"]


def test_oracle_mers_and_sml_golden():
    g = GOLD["mers"]
    assert [int(x) for x in O.mers(g["seq"], g["pattern"])] == g["mers"]
    assert [int(x) for x in O.sml(g["seq"], g["pattern"])] == g["sml"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_cuda_reproduces_golden(case):
    import mauvealigner_b200 as mb
    ctx = mb.Context(0)
    try:
        for s in case["seqs"]:
            ctx.add_sequence(s)
        ctx.set_seed(case["pattern"])
        if case["mode"] == O.MODE_SEED_ENUM and len(case["seqs"]) != 1:
            pytest.skip("SeedMatchEnumerator takes one sequence")
        res = ctx.find(case["mode"], **case["params"])
        _check(res, case)
        if case["mode"] == O.MODE_UNIQUE_COUNT:
            assert [int(x) for x in res["unique_mers_per_seq"]] == case["unique_mers_per_seq"]
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_mers_and_sml_golden():
    import mauvealigner_b200 as mb
    g = GOLD["mers"]
    ctx = mb.Context(0)
    try:
        ctx.add_sequence(g["seq"])
        ctx.set_seed(g["pattern"])
        assert [int(x) for x in ctx.mers(0, len(g["seq"]))] == g["mers"]
        ctx.find(mb.MODE_UNIQUE_COUNT)
        assert [int(x) for x in ctx.sml(0, len(g["seq"]))] == g["sml"]
    finally:
        ctx.close()


def test_oracle_reproduces_fullsize_digest_c1():
    """tests/golden/fullsize_digests.json (the oracle at the FULL BASELINE sizes, pinned as SHA-256 digests for the CUDA
    tests in tests/test_gpu_zz_properties.py) is what the oracle gives today: C1 re-derived here (the larger ones take
    minutes; their generator is tests/golden/make_fullsize_digests.py)."""
    import importlib.util
    import mauvealigner_b200 as mb
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_fullsize_digests.py")
    spec = importlib.util.spec_from_file_location("make_fullsize_digests", path)
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    gold = json.load(open(os.path.join(os.path.dirname(path), "fullsize_digests.json")))
    assert sorted(gold) == ["1", "2", "3", "4", "5"]
    pattern, mode, kw = tool.config_params(mb, 1)
    res = O.find(mb.synth_genomes(1, 1), pattern, mode, **kw)
    assert (res["n_matches"], res["n_comps"]) == (gold["1"]["n_matches"], gold["1"]["n_comps"])
    assert tool.digest(res) == gold["1"]["sha256"]


def test_generator_reproduces_genome_digests():
    """the synthetic genomes are pinned by SHA-256 (tests/golden/genome_sha256.json, tools/genome_hashes.py)."""
    import hashlib
    import json
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tools.synth import synth_genomes
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "genome_sha256.json")))
    for c in (1, 2, 5):
        seqs = synth_genomes(c, 1)
        assert [hashlib.sha256(s.tobytes()).hexdigest() for s in seqs] == want[f"C{c}"]["sha256"], c
        assert [len(s) for s in seqs] == want[f"C{c}"]["lengths"]

"""Multi-GPU path, parity: every rank of a world runs inside this process on ONE B200 (LocalFabric: the exchanges
are tensor copies), through the same C-ABI stages and the same orchestration code the torchrun path uses.
The result must be bit-identical to the oracle (and so to the single-GPU path) for every world size."""
import numpy as np
import pytest

import oracle_lib as O
from toygen import family, rand_seq, revcomp

pytestmark = pytest.mark.gpu


def assert_same(got, want, what=""):
    assert got["n_matches"] == want["n_matches"], (what, got["n_matches"], want["n_matches"])
    for k in ("length", "comp_off", "comp_seq", "comp_start"):
        assert np.array_equal(np.asarray(got[k], dtype=np.int64), np.asarray(want[k], dtype=np.int64)), (what, k)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_emulated_world_matches_oracle(world):
    from mauvealigner_b200 import dist
    rng = np.random.default_rng(500 + world)
    seqs = family(rng, 20000, 5, sub=0.02, indel=0.003, inv=1)
    seqs[2] = revcomp(seqs[2])
    for pattern in (0b110111011, 0b1101110111110111011):
        got = dist.find_unique_emulated(seqs, pattern, world)
        assert_same(got, O.find(seqs, pattern, O.MODE_UNIQUE), f"world {world}")
        assert sum(i["seeds_received"] for i in got["info"]) == sum(max(0, len(s) - pattern.bit_length() + 1) for s in seqs)


def test_emulated_world_nccl_style_exchange1():
    """all-to-alls of local send buffers (0), exchange 1 fused into the partition pass (1), exchanges 2 / 3 fused into
    the pack kernels too (2): same result"""
    from mauvealigner_b200 import dist
    rng = np.random.default_rng(77)
    seqs = family(rng, 30000, 4, sub=0.03, indel=0.002, inv=1)
    want = O.find(seqs, 0b1101110111110111011, O.MODE_UNIQUE)
    for p2p in (0, 1, 2, 3):
        got = dist.find_unique_emulated(seqs, 0b1101110111110111011, 3, p2p=p2p)
        assert_same(got, want, f"p2p={p2p}")


SOLID31 = (1 << 31) - 1  # weight 31: 62 seed bits, so (seed, genome, position, strand) needs 16-byte records


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_emulated_unique_count(world):
    """MB_MODE_UNIQUE_COUNT over several ranks (the counts of the key ranges add up), 8-byte and 16-byte records"""
    import mauvealigner_b200 as mb
    from mauvealigner_b200 import dist
    rng = np.random.default_rng(640 + world)
    seqs = family(rng, 30000, 3, sub=0.05, indel=0.004, inv=1)
    seqs[0] = seqs[0] + seqs[0][2000:9000] + revcomp(seqs[0][100:3000])  # repeated mers inside one genome
    for pattern, rec in ((0b110111011, 8), (mb.get_seed(15, 0), 8), (SOLID31, 16)):
        got = dist.find_enum_emulated(seqs, pattern, world, mb.MODE_UNIQUE_COUNT)
        want = O.find(seqs, pattern, O.MODE_UNIQUE_COUNT)
        assert got["record_bytes"] == rec
        assert got["unique_mers"] == want["unique_mers"], (world, pattern)
        assert np.array_equal(np.asarray(got["unique_mers_per_seq"], dtype=np.int64), np.asarray(want["unique_mers_per_seq"], dtype=np.int64))
    # one genome (the uniqueMerCount tool), fewer tiles than ranks
    got = dist.find_enum_emulated([seqs[0][:700]], 0b1011101, world, mb.MODE_UNIQUE_COUNT)
    assert got["unique_mers"] == O.find([seqs[0][:700]], 0b1011101, O.MODE_UNIQUE_COUNT)["unique_mers"]


@pytest.mark.parametrize("world", [1, 2, 5])
def test_emulated_seed_enum(world):
    """MB_MODE_SEED_ENUM over several ranks: the ranks' pieces (key ranges) are disjoint; merged by first position they are
    the single-GPU list — multiplicity windows, direct_only, both record formats"""
    import mauvealigner_b200 as mb
    from mauvealigner_b200 import dist
    from toygen import mutate
    rng = np.random.default_rng(760 + world)
    unit = rand_seq(rng, 400)
    s = (rand_seq(rng, 20000) + unit + rand_seq(rng, 900) + mutate(rng, unit, sub=0.04, indel=0) + revcomp(unit) + rand_seq(rng, 5000) + "AT" * 200 + unit +
         rand_seq(rng, 3000) + "ACG" * 100)
    for pattern, kw in ((0b110111011, dict(min_multi=2, max_multi=500)), (mb.get_seed(11, 1), dict(min_multi=3, max_multi=1000)),
                        (mb.get_seed(9, 0), dict(min_multi=2, max_multi=4, direct_only=True)), (SOLID31, dict(min_multi=2, max_multi=1000))):
        got = dist.find_enum_emulated([s], pattern, world, mb.MODE_SEED_ENUM, **kw)
        want = O.find([s], pattern, O.MODE_SEED_ENUM, **kw)
        assert_same(got, want, f"world {world} pattern {pattern:b}")
        assert got["n_matches"] > 0
        assert sum(i["seeds_received"] for i in got["info"]) == len(s) - pattern.bit_length() + 1
    # nothing to report, and a sequence shorter than the seed
    for short in (rand_seq(rng, 300), "ACGTAC"):
        got = dist.find_enum_emulated([short], mb.get_seed(15, 0), world, mb.MODE_SEED_ENUM, min_multi=2, max_multi=10)
        assert_same(got, O.find([short], mb.get_seed(15, 0), O.MODE_SEED_ENUM, min_multi=2, max_multi=10))


def test_emulated_edge_cases():
    from mauvealigner_b200 import dist
    rng = np.random.default_rng(9)
    # fewer tiles than ranks, empty and short sequences, identical genomes (one long match)
    s = rand_seq(rng, 3000)
    for seqs in ([s, "", "ACG", s], [s, revcomp(s)], ["ACGT", "ACGTA"]):
        got = dist.find_unique_emulated(seqs, 0b11111, 4)
        assert_same(got, O.find(seqs, 0b11111, O.MODE_UNIQUE))


@pytest.mark.parametrize("config,scale,world", [(5, 100, 8), (2, 50, 4), (1, 50, 2)])
def test_emulated_baseline_configs(config, scale, world):
    import mauvealigner_b200 as mb
    from mauvealigner_b200 import dist
    seqs = mb.synth_genomes(config, scale)
    pattern = mb.get_seed(15, 0) if config == 1 else mb.get_seed(15, mb.CODING_SEED)
    got = dist.find_unique_emulated(seqs, pattern, world, p2p=2 if config == 5 else None)  # C5: every exchange fused
    assert_same(got, O.find(seqs, pattern, O.MODE_UNIQUE), f"C{config}")
    # the key-range partition is balanced to a few percent on these inputs
    recv = [i["seeds_received"] for i in got["info"]]
    assert max(recv) < 1.25 * (sum(recv) / world)
    # ... and so is the range partition of the result
    out = [i["matches"] for i in got["info"]]
    assert sum(out) == got["n_matches"] and max(out) < 1.5 * (sum(out) / world) + 64


@pytest.mark.parametrize("world", [1, 2, 5])
def test_find_multi_cpp_driver(world):
    """mb_find_multi: the same stages driven by the library's own host threads (one per context), every exchange device
    to device — here all contexts on the GPUs present (device r % n_devices), so a single-GPU box runs it too."""
    import mauvealigner_b200 as mb
    ndev = mb.lib().mb_device_count()
    rng = np.random.default_rng(900 + world)
    seqs = family(rng, 25000, 4, sub=0.02, indel=0.003, inv=1)
    seqs[1] = revcomp(seqs[1])
    pattern = 0b1101110111110111011
    ctxs = [mb.Context(r % max(1, ndev)) for r in range(world)]
    try:
        for c in ctxs:
            for s in seqs:
                c.add_sequence(s)
            c.set_seed(pattern)
        got = mb.find_multi(ctxs)
        again = mb.find_multi(ctxs)  # buffers and peer mappings reused
    finally:
        for c in ctxs:
            c.close()
    want = O.find(seqs, pattern, O.MODE_UNIQUE)
    assert_same(got, want, f"find_multi world {world}")
    assert_same(again, want, f"find_multi world {world}, second run")


def test_find_multi_on_distinct_devices():
    """mb_find_multi with one context per PHYSICAL GPU (NVLink peer stores and copy-engine pushes between real devices);
    needs at least two GPUs on the box."""
    import mauvealigner_b200 as mb
    ndev = mb.lib().mb_device_count()
    if ndev < 2:
        pytest.skip("one GPU on this box")
    world = min(ndev, 4)
    seqs = mb.synth_genomes(5, 40)
    pattern = mb.get_seed(15, mb.CODING_SEED)
    ctxs = [mb.Context(r) for r in range(world)]
    try:
        for c in ctxs:
            for s in seqs:
                c.add_sequence(s)
            c.set_seed(pattern)
        got = mb.find_multi(ctxs)
        want = ctxs[0].find(mb.MODE_UNIQUE)
    finally:
        for c in ctxs:
            c.close()
    assert_same(got, want, f"find_multi on {world} devices")
    assert want["n_matches"] > 1000


def test_torchrun_nccl_path_on_two_gpus():
    """The one-process-per-GPU path of bench.py (torchrun + NCCL + CUDA IPC peer stores) on two real GPUs, bit for bit
    against the single-GPU path (tools/dist_check.py); needs at least two GPUs on the box."""
    import os
    import subprocess
    import sys
    import mauvealigner_b200 as mb
    if mb.lib().mb_device_count() < 2:
        pytest.skip("one GPU on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p2p in ("1", "0"):
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                              "--master-port", "29533", os.path.join(root, "tools", "dist_check.py"), "5", "20"],
                             capture_output=True, text=True, timeout=600, env=dict(os.environ, MB_DIST_P2P=p2p))
        assert out.returncode == 0, out.stderr[-2000:]
        assert "OK bit-exact vs single-GPU path" in out.stdout, out.stdout[-2000:]

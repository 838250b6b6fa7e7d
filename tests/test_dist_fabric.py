"""Host-side logic of the multi-GPU path on CPU: the two fabrics must deliver the same words.  TorchFabric runs as
two real processes over gloo (world_size 2); LocalFabric is the in-process emulation the GPU parity tests use."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from mauvealigner_b200.dist import LocalFabric, TorchFabric


def _payload(world):
    rng = np.random.default_rng(3)
    counts = [[int(rng.integers(0, 50)) for _ in range(world)] for _ in range(world)]
    counts[1][0] = 0
    sends = [torch.from_numpy(rng.integers(-2 ** 62, 2 ** 62, size=2 * sum(c), dtype=np.int64)) for c in counts]
    return counts, sends


def _expected(world, counts, sends, width):
    fab = LocalFabric(world)
    rc = fab.counts(counts)
    recvs = [torch.zeros(width * sum(k), dtype=torch.int64) for k in rc]
    fab.words(sends, counts, recvs, rc, width=width)
    return rc, recvs


def test_local_fabric_routes_by_destination():
    counts, sends = _payload(3)
    rc, recvs = _expected(3, counts, sends, 2)
    for d in range(3):
        parts = []
        for s in range(3):
            o = 2 * sum(counts[s][:d])
            parts.append(sends[s][o:o + 2 * counts[s][d]])
        assert torch.equal(recvs[d], torch.cat(parts))
        assert rc[d] == [counts[s][d] for s in range(3)]


def test_local_fabric_count_matrix_and_offsets():
    counts, _ = _payload(3)
    fab = LocalFabric(3)
    M = fab.gather_counts(counts)
    assert M == counts
    # block of source r inside destination d's receive array starts after the blocks of the lower ranks
    rc = fab.counts(counts)
    for d in range(3):
        for r in range(3):
            assert sum(M[s][d] for s in range(r)) == sum(rc[d][:r])


def test_local_fabric_allreduce():
    fab = LocalFabric(3)
    ts = [torch.arange(5, dtype=torch.int64) * (r + 1) for r in range(3)]
    fab.allreduce_sum(ts)
    for t in ts:
        assert t.tolist() == [0, 6, 12, 18, 24]


def test_orchestration_in_process_world3():
    """find_unique over stand-in contexts on the host (tests/fake_dist_ctx.py): every stage checks that exactly its
    peers' words arrived in source-rank order — seeds, 4-word rows, verdict bytes going back, summed histogram, match
    headers and components — for the all-to-all exchanges and for the three peer-memory levels (block offsets from the
    count matrix, stores / pushes into the peers' buffers)."""
    from fake_dist_ctx import FakeCtx, n_match, n_rows, n_seeds
    from mauvealigner_b200.dist import find_unique
    world = 3
    ctxs = [FakeCtx() for _ in range(world)]
    for p2p in (0, 0, 1, 2, 3):  # NCCL-style twice (buffers are re-requested every run), then the three peer-memory levels
        info = find_unique(ctxs, LocalFabric(world), torch.device("cpu"), p2p=p2p)
        assert all(c.done for c in ctxs)
        for r in range(world):
            assert info[r]["seeds_sent"] == sum(n_seeds(r, d) for d in range(world))
            assert info[r]["seeds_received"] == sum(n_seeds(s, r) for s in range(world))
            assert info[r]["candidates_local"] == sum(n_rows(r, d) for d in range(world))
            assert info[r]["candidates_owned"] == sum(n_rows(s, r) for s in range(world))
            assert info[r]["matches"] == sum(n_match(s, r) for s in range(world))


def test_enum_orchestration_in_process_world3():
    """find_enum (MODE_UNIQUE_COUNT / MODE_SEED_ENUM over several ranks) over the stand-in contexts: both record formats
    arrive in source-rank order, the counts are summed over the ranks"""
    from fake_dist_ctx import FakeCtx, n_seeds
    from mauvealigner_b200 import MODE_SEED_ENUM, MODE_UNIQUE_COUNT
    from mauvealigner_b200.dist import find_enum
    world = 3
    for wide in (False, True):
        ctxs = [FakeCtx() for _ in range(world)]
        for c in ctxs:
            c.wide = wide
        info = find_enum(ctxs, LocalFabric(world), torch.device("cpu"), MODE_UNIQUE_COUNT)
        assert all(c.done for c in ctxs)
        for r in range(world):
            assert info[r]["record_bytes"] == (16 if wide else 8)
            assert info[r]["seeds_received"] == sum(n_seeds(s, r) for s in range(world))
            assert info[r]["unique_mers"] == sum(10 * q + 1 for q in range(world)) and info[r]["unique_mers_local"] == 10 * r + 1
            assert info[r]["unique_mers_per_seq"] == [sum(q + 1 for q in range(world)), sum(2 * q for q in range(world))]
        info = find_enum(ctxs, LocalFabric(world), torch.device("cpu"), MODE_SEED_ENUM, min_multi=2, max_multi=9)
        assert all(c.done and c.mode == MODE_SEED_ENUM for c in ctxs) and "unique_mers" not in info[0]
    with pytest.raises(ValueError):
        find_enum(ctxs, LocalFabric(world), torch.device("cpu"), 0)


def test_merge_enum_results_restores_canonical_order():
    """the ranks' MODE_SEED_ENUM pieces (sorted by first position, disjoint) merge into one list sorted by first position"""
    from mauvealigner_b200.dist import merge_enum_results
    a = dict(n_matches=2, n_comps=5, length=np.array([9, 7]), comp_off=np.array([0, 2, 5], dtype=np.uint64), comp_seq=np.zeros(5, dtype=np.uint32),
             comp_start=np.array([3, -40, 20, 25, -90], dtype=np.int64))
    b = dict(n_matches=0, n_comps=0, length=np.zeros(0, dtype=np.uint32), comp_off=np.zeros(1, dtype=np.uint64), comp_seq=np.zeros(0, dtype=np.uint32),
             comp_start=np.zeros(0, dtype=np.int64))
    c = dict(n_matches=2, n_comps=4, length=np.array([5, 6]), comp_off=np.array([0, 2, 4], dtype=np.uint64), comp_seq=np.zeros(4, dtype=np.uint32),
             comp_start=np.array([1, 2, 10, -11], dtype=np.int64))
    m = merge_enum_results([a, b, c])
    assert m["n_matches"] == 4 and m["n_comps"] == 9
    assert m["length"].tolist() == [5, 9, 6, 7]
    assert m["comp_off"].tolist() == [0, 2, 4, 6, 9]
    assert m["comp_start"].tolist() == [1, 2, 3, -40, 10, -11, 20, 25, -90]
    assert merge_enum_results([b, b])["n_matches"] == 0


def _enum_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_dist_ctx import FakeCtx
        from mauvealigner_b200 import MODE_UNIQUE_COUNT
        from mauvealigner_b200.dist import find_enum
        ctx = FakeCtx()
        ctx.wide = True
        info = find_enum([ctx], TorchFabric(), torch.device("cpu"), MODE_UNIQUE_COUNT)
        ok = ctx.done and info[0]["unique_mers"] == sum(10 * q + 1 for q in range(world)) and info[0]["unique_mers_per_seq"] == [3, 2]
        out.put((rank, bool(ok)))
    except Exception as e:
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_enum_orchestration_gloo_world2():
    """find_enum over two real processes (gloo): 16-byte records through two all-to-alls, counts through the all-reduce"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_enum_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert res == {0: True, 1: True}


def _orchestration_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_dist_ctx import FakeCtx, n_match
        from mauvealigner_b200.dist import find_unique
        ctx = FakeCtx()
        info = find_unique([ctx], TorchFabric(), torch.device("cpu"), p2p=0)
        ok = ctx.done and info[0]["rank"] == rank and info[0]["matches"] == sum(n_match(s, rank) for s in range(world))
        out.put((rank, bool(ok)))
    except Exception as e:  # surfaced by the parent as a failed rank
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_orchestration_gloo_world2():
    """the same over two real processes and torch.distributed (gloo): the N > 1 host path of the bench, without a GPU"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_orchestration_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert res == {0: True, 1: True}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts, sends = _payload(world)
        fab = TorchFabric()
        rc = fab.counts([counts[rank]])
        recv = torch.zeros(2 * sum(rc[0]), dtype=torch.int64)
        fab.words([sends[rank]], [counts[rank]], [recv], rc, width=2)
        exp_rc, exp = _expected(world, counts, sends, 2)
        ok = rc[0] == exp_rc[rank] and torch.equal(recv, exp[rank])
        # interleaved pair counts, as exchanged between stage 2 and 3
        pairs = [x for p in zip(counts[rank], [c + 1 for c in counts[rank]]) for x in p]
        both = fab.counts([pairs])[0]
        ok = ok and both[0::2] == exp_rc[rank] and both[1::2] == [c + 1 for c in exp_rc[rank]]
        h = torch.arange(4, dtype=torch.int64) + rank
        fab.allreduce_sum([h])
        ok = ok and h.tolist() == [1, 3, 5, 7]
        # the count matrix of the fused exchange 1, and the verdict bytes of exchange 2b (uint8, reversed counts)
        ok = ok and fab.gather_counts([counts[rank]]) == counts
        fab.barrier()
        vb = [torch.arange(sum(exp_rc[r]), dtype=torch.uint8) + r for r in range(world)]
        back = torch.zeros(sum(counts[rank]), dtype=torch.uint8)
        fab.words([vb[rank]], [exp_rc[rank]], [back], [counts[rank]])
        exp_back = [torch.zeros(sum(counts[r]), dtype=torch.uint8) for r in range(world)]
        LocalFabric(world).words(vb, exp_rc, exp_back, counts)
        ok = ok and torch.equal(back, exp_back[rank])
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_torch_fabric_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert res == {0: True, 1: True}

"""Multi-GPU plumbing of the seed-match path: one process per GPU, torch.distributed (NCCL over
NVLink) for the three exchanges, the library (libmauve_b200.so, mb_dist_*) for every per-rank stage.

    stage 1  extract this rank's slice of seeds, partition by seed-key range      -> all-to-all of seed records
    stage 2  sort / runs / policy over the received key range -> candidate rows,
             partitioned by owner of their de-dup group                            -> all-to-all of rows
    stage 3  chains / extension / resolve over the owned groups; histogram of the accepted matches' canonical
             keys (all-reduce) -> match rows by range of the canonical order       -> all-to-all of rows
    stage 4  every rank: canonical order + CSR of its range (the pieces in rank order are the result)

The orchestration is written over a small "fabric" interface so the same code runs (a) one rank per process over
torch.distributed (TorchFabric) and (b) all ranks of a world inside ONE process on one GPU (LocalFabric: the
exchanges are tensor copies), which is how the parity tests check world sizes 2..8 on a single B200.
Nothing here touches sequence data on the host."""
from __future__ import annotations

import json
import os
import sys
import time

import torch

from . import _lib as L


class _DevWords:
    """Zero-copy view of library-owned device memory as an int64 torch tensor."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


class _DevBytes:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _is_host(device):
    return device is not None and torch.device(device).type == "cpu"


def _host_view(ptr, n, ctype, dtype):
    # host memory (the CPU tests of the orchestration drive it with stand-in contexts whose buffers live on the host)
    import ctypes
    import numpy as np
    return torch.from_numpy(np.ctypeslib.as_array((ctype * int(n)).from_address(int(ptr)))).view(dtype)


def dev_bytes(ptr, n, device):
    if n == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    if _is_host(device):
        import ctypes
        return _host_view(ptr, n, ctypes.c_uint8, torch.uint8)
    return torch.as_tensor(_DevBytes(ptr, n), device=device)


def dev_words(ptr, n, device):
    if n == 0 or not ptr:
        return torch.empty(0, dtype=torch.int64, device=device)
    if _is_host(device):
        import ctypes
        return _host_view(ptr, n, ctypes.c_int64, torch.int64)
    return torch.as_tensor(_DevWords(ptr, n), device=device)


class LocalFabric:
    """All ranks of a world inside one process (emulation on one GPU, or any CPU tensors)."""

    def __init__(self, world):
        self.world = world
        self.local_ranks = list(range(world))

    def counts(self, send_counts):
        # k numbers per destination: send_counts[src][dst*k:(dst+1)*k] -> recv_counts[dst][src*k:(src+1)*k]
        k = len(send_counts[0]) // self.world
        return [[x for s in range(self.world) for x in send_counts[s][d * k:(d + 1) * k]] for d in range(self.world)]

    def gather_counts(self, send_counts):
        # the whole count matrix M[src][dst], the same on every rank
        return [list(c) for c in send_counts]

    def barrier(self):
        pass  # one process, one stream

    def peer_buffers(self, ctxs, key, need):
        """The receive buffer `key` ("seed" records, "hdr" row words, "comp" component words) of every rank, addressable
        from every rank: here plain pointers of one process.  need[d] = units rank d must hold."""
        ptrs = [_alloc_recv(c, key, n) for c, n in zip(ctxs, need)]
        return [ptrs for _ in ctxs]

    def release_peer_buffers(self, ctxs):
        pass

    def p2p_ok(self, ctxs):
        return True

    def allreduce_sum(self, tensors):
        total = tensors[0].clone()
        for t in tensors[1:]:
            total += t
        for t in tensors:
            t.copy_(total)

    def words(self, sends, send_counts, recvs, recv_counts, width=1):
        for d in range(self.world):
            o = 0
            for s in range(self.world):
                n = send_counts[s][d] * width
                so = sum(send_counts[s][:d]) * width
                if n:
                    recvs[d][o:o + n].copy_(sends[s][so:so + n])
                o += n


class TorchFabric:
    """One rank per process over torch.distributed (NCCL for CUDA tensors, gloo for CPU tensors)."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local_ranks = [self.rank]
        self.device = device

    def counts(self, send_counts):
        t = torch.tensor(send_counts[0], dtype=torch.int64, device=self.device)
        k = t.numel() // self.world
        out = torch.empty_like(t)
        self.dist.all_to_all_single(out, t, [k] * self.world, [k] * self.world, group=self.group)
        return [out.tolist()]

    def gather_counts(self, send_counts):
        t = torch.tensor(send_counts[0], dtype=torch.int64, device=self.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return torch.stack(out).tolist()

    def bind_streams(self, ctxs):
        """The NCCL exchanges and the barrier below are ordered on torch's CURRENT stream; the library stages must run
        on that same stream, or a collective could read a send buffer before the pack kernel has finished (and a later
        stage a receive buffer before NCCL has filled it).  find_unique calls this first."""
        if self.device is None or torch.device(self.device).type != "cuda":
            return  # host stand-ins (CPU tests over gloo) have no stream
        cur = torch.cuda.current_stream(self.device).cuda_stream
        for c in ctxs:
            c.set_stream(cur)

    def barrier(self):
        # stream-ordered: completes on this rank only after every rank's earlier work on its stream (the peer stores
        # of its partition kernel) has completed
        if not hasattr(self, "_token"):
            self._token = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.dist.all_reduce(self._token, group=self.group)

    def peer_buffers(self, ctxs, key, need):
        """The receive buffer `key` ("seed" records, "hdr" row words, "comp" component words) of every rank, mapped into
        this process through CUDA IPC (NVLink peer access).  need[d] = units rank d must hold — known to every rank from
        the count matrix, so all ranks take the same decision without talking: the mapping is redone (collectively)
        only when some rank's need outgrows the capacity it exported."""
        pb = self.__dict__.setdefault("_pb", {})
        st = pb.setdefault(key, dict(cap=[-1] * self.world, ptrs=None))
        if st["ptrs"] is not None and all(n <= cp for n, cp in zip(need, st["cap"])):
            return [st["ptrs"]]
        c = ctxs[0]
        self._unmap(c, st)
        st["cap"] = [max(cp, n + n // 8 + 4096) for n, cp in zip(need, st["cap"])]
        mine = _alloc_recv(c, key, st["cap"][self.rank])
        h = torch.frombuffer(bytearray(c.ipc_export(mine)), dtype=torch.uint8).to(self.device)
        out = [torch.empty_like(h) for _ in range(self.world)]
        self.dist.all_gather(out, h, group=self.group)
        st["ptrs"] = [mine if r == self.rank else c.ipc_import(out[r].cpu().numpy().tobytes()) for r in range(self.world)]
        return [st["ptrs"]]

    def p2p_ok(self, ctxs):
        """Collective capability probe, once per fabric: can every rank map every other rank's device memory (CUDA IPC +
        peer access over NVLink)?  If not, the exchanges stay NCCL all-to-alls (still GPU to GPU; nothing moves to
        the host)."""
        if "_p2p_ok" not in self.__dict__:
            ok = 1
            try:
                self.peer_buffers(ctxs, "seed", [0] * self.world)
            except Exception as e:  # the other ranks may be waiting in the handle all-gather: it has completed by now
                print(f"[mauve_b200.dist] rank {self.rank}: peer mapping unavailable ({e}); exchanges stay on NCCL", file=sys.stderr, flush=True)
                ok = 0
            t = torch.tensor([ok], dtype=torch.int64, device=self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
            self._p2p_ok = bool(int(t.item()))
        return self._p2p_ok

    def _unmap(self, c, st):
        # unmap the peers' buffers before any rank frees or regrows its own
        if st["ptrs"] is not None:
            for r, p in enumerate(st["ptrs"]):
                if r != self.rank:
                    c.ipc_close(p)
            st["ptrs"] = None
        torch.cuda.synchronize(self.device)
        self.dist.barrier(group=self.group)

    def release_peer_buffers(self, ctxs):
        """Collective: unmap every peer buffer (before the contexts are destroyed)."""
        for key in sorted(self.__dict__.get("_pb", {})):
            st = self._pb[key]
            if st["ptrs"] is not None:
                self._unmap(ctxs[0], st)
            st["cap"] = [-1] * self.world

    def allreduce_sum(self, tensors):
        self.dist.all_reduce(tensors[0], group=self.group)

    def words(self, sends, send_counts, recvs, recv_counts, width=1):
        self.dist.all_to_all_single(recvs[0], sends[0], [c * width for c in recv_counts[0]], [c * width for c in send_counts[0]],
                                    group=self.group)


def _alloc_recv(ctx, key, n):
    if key == "seed":
        return ctx.dist_p2p_recv_array(n)
    return ctx.dist_recv_buffer(1 if key == "hdr" else 4, n)


def p2p_default():
    """MB_DIST_P2P: 0 = every exchange is an NCCL all-to-all of a local send buffer; 1 (default) = exchange 1 (seed
    records, the largest) is fused into the partition kernel as NVLink peer stores; 2 = exchanges 2 and 3 (candidate
    rows, match rows) too; 3 = exchange 1 fused, exchanges 2 and 3 packed locally and pushed into the peers' buffers by
    the copy engines (cudaMemcpyAsync device to device over NVLink).  Measured at 8 x B200 on C5: level 1 is the fastest — the partition pass keeps enough
    stores in flight to beat the all-to-all (0.9 vs 2.4 ms), the gather-bound row pack kernels do not (rows 1.05 vs
    0.6 ms, match rows 1.6 vs 1.2 ms), so level 2 stays an option, not the default."""
    return int(os.environ.get("MB_DIST_P2P", "1"))


def find_unique(ctxs, fabric, device, nway_mask=0, p2p=None):
    """MODE_UNIQUE over the ranks of `fabric`; ctxs[i] is the library context of fabric.local_ranks[i] (sequences and
    seed already set, identical on every rank).  Leaves every rank's piece of the canonical match CSR on its device
    (fetch it with ctxs[...].fetch()); returns per-local-rank info dicts.
    p2p (see p2p_default): 1 = exchange #1 fused into the partition pass, 2 = exchanges #2 and #3 fused into the row
    pack kernels as well (the kernels store straight into the destination ranks' receive buffers over NVLink; the
    ranks share the count matrix first and meet at a stream-ordered barrier afterwards), 0 = NCCL all-to-alls of
    local send buffers only."""
    W, R = fabric.world, fabric.local_ranks
    if hasattr(fabric, "bind_streams"):
        fabric.bind_streams(ctxs)  # library stages and NCCL exchanges on ONE stream (torch's current one)
    if p2p is None:
        p2p = p2p_default()
    p2p = int(p2p)
    if p2p and W > 1 and not fabric.p2p_ok(ctxs):
        p2p = 0
    info = [dict(rank=r) for r in R]
    trace = os.environ.get("MB_DIST_TRACE")
    marks = []

    def mark(name):
        if trace:
            torch.cuda.synchronize(device)
            marks.append((name, time.perf_counter()))

    mark("start")
    # ---- stage 1 + exchange 1: seed records by key range
    if p2p:
        # fused: every rank counts its slice per destination, the ranks share the count matrix, and the partition
        # kernel itself stores each record into its destination's receive array (source-rank order there)
        sc = [c.dist_extract_count(r, W) for c, r in zip(ctxs, R)]
        mark("stage1a extract+count")
        M = fabric.gather_counts(sc)
        peers = fabric.peer_buffers(ctxs, "seed", [sum(M[s][d] for s in range(W)) for d in range(W)])
        for c, r, pp in zip(ctxs, R, peers):
            c.dist_partition_p2p(pp, [sum(M[s][d] for s in range(r)) for d in range(W)])
        fabric.barrier()
        rc = [[M[s][r] for s in range(W)] for r in R]
        for c in ctxs:
            c.dist_use_p2p_recv(True)
        mark("stage1b partition with peer stores")
    else:
        s1 = [c.dist_extract(r, W) for c, r in zip(ctxs, R)]
        sc = [cnt for _, cnt in s1]
        mark("stage1 extract+partition")
        rc = fabric.counts(sc)
        sends = [dev_words(p, sum(cnt), device) for p, cnt in s1]
        recvs = [dev_words(c.dist_recv_buffer(0, sum(k)), sum(k), device) for c, k in zip(ctxs, rc)]
        fabric.words(sends, sc, recvs, rc)
        for c in ctxs:
            c.dist_use_p2p_recv(False)
        mark("exchange1 seeds")
    for i, k in enumerate(rc):
        info[i]["seeds_sent"], info[i]["seeds_received"] = sum(sc[i]), sum(k)
    # ---- stage 2 + exchange 2: every candidate extended at its source; 4-word rows to the owner of the de-dup group
    scc = [c.dist_local(W, sum(k), nway_mask=nway_mask) for c, k in zip(ctxs, rc)]
    mark("stage2 sort+buckets+extend")
    if p2p >= 2:
        M2 = fabric.gather_counts(scc)
        peers = fabric.peer_buffers(ctxs, "hdr", [4 * sum(M2[s][d] for s in range(W)) for d in range(W)])
        for c, r, pp, cc in zip(ctxs, R, peers, scc):
            offs = [sum(M2[s][d] for s in range(r)) for d in range(W)]
            if p2p == 2:
                c.dist_rows_pack(pp, offs)            # the pack kernel stores into the owners' buffers
            else:
                c.dist_push(c.dist_rows_pack(), cc, 4, pp, offs)  # local pack, then copy-engine pushes
        fabric.barrier()
        rcc = [[M2[s][r] for s in range(W)] for r in R]
        mark("rows packed into the owners' buffers")
    else:
        rs = [dev_words(c.dist_rows_pack(), 4 * sum(cc), device) for c, cc in zip(ctxs, scc)]
        rcc = fabric.counts(scc)
        rr = [dev_words(c.dist_recv_buffer(1, 4 * sum(k)), 4 * sum(k), device) for c, k in zip(ctxs, rcc)]
        fabric.words(rs, scc, rr, rcc, width=4)
        mark("rows + exchange2")
    for i in range(len(R)):
        info[i]["candidates_local"], info[i]["candidates_owned"] = sum(scc[i]), sum(rcc[i])
    # ---- stage 3a (owner): chains / resolve; one verdict byte per row goes back to the row's source
    vs = [dev_bytes(c.dist_resolve(sum(k)), sum(k), device) for c, k in zip(ctxs, rcc)]
    mark("stage3a chains+resolve")
    vr = [dev_bytes(c.dist_recv_buffer(5, sum(cc)), sum(cc), device) for c, cc in zip(ctxs, scc)]
    fabric.words(vs, rcc, vr, scc)
    mark("exchange2b verdicts")
    # ---- stage 3b (source): accepted candidates = matches; they go to the rank that owns their range of the canonical order
    hists = [c.dist_accept() for c in ctxs]
    fabric.allreduce_sum([dev_words(p, 4096, device) for p in hists])
    s3 = [c.dist_match_partition(W) for c in ctxs]
    mark("stage3b matches")
    pairs = [[x for pair in zip(cc, mc) for x in pair] for cc, mc in s3]
    if p2p >= 2:
        M3 = fabric.gather_counts(pairs)
        ph = fabric.peer_buffers(ctxs, "hdr", [2 * sum(M3[s][2 * d] for s in range(W)) for d in range(W)])
        pc = fabric.peer_buffers(ctxs, "comp", [sum(M3[s][2 * d + 1] for s in range(W)) for d in range(W)])
        for c, r, hh, cc_, (nh, nc) in zip(ctxs, R, ph, pc, s3):
            oh = [sum(M3[s][2 * d] for s in range(r)) for d in range(W)]
            oc = [sum(M3[s][2 * d + 1] for s in range(r)) for d in range(W)]
            if p2p == 2:
                c.dist_match_pack(hh, oh, cc_, oc)
            else:
                hsrc, csrc = c.dist_match_pack()
                c.dist_push(hsrc, nh, 2, hh, oh)
                c.dist_push(csrc, nc, 1, cc_, oc)
        fabric.barrier()
        gcc = [[M3[s][2 * r] for s in range(W)] for r in R]
        gmc = [[M3[s][2 * r + 1] for s in range(W)] for r in R]
        mark("match rows packed into the destinations' buffers")
    else:
        both = fabric.counts(pairs)
        gcc = [b[0::2] for b in both]
        gmc = [b[1::2] for b in both]
        packed = [c.dist_match_pack() for c in ctxs]
        hs = [dev_words(h, 2 * sum(cc), device) for (h, _), (cc, _mc) in zip(packed, s3)]
        ms = [dev_words(m, sum(mc), device) for (_, m), (_cc, mc) in zip(packed, s3)]
        hr = [dev_words(c.dist_recv_buffer(3, 2 * sum(k)), 2 * sum(k), device) for c, k in zip(ctxs, gcc)]
        mr = [dev_words(c.dist_recv_buffer(4, sum(k)), sum(k), device) for c, k in zip(ctxs, gmc)]
        fabric.words(hs, [cc for cc, _ in s3], hr, gcc, width=2)
        fabric.words(ms, [mc for _, mc in s3], mr, gmc)
        mark("match rows + exchange3")
    # ---- stage 4: every rank builds the canonical CSR of its range
    for i in range(len(R)):
        info[i]["matches_accepted"] = sum(s3[i][0])
        ctxs[i].dist_output(sum(gcc[i]), sum(gmc[i]))
        info[i]["matches"] = sum(gcc[i])
    mark("stage4 output")
    if trace and R[0] == 0:
        print("[dist-trace] " + ", ".join(f"{n} {1e3 * (t - marks[k][1]):.2f} ms" for k, (n, t) in enumerate(marks[1:])), file=sys.stderr, flush=True)
    return info


def find_enum(ctxs, fabric, device, mode, min_multi=2, max_multi=1000, direct_only=False):
    """MODE_UNIQUE_COUNT / MODE_SEED_ENUM over the ranks of `fabric` (8-byte or 16-byte seed records): one exchange — the
    seed records by key range — then every rank runs the mode's tail over its range.  Equal seeds meet on one rank, so
    the counts add up (summed here: every rank's info carries the totals) and the match lists are disjoint (fetch every
    rank's piece; merge_enum_results gives the canonical list).  Returns per-local-rank info dicts."""
    from . import _lib as L
    W, R = fabric.world, fabric.local_ranks
    if mode not in (L.MODE_UNIQUE_COUNT, L.MODE_SEED_ENUM):
        raise ValueError("find_enum: MODE_UNIQUE_COUNT or MODE_SEED_ENUM")
    if hasattr(fabric, "bind_streams"):
        fabric.bind_streams(ctxs)
    s1 = [c.dist_extract_records(r, W) for c, r in zip(ctxs, R)]
    sc = [cnt for _, _, cnt in s1]
    rc = fabric.counts(sc)
    wide = bool(s1[0][1])
    for which, col in ((0, 0), (2, 1)) if wide else ((0, 0),):
        sends = [dev_words(t[col], sum(t[2]), device) for t in s1]
        recvs = [dev_words(c.dist_recv_buffer(which, sum(k)), sum(k), device) for c, k in zip(ctxs, rc)]
        fabric.words(sends, sc, recvs, rc)
    for c, k in zip(ctxs, rc):
        c.dist_enum_local(sum(k), mode, min_multi=min_multi, max_multi=max_multi, direct_only=direct_only)
    info = [dict(rank=r, seeds_sent=sum(sc[i]), seeds_received=sum(rc[i]), record_bytes=16 if wide else 8) for i, r in enumerate(R)]
    if mode == L.MODE_UNIQUE_COUNT:
        import numpy as np
        pieces = [c.fetch() for c in ctxs]
        nseq = len(pieces[0]["unique_mers_per_seq"])
        host = _is_host(device)
        ts = []
        for p in pieces:
            v = np.concatenate([[p["unique_mers"]], np.asarray(p["unique_mers_per_seq"], dtype=np.int64)]).astype(np.int64)
            ts.append(torch.from_numpy(v) if host else torch.from_numpy(v).to(device))
        fabric.allreduce_sum(ts)
        for i, t in enumerate(ts):
            tot = t.cpu().numpy()
            info[i]["unique_mers"] = int(tot[0])
            info[i]["unique_mers_per_seq"] = [int(x) for x in tot[1:1 + nseq]]
            info[i]["unique_mers_local"] = int(pieces[i]["unique_mers"])
    return info


def merge_enum_results(pieces):
    """The ranks' MODE_SEED_ENUM pieces (each in canonical order, disjoint) -> the canonical list of the single-GPU search:
    a stable merge by first position, which is unique per match."""
    import numpy as np
    cat = concat_results(pieces)
    n = cat["n_matches"]
    if n == 0:
        return cat
    off = np.asarray(cat["comp_off"], dtype=np.int64)
    first = np.abs(np.asarray(cat["comp_start"], dtype=np.int64)[off[:-1]])
    order = np.argsort(first, kind="stable")
    cnt = (off[1:] - off[:-1])[order]
    new_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    idx = np.repeat(off[:-1][order] - new_off[:-1], cnt) + np.arange(int(new_off[-1]), dtype=np.int64)
    out = dict(n_matches=n, n_comps=cat["n_comps"])
    out["length"] = np.asarray(cat["length"])[order]
    out["comp_off"] = new_off.astype(np.uint64)
    out["comp_seq"] = np.asarray(cat["comp_seq"])[idx]
    out["comp_start"] = np.asarray(cat["comp_start"])[idx]
    return out


def find_enum_emulated(seqs, pattern, world, mode, device=0, **kw):
    """All `world` ranks of find_enum inside this process on one GPU (parity tests): MODE_UNIQUE_COUNT -> rank 0's info;
    MODE_SEED_ENUM -> the merged result dict."""
    from . import _lib as L
    from .finder import Context
    dev = torch.device("cuda", device)
    ctxs = [Context(device) for _ in range(world)]
    stream = torch.cuda.Stream(dev)
    try:
        with torch.cuda.stream(stream):
            for c in ctxs:
                c.set_stream(stream.cuda_stream)
                for s in seqs:
                    c.add_sequence(s)
                c.set_seed(pattern)
            info = find_enum(ctxs, LocalFabric(world), dev, mode, **kw)
            stream.synchronize()
            if mode == L.MODE_UNIQUE_COUNT:
                return info[0]
            res = merge_enum_results([c.fetch() for c in ctxs])
        res["info"] = info
        return res
    finally:
        for c in ctxs:
            c.close()


def concat_results(pieces):
    """The ranks' CSR pieces in rank order -> one result dict (numpy, host)."""
    import numpy as np
    out = dict(n_matches=sum(p["n_matches"] for p in pieces), n_comps=sum(p["n_comps"] for p in pieces))
    out["length"] = np.concatenate([p["length"] for p in pieces])
    out["comp_seq"] = np.concatenate([p["comp_seq"] for p in pieces])
    out["comp_start"] = np.concatenate([p["comp_start"] for p in pieces])
    offs, base = [], 0
    for p in pieces:
        offs.append(np.asarray(p["comp_off"][:-1], dtype=np.uint64) + np.uint64(base))
        base += p["n_comps"]
    out["comp_off"] = np.concatenate(offs + [np.array([base], dtype=np.uint64)])
    return out


def find_unique_emulated(seqs, pattern, world, device=0, nway_mask=0, p2p=None):
    """All `world` ranks inside this process on one GPU (parity tests): returns rank 0's result dict."""
    from .finder import Context
    dev = torch.device("cuda", device)
    ctxs = [Context(device) for _ in range(world)]
    stream = torch.cuda.Stream(dev)  # one explicit stream for the library and for the exchange copies
    try:
        with torch.cuda.stream(stream):
            for c in ctxs:
                c.set_stream(stream.cuda_stream)
                for s in seqs:
                    c.add_sequence(s)
                c.set_seed(pattern)
            info = find_unique(ctxs, LocalFabric(world), dev, nway_mask=nway_mask, p2p=p2p)
            stream.synchronize()
            res = concat_results([c.fetch() for c in ctxs])
        res["info"] = info
        return res
    finally:
        for c in ctxs:
            c.close()


# ------------------------------------------------------------------------------------------ bench (N > 1)
def add_sequences_shared(ctx, packer, host_seqs, world, rank, dev):
    """Replicated genomes without N x PCIe: rank r uploads and packs only the genomes g with g % world == r (in `packer`, a
    second context on the same device and stream), the ranks all-gather the packed words over NCCL / NVLink (80 MB at C5),
    and every rank adds all genomes to `ctx` from device memory.  host_seqs: pinned uint8 tensors, the same on every rank."""
    import torch.distributed as dist
    lens = [int(t.numel()) for t in host_seqs]
    nwords = [(n + 31) // 32 for n in lens]
    maxw = max(nwords) if nwords else 0
    per = (len(host_seqs) + world - 1) // world
    send = torch.zeros(max(1, per * maxw), dtype=torch.int64, device=dev)
    packer.clear_sequences()
    for j, g in enumerate(range(rank, len(host_seqs), world)):
        packer.add_sequence_ptr(host_seqs[g].data_ptr(), lens[g])
        packer.copy_packed_device(j, send.data_ptr() + j * maxw * 8)
    recv = torch.empty(world * send.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(recv, send)
    ctx.clear_sequences()
    for g in range(len(host_seqs)):
        owner, j = g % world, g // world
        ctx.add_sequence_device_packed(recv.data_ptr() + (owner * send.numel() + j * maxw) * 8, lens[g])
    return recv  # keep it alive until the stream has passed the copies


def gather_result_digest(res, world, rank, B, config, mode, scale, merge=False):
    """Rank 0 receives every rank's CSR piece over a gloo group (host memory, outside every timed region), joins them in
    rank order (= canonical order) and hashes the whole result exactly like the single-GPU line does: the only proof that
    the NCCL + peer-memory path is bit-exact on real GPUs."""
    import numpy as np
    import torch.distributed as dist
    g = dist.new_group(backend="gloo")
    keys = (("length", np.uint32), ("comp_off", np.uint64), ("comp_seq", np.uint32), ("comp_start", np.int64))
    mine = {k: np.ascontiguousarray(np.asarray(res[k]).astype(dt, copy=False)) for k, dt in keys}
    sizes = torch.tensor([mine[k].size for k, _ in keys], dtype=torch.int64)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=g)
    pieces = [dict(n_matches=int(all_sizes[r][0]), n_comps=int(all_sizes[r][2])) for r in range(world)]
    for j, (k, dt) in enumerate(keys):
        t = torch.from_numpy(mine[k].view(np.uint8))
        if rank == 0:
            bufs = [torch.empty(int(all_sizes[r][j]) * np.dtype(dt).itemsize, dtype=torch.uint8) for r in range(world)]
            bufs[0] = t
            reqs = [dist.irecv(bufs[r], src=r, group=g) for r in range(1, world)]
            for q in reqs:
                q.wait()
            for r in range(world):
                pieces[r][k] = bufs[r].numpy().view(dt)
        else:
            dist.send(t, dst=0, group=g)
    out = None
    if rank == 0:
        out = B.parity_of(merge_enum_results(pieces) if merge else concat_results(pieces), config, mode, scale)
    dist.barrier(group=g)
    return out


def bench_main(args, B):
    """bench.py --gpus N under torchrun: strong scaling of one configuration (C5 by default, the same as N = 1) over N
    ranks.  B = the bench module (config tables, parity check, rooflines, CPU leg)."""
    import torch.distributed as dist
    import mauvealigner_b200 as mb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl", device_id=dev)
    W = max(3, args.warmup)
    config = args.config
    pattern, mode, kw = B.config_params(None, config)
    enum = mode != mb.MODE_UNIQUE  # C3 (MODE_UNIQUE_COUNT, 16-byte records) and C4 (MODE_SEED_ENUM): find_enum

    def parity_check(res, info):
        if mode == mb.MODE_UNIQUE_COUNT:  # the summed count is the result
            return B.parity_of(dict(n_matches=0, n_comps=0, unique_mers=info[0]["unique_mers"]), config, mode, args.scale) if rank == 0 else None
        return gather_result_digest(res, world, rank, B, config, mode, args.scale, merge=enum)
    seqs = B.synth.synth_genomes(config, args.scale)
    bp = sum(len(s) for s in seqs)
    ctx = mb.Context(local)
    # one explicit stream for the library kernels, the NCCL exchanges and the timing events (a NULL handle would make
    # the library create a private stream the exchanges are not ordered against)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_seed(pattern)
    fabric = TorchFabric(device=dev)

    dev_ascii = [torch.from_numpy(s).to(dev) for s in seqs]
    ctx.clear_sequences()
    for t in dev_ascii:
        ctx.add_sequence_device(t.data_ptr(), t.numel())
    del dev_ascii
    torch.cuda.synchronize()

    def step():
        return find_enum([ctx], fabric, dev, mode, **kw) if enum else find_unique([ctx], fabric, dev)

    for _ in range(W):
        step()
    sampler = None
    if rank == 0:  # one sampler for the whole job: all GPUs in use, through NVML
        sampler = B.ClockSampler(list(range(world)))
        sampler.start()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        info = step()
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = torch.tensor([max(ev0.elapsed_time(ev1), 0.0)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    st = ctx.stats()
    stage = ctx.dist_stage_ms()
    launches = torch.tensor([float(st["kernel_launches"])], device=dev)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    res = ctx.fetch(copy=False)
    parity = parity_check(res, info)

    # ---- e2e: pinned host ASCII on every rank -> C ABI stages + exchanges -> every rank's piece of the CSR in host memory
    pinned = [torch.from_numpy(s).pin_memory() for s in seqs]

    packer = mb.Context(local)
    packer.set_stream(stream.cuda_stream)
    keep = []

    def e2e_step():
        # every rank uploads 1 / world of the genomes; the packed words travel over NVLink
        t_ = [time.perf_counter()]

        def lap():
            if e2e_trace:
                torch.cuda.synchronize()
                t_.append(time.perf_counter())

        keep[:] = [add_sequences_shared(ctx, packer, pinned, world, rank, dev)]
        lap()
        e2e_info[:] = step()
        lap()
        out = ctx.fetch(copy=False, compact=True)  # the compact result form (5 B / component over PCIe)
        lap()
        if e2e_trace and rank == 0:
            print("[e2e-trace] upload+share %.2f ms, search %.2f ms, fetch %.2f ms" % tuple(1e3 * (b - a) for a, b in zip(t_, t_[1:])), file=sys.stderr, flush=True)
        return out

    e2e_info = []
    e2e_trace = bool(os.environ.get("MB_E2E_TRACE"))
    for _ in range(W):  # first calls: pinned result buffers, the packer's workspaces, NCCL's set-up for the new message sizes
        e2e_step()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        r = e2e_step()  # ends with the fetch, which waits for the stream
        e2e_steps.append(round((time.perf_counter() - t1) * 1e3, 2))
    torch.cuda.synchronize()
    dist.barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], device=dev)
    dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    st2 = ctx.stats()
    h2d_t = torch.tensor([float(packer.stats()["h2d_bytes"])], device=dev)  # this rank's share of the genomes
    dist.all_reduce(h2d_t, op=dist.ReduceOp.SUM)
    h2d_total = float(h2d_t.item())
    d2h = torch.tensor([float(st2["d2h_bytes"])], device=dev)
    dist.all_reduce(d2h, op=dist.ReduceOp.SUM)
    e2e_parity = parity_check(r, e2e_info)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    if rank == 0:
        peak, peak_src = B.peaks()
        R = st["record_bytes"]
        radix_ms, radix_launches = st["ms_radix_kernels"], st["radix_launches"]
        n_local = info[0]["seeds_received"]
        roofline = None
        if radix_launches:
            avg_ms = radix_ms / radix_launches
            alg = 2.0 * R * n_local
            ach = alg / (avg_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "k_onesweep (one LSD radix pass over this rank's seed records; rank 0)", "achieved": ach,
                        "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": alg, "avg_launch_ms": avg_ms, "launches_per_step": radix_launches}
        P = (2 * mb.seed_weight(pattern) + 7) // 8
        b_alg = 0.25 + R * (3 + 2 * P)
        path = {"b_alg_bytes_per_bp": b_alg, "achieved": b_alg * bp / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s (all GPUs)"}
        path["frac"] = path["achieved"] / (peak * world)
        cpu = None if args.no_cpu_baseline else B.cpu_baseline(config, args.scale)
        cfg = B.config_dict(config, args.scale)
        line = {
            "metric": B.METRIC, "value": bp / (ms_per_step * 1e-3) / 1e9, "unit": "Gbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": cfg,
            "parallelism": (f"key-range x{world} (seeds); result = the sum of the ranks' counts / the ranks' disjoint match lists, merged by first "
                            "position for the digest" if enum else
                            f"key-range x{world} (seeds), group-hash x{world} (de-dup), canonical-range x{world} (output); result = the ranks' CSR "
                            "pieces in rank order (BASELINE.md §3)"),
            "parity": parity, "roofline": roofline, "path_roofline": path, "cpu_baseline": cpu,
            "e2e": {"value": bp / (e2e_ms * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d_total),
                    "d2h_bytes_per_step": int(d2h.item()), "digest_ok": e2e_parity["digest_ok"], "rank0_ms_steps": e2e_steps},
            "gpu_launches": int(launches.item()) * args.steps, "stages_ms_rank0": {k: round(v, 4) for k, v in stage.items()},
            "rank0": {k: v for k, v in info[0].items()}, "wall_ms_per_step": wall_ms / args.steps, "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    fabric.release_peer_buffers([ctx])
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()

"""A stand-in for mauvealigner_b200.Context on the HOST, for the CPU tests of the multi-GPU orchestration
(mauvealigner_b200/dist.py::find_unique).  It has no sequence semantics: every stage emits synthetic, self-describing
words — (kind, source rank, destination rank, index) — and every stage that consumes an exchange checks that exactly the
words its peers addressed to it arrived, in source-rank order.  That pins the count / offset / width book-keeping of the
orchestration (the part that runs on the host) without a GPU; the CUDA stages themselves are covered by the GPU tests."""
import numpy as np


def n_seeds(src, dst):
    return 3 + (src * 7 + dst * 3) % 5


def n_rows(src, dst):
    return (src + 2 * dst) % 4  # some pairs exchange nothing


def n_match(src, dst):
    return 1 + (3 * src + dst) % 3


def n_mcomp(src, dst):
    return 2 * n_match(src, dst) + (src + dst) % 2


def word(kind, src, dst, i):
    return (kind << 56) | (src << 40) | (dst << 24) | i


def _store(base, byte_offset, arr):
    import ctypes
    if arr.size:
        ctypes.memmove(int(base) + int(byte_offset), arr.ctypes.data, arr.nbytes)


class FakeCtx:
    def __init__(self, log=None):
        self.bufs = {}      # receive buffers by `which`
        self.keep = []      # send buffers stay alive until the next run
        self.rank = self.world = None
        self.done = False
        self.log = log if log is not None else []

    # ---- helpers
    def _send(self, arr):
        self.keep.append(arr)
        return arr.ctypes.data

    def _expect(self, which, kind, count_fn, width=1, what=""):
        buf = self.bufs[which]
        exp = []
        for s in range(self.world):
            for i in range(count_fn(s, self.rank)):
                exp += [word(kind, s, self.rank, i * width + k) for k in range(width)]
        got = buf[:len(exp)].tolist()
        assert got == exp, (what, self.rank, got[:6], exp[:6])

    # ---- the stage interface find_unique drives (NCCL-style exchanges, p2p = 0)
    def dist_extract(self, rank, world):
        self.rank, self.world, self.keep, self.done = rank, world, [], False
        counts = [n_seeds(rank, d) for d in range(world)]
        arr = np.array([word(1, rank, d, i) for d in range(world) for i in range(counts[d])], dtype=np.int64)
        return self._send(arr), counts

    # ---- find_enum (MODE_UNIQUE_COUNT / MODE_SEED_ENUM): one exchange of 8-byte or 16-byte records, then a local tail
    wide = False

    def dist_extract_records(self, rank, world):
        kp, counts = self.dist_extract(rank, world)
        vp = 0
        if self.wide:
            vp = self._send(np.array([word(6, rank, d, i) for d in range(world) for i in range(counts[d])], dtype=np.int64))
        return kp, vp, counts

    def dist_enum_local(self, n_recv, mode, min_multi=2, max_multi=1000, direct_only=False):
        assert n_recv == sum(n_seeds(s, self.rank) for s in range(self.world))
        self._expect(0, 1, n_seeds, what="seed records")
        if self.wide:
            self._expect(2, 6, n_seeds, what="second words of the seed records")
        self.mode, self.done = mode, True

    def fetch(self):
        # this rank's share of the counts: 10 r + 1 distinct seeds, (r + 1, 2 r) per sequence
        return dict(unique_mers=10 * self.rank + 1, unique_mers_per_seq=np.array([self.rank + 1, 2 * self.rank], dtype=np.uint64))

    def dist_recv_buffer(self, which, n):
        self.bufs[which] = np.full(int(n) + 4, 255, dtype=np.uint8) if which == 5 else np.full(int(n) + 4, -1, dtype=np.int64)
        return self.bufs[which].ctypes.data

    def dist_use_p2p_recv(self, on):
        if on:
            self.bufs[0] = self.p2p  # stage 2 reads the array the peers stored into

    def dist_local(self, world, n_recv, nway_mask=0):
        assert n_recv == sum(n_seeds(s, self.rank) for s in range(world))
        self._expect(0, 1, n_seeds, what="seed records")
        self.log.append("local")
        return [n_rows(self.rank, d) for d in range(world)]

    def dist_rows_pack(self, peer_ptrs=None, peer_row_offsets=None):
        blocks = [np.array([word(2, self.rank, d, 4 * i + k) for i in range(n_rows(self.rank, d)) for k in range(4)], dtype=np.int64)
                  for d in range(self.world)]
        if peer_ptrs is not None:  # fused exchange 2: rows straight into the owners' buffers
            for d, b in enumerate(blocks):
                _store(peer_ptrs[d], 32 * peer_row_offsets[d], b)
            return None
        arr = np.concatenate(blocks) if sum(b.size for b in blocks) else np.zeros(1, dtype=np.int64)
        return self._send(arr)

    # ---- peer-memory variants (p2p levels 1..3; in-process pointers)
    def dist_extract_count(self, rank, world):
        self.rank, self.world, self.keep, self.done = rank, world, [], False
        return [n_seeds(rank, d) for d in range(world)]

    def dist_p2p_recv_array(self, capacity):
        if getattr(self, "p2p", None) is None or self.p2p.size < capacity + 4:
            self.p2p = np.full(int(capacity) + 4, -1, dtype=np.int64)
        else:
            self.p2p[:] = -1
        return self.p2p.ctypes.data

    def dist_partition_p2p(self, peer_ptrs, peer_offsets):
        for d in range(self.world):
            _store(peer_ptrs[d], 8 * peer_offsets[d], np.array([word(1, self.rank, d, i) for i in range(n_seeds(self.rank, d))], dtype=np.int64))

    def dist_push(self, src_ptr, counts, unit_words, peer_ptrs, dst_offsets):
        import ctypes
        off = 0
        for d in range(self.world):
            nbytes = counts[d] * unit_words * 8
            if nbytes:
                ctypes.memmove(int(peer_ptrs[d]) + dst_offsets[d] * unit_words * 8, int(src_ptr) + off, nbytes)
            off += nbytes

    def dist_resolve(self, n):
        assert n == sum(n_rows(s, self.rank) for s in range(self.world))
        self._expect(1, 2, n_rows, width=4, what="candidate rows")
        # verdict of the i-th row that came from source s
        v = np.array([(s + i + self.rank) % 2 for s in range(self.world) for i in range(n_rows(s, self.rank))], dtype=np.uint8)
        return self._send(v if v.size else np.zeros(1, dtype=np.uint8))

    def dist_accept(self):
        exp = [(self.rank + i + d) % 2 for d in range(self.world) for i in range(n_rows(self.rank, d))]  # by owner, row order
        assert self.bufs[5][:len(exp)].tolist() == exp, ("verdicts", self.rank)
        self.hist = np.zeros(4096, dtype=np.int64)
        self.hist[:4] = self.rank + 1
        return self.hist.ctypes.data

    def dist_match_partition(self, world):
        assert self.hist[:4].tolist() == [sum(r + 1 for r in range(world))] * 4 and not self.hist[4:].any(), "histogram all-reduce"
        return [n_match(self.rank, d) for d in range(world)], [n_mcomp(self.rank, d) for d in range(world)]

    def dist_match_pack(self, hdr_ptrs=None, hdr_offsets=None, comp_ptrs=None, comp_offsets=None):
        hb = [np.array([word(3, self.rank, d, 2 * i + k) for i in range(n_match(self.rank, d)) for k in range(2)], dtype=np.int64)
              for d in range(self.world)]
        cb = [np.array([word(4, self.rank, d, i) for i in range(n_mcomp(self.rank, d))], dtype=np.int64) for d in range(self.world)]
        if hdr_ptrs is not None:  # fused exchange 3
            for d in range(self.world):
                _store(hdr_ptrs[d], 16 * hdr_offsets[d], hb[d])
                _store(comp_ptrs[d], 8 * comp_offsets[d], cb[d])
            return None
        return self._send(np.concatenate(hb)), self._send(np.concatenate(cb))

    def dist_output(self, n_m, n_c):
        assert n_m == sum(n_match(s, self.rank) for s in range(self.world)) and n_c == sum(n_mcomp(s, self.rank) for s in range(self.world))
        self._expect(3, 3, n_match, width=2, what="match headers")
        self._expect(4, 4, n_mcomp, what="match components")
        self.done = True

"""Host-side mirror of the reference's match-finder protocol over the C ABI.

Names and argument meaning follow the reference's plugin interface for this path:
  MatchList                      (libMems; members used at src/mauveAligner.cpp:450-466)
  SeedMatchEnumerator.FindMatches(match_list, min_multi=2, max_multi=1000, direct_repeats_only=False)
                                  /root/reference/src/SeedMatchEnumerator.h:19-33
  UniqueMatchFinder.FindMatches(match_list)      /root/reference/src/UniqueMatchFinder.h:21-32,
                                                 call site src/progressiveMauve.cpp:490-495
  MaskedMemHash.SetMask(mask)                    /root/reference/src/mauveAligner.cpp:523-531
  SortedMerList.UniqueMerCount()                 /root/reference/src/uniqueMerCount.cpp:39
The C++ twin of this file is include/mems_compat/mems_compat.h (same names, same calls into libmauve_b200.so).
All compute happens in the CUDA library; nothing here touches sequence data."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .seeds import get_seed, default_seed_weight, seed_length, seed_weight, seed_valid, SOLID_SEED, CODING_SEED  # noqa: F401

NO_MATCH = 0


def _check(ctx, rc):
    if rc != L.MB_OK:
        detail = L.lib().mb_last_cuda_error(ctx).decode() if ctx else ""
        raise L.MauveError(rc, detail)


class Context:
    """One CUDA context of the library (one per GPU / process)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = L.lib().mb_ctx_create(C.byref(self._h), device)
        if rc != L.MB_OK:
            raise L.MauveError(rc, "mb_ctx_create: a B200 (sm_100) device is required; there is no CPU path")
        self.device = device

    def close(self):
        if self._h:
            L.lib().mb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        _check(self._h, L.lib().mb_set_stream(self._h, C.c_void_p(cuda_stream)))

    def clear_sequences(self):
        _check(self._h, L.lib().mb_clear_sequences(self._h))

    def accumulate(self, on=True):
        """MemHash table across searches (mb_accumulate): on = later MODE_UNIQUE searches drop what earlier matches
        contain and add their own matches; off = forget the table (MemHash::Clear())."""
        _check(self._h, L.lib().mb_accumulate(self._h, 1 if on else 0))

    def add_sequence(self, data, packed=False):
        """data: bytes / str / numpy uint8 (ASCII) or numpy uint64 words (packed=True, with length attr)."""
        if packed:
            words, length = data
            words = np.ascontiguousarray(words, dtype=np.uint64)
            ptr, n = words.ctypes.data, int(length)
            self._keep = words
        else:
            if isinstance(data, str):
                data = data.encode()
            if isinstance(data, (bytes, bytearray)):
                data = np.frombuffer(bytes(data), dtype=np.uint8)
            data = np.ascontiguousarray(data, dtype=np.uint8)
            ptr, n = data.ctypes.data, int(data.size)
            self._keep = data
        sid = C.c_int(-1)
        _check(self._h, L.lib().mb_add_sequence(self._h, C.c_void_p(ptr), n, 1 if packed else 0, C.byref(sid)))
        return sid.value

    def add_sequence_ptr(self, host_ptr, length, packed=False):
        sid = C.c_int(-1)
        _check(self._h, L.lib().mb_add_sequence(self._h, C.c_void_p(host_ptr), int(length), 1 if packed else 0, C.byref(sid)))
        return sid.value

    def add_sequence_device(self, dev_ptr, length):
        sid = C.c_int(-1)
        _check(self._h, L.lib().mb_add_sequence_device(self._h, C.c_void_p(dev_ptr), int(length), C.byref(sid)))
        return sid.value

    def add_sequence_device_packed(self, dev_ptr, length):
        """A sequence given as 2-bit words already in device memory (mb_add_sequence_device_packed); length in bases."""
        _check(self._h, L.lib().mb_add_sequence_device_packed(self._h, C.c_void_p(dev_ptr), length, None))

    def copy_packed_device(self, seq, dst_dev_ptr):
        """copy the packed words of sequence `seq` to device memory at dst_dev_ptr, on the context stream"""
        _check(self._h, L.lib().mb_copy_packed_device(self._h, seq, C.c_void_p(dst_dev_ptr)))

    def packed_device(self, seq):
        """(device pointer, number of uint64 words) of the packed form of sequence `seq` (mb_get_packed_device)."""
        p, n = C.c_void_p(), C.c_uint64(0)
        _check(self._h, L.lib().mb_get_packed_device(self._h, seq, C.byref(p), C.byref(n)))
        return p.value, n.value

    def set_seed(self, pattern):
        _check(self._h, L.lib().mb_set_seed(self._h, int(pattern)))

    def _params(self, mode, min_multi, max_multi, direct_only, nway_mask):
        return L.MbParams(int(mode), int(bool(direct_only)), int(min_multi), int(max_multi), int(nway_mask))

    def find_device(self, mode, min_multi=2, max_multi=1000, direct_only=False, nway_mask=0):
        p = self._params(mode, min_multi, max_multi, direct_only, nway_mask)
        _check(self._h, L.lib().mb_find_device(self._h, C.byref(p)))

    def fetch(self, copy=True, compact=False):
        """compact: the layout the device keeps (comp_seq uint8, comp_start int32: 5 instead of 12 bytes per component
        over PCIe); same values either way."""
        if compact:
            out = C.POINTER(L.MbResultCompact)()
            _check(self._h, L.lib().mb_fetch_result_compact(self._h, C.byref(out)))
        else:
            out = C.POINTER(L.MbResult)()
            _check(self._h, L.lib().mb_fetch_result(self._h, C.byref(out)))
        return self._wrap(out.contents, copy)

    def find(self, mode, min_multi=2, max_multi=1000, direct_only=False, nway_mask=0, copy=True, compact=False):
        p = self._params(mode, min_multi, max_multi, direct_only, nway_mask)
        if compact:
            out = C.POINTER(L.MbResultCompact)()
            _check(self._h, L.lib().mb_find_compact(self._h, C.byref(p), C.byref(out)))
        else:
            out = C.POINTER(L.MbResult)()
            _check(self._h, L.lib().mb_find(self._h, C.byref(p), C.byref(out)))
        return self._wrap(out.contents, copy)

    def position_table(self):
        """mb_position_table: repeatoire's match position lookup table of the last single-sequence result
        (src/repeatoire.cpp:1944-1966): (match_of_pos, comp_of_pos), entry p = 1-based left end, 0xFFFFFFFF = none."""
        pm, pc, n = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.c_uint64(0)
        _check(self._h, L.lib().mb_position_table(self._h, C.byref(pm), C.byref(pc), C.byref(n)))
        return (np.ctypeslib.as_array(pm, shape=(n.value,)).copy(), np.ctypeslib.as_array(pc, shape=(n.value,)).copy())

    def find_batch(self, problems, mode=L.MODE_UNIQUE, min_multi=2, max_multi=1000, direct_only=False, nway_mask=0):
        """mb_find_batch: many small independent searches (recursive anchoring, src/mauveAligner.cpp:94,698) in ONE pass of
        the pipeline.  problems = list of lists of sequences (all lists of one length; an empty sequence = absent).
        Returns one result dict per problem (n_matches, length, comp_off, comp_seq, comp_start), each exactly what find()
        returns for that problem alone."""
        nprob = len(problems)
        nseq = len(problems[0]) if nprob else 1
        keep, ptrs, lens = [], (C.c_void_p * max(1, nprob * nseq))(), (C.c_uint64 * max(1, nprob * nseq))()
        for i, prob in enumerate(problems):
            if len(prob) != nseq:
                raise ValueError("every problem needs the same number of sequences")
            for g, s in enumerate(prob):
                a = np.frombuffer(s.encode() if isinstance(s, str) else bytes(s), dtype=np.uint8) if not isinstance(s, np.ndarray) else np.ascontiguousarray(s, dtype=np.uint8)
                keep.append(a)
                ptrs[i * nseq + g] = a.ctypes.data if a.size else None
                lens[i * nseq + g] = a.size
        out = C.POINTER(L.MbBatchResult)()
        p = self._params(mode, min_multi, max_multi, direct_only, nway_mask)
        _check(self._h, L.lib().mb_find_batch(self._h, C.byref(p), nprob, nseq, ptrs, lens, C.byref(out)))
        r = out.contents

        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dtype=dt)

        moff = arr(r.match_off, nprob + 1, np.uint64)
        length, coff = arr(r.length, r.n_matches, np.uint32), arr(r.comp_off, r.n_matches + 1, np.uint64)
        cseq, cstart = arr(r.comp_seq, r.n_comps, np.uint32), arr(r.comp_start, r.n_comps, np.int64)
        res = []
        for i in range(nprob):
            a, b = int(moff[i]), int(moff[i + 1])
            ca, cb = (int(coff[a]), int(coff[b])) if r.n_matches else (0, 0)
            res.append(dict(n_matches=b - a, n_comps=cb - ca, length=length[a:b], comp_off=coff[a:b + 1] - np.uint64(ca) if b > a else np.zeros(1, dtype=np.uint64),
                            comp_seq=cseq[ca:cb], comp_start=cstart[ca:cb]))
        return res

    @staticmethod
    def _wrap(r, copy):
        nm, nc, ns = int(r.n_matches), int(r.n_comps), int(r.nseq)

        def arr(ptr, n):
            if n == 0:
                return np.zeros(0, dtype=np.ctypeslib.as_array(ptr, shape=(1,)).dtype)
            a = np.ctypeslib.as_array(ptr, shape=(n,))
            return a.copy() if copy else a

        return dict(n_matches=nm, n_comps=nc, length=arr(r.length, nm), comp_off=arr(r.comp_off, nm + 1), comp_seq=arr(r.comp_seq, nc),
                    comp_start=arr(r.comp_start, nc), unique_mers=int(r.unique_mers),
                    unique_mers_per_seq=arr(r.unique_mers_per_seq, ns))

    # ---- multi-GPU stages (include/mauve_b200.h "multi-GPU path"; driven by mauvealigner_b200/dist.py)
    def dist_extract(self, rank, world):
        """-> (device pointer of the partitioned seed records, records per destination rank)"""
        ptr = C.c_void_p()
        counts = (C.c_uint64 * world)()
        _check(self._h, L.lib().mb_dist_extract(self._h, int(rank), int(world), C.byref(ptr), counts))
        return ptr.value or 0, [int(x) for x in counts]

    def dist_extract_records(self, rank, world):
        """stage 1 for either record format -> (pointer of the first words, pointer of the second words of 16-byte records or
        0, records per destination rank)"""
        kp, vp = C.c_void_p(), C.c_void_p()
        counts = (C.c_uint64 * world)()
        _check(self._h, L.lib().mb_dist_extract_records(self._h, int(rank), int(world), C.byref(kp), C.byref(vp), counts))
        return kp.value or 0, vp.value or 0, [int(x) for x in counts]

    def dist_enum_local(self, n_recv, mode, min_multi=2, max_multi=1000, direct_only=False):
        """stage 2 of MODE_UNIQUE_COUNT / MODE_SEED_ENUM over the received key range; fetch() afterwards"""
        p = self._params(mode, min_multi, max_multi, direct_only, 0)
        _check(self._h, L.lib().mb_dist_enum_local(self._h, C.byref(p), int(n_recv)))

    def dist_extract_count(self, rank, world):
        """stage 1a -> records of this rank's slice per destination rank (the slice stays in library memory)"""
        counts = (C.c_uint64 * world)()
        _check(self._h, L.lib().mb_dist_extract_count(self._h, int(rank), int(world), counts))
        return [int(x) for x in counts]

    def dist_partition_p2p(self, peer_ptrs, peer_offsets):
        """stage 1b fused with exchange #1: the partition pass stores every record into its destination rank's
        receive array (peer_ptrs[d], at record index peer_offsets[d] + ...) over NVLink; asynchronous"""
        n = len(peer_ptrs)
        pp = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        po = (C.c_uint64 * n)(*[int(o) for o in peer_offsets])
        _check(self._h, L.lib().mb_dist_partition(self._h, pp, po, None))

    def dist_p2p_recv_array(self, capacity_records):
        ptr = C.c_void_p()
        _check(self._h, L.lib().mb_dist_p2p_recv_array(self._h, int(capacity_records), C.byref(ptr)))
        return ptr.value or 0

    def dist_use_p2p_recv(self, on):
        _check(self._h, L.lib().mb_dist_use_p2p_recv(self._h, 1 if on else 0))

    def ipc_export(self, dev_ptr):
        buf = C.create_string_buffer(64)
        _check(self._h, L.lib().mb_ipc_export(self._h, C.c_void_p(int(dev_ptr)), buf))
        return buf.raw

    def ipc_import(self, handle):
        ptr = C.c_void_p()
        _check(self._h, L.lib().mb_ipc_import(self._h, bytes(handle), C.byref(ptr)))
        return ptr.value or 0

    def ipc_close(self, dev_ptr):
        _check(self._h, L.lib().mb_ipc_close(self._h, C.c_void_p(int(dev_ptr))))

    def dist_recv_buffer(self, which, n_words):
        ptr = C.c_void_p()
        _check(self._h, L.lib().mb_dist_recv_buffer(self._h, int(which), int(n_words), C.byref(ptr)))
        return ptr.value or 0

    def dist_local(self, world, n_recv, mode=L.MODE_UNIQUE, nway_mask=0):
        """stage 2 -> candidate rows per owner rank (the rows are written by dist_rows_pack)"""
        p = self._params(mode, 2, 1000, False, nway_mask)
        cc = (C.c_uint64 * world)()
        _check(self._h, L.lib().mb_dist_local(self._h, C.byref(p), int(n_recv), cc))
        return [int(x) for x in cc]

    @staticmethod
    def _ptr_array(ptrs):
        return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])

    @staticmethod
    def _u64_array(vals):
        return (C.c_uint64 * len(vals))(*[int(v) for v in vals])

    def dist_rows_pack(self, peer_ptrs=None, peer_row_offsets=None):
        """-> device pointer of the 4-word rows in owner order (local send buffer), or None after storing them straight
        into the owners' receive buffers (peer_ptrs[d], at row peer_row_offsets[d]) over NVLink"""
        if peer_ptrs is None:
            rows = C.c_void_p()
            _check(self._h, L.lib().mb_dist_rows_pack(self._h, None, None, C.byref(rows)))
            return rows.value or 0
        _check(self._h, L.lib().mb_dist_rows_pack(self._h, self._ptr_array(peer_ptrs), self._u64_array(peer_row_offsets), None))
        return None

    def dist_push(self, src_ptr, counts, unit_words, peer_ptrs, dst_offsets):
        """copy-engine push of a local send buffer's destination blocks into the peers' mapped receive buffers"""
        _check(self._h, L.lib().mb_dist_push(self._h, C.c_void_p(int(src_ptr)), self._u64_array(counts), int(unit_words),
                                             self._ptr_array(peer_ptrs), self._u64_array(dst_offsets)))

    def dist_resolve(self, n_rows):
        """stage 3a (owner) -> device pointer of one verdict byte per received row (1 accepted)"""
        v = C.c_void_p()
        _check(self._h, L.lib().mb_dist_resolve(self._h, int(n_rows), C.byref(v)))
        return v.value or 0

    def dist_accept(self):
        """stage 3b (source) -> device pointer of the 4096-bin histogram (uint64) of this rank's matches' canonical keys"""
        hist = C.c_void_p()
        _check(self._h, L.lib().mb_dist_accept(self._h, C.byref(hist)))
        return hist.value or 0

    def dist_match_partition(self, world):
        """after the histogram has been summed over the ranks in place -> (match rows per destination, component words
        per destination); the rows are written by dist_match_pack"""
        cc, mc = (C.c_uint64 * world)(), (C.c_uint64 * world)()
        _check(self._h, L.lib().mb_dist_match_partition(self._h, cc, mc))
        return [int(x) for x in cc], [int(x) for x in mc]

    def dist_match_pack(self, hdr_ptrs=None, hdr_offsets=None, comp_ptrs=None, comp_offsets=None):
        """-> (headers pointer, components pointer) of the local send buffers, or None after storing the rows straight
        into the destinations' receive buffers over NVLink"""
        if hdr_ptrs is None:
            hdr, comps = C.c_void_p(), C.c_void_p()
            _check(self._h, L.lib().mb_dist_match_pack(self._h, None, None, None, None, C.byref(hdr), C.byref(comps)))
            return hdr.value or 0, comps.value or 0
        _check(self._h, L.lib().mb_dist_match_pack(self._h, self._ptr_array(hdr_ptrs), self._u64_array(hdr_offsets), self._ptr_array(comp_ptrs),
                                                   self._u64_array(comp_offsets), None, None))
        return None

    def dist_output(self, n_match, n_comp):
        _check(self._h, L.lib().mb_dist_output(self._h, int(n_match), int(n_comp)))

    def dist_stage_ms(self):
        out = (C.c_float * 8)()
        _check(self._h, L.lib().mb_dist_stage_ms(self._h, out))
        return dict(extract_partition=out[0], sort=out[1], buckets=out[2], extend=out[3], resolve=out[4])

    def stats(self):
        s = L.MbStats()
        _check(self._h, L.lib().mb_get_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in s._fields_}

    def mers(self, seq, length_hint):
        out = np.zeros(max(1, length_hint), dtype=np.uint64)
        n = C.c_uint64(0)
        _check(self._h, L.lib().mb_get_mers(self._h, seq, C.c_void_p(out.ctypes.data), out.size, C.byref(n)))
        return out[: n.value]

    def sml(self, seq, length_hint):
        out = np.zeros(max(1, length_hint), dtype=np.uint32)
        n = C.c_uint64(0)
        _check(self._h, L.lib().mb_get_sml(self._h, seq, C.c_void_p(out.ctypes.data), out.size, C.byref(n)))
        return out[: n.value]


class Match:
    """mems::Match as the callers use it (src/SeedMatchEnumerator.h:75-119, src/repeatoire.cpp:1926-1936)."""

    __slots__ = ("_start", "_length")

    def __init__(self, seq_count):
        self._start = [NO_MATCH] * seq_count
        self._length = 0

    def SetLength(self, n):
        self._length = int(n)

    def SetStart(self, i, s):
        self._start[i] = int(s)

    def __getitem__(self, i):
        return self._start[i]

    def Start(self, i):
        return self._start[i]

    def Length(self, i=0):
        return self._length

    def SeqCount(self):
        return len(self._start)

    def Multiplicity(self):
        return sum(1 for s in self._start if s != NO_MATCH)

    def Orientation(self, i):
        return 0 if self._start[i] > 0 else (1 if self._start[i] < 0 else 2)  # forward, reverse, undefined

    def LeftEnd(self, i):
        return abs(self._start[i])

    def RightEnd(self, i):
        return NO_MATCH if self._start[i] == NO_MATCH else abs(self._start[i]) + self._length - 1

    def CropStart(self, n):
        """Match::CropStart: drop n columns at the match start (forward components move, reverse ones keep their left end)."""
        self._start = [s + n if s > 0 else s for s in self._start]
        self._length -= n

    def CropEnd(self, n):
        """Match::CropEnd: drop n columns at the match end (the LEFT end of a reverse component is the match end)."""
        self._start = [s - n if s < 0 else s for s in self._start]
        self._length -= n

    def Copy(self):
        m = Match(len(self._start))
        m._start = list(self._start)
        m._length = self._length
        return m

    def __repr__(self):
        return f"{self._length}\t" + "\t".join(str(s) for s in self._start)


class MatchList(list):
    """mems::MatchList: a vector<Match*> plus seq_filename / sml_filename / seq_table / sml_table."""

    def __init__(self):
        super().__init__()
        self.seq_filename, self.sml_filename, self.seq_table, self.sml_table = [], [], [], []
        self.seed_pattern = 0

    def MultiplicityFilter(self, mult):
        """MatchList::MultiplicityFilter (src/mauveAligner.cpp:600): keep the matches present in exactly `mult` sequences."""
        self[:] = [m for m in self if m.Multiplicity() == mult]

    def CreateMemorySMLs(self, seed_weight, log=None, seed_rank=0):
        """MatchList::CreateMemorySMLs(mer_size, ostream*, seed_rank) (src/mauveAligner.cpp:456): here it only fixes
        the seed; the sorted mer lists are built on the device inside FindMatches."""
        if seed_weight == 0:
            avg = sum(len(s) for s in self.seq_table) // max(1, len(self.seq_table))
            seed_weight = default_seed_weight(avg)
        self.seed_pattern = get_seed(seed_weight, seed_rank)
        self.sml_table = [SortedMerList(self, i) for i in range(len(self.seq_table))]

    LoadSMLs = CreateMemorySMLs


def EliminateOverlaps(ml):
    """EliminateOverlaps(MatchList&) (src/mauveAligner.cpp:594-596,611-612): afterwards no two matches overlap in any
    sequence.  Rule (libMems' body is not in the tree): sequence by sequence, the matches present in it
    are swept by (left end, position in the list); what earlier matches of the sweep already cover — a prefix in that
    sequence's coordinates — is cropped off, a match covered entirely is dropped."""
    nseq = ml[0].SeqCount() if len(ml) else 0
    items = list(ml)
    for g in range(nseq):
        order = sorted((i for i, m in enumerate(items) if m is not None and m.Start(g) != NO_MATCH), key=lambda i: items[i].LeftEnd(g))
        covered = 0
        for i in order:
            m = items[i]
            l, r = m.LeftEnd(g), m.RightEnd(g)
            if r <= covered:
                items[i] = None
                continue
            if l <= covered:
                ov = covered - l + 1
                if m.Orientation(g) == 0:
                    m.CropStart(ov)
                else:
                    m.CropEnd(ov)
            covered = r
    ml[:] = [m for m in items if m is not None]


def transposeMatches(ml, seqI, seq_regions):
    """transposeMatches(MatchList&, seqI, seq_regions) (src/mauveAligner.cpp:628-637, src/transposeCoordinates.cpp:46-65):
    sequence seqI was searched in FILTERED form — the concatenation of its used regions (first_0, last_0, first_1, ...;
    1-based, inclusive) — so the matches carry filtered coordinates; back to the original ones, splitting a match that
    runs across a region boundary."""
    import bisect
    if len(seq_regions) < 2:
        return
    nreg = len(seq_regions) // 2
    cum = [0]
    for k in range(nreg):
        cum.append(cum[-1] + seq_regions[2 * k + 1] - seq_regions[2 * k] + 1)
    out = []
    for m in ml:
        if m.Start(seqI) == NO_MATCH:
            out.append(m)
            continue
        while m is not None:
            l = m.LeftEnd(seqI)
            k = min(bisect.bisect_right(cum, l - 1) - 1, nreg - 1)
            room = cum[k + 1] - (l - 1)
            rest = None
            if k + 1 < nreg and m.Length() > room:
                rest = m.Copy()
                tail = m.Length() - room
                if m.Orientation(seqI) == 0:
                    m.CropEnd(tail)
                    rest.CropStart(room)
                else:
                    m.CropStart(tail)
                    rest.CropEnd(room)
            nl = seq_regions[2 * k] + (m.LeftEnd(seqI) - 1 - cum[k])
            m.SetStart(seqI, -nl if m.Start(seqI) < 0 else nl)
            out.append(m)
            m = rest
    ml[:] = out


class SortedMerList:
    """Façade over the device-built sorted mer list of one sequence (libMems SortedMerList)."""

    def __init__(self, ml, index):
        self._ml, self._i = ml, index

    def Seed(self):
        return self._ml.seed_pattern

    def SeedLength(self):
        return seed_length(self._ml.seed_pattern)

    def SeedWeight(self):
        return seed_weight(self._ml.seed_pattern)

    def Length(self):
        return len(self._ml.seq_table[self._i])

    def UniqueMerCount(self, ctx=None):
        own = ctx is None
        ctx = ctx or Context()
        try:
            ctx.clear_sequences()
            ctx.add_sequence(self._ml.seq_table[self._i])
            ctx.set_seed(self._ml.seed_pattern)
            return ctx.find(L.MODE_UNIQUE_COUNT)["unique_mers"]
        finally:
            if own:
                ctx.close()


class MatchFinder:
    def __init__(self, ctx=None):
        self._ctx = ctx
        self._own = ctx is None
        self.seq_count = 0
        self._log = None
        self.last = None

    def _context(self):
        if self._ctx is None:
            self._ctx = Context()
        return self._ctx

    def LogProgress(self, stream):
        self._log = stream

    def Clear(self):
        self.last = None

    def ClearSequences(self):
        self.seq_count = 0
        if self._ctx is not None:
            self._ctx.clear_sequences()

    def _load(self, match_list):
        ctx = self._context()
        ctx.clear_sequences()
        for s in match_list.seq_table:
            ctx.add_sequence(s)
        self.seq_count = len(match_list.seq_table)
        if not seed_valid(match_list.seed_pattern):
            raise L.MauveError(-2)
        ctx.set_seed(match_list.seed_pattern)
        return ctx

    @staticmethod
    def _fill(match_list, res, seq_count, dense):
        del match_list[:]
        off, seqs, starts, lens = res["comp_off"], res["comp_seq"], res["comp_start"], res["length"]
        for i in range(res["n_matches"]):
            a, b = int(off[i]), int(off[i + 1])
            m = Match(seq_count if dense else b - a)
            m.SetLength(int(lens[i]))
            for k in range(a, b):
                m.SetStart(int(seqs[k]) if dense else k - a, int(starts[k]))
            match_list.append(m)


def _canonical_key(m):
    """D18 canonical order of MODE_UNIQUE matches: |start| per sequence, then the sign vector, then the length."""
    return (tuple(abs(x) for x in m._start), tuple(1 if x < 0 else 0 for x in m._start), m._length)


class UniqueMatchFinder(MatchFinder):
    """UniqueMatchFinder / MemHash: multi-MUMs with unique seeds (src/UniqueMatchFinder.cpp:36-60).  As in libMems the
    table persists across FindMatches calls until Clear(): a later call (the seed-family search of
    src/progressiveMauve.cpp:503-548 makes three, with ClearSequences() in between) drops what the matches found so far
    contain, and GetMatchList returns the union in canonical order."""

    def __init__(self, ctx=None):
        super().__init__(ctx)
        self._mask = 0
        self._found = []

    def FindMatches(self, match_list):
        ctx = self._load(match_list)
        ctx.accumulate(True)
        self.last = ctx.find(L.MODE_UNIQUE, nway_mask=self._mask)
        new = MatchList()
        self._fill(new, self.last, self.seq_count, dense=True)
        first = not self._found
        self._found.extend(new)
        if not first:
            self._found.sort(key=_canonical_key)
        self.GetMatchList(match_list)
        return True

    def GetMatchList(self, match_list):
        del match_list[:]
        match_list.extend(m.Copy() for m in self._found)

    def Clear(self):
        super().Clear()
        self._found = []
        if self._ctx is not None:
            self._ctx.accumulate(False)

    def Clone(self):
        c = type(self)(self._ctx)
        c._mask = self._mask
        return c


MemHash = UniqueMatchFinder


class PairwiseMatchFinder(UniqueMatchFinder):
    """libMems PairwiseMatchFinder (src/progressiveMauve.cpp:496-501: used when <= 4 genomes are aligned): the unique
    filter of MemHash, then one two-genome match per PAIR of the bucket's unique genomes."""

    def FindMatches(self, match_list):
        ctx = self._load(match_list)
        self.last = ctx.find(L.MODE_PAIRWISE)
        self._found = []
        self._fill(self._found, self.last, self.seq_count, dense=True)
        self.GetMatchList(match_list)
        return True


class MaskedMemHash(UniqueMatchFinder):
    def SetMask(self, mask):
        self._mask = int(mask)


class SeedMatchEnumerator(MatchFinder):
    """Every seed match becomes a full match without extension (src/SeedMatchEnumerator.h:11-141)."""

    def FindMatches(self, match_list, min_multi=2, max_multi=1000, direct_repeats_only=False):
        del match_list[:]
        if len(match_list.seq_table) != 1:
            # CreateMatches() is a no-op unless seq_count == 1 (:59-65); the list stays empty
            return
        ctx = self._load(match_list)
        self.last = ctx.find(L.MODE_SEED_ENUM, min_multi=min_multi, max_multi=max_multi, direct_only=direct_repeats_only)
        self._fill(match_list, self.last, 1, dense=False)

    def Clone(self):
        return SeedMatchEnumerator(self._ctx)


class RepeatHash(MatchFinder):
    """libMems RepeatHash (mauveAligner --repeats, src/mauveAligner.cpp:480-487): repeats inside ONE sequence — every
    bucket of min_multi..max_multi occurrences becomes one match with a column per occurrence, extended and
    de-duplicated like a MemHash entry."""

    def FindMatches(self, match_list, min_multi=2, max_multi=255):
        del match_list[:]
        if len(match_list.seq_table) != 1:
            return False
        ctx = self._load(match_list)
        self.last = ctx.find(L.MODE_REPEAT, min_multi=min_multi, max_multi=max_multi)
        self._fill(match_list, self.last, 1, dense=False)
        return True

    def Clone(self):
        return RepeatHash(self._ctx)


def find_multi(ctxs, nway_mask=0):
    """mb_find_multi: MODE_UNIQUE over one context per GPU (or several per GPU), driven by the library's own host threads
    with every exchange device to device over NVLink peer access — no torch.distributed, no NCCL.  The same sequences
    and seed must be set on every context.  -> the whole result (the ranks' pieces concatenated in rank order)."""
    from .dist import concat_results
    p = ctxs[0]._params(L.MODE_UNIQUE, 2, 1000, False, nway_mask)
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    _check(ctxs[0]._h, L.lib().mb_find_multi(arr, len(ctxs), C.byref(p)))
    return concat_results([c.fetch() for c in ctxs])


class ContextPool:
    """Many small problems (the aligners re-run the MUM search inside every inter-anchor gap,
    src/mauveAligner.cpp:94,698, src/progressiveMauve.cpp:661-664): a pool of library contexts, each with its own CUDA
    stream and workspace, driven by one host thread each.  A small search is launch- and latency-bound (about 40 short
    kernels and 5 scalar read-backs), so independent searches overlap on the device; the C ABI calls release the GIL.
    Every problem runs through the same mb_find as a single search, so the results are the same bit for bit."""

    def __init__(self, n_contexts=8, device=0):
        self.ctxs = [Context(device) for _ in range(n_contexts)]

    def close(self):
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    def find_many(self, problems, pattern, mode=L.MODE_UNIQUE, **kw):
        """problems: iterable of sequence lists -> list of result dicts, in order."""
        import queue
        import threading
        problems = list(problems)
        out = [None] * len(problems)
        work = queue.SimpleQueue()
        for i in range(len(problems)):
            work.put(i)
        errors = []

        def run(ctx):
            try:
                ctx.set_seed(pattern)
                while True:
                    try:
                        i = work.get_nowait()
                    except queue.Empty:
                        return
                    ctx.clear_sequences()
                    for s in problems[i]:
                        ctx.add_sequence(s)
                    out[i] = ctx.find(mode, **kw)
            except Exception as e:  # surfaced to the caller below
                errors.append(e)

        threads = [threading.Thread(target=run, args=(c,)) for c in self.ctxs[:max(1, min(len(self.ctxs), len(problems)))]]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out

"""Naive, independent restatement of SURVEY.md Appendix A (D1-D18) used to pin the
C++ oracle on small inputs.  TEST INFRASTRUCTURE ONLY.

No sorting tricks, no hashing by diagonal, no Invert(): every rule is applied the
slow, literal way straight from the ASCII sequences, so an agreement between this
file and oracle/oracle.cpp is evidence that both implement the written rules.

Policy code follows /root/reference/src/UniqueMatchFinder.cpp:36-60 and
/root/reference/src/SeedMatchEnumerator.h:71-141.
"""
from __future__ import annotations

MODE_UNIQUE, MODE_SEED_ENUM, MODE_UNIQUE_COUNT, MODE_PAIRWISE, MODE_REPEAT = 0, 1, 2, 3, 4
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "a": 0, "c": 1, "g": 2, "t": 3}


def seed_len(pattern: int) -> int:
    return pattern.bit_length()


def care_offsets(pattern: int):
    L = seed_len(pattern)
    return [j for j in range(L) if (pattern >> (L - 1 - j)) & 1]


def mer_at(seq: str, p: int, pattern: int):
    """D4: (key, strand) of the L-window at 0-based p."""
    L = seed_len(pattern)
    win = [_CODE.get(ch, 0) for ch in seq[p:p + L]]
    rc = [3 - b for b in reversed(win)]
    offs = care_offsets(pattern)
    fwd = 0
    rev = 0
    for j in offs:
        fwd = fwd * 4 + win[j]
        rev = rev * 4 + rc[j]
    if rev < fwd:
        return rev, 1
    return fwd, 0


def all_mers(seqs, pattern):
    L = seed_len(pattern)
    return [[mer_at(s, p, pattern) for p in range(len(s) - L + 1)] if len(s) >= L else [] for s in seqs]


def _wok(mers, comps, L, length, grow_left):
    """D13 window predicate for the newly covered window.
    comps: list of (g, start1 (1-based left end), reverse?) with the *already grown* coordinates."""
    ref = None
    for g, start, rev in comps:
        if grow_left:
            pos = start if not rev else start + length - L  # match-left = right end of a reverse comp
        else:
            pos = start + length - L if not rev else start
        key, strand = mers[g][pos - 1]
        par = strand ^ (1 if rev else 0)
        if ref is None:
            ref = (key, par)
        elif ref != (key, par):
            return False
    return True


def extend(mers, lens, st, L):
    """D14.  st: dict g -> signed 1-based start.  Returns (st, length)."""
    comps = [[g, abs(s), s < 0] for g, s in sorted(st.items())]
    length = L

    def room(left):
        r = None
        for g, start, rev in comps:
            lroom = start - 1
            rroom = lens[g] - (start + length - 1)
            v = (lroom if not rev else rroom) if left else (rroom if not rev else lroom)
            r = v if r is None else min(r, v)
        return r

    def grow(left, step):
        nonlocal length
        # tentatively grow
        for c in comps:
            if (left and not c[2]) or (not left and c[2]):
                c[1] -= step
        length += step
        if _wok(mers, comps, L, length, left):
            return True
        length -= step
        for c in comps:
            if (left and not c[2]) or (not left and c[2]):
                c[1] += step
        return False

    for left, step, cap in ((True, L, None), (False, L, None), (True, 1, L), (False, 1, L)):
        n = 0
        while (cap is None or n < cap) and room(left) >= step:
            if not grow(left, step):
                break
            n += 1
    return {g: (-s if rev else s) for g, s, rev in comps}, length


def _same_group_contains(e, c, L):
    """D16: accepted extended match e contains length-L candidate c."""
    est, elen = e
    if set(est) != set(c):
        return False
    first = min(c)
    for g in c:
        if (c[g] < 0) != (est[g] < 0):
            return False
    d = c[first] - est[first]
    if d < 0 or d + L > elen:
        return False
    for g in c:
        if c[g] > 0:
            if c[g] - est[g] != d:
                return False
        else:
            if abs(c[g]) - abs(est[g]) != elen - L - d:
                return False
    return True


def find(seqs, pattern, mode, min_multi=2, max_multi=1000, direct_only=False, nway_mask=0):
    """Returns dict(matches=[(length, [(g, start), ...])], unique_mers=int, unique_per_seq=[...])."""
    L = seed_len(pattern)
    N = len(seqs)
    lens = [len(s) for s in seqs]
    mers = all_mers(seqs, pattern)
    buckets = {}
    for g in range(N):
        for p, (key, strand) in enumerate(mers[g]):
            buckets.setdefault(key, []).append((g, p, strand))
    res = dict(unique_mers=len(buckets), unique_per_seq=[len({k for k, _ in mers[g]}) for g in range(N)], matches=[])
    if mode == MODE_UNIQUE_COUNT:
        return res
    if mode == MODE_SEED_ENUM:
        if N != 1:
            return res
        out = []
        for key in sorted(buckets):
            b = sorted(buckets[key], key=lambda t: t[1])
            m = len(b)
            if m < 2:
                continue
            st = [p + 1 for _, p, _ in b]
            ref = b[0][2]
            st = [s if b[i][2] == ref else -s for i, s in enumerate(st)]
            found_reverse = any(s < 0 for s in st)
            if m > max_multi or m < min_multi:
                continue
            if direct_only and found_reverse:
                fw = [s for s in st if s > 0]
                if len(fw) > 1:
                    out.append((L, fw))
            else:
                out.append((L, st))
        out.sort(key=lambda r: (r[1][0], len(r[1]), r[1][1:]))
        res["matches"] = [(ln, [(0, s) for s in st]) for ln, st in out]
        return res

    if mode == MODE_REPEAT:
        # RepeatHash: one sequence, every occurrence of a bucket a component (columns = occurrences in position order),
        # then containment de-dup + extension exactly as for a multi-genome entry
        if N != 1:
            return res
        accepted = []
        for key in sorted(buckets):
            b = sorted(buckets[key], key=lambda t: t[1])
            m = len(b)
            if m < 2 or m < min_multi or m > max_multi or m > 255:
                continue
            st = {i: (p + 1) if strand == b[0][2] else -(p + 1) for i, (_, p, strand) in enumerate(b)}
            if any(_same_group_contains(e, st, L) for e in accepted):
                continue
            accepted.append(extend([mers[0]] * m, [lens[0]] * m, st, L))
        out = [(length, [st[i] for i in range(len(st))]) for st, length in accepted]
        out.sort(key=lambda r: (r[1][0], len(r[1]), r[1][1:], r[0]))
        res["matches"] = [(ln, [(0, s) for s in st]) for ln, st in out]
        return res

    accepted = []  # (st dict, length)

    def hash_match(entries):
        first_strand = entries[0][2]
        st = {}
        for i, (g, p, strand) in enumerate(entries):
            st[g] = (p + 1) if strand == first_strand else -(p + 1)
        for e in accepted:
            if _same_group_contains(e, st, L):
                return
        st2, length = extend(mers, lens, st, L)
        accepted.append((st2, length))

    for key in sorted(buckets):
        b = sorted(buckets[key])
        if len(b) < 2:
            continue
        cnt = {}
        for g, _, _ in b:
            cnt[g] = cnt.get(g, 0) + 1
        uniq = [t for t in b if cnt[t[0]] == 1]
        if len(uniq) < 2:
            continue
        if mode == MODE_PAIRWISE:
            for i in range(len(uniq)):
                for j in range(i + 1, len(uniq)):
                    hash_match([uniq[i], uniq[j]])
            continue
        if nway_mask:
            present = 0
            for g, _, _ in uniq:
                present |= 1 << g
            if present != nway_mask:
                continue
        hash_match(uniq)

    def sort_key(m):
        st, length = m
        return ([abs(st.get(g, 0)) for g in range(N)], [1 if st.get(g, 0) < 0 else 0 for g in range(N)], length)

    accepted.sort(key=sort_key)
    res["matches"] = [(length, sorted(st.items())) for st, length in accepted]
    return res


def find_family(seqs, patterns, nway_mask=0):
    """Seed-family search (src/progressiveMauve.cpp:503-548), the naive way: MODE_UNIQUE once per pattern in the given
    order, with ONE list of accepted matches that persists across the patterns — a candidate of a later pattern is
    dropped when an accepted match of any earlier pattern (or an earlier one of its own pass) contains its seed window."""
    N = len(seqs)
    lens = [len(s) for s in seqs]
    accepted = []
    for pattern in patterns:
        L = seed_len(pattern)
        mers = all_mers(seqs, pattern)
        buckets = {}
        for g in range(N):
            for p, (key, strand) in enumerate(mers[g]):
                buckets.setdefault(key, []).append((g, p, strand))
        for key in sorted(buckets):
            b = sorted(buckets[key])
            if len(b) < 2:
                continue
            cnt = {}
            for g, _, _ in b:
                cnt[g] = cnt.get(g, 0) + 1
            uniq = [t for t in b if cnt[t[0]] == 1]
            if len(uniq) < 2:
                continue
            if nway_mask:
                present = 0
                for g, _, _ in uniq:
                    present |= 1 << g
                if present != nway_mask:
                    continue
            first_strand = uniq[0][2]
            st = {g: (p + 1) if strand == first_strand else -(p + 1) for g, p, strand in uniq}
            if any(_same_group_contains(e, st, L) for e in accepted):
                continue
            accepted.append(extend(mers, lens, st, L))

    def sort_key(m):
        st, length = m
        return ([abs(st.get(g, 0)) for g in range(N)], [1 if st.get(g, 0) < 0 else 0 for g in range(N)], length)

    accepted.sort(key=sort_key)
    return dict(matches=[(length, sorted(st.items())) for st, length in accepted])


# ---- match-list post-filters, restated per base: a match is a list of COLUMNS, every column one
# coordinate per sequence (None where the match is absent); cropping and splitting fall out of dropping / regrouping columns
def _columns(length, starts):
    cols = []
    for j in range(length):
        cols.append([None if s == 0 else (s + j if s > 0 else -s + (length - 1 - j)) for s in starts])
    return cols


def _from_columns(cols, starts):
    """columns (a contiguous run of the original match, in match order) -> (length, starts)"""
    n = len(cols)
    out = []
    for g, s in enumerate(starts):
        if s == 0:
            out.append(0)
        elif s > 0:
            out.append(cols[0][g])
        else:
            out.append(-cols[-1][g])
    return n, out


def eliminate_overlaps(matches):
    """matches: [(length, [start per sequence])].  Per sequence g, in order of (left end, list position): every base of
    g belongs to the first match of that order that covers it; a match keeps the columns whose base in g it owns."""
    items = [(_columns(l, st), list(st)) for l, st in matches]
    nseq = len(matches[0][1]) if matches else 0
    for g in range(nseq):
        order = sorted((i for i, (cols, st) in enumerate(items) if cols and st[g] != 0), key=lambda i: min(c[g] for c in items[i][0]))
        owned = set()
        for i in order:
            cols, st = items[i]
            keep = [c for c in cols if c[g] not in owned]
            owned.update(c[g] for c in cols)
            items[i] = (keep, st)
    return [_from_columns(cols, st) for cols, st in items if cols]


def transpose_matches(matches, seqI, regions):
    """filtered -> original coordinates of sequence seqI, per column; a match is cut wherever consecutive columns fall
    into different regions"""
    if len(regions) < 2:
        return [(l, list(st)) for l, st in matches]
    orig = []
    for k in range(len(regions) // 2):
        orig.extend((k, x) for x in range(regions[2 * k], regions[2 * k + 1] + 1))
    last_k, last_x = orig[-1]
    out = []
    for l, st in matches:
        if st[seqI] == 0:
            out.append((l, list(st)))
            continue
        cols = _columns(l, st)
        tagged = []
        for c in cols:
            f = c[seqI]
            k, x = orig[f - 1] if f - 1 < len(orig) else (last_k, last_x + (f - len(orig)))
            c = list(c)
            c[seqI] = x
            tagged.append((k, c))
        pieces, cur = [], [tagged[0]]
        for t in tagged[1:]:
            if t[0] != cur[-1][0]:
                pieces.append(cur)
                cur = []
            cur.append(t)
        pieces.append(cur)
        if st[seqI] < 0:
            pieces.reverse()  # pieces in ascending order of the seqI coordinate
        for pc in pieces:
            out.append(_from_columns([c for _, c in pc], st))
    return out

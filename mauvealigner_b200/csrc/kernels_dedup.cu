// kernels_dedup.cu — a9-a11 of SURVEY.md §8a: MemHash::AddHashEntry (containment de-dup in ascending
// seed order) and MatchFinder::ExtendMatch (ungapped extension), both absent from /root/reference
// (libMems); semantics = SURVEY.md Appendix A D13-D16 and the A.2 pseudo-code.  Reached in the
// reference from /root/reference/src/UniqueMatchFinder.cpp:58 (HashMatch(unique_list)).
//
// The reference decides candidates one at a time in ascending seed order (= candidate index here,
// the "rank"): a candidate is dropped iff an already ACCEPTED extended match of its group (same
// genome set, strands and diagonal, D16) contains it, otherwise it is extended and accepted.
// That order-dependent rule is evaluated here in three data-parallel steps:
//
//  1 chains   Slots = candidates ordered by (first genome, position) through a bitmap + popcount
//             ranks.  Candidates of one group at consecutive positions x, x+1, ... form a chain.
//             Every extent of a same-group match is bounded by a failing window (or a sequence end),
//             so a match that contains one member of a chain contains all of it; therefore only
//             the lowest-rank member of a chain (its "rep") can ever be accepted and all other
//             members are dropped without being extended.  Two segmented min-scans (forward,
//             backward; single pass, decoupled look-back) find the reps.
//  2 extend   every rep is extended once (pure function of the packed genomes): one thread on
//             64-base mismatch maps; long extensions by one warp, 32 chunks per step.
//  3 resolve  fix-point over the reps, equal to the sequential result: in rounds, every undecided
//             rep claims (atomicMin of its rank) the undecided higher-rank reps of its group inside
//             its extent; an unclaimed rep is accepted and marks those reps covered; a covered rep
//             is dropped.  The lowest undecided rank is never claimed, so every round decides.
#include "common.cuh"
#include "kernels.h"
#include "lookback.cuh"

#define INF32 0xFFFFFFFFu
#define INF64 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ u32 slot_rank(const u64* __restrict__ bitmap, const u32* __restrict__ bmrank, u64 gp) {
    u64 w = bitmap[gp >> 6];
    return bmrank[gp >> 6] + (u32)__popcll(w & ((1ull << (gp & 63)) - 1));
}

// D16 group test, exact: same genome set, same strands, same diagonal.  Four components per step with all loads issued
// before the first comparison (the loop is latency-bound: a one-at-a-time walk with an early exit per component took
// 1.3 ms in k_chain and 1.2 ms in k_resolve on C5).
__device__ __forceinline__ bool same_group(const DedupArgs& a, u32 j, u32 e) {
    const u32 offj = a.cand_off[j], offe = a.cand_off[e];
    const u32 m = a.cand_off[e + 1] - offe;
    if (a.cand_off[j + 1] - offj != m) return false;
    const u32 xj = a.comp_pos[offj], xe = a.comp_pos[offe];
    const u32* __restrict__ pj = a.comp_pos + offj;
    const u32* __restrict__ pe = a.comp_pos + offe;
    const u8* __restrict__ gj = a.comp_gs + offj;
    const u8* __restrict__ ge = a.comp_gs + offe;
    u32 k = 0;
    for (; k + 4 <= m; k += 4) {
        u32 vj[4], ve[4], sj[4], se[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { vj[t] = pj[k + t]; ve[t] = pe[k + t]; sj[t] = gj[k + t]; se[t] = ge[k + t]; }
        bool ok = true;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            ok &= sj[t] == se[t];
            ok &= (se[t] & 0x80u) ? (vj[t] + xj == ve[t] + xe) : (vj[t] - xj == ve[t] - xe);
        }
        if (!ok) return false;
    }
    for (; k < m; ++k) {
        const u32 s1 = gj[k], s2 = ge[k];
        if (s1 != s2) return false;
        if (s2 & 0x80u) { if (pj[k] + xj != pe[k] + xe) return false; }
        else if (pj[k] - xj != pe[k] - xe) return false;
    }
    return true;
}

// Every decision that DROPS a candidate (a chain link, a containment) compares the two group hashes first — the fast
// reject — and then the groups themselves, component by component (same_group): bit-exact results do not rest on a hash.
// Measured cost of the exact comparison: C5 33.3 -> 35.7 ms (+ 7 %; tools/bench_variants.sh with -DMB_HASH_ONLY).
// Hash equality alone (two independent 64-bit hashes, 120 bits compared) decides only where the component lists are not
// at hand: on the owner side of the multi-GPU path (a.rows: the rows carry the hashes, the lists stay at their source)
// and in builds with -DMB_HASH_ONLY.
__device__ __forceinline__ bool groups_equal(const DedupArgs& a, u32 c1, u32 c2) {
#ifndef MB_HASH_ONLY
    return a.rows ? true : same_group(a, c1, c2);
#else
    (void)a; (void)c1; (void)c2;
    return true;
#endif
}

// ---- slots ---------------------------------------------------------------------------------------
// candidate -> slot (rank of its first-genome position among all candidates).  One 32-byte record per slot
// (one full-sector scattered store per candidate; everything later steps need about the candidate, so that
// they never chase the candidate CSR again):
//   a.x = group hash with bit 0 replaced by "the previous base of the genome holds a candidate too" (then
//         that candidate is slot s-1),  a.y = candidate | components << 32 | first genome << 40 | virtual genome << 48
//   b.x = second group hash,             b.y = first component row | first position << 32
#define HASH_MASK (~1ull)
__global__ void __launch_bounds__(256) k_slot_scatter(DedupArgs a, GenomeTable gt) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_cand) return;
    u32 off, m, g, p, vg;
    u64 h1, h2;
    if (a.rows) { // multi-GPU owner: the row carries what the CSR would tell
        const ulonglong2 r0 = reinterpret_cast<const ulonglong2*>(a.rows)[2 * (size_t)c], r1 = reinterpret_cast<const ulonglong2*>(a.rows)[2 * (size_t)c + 1];
        h1 = r0.x; h2 = r0.y;
        g = (u32)r1.x & 0xFFu; vg = (u32)(r1.x >> 8) & 0xFFFFu; m = (u32)(r1.x >> 24) & 0xFFu; p = (u32)(r1.x >> 32);
        off = 0;
    } else {
        off = a.cand_off[c]; m = a.cand_off[c + 1] - off;
        g = a.comp_gs[off] & 0x7F; p = a.comp_pos[off];
        vg = vgenome(gt, g, a.comp_gs[off + 1] & 0x7F); // a candidate has >= 2 components
        h1 = a.ghash[c]; h2 = a.ghash2[c];
    }
    u64 gp = gt.vbase[vg] + p;
    u64 w = a.bitmap[gp >> 6];
    u32 s = a.bmrank[gp >> 6] + (u32)__popcll(w & ((1ull << (gp & 63)) - 1));
    // a candidate never sits on the last L-1 bases of a genome, so base gp-1 belongs to the same genome
    // whenever it holds a candidate
    u64 adj = 0;
    if (gp > 0) adj = (gp & 63) ? (w >> ((gp & 63) - 1)) & 1 : (a.bitmap[(gp >> 6) - 1] >> 63) & 1;
    a.slot_rec[2 * (size_t)s] = make_ulonglong2((h1 & HASH_MASK) | adj, (u64)c | ((u64)m << 32) | ((u64)g << 40) | ((u64)vg << 48));
    a.slot_rec[2 * (size_t)s + 1] = make_ulonglong2(h2, (u64)off | ((u64)p << 32));
}

// ---- chains: segmented min-scans over the slots ---------------------------------------------------
// Element = (head flag, rank).  Exclusive segmented prefix minimum; REV = false scans towards higher
// slots (segment heads = chain heads) and records the links, REV = true scans towards lower slots
// (segment heads = chain tails), combines both directions and emits the reps.
#define CH_NT 256
#define CH_IPT 8
#define CH_TILE (CH_NT * CH_IPT)

// Tile carry by decoupled look-back.  Status word: 2 flag bits | head-seen bit (32) | open minimum (32).
// A predecessor that has seen a head (or holds an inclusive value) ends the walk.
__device__ __forceinline__ u32 segmin_lookback(u64* status, u32 tile, u32 F, u32 M) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) lb_st(status + tile, (tile == 0 ? LB_FLAG_INC : LB_FLAG_AGG) | ((u64)F << 32) | M);
    if (tile == 0) return INF32;
    u32 carry = INF32;
    i64 base = (i64)tile - 1;
    while (true) {
        i64 t = base - lane;
        u64 s = LB_FLAG_INC | INF32;
        if (t >= 0) {
            do { s = lb_ld(status + t); } while ((s >> 62) == 0);
        }
        bool stop = (s & LB_FLAG_INC) != 0 || ((s >> 32) & 1);
        u32 stops = __ballot_sync(0xFFFFFFFFu, stop);
        int first = __ffs(stops) - 1;
        u32 v = (first < 0 || lane <= first) ? (u32)s : INF32;
        carry = min(carry, __reduce_min_sync(0xFFFFFFFFu, v));
        if (first >= 0) break;
        base -= 32;
    }
    if (lane == 0) lb_st(status + tile, LB_FLAG_INC | ((u64)F << 32) | (F ? M : min(carry, M)));
    return carry;
}

template <bool REV>
__global__ void __launch_bounds__(CH_NT) k_chain(DedupArgs a, u64* status, u32* ticket) {
    __shared__ u32 sF[CH_NT / 32], sM[CH_NT / 32];
    __shared__ u32 sTile, sCarry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 n = a.n_cand;
    const u32 ntiles = (n + CH_TILE - 1) / CH_TILE;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 ord = sTile;                                // position of this tile in scan order
    const u32 tile = REV ? ntiles - 1 - ord : ord;
    // the thread's 8 consecutive slots; scan order inside the thread: ascending (FWD) / descending (REV)
    const u32 sbase = tile * CH_TILE + (REV ? (CH_NT - 1 - tid) : tid) * CH_IPT;
    u32 v[CH_IPT];
    bool f[CH_IPT];
    u64 hs[CH_IPT + 1];
    if (!REV) {
        // link of slot s: same group as slot s-1 and exactly one base further
        const u32 s0 = sbase;
        hs[0] = (s0 > 0 && s0 - 1 < n) ? a.slot_rec[2 * (size_t)(s0 - 1)].x : 0;
        u32 linkbits = 0;
#pragma unroll
        for (int k = 0; k < CH_IPT; ++k) {
            u32 s = s0 + k;
            bool valid = s < n;
            ulonglong2 r = valid ? a.slot_rec[2 * (size_t)s] : make_ulonglong2(0, INF32);
            hs[k + 1] = r.x;
            v[k] = (u32)r.y;
            bool link = valid && s > 0 && (r.x & 1) && (hs[k] & HASH_MASK) == (r.x & HASH_MASK);
#ifndef MB_HASH_ONLY
            if (link && !a.rows) link = same_group(a, (u32)a.slot_rec[2 * (size_t)(s - 1)].y, v[k]);
#endif
            f[k] = !link;
            linkbits |= (link ? 1u : 0u) << k;
        }
        if (s0 < n) a.link_bits[s0 / CH_IPT] = (u8)linkbits; // bit k: slot s0+k continues the chain of s0+k-1
    } else {
        // scan order k = 0..7 <-> slot sbase+7-k; head (in this direction) = the chain's last slot
        u32 lb = sbase < n ? a.link_bits[sbase / CH_IPT] : 0u;
        u32 lb_next = (sbase + CH_IPT < n) ? a.link_bits[sbase / CH_IPT + 1] : 0u;
        lb |= (lb_next & 1u) << CH_IPT;
#pragma unroll
        for (int k = 0; k < CH_IPT; ++k) {
            u32 s = sbase + (CH_IPT - 1 - k);
            bool valid = s < n;
            v[k] = valid ? (u32)a.slot_rec[2 * (size_t)s].y : INF32;
            bool link_next = valid && s + 1 < n && ((lb >> (CH_IPT - k)) & 1u); // slot s+1 continues s
            f[k] = !link_next;
        }
    }
    // thread aggregate: (any head, minimum after the last head)
    u32 F = 0, M = INF32;
#pragma unroll
    for (int k = 0; k < CH_IPT; ++k) {
        if (f[k]) { F = 1; M = v[k]; } else M = min(M, v[k]);
    }
    // inclusive scan of the aggregates over the block, then exclusive per thread
    u32 iF = F, iM = M;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 yF = __shfl_up_sync(0xFFFFFFFFu, iF, o), yM = __shfl_up_sync(0xFFFFFFFFu, iM, o);
        if (lane >= o) { iM = iF ? iM : min(yM, iM); iF |= yF; }
    }
    if (lane == 31) { sF[warp] = iF; sM[warp] = iM; }
    __syncthreads();
    // exclusive prefix of this thread inside the tile
    u32 eF = __shfl_up_sync(0xFFFFFFFFu, iF, 1), eM = __shfl_up_sync(0xFFFFFFFFu, iM, 1);
    if (lane == 0) { eF = 0; eM = INF32; }
    u32 wF = 0, wM = INF32; // all earlier warps
    for (int w = 0; w < warp; ++w) { u32 f2 = sF[w], m2 = sM[w]; wM = f2 ? m2 : min(wM, m2); wF |= f2; }
    u32 pM = eF ? eM : min(wM, eM), pF = wF | eF;
    if (warp == 0) {
        u32 tF = 0, tM = INF32;
        for (int w = 0; w < CH_NT / 32; ++w) { u32 f2 = sF[w], m2 = sM[w]; tM = f2 ? m2 : min(tM, m2); tF |= f2; }
        u32 carry = segmin_lookback(status, ord, tF, tM);
        if (lane == 0) sCarry = carry;
    }
    __syncthreads();
    u32 run = pF ? pM : min(sCarry, pM); // open minimum entering this thread's first element
    u32 repbits = 0;
#pragma unroll
    for (int k = 0; k < CH_IPT; ++k) {
        u32 excl = f[k] ? INF32 : run;
        u32 s = REV ? sbase + (CH_IPT - 1 - k) : sbase + k;
        if (s < n) {
            if (!REV) a.chain_min[s] = excl;
            else {
                u32 c = v[k];
                bool rep = c < excl && c < a.chain_min[s];
                if (rep) repbits |= 1u << (CH_IPT - 1 - k);
            }
        }
        run = f[k] ? v[k] : min(run, v[k]);
    }
    if (REV && sbase < n) reinterpret_cast<u8*>(a.rep_bits)[sbase / CH_IPT] = (u8)repbits;
}

// Reps in slot order (bitmap ranks) -> sort key (group colour = top 16 bits of the group hash, slot).
// Two stable radix passes on the colour (driver) then put the reps of one group next to each other,
// ordered along their diagonal, so "the reps of my group inside my extent" is a short index range.
// Single-GPU path: also the extension record of the rep, in SLOT order — the extension runs over the reps in
// (first genome, position) order, where neighbouring reps read the same genome sectors (L1 hits) instead of the
// (colour, slot) order of the resolve step, where every component window is a random L2 access.
// Extension order.  In slot order the reps of ONE first genome are neighbours, but a sector of genome 40 is also read by the
// reps whose first genome is 0, 1, ... 39 — one pass over the slot axis later each time, long after the sector has left
// the L2 (C5: 9 GB of DRAM reads for 80 MB of packed genomes).  Related genomes are roughly collinear, so the extension
// records are laid out by (block of 4096 positions, first genome, slot) instead: the reps around one position of ALL
// first genomes run together and find each other's sectors in the L2.  Only a layout of xrec: any order gives the same
// extents.  Cells (virtual genome v, block b), index t = b * V + v: cellR[t] = reps before the cell in slot order,
// cellNew[t] = reps before it in the new order (k_cell_bounds counts, k_cell_scan scans).
#ifndef XO_SHIFT
#define XO_SHIFT 12
#endif
__global__ void __launch_bounds__(256) k_cell_bounds(DedupArgs a, GenomeTable gt, u32 V, u32 NB, u64 axis_end, u32* __restrict__ cellR,
                                                      u32* __restrict__ cellCnt) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= V * NB) return;
    const u32 b = t / V, v = t % V;
    const u64 v0 = gt.vbase[v], v1 = v + 1 < V ? gt.vbase[v + 1] : axis_end;
    const u64 lo = min(v0 + ((u64)b << XO_SHIFT), v1), hi = min(v0 + ((u64)(b + 1) << XO_SHIFT), v1);
    u32 r[2];
    for (int k = 0; k < 2; ++k) {
        const u32 s = slot_rank(a.bitmap, a.bmrank, k ? hi : lo);
        r[k] = a.rep_rank[s >> 6] + (u32)__popcll(a.rep_bits[s >> 6] & ((1ull << (s & 63)) - 1));
    }
    cellR[t] = r[0];
    cellCnt[t] = r[1] - r[0];
}
// exclusive scan in place, one block (the table is small: axis length / 4096 entries)
__global__ void __launch_bounds__(1024) k_cell_scan(u32* __restrict__ v, u32 n) {
    __shared__ u32 swarp[32];
    __shared__ u32 carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += 4096) {
        const u32 i0 = base + threadIdx.x * 4;
        u32 x[4];
        u32 sum = 0;
        for (int k = 0; k < 4; ++k) { x[k] = i0 + k < n ? v[i0 + k] : 0; sum += x[k]; }
        u32 inc = sum;
        for (int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += y; }
        if (lane == 31) swarp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            u32 w = swarp[lane], winc = w;
            for (int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, winc, d); if (lane >= d) winc += y; }
            swarp[lane] = winc - w;
        }
        __syncthreads();
        u32 ex = carry + swarp[warp] + inc - sum;
        for (int k = 0; k < 4; ++k) { if (i0 + k < n) v[i0 + k] = ex; ex += x[k]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = ex;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_rep_keys(DedupArgs a, u64* __restrict__ skey, const u32* __restrict__ cellR, const u32* __restrict__ cellNew, u32 V) {
    u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_cand) return;
    u64 w = a.rep_bits[s >> 6];
    if (!((w >> (s & 63)) & 1)) return;
    u32 idx = a.rep_rank[s >> 6] + (u32)__popcll(w & ((1ull << (s & 63)) - 1));
    const ulonglong2 r = a.slot_rec[2 * (size_t)s];
    skey[idx] = ((r.x >> 48) << 32) | s;
    if (!a.rows) {
        const ulonglong2 q = a.slot_rec[2 * (size_t)s + 1];
        const u32 c = (u32)r.y, m = (u32)(r.y >> 32) & 0xFFu, g0 = (u32)(r.y >> 40) & 0xFFu, vg = (u32)(r.y >> 48) & 0xFFu;
        const u32 p0 = (u32)(q.y >> 32);
        u32 at = idx;
        if (cellR) {
            const u32 t = (p0 >> XO_SHIFT) * V + vg;
            at = cellNew[t] + (idx - cellR[t]);
        }
        a.xrec[at] = make_uint4(c, (u32)q.y, m | (g0 << 8) | (vg << 16), p0);
    }
}

// ---- extension -----------------------------------------------------------------------------------
// oriented masked window of component (g, p): forward comps as stored, reverse comps reverse-complemented
__device__ __forceinline__ void oriented_window(const u64* __restrict__ packed, const GenomeTable& gt, const SeedDev& sd, u32 g, u32 p,
                                                bool rev, u64& hi, u64& lo) {
    u64 h, l;
    load_window(packed + gt.word_base[g], p, sd.wide, h, l);
    if (rev) {
        // bases after the window sit in the low bits; rc_window shifts them out
        u64 wh, wl;
        rc_window(sd.L, h, l, wh, wl);
        hi = wh & sd.mask_hi; lo = wl & sd.mask_lo; // palindromic mask: same care columns on both strands
    } else {
        hi = h & sd.mask_hi; lo = l & sd.mask_lo;
    }
}

// Warp version for L > 32: number of consecutive steps t = 1..maxcount whose windows agree across all
// components.  dir = -1: grow left in match coordinates, +1: grow right.  Offset of step t: o0 + t*stride.
__device__ u32 scan_steps(const u64* __restrict__ packed, const GenomeTable& gt, const SeedDev& sd, const u32* __restrict__ cpos,
                          const u8* __restrict__ cgs, u32 m, int dir, u32 o0, u32 stride, u32 maxcount) {
    const int lane = threadIdx.x & 31;
    const u32 m1 = m - 1;
    const u64 total = (u64)maxcount * m1;
    const u32 g0 = cgs[0] & 0x7F, p0 = cpos[0];
    for (u64 base = 0; base < total; base += 32) {
        u64 item = base + lane;
        bool ok = true;
        if (item < total) {
            u32 t = 1 + (u32)(item / m1), k = 1 + (u32)(item % m1);
            u32 off = o0 + t * stride;
            u64 ah, al, bh, bl;
            oriented_window(packed, gt, sd, g0, dir < 0 ? p0 - off : p0 + off, false, ah, al);
            u8 gs = cgs[k];
            bool rev = gs & 0x80;
            u32 pk = cpos[k];
            u32 q = ((dir < 0) != rev) ? pk - off : pk + off;
            oriented_window(packed, gt, sd, gs & 0x7F, q, rev, bh, bl);
            ok = (ah == bh) && (al == bl);
        }
        u32 fail = __ballot_sync(0xFFFFFFFFu, !ok);
        if (fail) {
            u64 first = base + (u32)(__ffs(fail) - 1);
            return (u32)(first / m1); // = t_fail - 1
        }
    }
    return maxcount;
}

__device__ __forceinline__ void candidate_room(const GenomeTable& gt, u32 L, const u32* cpos, const u8* cgs, u32 k, u32& room_l, u32& room_r) {
    u8 gs = cgs[k];
    u32 p = cpos[k], lo, hi;
    seg_range(gt, gs & 0x7F, p, lo, hi);
    u32 lroom = p - lo, rroom = hi - L - p;
    bool rev = gs & 0x80;
    room_l = min(room_l, rev ? rroom : lroom);
    room_r = min(room_r, rev ? lroom : rroom);
}

// D14: four phases.  One warp per candidate, window by window (L > 32).
__device__ void extend_candidate_warp(const u64* __restrict__ packed, const GenomeTable& gt, const SeedDev& sd, const u32* cpos,
                                      const u8* cgs, u32 m, u32& ext_l, u32& ext_r) {
    const int lane = threadIdx.x & 31;
    const u32 L = sd.L;
    u32 room_l = INF32, room_r = INF32;
    for (u32 k = lane; k < m; k += 32) candidate_room(gt, L, cpos, cgs, k, room_l, room_r);
    room_l = __reduce_min_sync(0xFFFFFFFFu, room_l);
    room_r = __reduce_min_sync(0xFFFFFFFFu, room_r);
    u32 a = scan_steps(packed, gt, sd, cpos, cgs, m, -1, 0, L, room_l / L);
    u32 b = scan_steps(packed, gt, sd, cpos, cgs, m, +1, 0, L, room_r / L);
    u32 c = scan_steps(packed, gt, sd, cpos, cgs, m, -1, a * L, 1, min(L, room_l - a * L));
    u32 d = scan_steps(packed, gt, sd, cpos, cgs, m, +1, b * L, 1, min(L, room_r - b * L));
    ext_l = a * L + c;
    ext_r = b * L + d;
}

// ---- extension on mismatch maps (L <= 32) ----------------------------------------------------------
// 64 bases of one component at match offsets [i0, i0+64) (relative to the seed start; the match strand
// is component 0's), first base in the top bits of `a`.  Reads may run up to 128 bases outside the
// genome: the packed buffer is padded on both sides of every genome and such bases never reach a
// tested window.
// The packed genomes (0.25 B/bp: 10 MB at C2, 80 MB at C5) are the only data of the extension that is re-read — at
// random, by every component of every rep — while the rep records, component rows and extents stream through once.
// Genome words are therefore loaded with an L2 evict-last policy (ncu before: 30 % L2 hit rate and 1.5 GB of DRAM
// reads in k_extend at C2).  Measured effect: small (extension 1025 -> 995 us at C2, none at C5); streaming hints
// (ld.cs / st.cs) on the rep records, component rows and extents on top of it made it slower and were dropped.
__device__ __forceinline__ u64 l2_keep_policy() {
    u64 pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ u64 ld_keep(const u64* p, u64 pol) {
    u64 v;
    asm("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void oriented_bases64(const u64* __restrict__ packed, const GenomeTable& gt, u32 L, u32 g, u32 pos, bool rev, i64 i0,
                                                 u64& a, u64& b) {
    i64 q = (i64)gt.word_base[g] * 32 + (i64)pos + (rev ? (i64)L - 64 - i0 : i0);
    const u64* w = packed + ((u64)q >> 5);
    int sh = (int)(q & 31) * 2;
    const u64 pol = l2_keep_policy();
    u64 w0 = ld_keep(w, pol), w1 = ld_keep(w + 1, pol), w2 = ld_keep(w + 2, pol);
    u64 fa = shl128_hi(w0, w1, sh), fb = shl128_hi(w1, w2, sh);
    if (rev) { a = rc_word(fb); b = rc_word(fa); }
    else { a = fa; b = fb; }
}
__device__ __forceinline__ u64 spread_nz(u64 x) { return (x | (x >> 1)) & 0x5555555555555555ull; }
// top 64 bits of the 128-bit map xa:xb shifted left by `idx` bases (0 <= idx < 64)
__device__ __forceinline__ u64 map_at(u64 xa, u64 xb, u32 idx) { return idx < 32 ? shl128_hi(xa, xb, 2 * (int)idx) : (xb << (2 * (idx - 32))); }

// Mismatch map (one flag per base, in the low bit of its 2-bit cell) of the 64 bases at match
// offsets [o_lo, o_lo + 64): flag set iff some component differs from component 0 there.
__device__ __forceinline__ void mismatch64(const u64* __restrict__ packed, const GenomeTable& gt, u32 L, const u32* __restrict__ cpos,
                                           const u8* __restrict__ cgs, u32 m, i64 o_lo, u64& xa, u64& xb) {
    u64 a0, b0;
    oriented_bases64(packed, gt, L, cgs[0] & 0x7F, cpos[0], false, o_lo, a0, b0);
    xa = 0; xb = 0;
    for (u32 k = 1; k < m; ++k) {
        u8 gs = cgs[k];
        u64 a, b;
        oriented_bases64(packed, gt, L, gs & 0x7F, cpos[k], gs & 0x80, o_lo, a, b);
        xa |= spread_nz(a ^ a0);
        xb |= spread_nz(b ^ b0);
    }
}

#define DD_THREAD_CHUNKS 8 // 64-base chunks one thread walks per direction before deferring to the warp phase

// chunk geometry of the directional walks.  Right (dir = +1): the chunk after b jumps starts at match
// offset bL+1, holds jump window b+1 at chunk index L-1 (and b+2 at 2L-1 when 3L <= 65); on a failing
// jump the single-step windows bL+s start at chunk index s-1 — all inside the same chunk.  Left is the
// mirror image: offsets [-bL-65+L, -bL+L-1), chunk index i <-> offset -bL-65+L+i.
__device__ __forceinline__ i64 chunk_lo(int dir, u32 b, u32 L) { return dir > 0 ? (i64)b * L + 1 : -(i64)b * L - 65 + (i64)L; }
// Window map of a chunk: cell q (flag in the low bit of the 2-bit cell, cell 0 at the top of hi) is set iff the
// L-window that starts at chunk index q covers a mismatch on a cared column = OR over the cared offsets o of
// the mismatch map shifted left by o cells.  Windows that run past the chunk see zeros there; callers only
// use windows that fit.  The loop is uniform (the pattern is a launch constant).
__device__ __forceinline__ void window_map(u64 xa, u64 xb, const SeedDev& sd, u64& bhi, u64& blo) {
    bhi = 0; blo = 0;
    for (int t = 0; t < sd.w; ++t) {
        const int s = 2 * (int)sd.care_off[t];
        bhi |= shl128_hi(xa, xb, s);
        blo |= xb << s;
    }
}
__device__ __forceinline__ bool win_bad(u64 bhi, u64 blo, u32 q) { return ((q < 32 ? bhi >> (62 - 2 * q) : blo >> (62 - 2 * (q - 32))) & 1) != 0; }
// number of consecutive good windows at chunk indices q, q+1, ... (at most 64 - q)
__device__ __forceinline__ u32 good_up(u64 bhi, u64 blo, u32 q) {
    u64 hi = q < 32 ? shl128_hi(bhi, blo, 2 * (int)q) : (blo << (2 * (q - 32)));
    u64 lo = q < 32 ? (blo << (2 * q)) : 0;
    if (hi) return (u32)__clzll((long long)hi) >> 1;
    if (lo) return 32 + ((u32)__clzll((long long)lo) >> 1);
    return 64 - q;
}
// number of consecutive good windows at chunk indices q, q-1, ... (at most q + 1)
__device__ __forceinline__ u32 good_down(u64 bhi, u64 blo, u32 q) {
    // shift right so that cell q becomes the lowest cell
    int s = 2 * (63 - (int)q);
    u64 lo = s >= 64 ? (bhi >> (s - 64)) : (s ? ((blo >> s) | (bhi << (64 - s))) : blo);
    u64 hi = s >= 64 ? 0 : (bhi >> s);
    if (lo) return ((u32)__ffsll((long long)lo) - 1) >> 1;
    if (hi) return 32 + (((u32)__ffsll((long long)hi) - 1) >> 1);
    return q + 1;
}

// evaluate one chunk that starts after b0 successful jumps: advances b over the jumps it holds; returns
// true when the walk ends inside this chunk (growth in `out`)
__device__ __forceinline__ bool walk_chunk(u64 bhi, u64 blo, int dir, u32 L, u32 room, u32 maxjumps, u32 per_chunk, u32& b, u32& out) {
    const u32 b0 = b;
    bool failed = false;
    for (u32 w = 0; w < per_chunk && b < maxjumps; ++w) {
        // right: jump window b+1 covers offsets [(b+1)L, (b+2)L)   -> chunk index (b-b0)L + L-1
        // left : jump window b+1 covers offsets [-(b+1)L, -bL)     -> chunk index 65 - 2L - (b-b0)L
        u32 idx = dir > 0 ? (b - b0) * L + L - 1 : 65 - 2 * L - (b - b0) * L;
        if (win_bad(bhi, blo, idx)) { failed = true; break; }
        ++b;
    }
    if (!(failed || b >= maxjumps)) return false;
    if (b - b0 == per_chunk && !failed) return false; // the chunk is used up: the single steps need a fresh one
    // single steps: right window s starts at offset bL+s (chunk index (b-b0)L + s - 1, ascending), left window s
    // at offset -bL-s (chunk index 65 - L - (b-b0)L - s, descending)
    u32 maxs = min(L, room - b * L);
    u32 run = dir > 0 ? good_up(bhi, blo, (b - b0) * L) : good_down(bhi, blo, 64 - L - (b - b0) * L);
    out = b * L + min(run, maxs);
    return true;
}

// Growth to the right (dir = +1) or left (dir = -1) of the seed: L-jumps, then at most L single steps
// (phases 1+3 resp. 0+2 of D14; the two directions do not interact).  One thread; returns the growth
// in bases or INF32 when the chunk budget ran out.
__device__ __forceinline__ u32 grow_thread(const u64* __restrict__ packed, const GenomeTable& gt, const SeedDev& sd, const u32* cpos,
                                           const u8* cgs, u32 m, int dir, u32 room) {
    const u32 L = sd.L;
    const u64 care = sd.mask_hi & 0x5555555555555555ull;
    const u32 maxjumps = room / L;
    const u32 per_chunk = (3 * L <= 65) ? 2 : 1;
    u32 b = 0, out = 0;
    for (int chunk = 0; chunk < DD_THREAD_CHUNKS; ++chunk) {
        u64 xa, xb;
        mismatch64(packed, gt, L, cpos, cgs, m, chunk_lo(dir, b, L), xa, xb);
        u64 bhi, blo;
        window_map(xa, xb, sd, bhi, blo);
        if (walk_chunk(bhi, blo, dir, L, room, maxjumps, per_chunk, b, out)) return out;
    }
    return INF32;
}

// Same walk by one warp: lane l examines the chunk that follows b_base + l*per_chunk jumps; the first
// lane whose chunk ends the walk has the answer.  Unbounded.
__device__ u32 grow_warp(const u64* __restrict__ packed, const GenomeTable& gt, const SeedDev& sd, const u32* cpos, const u8* cgs, u32 m,
                         int dir, u32 room) {
    const int lane = threadIdx.x & 31;
    const u32 L = sd.L;
    const u64 care = sd.mask_hi & 0x5555555555555555ull;
    const u32 maxjumps = room / L;
    const u32 per_chunk = (3 * L <= 65) ? 2 : 1;
    for (u64 b_base = 0;; b_base += 32 * per_chunk) {
        u64 b064 = b_base + (u64)lane * per_chunk;
        bool active = b064 <= maxjumps;
        u32 b = (u32)b064, out = 0;
        bool ends = false;
        if (active) {
            u64 xa, xb;
            mismatch64(packed, gt, L, cpos, cgs, m, chunk_lo(dir, b, L), xa, xb);
            u64 bhi, blo;
            window_map(xa, xb, sd, bhi, blo);
            ends = walk_chunk(bhi, blo, dir, L, room, maxjumps, per_chunk, b, out);
        }
        u32 em = __ballot_sync(0xFFFFFFFFu, ends);
        if (em) return __shfl_sync(0xFFFFFFFFu, out, __ffs(em) - 1);
    }
}

// One thread, both directions.  For 3L <= 64 one chunk centred on the seed (match offsets [-L, 63-L),
// chunk index i <-> offset i-L) holds the first jump window and all single-step windows of both
// directions, which settles every candidate whose first jumps fail (the usual case).
__device__ __forceinline__ bool extend_thread(const u64* __restrict__ packed, const GenomeTable& gt, const SeedDev& sd, const u32* cpos,
                                              const u8* cgs, u32 m, u32 room_l, u32 room_r, u32& el, u32& er) {
    const u32 L = sd.L;
    bool lj = true, rj = true;
    if (3 * L <= 64) {
        const u64 care = sd.mask_hi & 0x5555555555555555ull;
        u64 xa, xb;
        mismatch64(packed, gt, L, cpos, cgs, m, -(i64)L, xa, xb);
        lj = room_l >= L && !(map_at(xa, xb, 0) & care);
        rj = room_r >= L && !(map_at(xa, xb, 2 * L) & care);
        if (!lj) {
            u32 maxs = min(L, room_l), sdone = 0;
            for (u32 s = 1; s <= maxs; ++s) {
                if (map_at(xa, xb, L - s) & care) break;
                sdone = s;
            }
            el = sdone;
        }
        if (!rj) {
            u32 maxs = min(L, room_r), sdone = 0;
            for (u32 s = 1; s <= maxs; ++s) {
                if (map_at(xa, xb, L + s) & care) break;
                sdone = s;
            }
            er = sdone;
        }
    }
    if (lj) el = grow_thread(packed, gt, sd, cpos, cgs, m, -1, room_l);
    if (el == INF32) return false;
    if (rj) er = grow_thread(packed, gt, sd, cpos, cgs, m, +1, room_r);
    return er != INF32;
}

// ------------------------------------------------------------------------------------------------
// Work lists and their counters live in device memory and the convergence test of the resolve rounds
// is made on the device (one cooperative launch, grid-wide barriers between the phases), so the host
// never synchronises inside the resolve step.  All per-rep state is indexed by the rep's position i in
// the (colour, slot) order.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define DD_NT 256
#define DD_WIDE 2048 // index ranges longer than this are handled by one warp instead of one thread

// warp-aggregated append to a device work list (call with the whole warp converged)
__device__ __forceinline__ void wl_push(u32* list, u32* count, bool pred, u32 value) {
    u32 m = __ballot_sync(0xFFFFFFFFu, pred);
    if (!pred) return;
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    u32 base = 0;
    if (lane == leader) base = atomicAdd(count, (u32)__popc(m));
    base = __shfl_sync(m, base, leader);
    list[base + __popc(m & ((1u << lane) - 1))] = value;
}

// slot range [rlo, rhi) of all candidates whose first-genome position lies in the extent [gp - el, gp + er]
// (gp = global base index of the first component; two rank look-ups in the candidate bitmap)
__device__ __forceinline__ void extent_slots(const DedupArgs& a, u64 gp, u32 el, u32 er, u32& rlo, u32& rhi) {
    rlo = slot_rank(a.bitmap, a.bmrank, gp - el);
    rhi = slot_rank(a.bitmap, a.bmrank, gp + er + 1);
}

// Per-rep record in (colour, slot) order: x = colour:16 | slot:32 | hash[15:0], y = hash[47:16] | candidate:32
// (the colour IS hash[63:48], so the two words hold the whole group hash).  One 16-byte load per neighbour.
#define REC_COL(x) ((x) & 0xFFFF000000000000ull)
__device__ __forceinline__ bool rec_same_hash(const ulonglong2& p, const ulonglong2& q) {
    return ((p.x ^ q.x) & 0xFFFF00000000FFFFull) == 0 && (p.y >> 32) == (q.y >> 32);
}

// ---- per-rep records in (colour, slot) order: everything the extension and the resolve rounds need about a
// rep, gathered once by a plain throughput kernel so that those kernels start from one coalesced load
__global__ void __launch_bounds__(256) k_rep_setup(DedupArgs a, GenomeTable gt) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_rep) return;
    const u32 slot = (u32)a.s_key[i];
    const ulonglong2 r = a.slot_rec[2 * (size_t)slot], q = a.slot_rec[2 * (size_t)slot + 1];
    const u32 c = (u32)r.y, m = (u32)(r.y >> 32) & 0xFFu, g0 = (u32)(r.y >> 40) & 0xFFu, vg = (u32)(r.y >> 48) & 0xFFu;
    const u64 h = r.x & HASH_MASK;
    a.s_rec[i] = make_ulonglong2((h & 0xFFFF000000000000ull) | ((u64)slot << 16) | (h & 0xFFFFull), (((h >> 16) & 0xFFFFFFFFull) << 32) | c);
    a.s_cand[i] = c;
    a.s_h2[i] = q.x;
    a.minrank[i] = INF64; a.minrank[(size_t)a.n_rep + i] = INF64;
    a.reach[i] = 0;
    a.rstate[i] = (a.pre_drop && a.pre_drop[c]) ? 2 : 0;
    if (a.rows) a.xrec[i] = make_uint4(c, (u32)q.y, m | (g0 << 8) | (vg << 16), (u32)(q.y >> 32)); // owner side: k_extent_ranges follows
    else { // the extension (slot order) has run: slot range of the extent
        u32 rlo, rhi;
        extent_slots(a, gt.vbase[vg] + (u32)(q.y >> 32), a.ext_l[c], a.ext_r[c], rlo, rhi);
        a.rng_lo[i] = rlo; a.rng_hi[i] = rhi;
    }
}

// multi-GPU source side: extension record of every candidate, straight from the CSR (rep index = candidate)
__global__ void __launch_bounds__(256) k_cand_xrec(DedupArgs a, GenomeTable gt) {
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_cand) return;
    const u32 off = a.cand_off[c], m = a.cand_off[c + 1] - off;
    const u32 g0 = a.comp_gs[off] & 0x7F;
    const u32 vg = vgenome(gt, g0, a.comp_gs[off + 1] & 0x7F);
    a.xrec[c] = make_uint4(c, off, m | (g0 << 8) | (vg << 16), a.comp_pos[off]);
}
// multi-GPU owner side: the extents came with the rows; only the slot ranges are left to derive
__global__ void __launch_bounds__(256) k_extent_ranges(DedupArgs a, GenomeTable gt) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_rep) return;
    const uint4 q = a.xrec[i];
    const u64 e = a.rows[4 * (size_t)q.x + 3];
    const u32 el = (u32)e, er = (u32)(e >> 32); // (the owner never needs per-candidate extent arrays: the rows keep them)
    u32 rlo, rhi;
    extent_slots(a, gt.vbase[q.z >> 16] + q.w, el, er, rlo, rhi);
    a.rng_lo[i] = rlo; a.rng_hi[i] = rhi;
}

// ---- extend: a warp takes 32 reps.  Lane j owns rep j's state; the (rep, component) pairs of the warp
// are spread over all lanes, so every lane builds exactly one component-vs-first mismatch map per step
// whatever the multiplicities are; maps are OR-ed per rep in shared memory.  Rounds: centred chunk, then
// further chunks to the left, then to the right.  k_extend runs the first rounds over all reps; the reps
// that need more are compacted into a work list and taken up again, densely packed, by k_extend_more
// (a few launches); what is still unfinished then goes to the warp-per-rep kernel.
enum { ST_CENTER = 0, ST_LEFT = 1, ST_RIGHT = 2, ST_DONE = 3 };
struct XLane {
    u32 c, off, m, p0, g0, vg;  // candidate, its component rows, first component, virtual genome of the bitmap axis
    int st; bool rj;            // walk state; rj: the right side still has to be walked
    u32 b, el, er, room_l, room_r;
};

// The genome table arrives as a kernel argument, i.e. in the constant bank: gt.word_base[g] with a per-lane genome g is an
// indexed constant load that replays once per distinct g of the warp (a rep's pairs sit on ~20 lanes with ~20 genomes; the
// ncu source page had a quarter of the kernel's stall samples behind these loads).  The hot loop reads a shared-memory copy.
struct GShared { u32 wbase[MB_MAX_SEQ]; u32 len[MB_MAX_SEQ]; }; // (packed words of all genomes < 2^32: at most 2^31 bases)
__device__ __forceinline__ void gshared_fill(GShared& gs, const GenomeTable& gt) {
    for (u32 g = threadIdx.x; g < MB_MAX_SEQ; g += blockDim.x) { gs.wbase[g] = (u32)gt.word_base[g]; gs.len[g] = gt.len[g]; }
    __syncthreads();
}
// 64 oriented bases of component (g, pos) at match offset i0, as oriented_bases64 — with 32-bit index arithmetic (the
// offset may be negative: every genome is preceded by padding words), the L2 policy created once by the caller, and the
// reverse complement only computed when `any_rev` (warp-uniform: some lane of the warp holds a reverse component)
__device__ __forceinline__ void oriented_bases64_s(const u64* __restrict__ packed, const GShared& gs, u32 L, u32 g, u32 pos, bool rev, int i0, u64 pol,
                                                   bool any_rev, u64& a, u64& b) {
    const int t = (int)pos + (rev ? (int)L - 64 - i0 : i0);
    const u64* w = packed + (gs.wbase[g] + (u32)(t >> 5));
    const int sh = (t & 31) * 2;
    u64 w0 = ld_keep(w, pol), w1 = ld_keep(w + 1, pol), w2 = ld_keep(w + 2, pol);
    a = shl128_hi(w0, w1, sh); b = shl128_hi(w1, w2, sh);
    if (any_rev && rev) { const u64 fa = a; a = rc_word(b); b = rc_word(fa); }
}
__device__ __forceinline__ void seg_range_s(const GenomeTable& gt, const GShared& gs, u32 g, u32 p, u32& lo, u32& hi) {
    if (!gt.n_seg) { lo = 0; hi = gs.len[g]; }
    else seg_range(gt, g, p, lo, hi);
}

template <bool FIRST> // FIRST: round 0 also reduces the rooms of the components
__device__ __forceinline__ void extend_rounds(const DedupArgs& a, const GenomeTable& gt, const GShared& gsh, const SeedDev& sd, XLane& x, int max_rounds,
                                              u32 (*sMap)[4], u32 (*sRoom)[2], ulonglong2* sRepA, uint2* sRepB) {
    const int lane = threadIdx.x & 31;
    const u32 L = sd.L;
    const u64 pol = l2_keep_policy();
    const u32 per_chunk = (3 * L <= 65) ? 2 : 1;
    if (FIRST) {
        // component 0 is forward
        u32 lo0 = 0, hi0 = 0;
        if (x.st != ST_DONE) seg_range_s(gt, gsh, x.g0, x.p0, lo0, hi0);
        sRoom[lane][0] = x.st != ST_DONE ? x.p0 - lo0 : INF32;
        sRoom[lane][1] = x.st != ST_DONE ? hi0 - L - x.p0 : INF32;
    }
    for (int round = 0; round < max_rounds; ++round) {
        const bool want = x.st != ST_DONE;
        if (!__any_sync(0xFFFFFFFFu, want)) break;
        const int o_lo = x.st == ST_CENTER ? -(int)L : (int)chunk_lo(x.st == ST_LEFT ? -1 : +1, x.b, L);
        const u32 np = want ? x.m - 1 : 0;
        u32 incl = np;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += y;
        }
        const u32 start = incl - np, T = __shfl_sync(0xFFFFFFFFu, incl, 31);
        u64 a0 = 0, b0 = 0;
        if (want) oriented_bases64_s(a.packed, gsh, L, x.g0, x.p0, false, o_lo, pol, false, a0, b0);
        sMap[lane][0] = 0; sMap[lane][1] = 0; sMap[lane][2] = 0; sMap[lane][3] = 0;
        // what the lanes that work on a rep's pairs need about it: one 16-byte and one 8-byte shared load instead of six shuffles
        sRepA[lane] = make_ulonglong2(a0, b0);
        sRepB[lane] = make_uint2(x.off, (u32)o_lo);
        __syncwarp();
        for (u32 base = 0; base < T; base += 32) {
            const u32 p = base + lane;
            const bool act = p < T;
            u32 j = 0; // owner of pair p: the last lane whose first pair index is <= p
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                u32 sv = __shfl_sync(0xFFFFFFFFu, start, (j + step) & 31);
                if (j + step < 32 && sv <= p) j += step;
            }
            const u32 k = p - __shfl_sync(0xFFFFFFFFu, start, j) + 1;
            const uint2 rb = sRepB[j];
            const u32 offj = rb.x;
            const int oj = (int)rb.y;
            u32 gs = 0, pk = 0;
            if (act) { gs = a.comp_gs[offj + k]; pk = a.comp_pos[offj + k]; }
            const bool any_rev = __any_sync(0xFFFFFFFFu, (gs & 0x80u) != 0);
            if (act) {
                const ulonglong2 ra = sRepA[j];
                u64 A, B;
                oriented_bases64_s(a.packed, gsh, L, gs & 0x7F, pk, gs & 0x80, oj, pol, any_rev, A, B);
                u64 xa = spread_nz(A ^ ra.x), xb = spread_nz(B ^ ra.y);
                if (xa >> 32) atomicOr(&sMap[j][0], (u32)(xa >> 32));
                if ((u32)xa) atomicOr(&sMap[j][1], (u32)xa);
                if (xb >> 32) atomicOr(&sMap[j][2], (u32)(xb >> 32));
                if ((u32)xb) atomicOr(&sMap[j][3], (u32)xb);
                if (FIRST && round == 0) {
                    u32 lo, hi;
                    seg_range_s(gt, gsh, gs & 0x7F, pk, lo, hi);
                    u32 lroom = pk - lo, rroom = hi - L - pk;
                    bool rev = gs & 0x80;
                    atomicMin(&sRoom[j][0], rev ? rroom : lroom);
                    atomicMin(&sRoom[j][1], rev ? lroom : rroom);
                }
            }
        }
        __syncwarp();
        if (FIRST && round == 0) { x.room_l = sRoom[lane][0]; x.room_r = sRoom[lane][1]; }
        if (want) {
            const u64 xa = ((u64)sMap[lane][0] << 32) | sMap[lane][1], xb = ((u64)sMap[lane][2] << 32) | sMap[lane][3];
            u64 bhi, blo;
            window_map(xa, xb, sd, bhi, blo); // one call site for the three states: the lanes of a warp differ in state
            if (x.st == ST_CENTER) {
                // chunk index i <-> match offset i - L: first jump windows at 0 (left) and 2L (right), single-step
                // windows s at L - s (left, descending from L - 1) and L + s (right, ascending from L + 1)
                const bool lj = x.room_l >= L && !win_bad(bhi, blo, 0);
                x.rj = x.room_r >= L && !win_bad(bhi, blo, 2 * L);
                if (!lj) x.el = min(good_down(bhi, blo, L - 1), min(L, x.room_l));
                if (!x.rj) x.er = min(good_up(bhi, blo, L + 1), min(L, x.room_r));
                x.b = 1; // the first jump of the side walked next is known to succeed
                x.st = lj ? ST_LEFT : (x.rj ? ST_RIGHT : ST_DONE);
            } else {
                const bool left = x.st == ST_LEFT;
                const u32 room = left ? x.room_l : x.room_r;
                u32 out = 0;
                if (walk_chunk(bhi, blo, left ? -1 : +1, L, room, room / L, per_chunk, x.b, out)) {
                    if (left) { x.el = out; x.b = (3 * L <= 64) ? 1 : 0; x.st = x.rj ? ST_RIGHT : ST_DONE; }
                    else { x.er = out; x.st = ST_DONE; }
                }
            }
        }
        __syncwarp();
    }
}

// done: extents + slot range; unfinished: state saved, rep appended to `out_list` (or to the long list when `last`)
__device__ __forceinline__ void extend_finish(const DedupArgs& a, const GenomeTable& gt, bool valid, u32 i, const XLane& x, u32* out_list,
                                              u32* out_count, bool last) {
    const bool more = valid && x.st != ST_DONE;
    if (valid && !more) {
        a.ext_l[x.c] = x.el; a.ext_r[x.c] = x.er;
        if (a.bitmap) { // (the multi-GPU source side extends without the slot axis: the owner derives the ranges)
            u32 rlo, rhi;
            extent_slots(a, gt.vbase[x.vg] + x.p0, x.el, x.er, rlo, rhi);
            a.rng_lo[i] = rlo; a.rng_hi[i] = rhi;
        }
    }
    if (more && !last) {
        a.xstate[i] = make_uint4((u32)x.st | (x.rj ? 4u : 0u), x.b, x.room_l, x.room_r);
        a.ext_l[x.c] = x.el; a.ext_r[x.c] = x.er;
    }
    wl_push(last ? a.wl_long : out_list, last ? a.ctr + 6 : out_count, more, i);
}

#ifndef DD_EXT_FIRST_ROUNDS
#define DD_EXT_FIRST_ROUNDS 6
#endif
#ifndef DD_EXT_MORE_ROUNDS
#define DD_EXT_MORE_ROUNDS 8
#endif

#ifndef DD_EXT_MINB
#define DD_EXT_MINB 5 // 48 registers, no spills.  6 / 8 blocks (40 / 32 registers, a few spilled words) measured within 0.1 ms on C5: occupancy is
#endif                // not what k_extend waits for; 1 lets ptxas take 68 registers and costs 1.2 ms

__global__ void __launch_bounds__(DD_NT, DD_EXT_MINB) k_extend(DedupArgs a, GenomeTable gt, SeedDev sd, u32* out_list, u32* out_count) {
    __shared__ u32 sMap[DD_NT / 32][32][4];
    __shared__ u32 sRoom[DD_NT / 32][32][2];
    __shared__ ulonglong2 sRepA[DD_NT / 32][32];
    __shared__ uint2 sRepB[DD_NT / 32][32];
    __shared__ GShared gsh;
    gshared_fill(gsh, gt);
    const int warp = threadIdx.x >> 5;
    const u32 i = blockIdx.x * DD_NT + threadIdx.x;
    const bool valid = i < a.n_rep;
    const u32 L = sd.L;
    XLane x{0, 0, 1, 0, 0, 0, ST_DONE, true, 0, 0, 0, INF32, INF32};
    if (valid) {
        uint4 q = a.xrec[i];
        x.c = q.x; x.off = q.y; x.m = q.z & 0xFFu; x.g0 = (q.z >> 8) & 0xFFu; x.vg = q.z >> 16; x.p0 = q.w;
        x.st = 3 * L <= 64 ? ST_CENTER : ST_LEFT;
    }
    if (L > 32) { // uniform: window-by-window warp kernel only
        wl_push(a.wl_long, a.ctr + 6, valid, i);
        return;
    }
    extend_rounds<true>(a, gt, gsh, sd, x, DD_EXT_FIRST_ROUNDS, sMap[warp], sRoom[warp], sRepA[warp], sRepB[warp]);
    extend_finish(a, gt, valid, i, x, out_list, out_count, false);
}

// the unfinished reps of the previous launch, 32 per warp
__global__ void __launch_bounds__(DD_NT) k_extend_more(DedupArgs a, GenomeTable gt, SeedDev sd, const u32* in_list, const u32* in_count, u32* out_list,
                                                       u32* out_count, int last) {
    __shared__ u32 sMap[DD_NT / 32][32][4];
    __shared__ ulonglong2 sRepA[DD_NT / 32][32];
    __shared__ uint2 sRepB[DD_NT / 32][32];
    __shared__ GShared gsh;
    gshared_fill(gsh, gt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 n = *in_count;
    const u32 gwarp = (blockIdx.x * DD_NT + threadIdx.x) >> 5, nwarps = (gridDim.x * DD_NT) >> 5;
    for (u32 base = gwarp * 32; base < n; base += nwarps * 32) {
        const bool valid = base + lane < n;
        u32 i = 0;
        XLane x{0, 0, 1, 0, 0, 0, ST_DONE, true, 0, 0, 0, INF32, INF32};
        if (valid) {
            i = in_list[base + lane];
            uint4 q = a.xrec[i];
            x.c = q.x; x.off = q.y; x.m = q.z & 0xFFu; x.g0 = (q.z >> 8) & 0xFFu; x.vg = q.z >> 16; x.p0 = q.w;
            uint4 s = a.xstate[i];
            x.st = (int)(s.x & 3u); x.rj = (s.x & 4u) != 0; x.b = s.y; x.room_l = s.z; x.room_r = s.w;
            x.el = a.ext_l[x.c]; x.er = a.ext_r[x.c];
        }
        extend_rounds<false>(a, gt, gsh, sd, x, DD_EXT_MORE_ROUNDS, sMap[warp], nullptr, sRepA[warp], sRepB[warp]);
        extend_finish(a, gt, valid, i, x, out_list, out_count, last != 0);
    }
}

__global__ void __launch_bounds__(DD_NT) k_extend_long(DedupArgs a, GenomeTable gt, SeedDev sd) {
    const int lane = threadIdx.x & 31;
    const u32 gwarp = (blockIdx.x * DD_NT + threadIdx.x) >> 5, nwarps = (gridDim.x * DD_NT) >> 5;
    const u32 n = a.ctr[6];
    for (u32 t = gwarp; t < n; t += nwarps) {
        u32 i = a.wl_long[t];
        u32 c = a.xrec[i].x;
        u32 off = a.cand_off[c], m = a.cand_off[c + 1] - off;
        const u32* cpos = a.comp_pos + off;
        const u8* cgs = a.comp_gs + off;
        u32 el, er;
        if (sd.L > 32) extend_candidate_warp(a.packed, gt, sd, cpos, cgs, m, el, er);
        else {
            u32 room_l = INF32, room_r = INF32;
            for (u32 k = lane; k < m; k += 32) candidate_room(gt, sd.L, cpos, cgs, k, room_l, room_r);
            room_l = __reduce_min_sync(0xFFFFFFFFu, room_l);
            room_r = __reduce_min_sync(0xFFFFFFFFu, room_r);
            el = grow_warp(a.packed, gt, sd, cpos, cgs, m, -1, room_l);
            er = grow_warp(a.packed, gt, sd, cpos, cgs, m, +1, room_r);
        }
        if (lane == 0) {
            a.ext_l[c] = el; a.ext_r[c] = er;
            if (a.bitmap) {
                u32 rlo, rhi;
                extent_slots(a, gt.vbase[vgenome(gt, cgs[0] & 0x7F, cgs[1] & 0x7F)] + cpos[0], el, er, rlo, rhi);
                a.rng_lo[i] = rlo; a.rng_hi[i] = rhi;
            }
        }
    }
}

// ---- long extensions, shared inside a group -----------------------------------------------------------------------
// The sequential reference extends only candidates no accepted match contains; here every rep is extended first, so a
// long window-consistent run that holds thousands of reps of ONE group (two identical genomes, a recent duplication) would
// be walked from end to end once per rep (the "cliff").  But two reps of one group whose seeds lie k L bases
// apart inside one run walk the same lattice of jump windows: they stop at the same windows on both sides and so reach
// the SAME extent.  The reps that are still unfinished after the bounded rounds are therefore sorted by (hash of group and
// position mod L, rep index); one warp takes a class, walks the first member to the end, and every later member of the
// same group (hashes, then the exact comparison) whose seed lies inside that extent takes it over; a member outside
// starts a new walk.  L classes per group instead of one walk per rep.
__global__ void __launch_bounds__(256) k_long_keys(DedupArgs a, u32 L, u64* __restrict__ keys) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.ctr[6]) return;
    const u32 i = a.wl_long[t];
    const uint4 q = a.xrec[i];
    u64 h = a.ghash[q.x] ^ ((u64)(q.w % L) * 0x9E3779B97F4A7C15ull);
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
    keys[t] = (h << 32) | i;
}

__global__ void __launch_bounds__(DD_NT) k_extend_long_classes(DedupArgs a, GenomeTable gt, SeedDev sd, const u64* __restrict__ keys, u32 n) {
    const int lane = threadIdx.x & 31;
    const u32 t0 = (blockIdx.x * DD_NT + threadIdx.x) >> 5;
    if (t0 >= n) return;
    const u32 cls = (u32)(keys[t0] >> 32);
    if (t0 > 0 && (u32)(keys[t0 - 1] >> 32) == cls) return; // not the first member of its class
    const u32 L = sd.L;
    // the last finished walk of this class: its rep, seed position, hashes, extent [lend, rend), and the lowest candidate
    // index among the members that share it so far (members dropped before the de-dup do not count: they claim nothing)
    bool has = false;
    u32 c_lead = 0, x_lead = 0, lend = 0, rend = 0, c_min = INF32;
    u64 h1_lead = 0, h2_lead = 0;
    for (u32 base = t0; base < n; base += 32) { // 32 members at a time, one per lane
        const u32 t = base + lane;
        const u64 key = t < n ? keys[t] : 0;
        const bool mine = t < n && (u32)(key >> 32) == cls;
        const u32 i = (u32)key;
        u32 c = 0, x = 0, vg = 0;
        u64 h1 = 0, h2 = 0;
        bool pd = false;
        if (mine) {
            const uint4 q = a.xrec[i];
            c = q.x; x = q.w; vg = q.z >> 16;
            h1 = a.ghash[c]; h2 = a.ghash2[c];
            pd = a.pre_drop && a.pre_drop[c];
        }
        bool pending = mine;
        for (;;) {
            bool reuse = false;
            if (pending && has && x >= lend && x + L <= rend && (x >= x_lead ? x - x_lead : x_lead - x) % L == 0 && h1 == h1_lead && h2 == h2_lead)
                reuse = same_group(a, c, c_lead);
            c_min = min(c_min, __reduce_min_sync(0xFFFFFFFFu, reuse && !pd ? c : INF32));
            if (reuse) {
                const u32 el = x - lend, er = rend - (x + L);
                a.ext_l[c] = el; a.ext_r[c] = er;
                if (a.shadow && c > c_min) a.shadow[c] = 1;
                if (a.bitmap) {
                    u32 rlo, rhi;
                    extent_slots(a, gt.vbase[vg] + x, el, er, rlo, rhi);
                    a.rng_lo[i] = rlo; a.rng_hi[i] = rhi;
                }
                pending = false;
            }
            const unsigned rest = __ballot_sync(0xFFFFFFFFu, pending);
            if (!rest) break;
            // the first member left starts a new walk, by the whole warp
            const int src = __ffs(rest) - 1;
            const u32 wc = __shfl_sync(0xFFFFFFFFu, c, src), wx = __shfl_sync(0xFFFFFFFFu, x, src), wi = __shfl_sync(0xFFFFFFFFu, i, src);
            const bool wpd = __shfl_sync(0xFFFFFFFFu, pd ? 1 : 0, src) != 0;
            const u32 off = a.cand_off[wc], m = a.cand_off[wc + 1] - off;
            const u32* cpos = a.comp_pos + off;
            const u8* cgs = a.comp_gs + off;
            u32 el, er;
            if (L > 32) extend_candidate_warp(a.packed, gt, sd, cpos, cgs, m, el, er);
            else {
                u32 room_l = INF32, room_r = INF32;
                for (u32 k = lane; k < m; k += 32) candidate_room(gt, L, cpos, cgs, k, room_l, room_r);
                room_l = __reduce_min_sync(0xFFFFFFFFu, room_l);
                room_r = __reduce_min_sync(0xFFFFFFFFu, room_r);
                el = grow_warp(a.packed, gt, sd, cpos, cgs, m, -1, room_l);
                er = grow_warp(a.packed, gt, sd, cpos, cgs, m, +1, room_r);
            }
            has = true; c_lead = wc; x_lead = wx; lend = wx - el; rend = wx + L + er;
            h1_lead = __shfl_sync(0xFFFFFFFFu, h1, src); h2_lead = __shfl_sync(0xFFFFFFFFu, h2, src);
            c_min = wpd ? INF32 : wc;
            if (lane == src) {
                a.ext_l[wc] = el; a.ext_r[wc] = er;
                if (a.bitmap) {
                    u32 rlo, rhi;
                    extent_slots(a, gt.vbase[vgenome(gt, cgs[0] & 0x7F, cgs[1] & 0x7F)] + cpos[0], el, er, rlo, rhi);
                    a.rng_lo[wi] = rlo; a.rng_hi[wi] = rhi;
                }
                pending = false;
            }
            __syncwarp();
        }
        if (!__all_sync(0xFFFFFFFFu, mine)) break; // the class ended inside this batch
    }
}

// Rep state: 0 undecided, 1 accepted, 2 dropped; flag bit 6: lives on the wide list.
#define RS_MASK 0x0Fu
#define RS_WIDE 0x40u
#ifndef DD_WALK_MAX
#define DD_WALK_MAX 192 // neighbours one thread visits per direction before the rep is handed to a warp
#endif

// Visit the neighbours of rep i in the (colour, slot) order.  `first`/`stride` = 0/1 for one thread, lane/32 for
// a warp.
// CLAIM: over the same-colour reps whose slot lies in i's extent [rlo, rhi): every undecided higher-rank rep of
//   the same group hash gets atomicMin(rank:index of i) and atomicMax(index distance).  A claim only delays its
//   target, so the 64-bit hash is enough: a false claim lapses when the claimer is decided.
// PULL: over the reps within `reach[i]` indices (no claimer of i was ever further away): is there an ACCEPTED
//   lower-rank rep of i's group (hash + second hash) whose extent contains i's slot?  Then i is contained (D16).
// Returns the number of hits, or -1 when a single thread ran out of its walk budget.
template <bool PULL, bool WARP>
__device__ __forceinline__ int walk_neighbours(const DedupArgs& a, u32 i, const ulonglong2& me, u32 rlo, u32 rhi, u64* __restrict__ mr) {
    const u64 klo = REC_COL(me.x) | ((u64)rlo << 16), khi = REC_COL(me.x) | ((u64)rhi << 16);
    const u32 c = (u32)me.y, myslot = (u32)(me.x >> 16);
    const u32 first = WARP ? (threadIdx.x & 31) : 0, stride = WARP ? 32 : 1;
    const u32 reach = PULL ? a.reach[i] : 0;
    const u64 myh2 = PULL ? a.s_h2[i] : 0;
    int found = 0;
    for (int dir = 0; dir < 2; ++dir) {
        u32 steps = 0;
        for (u32 d = 1 + first;; d += stride) {
            bool in = dir == 0 ? (u64)i + d < a.n_rep : d <= i;
            if (PULL) in = in && d <= reach;
            ulonglong2 r = make_ulonglong2(0, 0);
            const u32 t = dir == 0 ? i + d : i - d;
            if (in) {
                r = a.s_rec[t];
                if (!PULL) in = dir == 0 ? r.x < khi : r.x >= klo;
            }
            if (in && rec_same_hash(me, r)) {
                if (PULL) {
                    if ((u32)r.y < c && (a.rstate[t] & RS_MASK) == 1 && a.s_h2[t] == myh2 && myslot >= a.rng_lo[t] && myslot < a.rng_hi[t] && groups_equal(a, c, (u32)r.y)) ++found;
                } else if ((u32)r.y > c && (a.rstate[t] & RS_MASK) == 0) {
                    atomicMin((unsigned long long*)&mr[t], ((u64)c << 32) | i);
                    atomicMax(&a.reach[t], d);
                    ++found;
                }
            }
            if (WARP) { if (!__any_sync(0xFFFFFFFFu, in)) break; }
            else {
                if (!in) break;
                if (++steps > DD_WALK_MAX) return -1;
            }
        }
    }
    return found;
}

__device__ __forceinline__ u64 gtimer() { u64 t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// optional phase trace (a.trace != null): thread 0 appends (tag, globaltimer ns) pairs after each barrier
#define DD_TRACE(tag) do { if (t0 && a.trace) { u64 k = a.trace[0]; if (k < 4000) { a.trace[2 + 2 * k] = (tag); a.trace[3 + 2 * k] = gtimer(); a.trace[0] = k + 1; } } } while (0)

// One decide step of an undecided rep i (claims of this round are complete): unclaimed -> accepted; the lowest
// claimer is itself unclaimed (so it is accepted in this very phase) and of the same group -> contained, dropped;
// otherwise still undecided.  Returns true when i stays undecided.
__device__ __forceinline__ bool decide_rep(const DedupArgs& a, u32 i, const u64* __restrict__ mr_cur, u64* __restrict__ mr_nxt) {
    const u64 m = mr_cur[i];
    if (m == INF64) { a.rstate[i] = (a.rstate[i] & RS_WIDE) | 1; return false; }
    const u32 mi = (u32)m;
    const u32 ms = a.rstate[mi] & RS_MASK; // 0, or 1 when that rep has just been decided in this phase
    if (ms <= 1 && mr_cur[mi] == INF64 && a.s_h2[mi] == a.s_h2[i] && groups_equal(a, a.s_cand[mi], a.s_cand[i])) { a.rstate[i] = (a.rstate[i] & RS_WIDE) | 2; return false; }
    mr_nxt[i] = INF64;
    return true;
}

// ctr layout (u32): [0..2] narrow-list counters (rotating; round 0 takes every rep), [3..5] wide-list
// counters (rotating), [6] long list, [9] rounds, [10] wide items.  A counter is reset one round before
// it is written and never while it may still be read.
__global__ void __launch_bounds__(DD_NT) k_resolve(DedupArgs a) {
    cg::grid_group grid = cg::this_grid();
    const bool t0 = blockIdx.x == 0 && threadIdx.x == 0;
    const u32 gtid = blockIdx.x * DD_NT + threadIdx.x, gsz = gridDim.x * DD_NT;
    const int lane = threadIdx.x & 31;
    const u32 gwarp = gtid >> 5, nwarps = gsz >> 5;
    u32* ctr = a.ctr;
    u32* nl[3] = {a.wl0, a.wl1, a.wl2};
    u32* wd[3] = {a.wd0, a.wd1, a.wd2};
    DD_TRACE(1);
    for (u32 r = 0;; ++r) {
        const u32 cur = r % 3, nxt = (r + 1) % 3, spare = (r + 2) % 3;
        u64* mr_cur = a.minrank + (size_t)(r & 1) * a.n_rep;
        u64* mr_nxt = a.minrank + (size_t)((r + 1) & 1) * a.n_rep;
        if (t0) { ctr[spare] = 0; ctr[3 + spare] = 0; ctr[9] += 1; }
        const u32 n_narrow = r == 0 ? a.n_rep : ctr[cur];
        // ---- round 0: claim, one thread per rep; reps with too many neighbours go to the wide list.
        // Later rounds: few reps are left, each takes a whole warp (pull, then claim).
        if (r == 0) {
            for (u32 i = gtid; i < n_narrow; i += gsz) {
                u32 rlo = a.rng_lo[i], rhi = a.rng_hi[i];
                if (rhi - rlo < 2) continue; // alone in its extent
                if (a.pre_drop && a.rstate[i] != 0) continue; // dropped before the de-dup: claims nothing
                ulonglong2 me = a.s_rec[i];
                // a lower-rank rep of the same group with the same extent makes every claim of this one, with a lower value;
                // this rep is itself claimed, so it cannot be accepted in this round, and claims again in the next if it is left
                if (a.shadow && a.shadow[(u32)me.y]) continue;
                if (walk_neighbours<false, false>(a, i, me, rlo, rhi, mr_cur) < 0) { a.rstate[i] = RS_WIDE; wd[cur][atomicAdd(ctr + 3 + cur, 1u)] = i; }
            }
        } else {
            for (u32 t = gwarp; t < n_narrow; t += nwarps) {
                u32 i = nl[cur][t];
                ulonglong2 me = a.s_rec[i];
                u32 rlo = a.rng_lo[i], rhi = a.rng_hi[i];
                int f = walk_neighbours<true, true>(a, i, me, rlo, rhi, mr_cur);
                if (__any_sync(0xFFFFFFFFu, f > 0)) { if (lane == 0) a.rstate[i] = 2; continue; } // contained in an accepted match
                if (rhi - rlo >= 2) walk_neighbours<false, true>(a, i, me, rlo, rhi, mr_cur);
            }
        }
        grid.sync(); // the wide list of this round is complete only now
        if (t0 && r == 0) ctr[10] = ctr[3];
        const u32 n_wide = ctr[3 + cur];
        // ---- pull + claim, wide: one warp per rep
        for (u32 t = gwarp; t < n_wide; t += nwarps) {
            u32 i = wd[cur][t];
            ulonglong2 me = a.s_rec[i];
            u32 rlo = a.rng_lo[i], rhi = a.rng_hi[i];
            if (r > 0) {
                int f = walk_neighbours<true, true>(a, i, me, rlo, rhi, mr_cur);
                if (__any_sync(0xFFFFFFFFu, f > 0)) { if (lane == 0) a.rstate[i] = RS_WIDE | 2; continue; }
            }
            walk_neighbours<false, true>(a, i, me, rlo, rhi, mr_cur);
        }
        if (n_wide) grid.sync(); // uniform (read after the barrier above); without wide reps all claims are already complete
        DD_TRACE(5);
        // ---- decide, narrow
        for (u32 base = blockIdx.x * DD_NT; base < n_narrow; base += gsz) {
            u32 t = base + threadIdx.x;
            bool keep = false;
            u32 i = 0;
            if (t < n_narrow) {
                i = r == 0 ? t : nl[cur][t];
                if (a.rstate[i] == 0) keep = decide_rep(a, i, mr_cur, mr_nxt); // undecided and not on the wide list
            }
            wl_push(nl[nxt], ctr + nxt, keep, i);
        }
        // ---- decide, wide
        for (u32 t = gwarp * 32 + lane; t < n_wide; t += nwarps * 32) {
            u32 i = wd[cur][t];
            if ((a.rstate[i] & RS_MASK) == 0 && decide_rep(a, i, mr_cur, mr_nxt)) wd[nxt][atomicAdd(ctr + 3 + nxt, 1u)] = i;
        }
        grid.sync();
        DD_TRACE(6);
        if (ctr[nxt] + ctr[3 + nxt] == 0) break;
    }
}

void launch_slot_scatter(const DedupArgs& a, const GenomeTable& gt, cudaStream_t st) {
    if (a.n_cand) k_slot_scatter<<<div_up(a.n_cand, 256), 256, 0, st>>>(a, gt);
}
u32 chain_tile() { return CH_TILE; }
void launch_chains(const DedupArgs& a, u64* status_fwd, u32* ticket_fwd, u64* status_bwd, u32* ticket_bwd, cudaStream_t st) {
    if (a.n_cand == 0) return;
    u32 tiles = div_up(a.n_cand, CH_TILE);
    k_chain<false><<<tiles, CH_NT, 0, st>>>(a, status_fwd, ticket_fwd);
    k_chain<true><<<tiles, CH_NT, 0, st>>>(a, status_bwd, ticket_bwd);
}
void launch_rep_keys(const DedupArgs& a, u64* skey, cudaStream_t st, const u32* cellR, const u32* cellNew, u32 V) {
    if (a.n_cand) k_rep_keys<<<div_up(a.n_cand, 256), 256, 0, st>>>(a, skey, cellR, cellNew, V);
}
u32 extension_cells(u32 V, u64 maxlen) { return V * (u32)((maxlen >> XO_SHIFT) + 1); }
void launch_extension_cells(const DedupArgs& a, const GenomeTable& gt, u32 V, u64 maxlen, u64 axis_end, u32* cellR, u32* cellNew, cudaStream_t st) {
    const u32 NB = (u32)((maxlen >> XO_SHIFT) + 1), cells = V * NB;
    k_cell_bounds<<<div_up(cells, 256), 256, 0, st>>>(a, gt, V, NB, axis_end, cellR, cellNew);
    k_cell_scan<<<1, 1024, 0, st>>>(cellNew, cells);
}
// k_extend over all reps, then DD_EXT_MORE launches over the shrinking list of unfinished reps (ping-pong lists
// wd0 / wd1 with counters ctr[12] / ctr[13]), then the warp-per-rep kernel for what is left
#ifndef DD_EXT_MORE
#define DD_EXT_MORE 2
#endif
#ifndef DD_MORE_BLOCKS
#define DD_MORE_BLOCKS 4 // grid of the continuation kernels, blocks per SM
#endif
#ifndef DD_LONG_BLOCKS
#define DD_LONG_BLOCKS 4
#endif
int extend_launches() { return 1 + DD_EXT_MORE; }
void launch_rep_setup(const DedupArgs& a, const GenomeTable& gt, cudaStream_t st) {
    if (a.n_rep) k_rep_setup<<<div_up(a.n_rep, 256), 256, 0, st>>>(a, gt);
}
void launch_cand_xrec(const DedupArgs& a, const GenomeTable& gt, cudaStream_t st) {
    if (a.n_cand) k_cand_xrec<<<div_up(a.n_cand, 256), 256, 0, st>>>(a, gt);
}
void launch_extent_ranges(const DedupArgs& a, const GenomeTable& gt, cudaStream_t st) {
    if (a.n_rep) k_extent_ranges<<<div_up(a.n_rep, 256), 256, 0, st>>>(a, gt);
}
void launch_extend(const DedupArgs& a, const GenomeTable& gt, const SeedDev& sd, cudaStream_t st) {
    if (a.n_rep == 0) return;
    u32* list[2] = {a.wd0, a.wd1};
    u32* cnt[2] = {a.ctr + 12, a.ctr + 13};
    k_extend<<<div_up(a.n_rep, DD_NT), DD_NT, 0, st>>>(a, gt, sd, list[0], cnt[0]);
    for (int r = 0; r < DD_EXT_MORE; ++r) {
        int in = r & 1, out = in ^ 1;
        cudaMemsetAsync(cnt[out], 0, 4, st);
        k_extend_more<<<148 * DD_MORE_BLOCKS, DD_NT, 0, st>>>(a, gt, sd, list[in], cnt[in], list[out], cnt[out], r == DD_EXT_MORE - 1);
    }
}
// what the bounded rounds left unfinished (a.ctr[6] reps on a.wl_long): one warp per rep ...
void launch_extend_long(const DedupArgs& a, const GenomeTable& gt, const SeedDev& sd, cudaStream_t st) {
    if (a.n_rep) k_extend_long<<<148 * DD_LONG_BLOCKS, DD_NT, 0, st>>>(a, gt, sd);
}
// ... or, for long lists, classes of reps that share their walks: keys (class hash << 32 | rep), sorted by the caller
void launch_long_keys(const DedupArgs& a, u32 L, u32 n_long, u64* keys, cudaStream_t st) {
    if (n_long) k_long_keys<<<div_up(n_long, 256), 256, 0, st>>>(a, L, keys);
}
void launch_extend_long_classes(const DedupArgs& a, const GenomeTable& gt, const SeedDev& sd, const u64* sorted_keys, u32 n_long, cudaStream_t st) {
    if (n_long) k_extend_long_classes<<<div_up((u64)n_long * 32, DD_NT), DD_NT, 0, st>>>(a, gt, sd, sorted_keys, n_long);
}
cudaError_t launch_resolve(const DedupArgs& a, cudaStream_t st) {
    if (a.n_rep == 0) return cudaSuccess;
    static int grid_blocks_of[64] = {}; // per device: co-resident blocks of the cooperative launch
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int local_blocks = 0;
    int& grid_blocks = (dev >= 0 && dev < 64) ? grid_blocks_of[dev] : local_blocks;
    if (grid_blocks == 0) {
        int sms = 0, per_sm = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_resolve, DD_NT, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        grid_blocks = sms * per_sm;
    }
    DedupArgs aa = a;
    void* args[] = {&aa};
    return cudaLaunchCooperativeKernel((const void*)k_resolve, dim3(grid_blocks), dim3(DD_NT), args, 0, st);
}

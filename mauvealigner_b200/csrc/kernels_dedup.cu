// kernels_dedup.cu — a9-a11 of SURVEY.md §8a: MemHash::AddHashEntry (containment de-dup in ascending
// seed order) and MatchFinder::ExtendMatch (ungapped extension), both absent from /root/reference
// (libMems); semantics = SURVEY.md Appendix A D13-D16 and the A.2 pseudo-code.  Reached in the
// reference from /root/reference/src/UniqueMatchFinder.cpp:58 (HashMatch(unique_list)).
//
// The reference decides candidates one at a time in ascending seed order: a candidate is dropped
// iff an already ACCEPTED extended match on its diagonal contains it, otherwise it is extended and
// accepted.  Here candidates (already in ascending seed order = rank) are taken in doubling batches:
//   begin    every not-yet-covered candidate of the batch is extended (pure function of the genomes)
//   claim    every undecided candidate writes atomicMin(rank) on the slots of all same-group
//            candidates its extent contains
//   decide   covered -> dropped;  min claimer == self -> accepted, marks its slots covered;
//            min claimer dropped -> reset slot and retry
// until the batch has no undecided candidate.  The fix-point equals the sequential result because a
// candidate is accepted exactly when every lower-rank container of it has been dropped.
// Slots: candidates ordered by (first genome, position) via a bitmap + popcount ranks, so "all
// candidates inside an extent" is a contiguous slot range found in O(1).
// Groups: candidates with the same genome set, strands and diagonal (D16) get the same exact group id
// through a hash table whose hits are verified component by component, so the walks compare integers.
#include "common.cuh"
#include "kernels.h"

__device__ __forceinline__ u32 slot_rank(const u64* __restrict__ bitmap, const u32* __restrict__ bmrank, u64 gp) {
    u64 w = bitmap[gp >> 6];
    return bmrank[gp >> 6] + (u32)__popcll(w & ((1ull << (gp & 63)) - 1));
}

// D16 group test, exact: same genome set, same strands, same diagonal
__device__ __forceinline__ bool same_group(const DedupArgs& a, u32 j, u32 e) {
    u32 offj = a.cand_off[j], offe = a.cand_off[e];
    u32 m = a.cand_off[e + 1] - offe;
    if (a.cand_off[j + 1] - offj != m) return false;
    u32 xj = a.comp_pos[offj], xe = a.comp_pos[offe];
    for (u32 k = 0; k < m; ++k) {
        u8 gj = a.comp_gs[offj + k], ge = a.comp_gs[offe + k];
        if (gj != ge) return false;
        u32 pj = a.comp_pos[offj + k], pe = a.comp_pos[offe + k];
        if (ge & 0x80) { if (pj + xj != pe + xe) return false; }
        else if (pj - xj != pe - xe) return false;
    }
    return true;
}

// exact group ids: open-addressing table keyed by the group hash; a hit counts only after the full
// component-wise comparison with the representative, so hash collisions just probe on.
__global__ void __launch_bounds__(256) k_group_ids(DedupArgs a) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_cand) return;
    u64 h = a.ghash[c];
    u32 slot = (u32)(h ^ (h >> 32)) & a.gid_table_mask;
    while (true) {
        u32 old = atomicCAS(&a.gid_table[slot], 0u, c + 1);
        if (old == 0) { a.gid[c] = c; return; }
        u32 r = old - 1;
        if (a.ghash[r] == h && same_group(a, r, c)) { a.gid[c] = r; return; }
        slot = (slot + 1) & a.gid_table_mask;
    }
}

// Slot order = (group id, position in the first genome): the candidates of one group lie next to each
// other, ordered along their diagonal, so "everything of my group inside my extent" is a short contiguous
// run of slots around my own.  Step 1 lists (group id, candidate) in (first genome, position) order
// through the bitmap ranks; a stable radix sort by group id (driver) then yields the final order.
__global__ void __launch_bounds__(256) k_slot_keys(DedupArgs a, GenomeTable gt, u64* __restrict__ skey, u64* __restrict__ sval) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_cand) return;
    u32 off = a.cand_off[c];
    u32 g = a.comp_gs[off] & 0x7F;
    u64 gp = gt.base_base[g] + a.comp_pos[off];
    u32 s = slot_rank(a.bitmap, a.bmrank, gp);
    skey[s] = a.gid[c];
    sval[s] = c;
}
__global__ void __launch_bounds__(256) k_slot_finish(DedupArgs a, const u64* __restrict__ skey, const u64* __restrict__ sval) {
    u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_cand) return;
    u32 c = (u32)sval[s];
    a.slot_of[c] = s;
    a.cand_at[s] = c;
    a.slot_gid[s] = (u32)skey[s];
    a.slot_x[s] = a.comp_pos[a.cand_off[c]];
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

// Warp version: number of consecutive steps t = 1..maxcount whose windows agree across all components.
// dir = -1: grow left in match coordinates, +1: grow right.  Offset of step t: o0 + t*stride.
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
    u32 p = cpos[k], len = gt.len[gs & 0x7F];
    u32 lroom = p, rroom = len - L - p;
    bool rev = gs & 0x80;
    room_l = min(room_l, rev ? rroom : lroom);
    room_r = min(room_r, rev ? lroom : rroom);
}

// D14: four phases.  One warp per candidate (long matches, many genomes, L > 32).
__device__ void extend_candidate_warp(const u64* __restrict__ packed, const GenomeTable& gt, const SeedDev& sd, const u32* cpos,
                                      const u8* cgs, u32 m, u32& ext_l, u32& ext_r) {
    const int lane = threadIdx.x & 31;
    const u32 L = sd.L;
    u32 room_l = 0xFFFFFFFFu, room_r = 0xFFFFFFFFu;
    for (u32 k = lane; k < m; k += 32) candidate_room(gt, L, cpos, cgs, k, room_l, room_r);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        room_l = min(room_l, __shfl_xor_sync(0xFFFFFFFFu, room_l, o));
        room_r = min(room_r, __shfl_xor_sync(0xFFFFFFFFu, room_r, o));
    }
    u32 a = scan_steps(packed, gt, sd, cpos, cgs, m, -1, 0, L, room_l / L);
    u32 b = scan_steps(packed, gt, sd, cpos, cgs, m, +1, 0, L, room_r / L);
    u32 c = scan_steps(packed, gt, sd, cpos, cgs, m, -1, a * L, 1, min(L, room_l - a * L));
    u32 d = scan_steps(packed, gt, sd, cpos, cgs, m, +1, b * L, 1, min(L, room_r - b * L));
    ext_l = a * L + c;
    ext_r = b * L + d;
}

// ---- one-thread extension on mismatch maps (L <= 32) ----------------------------------------------
// 32 bases of one component at match offsets [i0, i0+32) (relative to the seed start; the match strand
// is component 0's), first base in the top bits.  Reads may run up to 128 bases outside the genome:
// the packed buffer is padded on both sides of every genome and such bases never reach a tested window.
__device__ __forceinline__ u64 oriented_bases32(const u64* __restrict__ packed, const GenomeTable& gt, u32 L, u32 g, u32 pos, bool rev, int i0) {
    i64 q = (i64)gt.word_base[g] * 32 + (i64)pos + (rev ? (i64)L - 32 - i0 : (i64)i0);
    u64 i = (u64)q >> 5;
    int sh = (int)(q & 31) * 2;
    u64 w = shl128_hi(packed[i], packed[i + 1], sh);
    return rev ? rc_word(w) : w;
}
__device__ __forceinline__ u64 spread_nz(u64 x) { return (x | (x >> 1)) & 0x5555555555555555ull; }
// top 64 bits of the 128-bit map xa:xb shifted left by `idx` bases (0 <= idx < 64)
__device__ __forceinline__ u64 map_at(u64 xa, u64 xb, u32 idx) { return idx < 32 ? shl128_hi(xa, xb, 2 * (int)idx) : (xb << (2 * (idx - 32))); }

// Mismatch map (one flag per base, in the low bit of its 2-bit cell) of the 64 bases at match
// offsets [o_lo, o_lo + 64): flag set iff some component differs from component 0 there.
__device__ __forceinline__ void mismatch64(const u64* __restrict__ packed, const GenomeTable& gt, u32 L, const u32* __restrict__ cpos,
                                           const u8* __restrict__ cgs, u32 m, int o_lo, u64& xa, u64& xb) {
    u32 g0 = cgs[0] & 0x7F, p0 = cpos[0];
    u64 a0 = oriented_bases32(packed, gt, L, g0, p0, false, o_lo), b0 = oriented_bases32(packed, gt, L, g0, p0, false, o_lo + 32);
    xa = 0; xb = 0;
    for (u32 k = 1; k < m; ++k) {
        u8 gs = cgs[k];
        xa |= spread_nz(oriented_bases32(packed, gt, L, gs & 0x7F, cpos[k], gs & 0x80, o_lo) ^ a0);
        xb |= spread_nz(oriented_bases32(packed, gt, L, gs & 0x7F, cpos[k], gs & 0x80, o_lo + 32) ^ b0);
    }
}

#define DD_THREAD_CHUNKS 12 // 64-base chunks one thread walks per direction before deferring to the warp kernel

// Growth to the right (dir = +1) or left (dir = -1) of the seed: L-jumps, then at most L single steps
// (phases 1+3 resp. 0+2 of D14; the two directions do not interact).  Returns the growth in bases or
// 0xFFFFFFFF when the chunk budget ran out.
// Right: the chunk starts at match offset bL+1 (b = jumps so far), holds jump window b+1 at chunk
// index L-1 (and b+2 at 2L-1 when 3L <= 65); on a failing jump the single-step windows bL+s start at
// chunk index s-1 — all inside the same chunk.  Left is the mirror image.
__device__ __forceinline__ u32 grow_thread(const u64* __restrict__ packed, const GenomeTable& gt, const SeedDev& sd, const u32* cpos,
                                           const u8* cgs, u32 m, int dir, u32 room) {
    const u32 L = sd.L;
    const u64 care = sd.mask_hi & 0x5555555555555555ull;
    const u32 maxjumps = room / L;
    const u32 per_chunk = (3 * L <= 65) ? 2 : 1;
    u32 b = 0;
    for (int chunk = 0; chunk < DD_THREAD_CHUNKS; ++chunk) {
        u64 xa, xb;
        // right: chunk = offsets [bL+1, bL+65);  left: offsets [-bL-64+L-1, -bL+L-1) (chunk index i <-> offset -bL-65+L+i)
        int o_lo = dir > 0 ? (int)(b * L) + 1 : -(int)(b * L) - 65 + (int)L;
        mismatch64(packed, gt, L, cpos, cgs, m, o_lo, xa, xb);
        u32 b0 = b;
        bool failed = false;
        for (u32 w = 0; w < per_chunk && b < maxjumps; ++w) {
            // right: jump window b+1 covers offsets [(b+1)L, (b+2)L)   -> chunk index (b-b0)L + L-1
            // left : jump window b+1 covers offsets [-(b+1)L, -bL)     -> chunk index 65 - 2L - (b-b0)L
            u32 idx = dir > 0 ? (b - b0) * L + L - 1 : 65 - 2 * L - (b - b0) * L;
            if (map_at(xa, xb, idx) & care) { failed = true; break; }
            ++b;
        }
        if (failed || b >= maxjumps) {
            if (b - b0 == per_chunk && !failed) continue; // the chunk is used up: the single steps need a fresh one
            // single steps: right window s starts at offset bL+s, left window s at offset -bL-s
            u32 maxs = min(L, room - b * L), sdone = 0;
            for (u32 s = 1; s <= maxs; ++s) {
                u32 idx = dir > 0 ? (b - b0) * L + s - 1 : 65 - L - (b - b0) * L - s;
                if (map_at(xa, xb, idx) & care) break;
                sdone = s;
            }
            return b * L + sdone;
        }
    }
    return 0xFFFFFFFFu;
}

// ------------------------------------------------------------------------------------------------
// The whole batch loop runs inside ONE cooperative kernel (grid = all co-resident blocks): phases are
// separated by grid-wide barriers, work lists and their counters live in device memory, and the
// convergence test of a batch is made on the device, so the host enqueues one launch and never
// synchronises inside the de-dup stage.
//
// Because slots are ordered by (group, position), the candidates an extent contains are ONE contiguous
// slot range [rng_lo, rng_hi), found once per candidate by a galloping search around its own slot.
// "claim" then only looks at the set bits of the batch bitmap inside the range (live candidates of this
// batch), and "cover" ORs range masks into the covered bitmap.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define DD_NT 256
#define DD_THREAD_COMPS 8
#define DD_WIDE_SLOTS 1024 // slot ranges longer than this are handled by one warp instead of one thread

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

__device__ __forceinline__ u64 slot_key(const DedupArgs& a, u32 s) { return ((u64)a.slot_gid[s] << 33) | a.slot_x[s]; }
// first slot in (lo, hi] ... standard lower bound on slot_key over [lo, hi)
__device__ __forceinline__ u32 slot_lower_bound(const DedupArgs& a, u32 lo, u32 hi, u64 key) {
    while (lo < hi) {
        u32 mid = lo + (hi - lo) / 2;
        if (slot_key(a, mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// slot range of c's group with seed starts in [x - ext_l, x + ext_r]; galloping outwards from c's slot
__device__ __forceinline__ void extent_range(const DedupArgs& a, u32 c, u32 el, u32 er, u32& rlo, u32& rhi) {
    const u32 s0 = a.slot_of[c];
    const u32 x = a.comp_pos[a.cand_off[c]];
    const u64 g = (u64)a.gid[c] << 33;
    const u64 key_lo = g | (x - el), key_hi = g | ((u64)x + er + 1);
    u32 step = 1, hi = s0; // invariant: key(hi) >= key_lo
    while (true) {
        if (step > hi) { rlo = slot_lower_bound(a, 0, hi, key_lo); break; }
        u32 probe = hi - step;
        if (slot_key(a, probe) < key_lo) { rlo = slot_lower_bound(a, probe + 1, hi, key_lo); break; }
        hi = probe; step *= 2;
    }
    u32 lo = s0 + 1; // invariant: key(lo - 1) < key_hi
    step = 1;
    while (true) {
        if (lo + step > a.n_cand) { rhi = slot_lower_bound(a, lo, a.n_cand, key_hi); break; }
        u32 probe = lo + step - 1;
        if (slot_key(a, probe) >= key_hi) { rhi = slot_lower_bound(a, lo, probe, key_hi); break; }
        lo = probe + 1; step *= 2;
    }
}
__device__ __forceinline__ u64 range_mask(u32 word, u32 lo, u32 hi) { // bits of `word` inside [lo, hi)
    u64 m = ~0ull;
    if (lo > word * 64) m &= ~0ull << (lo - word * 64);
    if (hi < word * 64 + 64) m &= (1ull << (hi - word * 64)) - 1;
    return m;
}
__device__ __forceinline__ bool bit_of(const u64* bits, u32 s) { return (bits[s >> 6] >> (s & 63)) & 1; }

// claim the undecided higher-rank candidates of this batch among the slots of one bitmap word
__device__ __forceinline__ void claim_word(const DedupArgs& a, u32 w, u32 rlo, u32 rhi, u32 c) {
    u64 bits = a.batch_bits[w] & range_mask(w, rlo, rhi);
    while (bits) {
        u32 s = w * 64 + (u32)__ffsll((long long)bits) - 1;
        bits &= bits - 1;
        u32 j = a.cand_at[s];
        if (j > c && a.cstate[j] == 0) atomicMin(&a.minrank[s], c);
    }
}
// c is accepted: mark every slot of the range covered, except live lower-rank candidates of this batch
// (a match only contains candidates of higher rank, D16)
__device__ __forceinline__ void cover_word(const DedupArgs& a, u32 w, u32 rlo, u32 rhi, u32 c) {
    u64 m = range_mask(w, rlo, rhi), live = a.batch_bits[w] & m, excl = 0;
    while (live) {
        u64 b = live & (~live + 1);
        live ^= b;
        u32 s = w * 64 + (u32)__ffsll((long long)b) - 1;
        if (a.cand_at[s] < c) excl |= b;
    }
    atomicOr((unsigned long long*)&a.cov_bits[w], m & ~excl);
}
// 0 dropped, 1 accepted, 2 still undecided.  Claims of this round are complete (barrier).  One hop:
// if the lowest claimer of c is itself unclaimed and uncovered it is accepted in this very phase, so c
// is contained in an accepted match of lower rank.
__device__ __forceinline__ int decide_state(const DedupArgs& a, u32 c, u32 s) {
    if (bit_of(a.cov_bits, s)) return 0;
    u32 mr = a.minrank[s];
    if (mr == c) return 1;
    u32 se = a.slot_of[mr];
    bool mr_cov = bit_of(a.cov_bits, se);
    if (!mr_cov && a.minrank[se] == mr) return 0;
    if (mr_cov || a.cstate[mr] == 2) a.minrank[s] = 0xFFFFFFFFu; // that claimer is out: claim again next round
    return 2;
}

__device__ __forceinline__ u64 gtimer() { u64 t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// optional phase trace (a.trace != null): thread 0 appends (tag, globaltimer ns) pairs after each barrier
#define DD_TRACE(tag) do { if (t0 && a.trace) { u64 k = a.trace[0]; if (k < 4000) { a.trace[2 + 2 * k] = (tag); a.trace[3 + 2 * k] = gtimer(); a.trace[0] = k + 1; } } } while (0)

// ctr layout (u32): [0..2] narrow-list counters (rotating), [3..5] wide-list counters (rotating), [6] long list,
// [8] batches, [9] rounds, [10] wide items, [11] long items.  A counter is reset one round before it is
// written and never while it may still be read.
__global__ void __launch_bounds__(DD_NT) k_dedup_all(DedupArgs a, GenomeTable gt, SeedDev sd, u32 batch0) {
    cg::grid_group grid = cg::this_grid();
    const bool t0 = blockIdx.x == 0 && threadIdx.x == 0;
    const u32 gtid = blockIdx.x * DD_NT + threadIdx.x, gsz = gridDim.x * DD_NT;
    const int lane = threadIdx.x & 31;
    const u32 gwarp = gtid >> 5, nwarps = gsz >> 5;
    u32* ctr = a.ctr;
    u32* nl[3] = {a.wl0, a.wl1, a.wl2};
    u32* wd[3] = {a.wd0, a.wd1, a.wd2};
    const u32 bb_words = (a.n_cand + 63) / 64;
    u32 p = 0, batch = batch0;
    while (p < a.n_cand) {
        const u32 q = (u32)min((u64)a.n_cand, (u64)p + batch);
        if (t0) { for (int i = 0; i < 7; ++i) ctr[i] = 0; ctr[8] += 1; }
        for (u32 w = gtid; w < bb_words; w += gsz) a.batch_bits[w] = 0;
        grid.sync();
        DD_TRACE(1);
        // ---- begin: drop covered candidates, list the live ones, flag their slots
        for (u32 base = p + blockIdx.x * DD_NT; base < q; base += gsz) {
            u32 c = base + threadIdx.x;
            bool live = false;
            if (c < q) {
                u32 s = a.slot_of[c];
                if (bit_of(a.cov_bits, s)) a.cstate[c] = 2;
                else {
                    live = true;
                    a.cstate[c] = 0;
                    a.minrank[s] = 0xFFFFFFFFu;
                    atomicOr((unsigned long long*)&a.batch_bits[s >> 6], 1ull << (s & 63));
                }
            }
            wl_push(nl[0], ctr + 0, live, c);
        }
        grid.sync();
        DD_TRACE(2);
        // ---- extend: one thread per live candidate; long ones are parked for the warp phase
        {
            const u32 n = ctr[0];
            if (t0) atomicAdd(a.n_extended, n);
            for (u32 base = blockIdx.x * DD_NT; base < n; base += gsz) {
                u32 i = base + threadIdx.x;
                bool is_long = false;
                u32 c = 0;
                if (i < n) {
                    c = nl[0][i];
                    u32 off = a.cand_off[c], m = a.cand_off[c + 1] - off;
                    const u32* cpos = a.comp_pos + off;
                    const u8* cgs = a.comp_gs + off;
                    is_long = m > DD_THREAD_COMPS || sd.L > 32;
                    if (!is_long) {
                        u32 room_l = 0xFFFFFFFFu, room_r = 0xFFFFFFFFu;
                        for (u32 k = 0; k < m; ++k) candidate_room(gt, sd.L, cpos, cgs, k, room_l, room_r);
                        u32 el = grow_thread(a.packed, gt, sd, cpos, cgs, m, -1, room_l);
                        u32 er = el == 0xFFFFFFFFu ? el : grow_thread(a.packed, gt, sd, cpos, cgs, m, +1, room_r);
                        if (el == 0xFFFFFFFFu || er == 0xFFFFFFFFu) is_long = true;
                        else {
                            a.ext_l[c] = el; a.ext_r[c] = er;
                            u32 rlo, rhi;
                            extent_range(a, c, el, er, rlo, rhi);
                            a.rng_lo[c] = rlo; a.rng_hi[c] = rhi;
                        }
                    }
                }
                wl_push(a.wl_long, ctr + 6, is_long, c);
            }
        }
        grid.sync();
        DD_TRACE(3);
        if (ctr[6]) { // uniform across the grid
            const u32 n = ctr[6];
            for (u32 i = gwarp; i < n; i += nwarps) {
                u32 c = a.wl_long[i];
                u32 off = a.cand_off[c], m = a.cand_off[c + 1] - off;
                u32 el, er;
                extend_candidate_warp(a.packed, gt, sd, a.comp_pos + off, a.comp_gs + off, m, el, er);
                if (lane == 0) {
                    a.ext_l[c] = el; a.ext_r[c] = er;
                    u32 rlo, rhi;
                    extent_range(a, c, el, er, rlo, rhi);
                    a.rng_lo[c] = rlo; a.rng_hi[c] = rhi;
                }
            }
            if (t0) ctr[11] += n;
            grid.sync();
        }
        DD_TRACE(4);
        for (u32 r = 0;; ++r) {
            const u32 cur = r % 3, nxt = (r + 1) % 3, spare = (r + 2) % 3;
            if (t0) { ctr[spare] = 0; ctr[3 + spare] = 0; ctr[9] += 1; }
            const u32 n_narrow = ctr[cur];
            // ---- claim, one thread per candidate (round 0 also sorts out the wide ranges)
            for (u32 i = gtid; i < n_narrow; i += gsz) {
                u32 c = nl[cur][i];
                u32 rlo = a.rng_lo[c], rhi = a.rng_hi[c];
                if (r == 0 && rhi - rlo > DD_WIDE_SLOTS) { wd[0][atomicAdd(ctr + 3, 1u)] = c; continue; }
                for (u32 w = rlo >> 6; w <= (rhi - 1) >> 6; ++w) claim_word(a, w, rlo, rhi, c);
                atomicMin(&a.minrank[a.slot_of[c]], c);
            }
            if (r == 0) { grid.sync(); if (t0) ctr[10] += ctr[3]; } // the wide list is complete only now
            const u32 n_wide = ctr[3 + cur];
            // ---- claim, wide ranges: one warp each, a bitmap word per lane
            for (u32 i = gwarp; i < n_wide; i += nwarps) {
                u32 c = wd[cur][i];
                u32 rlo = a.rng_lo[c], rhi = a.rng_hi[c];
                for (u32 w = (rlo >> 6) + lane; w <= (rhi - 1) >> 6; w += 32) claim_word(a, w, rlo, rhi, c);
                if (lane == 0) atomicMin(&a.minrank[a.slot_of[c]], c);
            }
            grid.sync();
            DD_TRACE(5);
            // ---- decide, narrow
            for (u32 base = blockIdx.x * DD_NT; base < n_narrow; base += gsz) {
                u32 i = base + threadIdx.x;
                bool keep = false;
                u32 c = 0;
                if (i < n_narrow) {
                    c = nl[cur][i];
                    u32 rlo = a.rng_lo[c], rhi = a.rng_hi[c];
                    if (r > 0 || rhi - rlo <= DD_WIDE_SLOTS) {
                        int d = decide_state(a, c, a.slot_of[c]);
                        if (d == 0) a.cstate[c] = 2;
                        else if (d == 1) {
                            // accepted: everything of this group inside the extent is now contained
                            for (u32 w = rlo >> 6; w <= (rhi - 1) >> 6; ++w) cover_word(a, w, rlo, rhi, c);
                            a.cstate[c] = 1;
                        } else keep = true;
                    }
                }
                wl_push(nl[nxt], ctr + nxt, keep, c);
            }
            // ---- decide, wide
            for (u32 i = gwarp; i < n_wide; i += nwarps) {
                u32 c = wd[cur][i];
                int d = 0;
                if (lane == 0) d = decide_state(a, c, a.slot_of[c]);
                d = __shfl_sync(0xFFFFFFFFu, d, 0);
                if (d == 0) { if (lane == 0) a.cstate[c] = 2; }
                else if (d == 1) {
                    u32 rlo = a.rng_lo[c], rhi = a.rng_hi[c];
                    for (u32 w = (rlo >> 6) + lane; w <= (rhi - 1) >> 6; w += 32) cover_word(a, w, rlo, rhi, c);
                    if (lane == 0) a.cstate[c] = 1;
                } else if (lane == 0) wd[nxt][atomicAdd(ctr + 3 + nxt, 1u)] = c;
            }
            grid.sync();
            DD_TRACE(6);
            if (ctr[nxt] + ctr[3 + nxt] == 0) break;
        }
        p = q;
        batch = max(batch, p);
    }
}

void launch_group_ids(const DedupArgs& a, cudaStream_t st) {
    if (a.n_cand) k_group_ids<<<div_up(a.n_cand, 256), 256, 0, st>>>(a);
}
void launch_slot_keys(const DedupArgs& a, const GenomeTable& gt, u64* skey, u64* sval, cudaStream_t st) {
    if (a.n_cand) k_slot_keys<<<div_up(a.n_cand, 256), 256, 0, st>>>(a, gt, skey, sval);
}
void launch_slot_finish(const DedupArgs& a, const u64* skey, const u64* sval, cudaStream_t st) {
    if (a.n_cand) k_slot_finish<<<div_up(a.n_cand, 256), 256, 0, st>>>(a, skey, sval);
}
cudaError_t launch_dedup_all(const DedupArgs& a, const GenomeTable& gt, const SeedDev& sd, u32 batch0, cudaStream_t st) {
    if (a.n_cand == 0) return cudaSuccess;
    static int grid_blocks = 0;
    if (grid_blocks == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dedup_all, DD_NT, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        grid_blocks = sms * per_sm;
    }
    DedupArgs aa = a;
    GenomeTable g = gt;
    SeedDev s = sd;
    void* args[] = {&aa, &g, &s, &batch0};
    return cudaLaunchCooperativeKernel((const void*)k_dedup_all, dim3(grid_blocks), dim3(DD_NT), args, 0, st);
}

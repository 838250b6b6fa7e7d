// kernels_dedup.cu — a9-a11 of SURVEY.md §8a: MemHash::AddHashEntry (containment de-dup in ascending
// seed order) and MatchFinder::ExtendMatch (ungapped extension), both absent from /root/reference
// (libMems); semantics = SURVEY.md Appendix A D13-D16 and the A.2 pseudo-code.  Reached in the
// reference from /root/reference/src/UniqueMatchFinder.cpp:58 (HashMatch(unique_list)).
//
// The reference decides candidates one at a time in ascending seed order: a candidate is dropped
// iff an already ACCEPTED extended match on its diagonal contains it, otherwise it is extended and
// accepted.  Here candidates (already in ascending seed order = rank) are taken in doubling batches:
//   extend   every not-yet-covered candidate of the batch is extended (pure function of the genomes)
//   claim    every undecided candidate writes atomicMin(rank) on the slots of all same-diagonal
//            candidates its extent contains
//   decide   covered -> dropped;  min claimer == self -> accepted, marks its slots covered;
//            min claimer dropped -> reset slot and retry
// until the batch has no undecided candidate.  The fix-point equals the sequential result because a
// candidate is accepted exactly when every lower-rank container of it has been dropped.
// Slots: candidates ordered by (first genome, position) via a bitmap + popcount ranks, so "all
// candidates inside an extent" is a contiguous slot range found in O(1).
#include "common.cuh"
#include "kernels.h"

__device__ __forceinline__ u32 slot_rank(const u64* __restrict__ bitmap, const u32* __restrict__ bmrank, u64 gp) {
    u64 w = bitmap[gp >> 6];
    return bmrank[gp >> 6] + (u32)__popcll(w & ((1ull << (gp & 63)) - 1));
}

// candidate -> slot maps
__global__ void __launch_bounds__(256) k_build_slots(DedupArgs a, GenomeTable gt) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_cand) return;
    u32 off = a.cand_off[c];
    u32 g = a.comp_gs[off] & 0x7F;
    u64 gp = gt.base_base[g] + a.comp_pos[off];
    u32 s = slot_rank(a.bitmap, a.bmrank, gp);
    a.slot_of[c] = s;
    a.cand_at[s] = c;
}

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

// number of consecutive steps t = 1..maxcount whose windows agree across all components.
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

// D14: four phases.  One warp per candidate.
__device__ void extend_candidate(const u64* __restrict__ packed, const GenomeTable& gt, const SeedDev& sd, const u32* cpos, const u8* cgs,
                                 u32 m, u32& ext_l, u32& ext_r) {
    const int lane = threadIdx.x & 31;
    const u32 L = sd.L;
    u32 room_l = 0xFFFFFFFFu, room_r = 0xFFFFFFFFu;
    for (u32 k = lane; k < m; k += 32) {
        u8 gs = cgs[k];
        u32 p = cpos[k], len = gt.len[gs & 0x7F];
        u32 lroom = p, rroom = len - L - p;
        bool rev = gs & 0x80;
        room_l = min(room_l, rev ? rroom : lroom);
        room_r = min(room_r, rev ? lroom : rroom);
    }
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

// D16 group test: candidate j lies on candidate e's diagonal with the same genome set and strands
__device__ __forceinline__ bool same_group(const DedupArgs& a, u32 j, u32 offe, u32 m) {
    u32 offj = a.cand_off[j];
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

// slot range of the candidates whose seed window lies inside e's extent
__device__ __forceinline__ void extent_slots(const DedupArgs& a, const GenomeTable& gt, u32 c, u32& lo, u32& hi) {
    u32 off = a.cand_off[c];
    u32 g = a.comp_gs[off] & 0x7F;
    u64 x = gt.base_base[g] + a.comp_pos[off];
    u64 first = x - a.ext_l[c], last = x + a.ext_r[c]; // seed starts in [first, last]
    lo = slot_rank(a.bitmap, a.bmrank, first);
    hi = slot_rank(a.bitmap, a.bmrank, last + 1);
}

#define DD_WPB 8 // warps per block

// batch [p, q): drop covered candidates, extend the others, reset their claim slots
__global__ void __launch_bounds__(DD_WPB * 32) k_dd_extend(DedupArgs a, GenomeTable gt, SeedDev sd, u32 p, u32 q) {
    const int lane = threadIdx.x & 31;
    u32 c = p + blockIdx.x * DD_WPB + (threadIdx.x >> 5);
    if (c >= q) return;
    u32 s = a.slot_of[c];
    if (a.covered[s]) {
        if (lane == 0) a.cstate[c] = 2;
        return;
    }
    u32 off = a.cand_off[c], m = a.cand_off[c + 1] - off;
    u32 el, er;
    extend_candidate(a.packed, gt, sd, a.comp_pos + off, a.comp_gs + off, m, el, er);
    if (lane == 0) {
        a.ext_l[c] = el; a.ext_r[c] = er;
        a.minrank[s] = 0xFFFFFFFFu;
        a.cstate[c] = 0;
        atomicAdd(a.n_extended, 1u);
    }
}

__global__ void __launch_bounds__(DD_WPB * 32) k_dd_claim(DedupArgs a, GenomeTable gt, u32 p, u32 q) {
    const int lane = threadIdx.x & 31;
    u32 c = p + blockIdx.x * DD_WPB + (threadIdx.x >> 5);
    if (c >= q || a.cstate[c] != 0) return;
    u32 lo, hi;
    extent_slots(a, gt, c, lo, hi);
    u32 off = a.cand_off[c], m = a.cand_off[c + 1] - off;
    for (u32 s = lo + lane; s < hi; s += 32) {
        u32 j = a.cand_at[s];
        if (j >= p && j < q && a.cstate[j] == 0 && j > c && same_group(a, j, off, m)) atomicMin(&a.minrank[s], c);
    }
    if (lane == 0) atomicMin(&a.minrank[a.slot_of[c]], c);
}

__global__ void __launch_bounds__(DD_WPB * 32) k_dd_decide(DedupArgs a, GenomeTable gt, u32 p, u32 q) {
    const int lane = threadIdx.x & 31;
    u32 c = p + blockIdx.x * DD_WPB + (threadIdx.x >> 5);
    if (c >= q || a.cstate[c] != 0) return;
    u32 s = a.slot_of[c];
    if (a.covered[s]) {
        if (lane == 0) a.cstate[c] = 2;
        return;
    }
    u32 mr = a.minrank[s];
    if (mr == c) {
        // accepted: everything on this diagonal inside the extent is now contained
        u32 lo, hi;
        extent_slots(a, gt, c, lo, hi);
        u32 off = a.cand_off[c], m = a.cand_off[c + 1] - off;
        for (u32 t = lo + lane; t < hi; t += 32) {
            u32 j = a.cand_at[t];
            if (j > c && same_group(a, j, off, m)) a.covered[t] = 1;
        }
        if (lane == 0) { a.cstate[c] = 1; a.covered[s] = 1; }
    } else {
        if (lane == 0) {
            if (a.cstate[mr] == 2) a.minrank[s] = 0xFFFFFFFFu;
            atomicAdd(a.n_undecided, 1u);
        }
    }
}

void launch_build_slots(const DedupArgs& a, const GenomeTable& gt, cudaStream_t st) {
    if (a.n_cand) k_build_slots<<<div_up(a.n_cand, 256), 256, 0, st>>>(a, gt);
}
void launch_dd_extend(const DedupArgs& a, const GenomeTable& gt, const SeedDev& sd, u32 p, u32 q, cudaStream_t st) {
    if (q > p) k_dd_extend<<<div_up(q - p, DD_WPB), DD_WPB * 32, 0, st>>>(a, gt, sd, p, q);
}
void launch_dd_claim(const DedupArgs& a, const GenomeTable& gt, u32 p, u32 q, cudaStream_t st) {
    if (q > p) k_dd_claim<<<div_up(q - p, DD_WPB), DD_WPB * 32, 0, st>>>(a, gt, p, q);
}
void launch_dd_decide(const DedupArgs& a, const GenomeTable& gt, u32 p, u32 q, cudaStream_t st) {
    if (q > p) k_dd_decide<<<div_up(q - p, DD_WPB), DD_WPB * 32, 0, st>>>(a, gt, p, q);
}

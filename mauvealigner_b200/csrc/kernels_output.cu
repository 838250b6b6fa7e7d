// kernels_output.cu — a12 of SURVEY.md §8a: MemHash::GetMatchList -> MatchList, in the canonical
// order of SURVEY.md Appendix A D18 (call sites /root/reference/src/progressiveMauve.cpp:545,
// src/mauveAligner.cpp:583).  Accepted candidates are compacted, sorted by a 64-bit prefix of the
// D18 key with the radix sort, ties are finished with the full comparator, and the CSR is gathered.
#include "common.cuh"
#include "kernels.h"

__global__ void __launch_bounds__(256) k_uniq_flags(OutputArgs a) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < a.n_items) a.flags[c] = (a.state[c] & 15u) == 1u ? 1u : 0u;
}

// D18 for MODE_UNIQUE compares the dense vectors (|start[0]|..|start[N-1]|), 0 = absent.  For sparse
// matches this is: larger first genome index sorts first; then its start; then the next component...
__global__ void __launch_bounds__(256) k_uniq_keys(OutputArgs a, int sbits) {
    u32 it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= a.n_items || (a.state[it] & 15u) != 1u) return;
    u32 mi = a.match_idx[it];
    u32 c = a.item_cand[it];
    u32 off = a.cand_off[c];
    u32 f = a.comp_gs[off] & 0x7F;
    u64 st = (u64)a.comp_pos[off] - a.ext_l[c] + 1;
    a.sort_key[mi] = ((u64)(MB_MAX_SEQ - 1 - f) << sbits) | st;
    a.sort_val[mi] = c;
}

struct MView { u32 off, m, el, er; };
__device__ __forceinline__ MView mview(const OutputArgs& a, u32 c) {
    MView v; v.off = a.cand_off[c]; v.m = a.cand_off[c + 1] - v.off; v.el = a.ext_l[c]; v.er = a.ext_r[c];
    return v;
}
__device__ __forceinline__ u32 abs_start(const OutputArgs& a, const MView& v, u32 k) {
    u8 gs = a.comp_gs[v.off + k];
    return a.comp_pos[v.off + k] - ((gs & 0x80) ? v.er : v.el) + 1;
}
// full D18 comparator: A < B ?
__device__ bool d18_less(const OutputArgs& a, u32 ca, u32 cb, u32 L) {
    MView A = mview(a, ca), B = mview(a, cb);
    if (a.repeat) { // one sequence, columns = occurrences: (first start — equal here), multiplicity, signed starts, length
        if (A.m != B.m) return A.m < B.m;
        for (u32 k = 1; k < A.m; ++k) {
            i64 sa = abs_start(a, A, k), sb = abs_start(a, B, k);
            if (a.comp_gs[A.off + k] & 0x80) sa = -sa;
            if (a.comp_gs[B.off + k] & 0x80) sb = -sb;
            if (sa != sb) return sa < sb;
        }
        return (A.el + A.er) < (B.el + B.er);
    }
    u32 k = 0;
    for (;; ++k) {
        bool ea = k >= A.m, eb = k >= B.m;
        if (ea || eb) {
            if (ea && eb) break;
            return ea; // the one that still has components holds a non-zero where the other has 0
        }
        u32 ga = a.comp_gs[A.off + k] & 0x7F, gb = a.comp_gs[B.off + k] & 0x7F;
        if (ga != gb) return ga > gb; // smaller genome index = non-zero earlier = larger vector
        u32 sa = abs_start(a, A, k), sb = abs_start(a, B, k);
        if (sa != sb) return sa < sb;
    }
    for (k = 0; k < A.m; ++k) {
        bool ra = a.comp_gs[A.off + k] & 0x80, rb = a.comp_gs[B.off + k] & 0x80;
        if (ra != rb) return rb;
    }
    return (L + A.el + A.er) < (L + B.el + B.er);
}

__global__ void __launch_bounds__(256) k_uniq_tiefix(OutputArgs a, const u64* __restrict__ skey, u64* __restrict__ sval, u32 L) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    u32 n = (u32)*a.n_matches_ptr;
    if (j >= n) return;
    u64 kj = skey[j];
    if (j > 0 && skey[j - 1] == kj) return;
    u32 e = j + 1;
    while (e < n && skey[e] == kj) ++e;
    for (u32 i = j + 1; i < e; ++i) { // insertion sort of a (tiny) tie run
        u64 x = sval[i];
        u32 t = i;
        while (t > j && d18_less(a, (u32)x, (u32)sval[t - 1], L)) { sval[t] = sval[t - 1]; --t; }
        sval[t] = x;
    }
}

__global__ void __launch_bounds__(256) k_uniq_ncomp(OutputArgs a, const u64* __restrict__ sval) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    u32 n = (u32)*a.n_matches_ptr;
    if (j >= n) return;
    u32 c = (u32)sval[j];
    a.ncomp[j] = a.cand_off[c + 1] - a.cand_off[c];
}

// A warp takes 32 consecutive matches of the final order; their components are consecutive in the CSR, so
// the lanes walk the component index space together and every store is coalesced.
__global__ void __launch_bounds__(256) k_uniq_gather(OutputArgs a, const u64* __restrict__ sval, u32 L) {
    const int lane = threadIdx.x & 31;
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n = (u32)*a.n_matches_ptr;
    const bool valid = j < n;
    MView v{0, 0, 0, 0};
    u64 o = 0;
    if (valid) {
        u32 c = (u32)sval[j];
        v = mview(a, c);
        o = a.out_off[j];
        a.out_len[j] = L + v.el + v.er;
    }
    const u64 o_first = __shfl_sync(0xFFFFFFFFu, o, 0);
    u32 start = valid ? (u32)(o - o_first) : 0xFFFFFFFFu;
    const u32 nv = __popc(__ballot_sync(0xFFFFFFFFu, valid));
    const u32 last_start = __shfl_sync(0xFFFFFFFFu, start, nv ? nv - 1 : 0), last_m = __shfl_sync(0xFFFFFFFFu, v.m, nv ? nv - 1 : 0);
    const u32 T = nv ? last_start + last_m : 0;
    for (u32 base = 0; base < T; base += 32) {
        const u32 p = base + lane;
        u32 t = 0; // owner of component p: the last lane whose first component index is <= p
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            u32 sv = __shfl_sync(0xFFFFFFFFu, start, (t + step) & 31);
            if (t + step < 32 && sv <= p) t += step;
        }
        const u32 k = p - __shfl_sync(0xFFFFFFFFu, start, t);
        const u32 off = __shfl_sync(0xFFFFFFFFu, v.off, t), el = __shfl_sync(0xFFFFFFFFu, v.el, t), er = __shfl_sync(0xFFFFFFFFu, v.er, t);
        if (p < T) {
            u8 gs = a.comp_gs[off + k];
            const int32_t s = (int32_t)(a.comp_pos[off + k] - ((gs & 0x80) ? er : el) + 1);
            a.out_seq[o_first + p] = gs & 0x7F;
            a.out_start[o_first + p] = (gs & 0x80) ? -s : s;
        }
    }
}

// repeatoire's match position lookup table (/root/reference/src/repeatoire.cpp:1944-1966): for every position of the one
// sequence, which match (index in the result order = ascending LeftEnd(0), the order of repeatoire's seed_sort_list,
// :1920-1935) and which of its components starts there; 0xFFFFFFFF where none does.  Entry p = 1-based left end, entry 0 unused.
__global__ void __launch_bounds__(256) k_position_table(const u64* __restrict__ out_off, const int32_t* __restrict__ out_start, u32 n_matches,
                                                        unsigned long long* __restrict__ tab, u64 n_pos) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_matches) return;
    const u64 a = out_off[j], b = out_off[j + 1];
    for (u64 k = a; k < b; ++k) {
        const i64 s = out_start[k];
        const u64 p = (u64)(s < 0 ? -s : s);
        // components of two extended matches may share a left end: the later match of the list wins, as the last
        // assignment of the reference's loop over its sorted component list does
        if (p < n_pos) atomicMax(tab + p, ((unsigned long long)(j + 1) << 32) | (k - a));
    }
}
__global__ void __launch_bounds__(256) k_position_split(const unsigned long long* __restrict__ tab, u32* __restrict__ match_of, u32* __restrict__ comp_of, u64 n_pos) {
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pos) return;
    const unsigned long long v = tab[p];
    match_of[p] = v ? (u32)(v >> 32) - 1 : 0xFFFFFFFFu;
    comp_of[p] = v ? (u32)v : 0xFFFFFFFFu;
}
void launch_position_table(const u64* out_off, const int32_t* out_start, u32 n_matches, u64* tab, u32* match_of, u32* comp_of, u64 n_pos, cudaStream_t st) {
    if (n_matches) k_position_table<<<div_up(n_matches, 256), 256, 0, st>>>(out_off, out_start, n_matches, reinterpret_cast<unsigned long long*>(tab), n_pos);
    k_position_split<<<div_up(n_pos, 256), 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(tab), match_of, comp_of, n_pos);
}

// The device result keeps 5 bytes per component (1-byte sequence index, 4-byte signed 1-based start: every sequence is
// shorter than 2^31 bases); mb_fetch_result widens it to the u32 / i64 arrays of mb_result on request,
// mb_fetch_result_compact copies it as it is (less than half the bytes over PCIe).
__global__ void __launch_bounds__(256) k_expand_result(const u8* __restrict__ seq8, const int32_t* __restrict__ start32, u64 n, u32* __restrict__ seq,
                                                       i64* __restrict__ start) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { seq[i] = seq8[i]; start[i] = start32[i]; }
}
void launch_expand_result(const u8* seq8, const int32_t* start32, u64 n_comps, u32* seq, i64* start, cudaStream_t st) {
    if (n_comps) k_expand_result<<<div_up(n_comps, 256), 256, 0, st>>>(seq8, start32, n_comps, seq, start);
}

void launch_uniq_flags(const OutputArgs& a, cudaStream_t st) {
    if (a.n_items) k_uniq_flags<<<div_up(a.n_items, 256), 256, 0, st>>>(a);
}
void launch_uniq_keys(const OutputArgs& a, int sbits, cudaStream_t st) {
    if (a.n_items) k_uniq_keys<<<div_up(a.n_items, 256), 256, 0, st>>>(a, sbits);
}
void launch_uniq_tiefix(const OutputArgs& a, const u64* skey, u64* sval, u32 L, u32 n_upper, cudaStream_t st) {
    if (n_upper) k_uniq_tiefix<<<div_up(n_upper, 256), 256, 0, st>>>(a, skey, sval, L);
}
void launch_uniq_ncomp(const OutputArgs& a, const u64* sval, u32 n_upper, cudaStream_t st) {
    if (n_upper) k_uniq_ncomp<<<div_up(n_upper, 256), 256, 0, st>>>(a, sval);
}
void launch_uniq_gather(const OutputArgs& a, const u64* sval, u32 L, u32 n_upper, cudaStream_t st) {
    if (n_upper) k_uniq_gather<<<div_up(n_upper, 256), 256, 0, st>>>(a, sval, L);
}

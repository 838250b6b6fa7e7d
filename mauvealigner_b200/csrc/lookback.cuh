// lookback.cuh — single-pass ordered prefix across tiles (decoupled look-back), shared by the
// run detector, the candidate selector and the generic scans.  One 64-bit status word per tile:
// 2 flag bits + 62 value bits (callers pack two counters into the value when they need a pair).
#pragma once
#include "common.cuh"

#define LB_FLAG_AGG (1ull << 62)
#define LB_FLAG_INC (2ull << 62)
#define LB_MASK ((1ull << 62) - 1)

__device__ __forceinline__ u64 lb_ld(const u64* p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lb_st(u64* p, u64 v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Called by ONE full warp of the tile's block with the tile aggregate `agg`.  Tiles must have been
// taken in ticket order.  Publishes agg, waits for predecessors, publishes the inclusive prefix,
// returns the exclusive prefix (to every lane).
__device__ __forceinline__ u64 lookback_exclusive(u64* status, u32 tile, u64 agg) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) lb_st(status + tile, (tile == 0 ? LB_FLAG_INC : LB_FLAG_AGG) | agg);
    if (tile == 0) return 0;
    u64 excl = 0;
    i64 base = (i64)tile - 1;
    while (true) {
        i64 t = base - lane;
        u64 s = LB_FLAG_INC; // virtual tiles before 0: inclusive prefix 0
        if (t >= 0) {
            do { s = lb_ld(status + t); } while ((s >> 62) == 0);
        }
        u32 inc = __ballot_sync(0xFFFFFFFFu, (s & LB_FLAG_INC) != 0);
        int first = __ffs(inc) - 1; // nearest predecessor holding an inclusive prefix, or -1
        u64 v = (first < 0 || lane <= first) ? (s & LB_MASK) : 0;
        excl += warp_sum_u64(v);
        if (first >= 0) break;
        base -= 32;
    }
    if (lane == 0) lb_st(status + tile, LB_FLAG_INC | (excl + agg));
    return excl;
}

// Block-wide exclusive scan of one u64 per thread; returns the exclusive prefix and the block total.
// NT threads, scratch must hold NT/32 + 1 u64.
template <int NT>
__device__ __forceinline__ u64 block_excl_scan_u64(u64 v, u64* scratch, u64& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) scratch[warp] = x;
    __syncthreads();
    if (warp == 0) {
        u64 w = lane < NT / 32 ? scratch[lane] : 0;
        u64 xs = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u64 y = __shfl_up_sync(0xFFFFFFFFu, xs, o);
            if (lane >= o) xs += y;
        }
        if (lane < NT / 32) scratch[lane] = xs - w;
        if (lane == NT / 32 - 1) scratch[NT / 32] = xs;
    }
    __syncthreads();
    u64 r = scratch[warp] + x - v;
    total = scratch[NT / 32];
    __syncthreads();
    return r;
}

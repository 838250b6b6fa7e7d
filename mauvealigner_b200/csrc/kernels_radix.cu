// kernels_radix.cu — a4 (sorted mer list construction) + a6 (N-way merge) of SURVEY.md §8a as ONE
// stable LSD radix sort of all genomes' seed records.
//
// Replaces libMems DNAMemorySML::Create / FileSML::Create (per-genome sort) and the streaming merge
// of MatchFinder::FindMatchSeeds (call sites /root/reference/src/SeedMatchEnumerator.h:61,
// src/mauveAligner.cpp:465,585).  Records are emitted in (genome, position) order and every pass
// is stable, so equal seeds end up ordered by (genome, position) = Appendix A D6.
//
// One pass = one kernel ("onesweep"): tiles are taken in order through an atomic ticket; each warp counts its
// digits with shared atomics, the tile total is published for the decoupled look-back of later tiles, the ranking
// (atomicOr peer masks in shared memory alternating with ballot rounds — the two load different pipes) yields the
// tile slot of every record directly, and each digit's records leave the reordered tile as one contiguous burst.
// Traffic per pass: one read + one write of every record (2R bytes/record); measured limits: the shared-memory
// data pipe (~75 % busy, bank conflicts of the digit-indexed tables and of the key scatter) and the ALU pipe (~53 %).
#include <algorithm>
#include "common.cuh"
#include "kernels.h"

#ifndef RS_NT
#define RS_NT 512
#endif
#define RS_NW (RS_NT / 32)
#ifndef RS_IPT
#define RS_IPT 14
#endif
#define RS_TILE (RS_NT * RS_IPT)
#ifndef RS_MINB
#define RS_MINB 2
#endif
#ifndef RS_OR_EVERY
#define RS_OR_EVERY 2 // every RS_OR_EVERY-th ranking round finds its digit peers through shared-memory atomicOr, the others by ballots
#endif                // (0: ballots only, 1: atomicOr only); the two load different pipes (ALU / shared memory)
#ifndef RS_LB
#define RS_LB 4       // predecessors read per look-back step (independent loads in flight)
#endif
#ifndef RS_LB_FIRST
#define RS_LB_FIRST 1 // warps 0..7 resolve the look-back before they rank (it overlaps the ranking of the other warps)
#endif

#define LB_FLAG_AGG (1ull << 62)
#define LB_FLAG_INC (2ull << 62)
#define LB_MASK ((1ull << 62) - 1)

__device__ __forceinline__ u64 ld_volatile_u64(const u64* p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(u64* p, u64 v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void bar_sync_256() { asm volatile("bar.sync 1, 256;" ::: "memory"); } // warps 0..7 only

#define DIGIT(x) (USE_LUT ? (u32)sLut[(u32)((x) >> shift) & dmask] : ((u32)((x) >> shift) & dmask))

// One CTA = one tile, taken in order through an atomic ticket.  Shared memory: the tile (RS_TILE records: keys and then —
// HAS_VAL — the values through the same buffer), one table of 256 {lo, hi} entries per warp, 256 global bases.
//   entry of (warp, digit):  hi = first the warp's count of the digit (shared atomics, before anything else), then — after
//                                 the scan over digits and warps — the tile slot of the warp's next record of the digit;
//                            lo = mask of the lanes that hold the digit in the current ranking round (atomicOr rounds)
// so that the ranking yields the tile slot of every record directly and the keys leave the registers at once.
// the reordered tile in shared memory.  RS_SPLIT=1 keeps the two 32-bit halves of a record in two arrays (two 4-byte
// scatter stores instead of one 8-byte store); measured slightly slower (0.224 against 0.221 ms per pass), so it is off
#ifndef RS_SPLIT
#define RS_SPLIT 0
#endif
#if RS_SPLIT
#define ST_TILE(slot, v) do { reinterpret_cast<u32*>(sKeys)[(slot)] = (u32)(v); reinterpret_cast<u32*>(sKeys)[RS_TILE + (slot)] = (u32)((v) >> 32); } while (0)
#define LD_TILE(slot) ((u64)reinterpret_cast<const u32*>(sKeys)[(slot)] | ((u64)reinterpret_cast<const u32*>(sKeys)[RS_TILE + (slot)] << 32))
#else
#define ST_TILE(slot, v) (sKeys[(slot)] = (v))
#define LD_TILE(slot) (sKeys[(slot)])
#endif

struct RsShared { u64* sKeys; u64* sMH; u32* sBase; u32* sWarpSums; const u8* sLut; u64* const* sPeer; };

// One tile.  FULL (every slot of the tile holds a record — all tiles but the last) drops the per-record validity tests.
template <bool HAS_VAL, bool USE_LUT, bool FULL>
__device__ __forceinline__ void onesweep_tile(const RsShared sh, const u64* __restrict__ kin, u64* __restrict__ kout, const u64* __restrict__ vin,
                                              u64* __restrict__ vout, const u32* __restrict__ digit_base, u64* lookback, int shift, u32 dmask, u32 tile,
                                              u64 tile_base, u32 tile_n) {
    u64* const sKeys = sh.sKeys; u64* const sMH = sh.sMH; u32* const sBase = sh.sBase; u32* const sWarpSums = sh.sWarpSums;
    const u8* const sLut = sh.sLut; u64* const* const sPeer = sh.sPeer;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool full = FULL;
    const u32 lt_mask = (1u << lane) - 1;
    const u32 lane_bit = 1u << lane;
    u64* myMH = sMH + warp * 256;
    const u32 wbase = warp * (RS_IPT * 32);

    // warp-striped load; the warp's digit counts by shared atomics
    u64 key[RS_IPT];
    u32 slot[HAS_VAL ? RS_IPT : 1];
    {
        const u64* src = kin + tile_base + wbase + lane;
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k) key[k] = (full || wbase + k * 32 + lane < tile_n) ? src[k * 32] : ~0ull;
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k)
            if (full || wbase + k * 32 + lane < tile_n) atomicAdd(reinterpret_cast<u32*>(myMH + DIGIT(key[k])) + 1, 1u);
    }
    __syncthreads();

    // warps 0..7, thread = digit: tile count (published at once), exclusive scan over the digits, slot bases of the warps
    u32 tile_count = 0, tile_prefix = 0;
    if (tid < 256) {
        u32 c[RS_NW];
#pragma unroll
        for (int w = 0; w < RS_NW; ++w) { c[w] = reinterpret_cast<const u32*>(sMH + w * 256 + tid)[1]; tile_count += c[w]; }
        st_volatile_u64(lookback + (u64)tile * 256 + tid, (tile == 0 ? LB_FLAG_INC : LB_FLAG_AGG) | (u64)tile_count);
        u32 x = tile_count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) sWarpSums[warp] = x;
        bar_sync_256();
        u32 pre = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) pre += (w < warp) ? sWarpSums[w] : 0u;
        tile_prefix = pre + x - tile_count;
        u32 acc = tile_prefix;
#pragma unroll
        for (int w = 0; w < RS_NW; ++w) { sMH[w * 256 + tid] = (u64)acc << 32; acc += c[w]; }
    }
    __syncthreads();

    auto look_back = [&]() {
        u64 excl = 0;
        if (tile > 0) {
            i64 t = (i64)tile - 1;
            bool done = false;
            while (!done) {
                u64 s[RS_LB];
#pragma unroll
                for (int j = 0; j < RS_LB; ++j) s[j] = ld_volatile_u64(lookback + (u64)(t - j > 0 ? t - j : 0) * 256 + tid);
#pragma unroll
                for (int j = 0; j < RS_LB; ++j) {
                    if (done || (s[j] >> 62) == 0) break; // not published yet: read again from there
                    excl += s[j] & LB_MASK;
                    if (s[j] & LB_FLAG_INC) done = true;
                    --t;
                }
            }
            st_volatile_u64(lookback + (u64)tile * 256 + tid, LB_FLAG_INC | (excl + tile_count));
        }
        sBase[tid] = digit_base[tid] + (u32)excl - tile_prefix; // n < 2^31: 32-bit wrap-around arithmetic is exact
    };
    if (RS_LB_FIRST && tid < 256) look_back();

    // ---- rank within the warp, per digit, in (k, lane) order.  Peers of a record = lanes holding the same digit in this
    // round, found either by an atomicOr of the lane bit into the entry of the digit (shared-memory pipe) or by 8 ballots
    // (ALU pipe; both measured in tools/ubench/ub_rank.cu); the lowest peer advances the warp's next slot of the digit.
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        const u32 d = DIGIT(key[k]);
        const bool valid = full || wbase + k * 32 + lane < tile_n;
        u32* e32 = reinterpret_cast<u32*>(myMH + d);
        u32 pm, old;
        if (RS_OR_EVERY != 0 && (k % RS_OR_EVERY) == RS_OR_EVERY - 1) {
            if (valid) atomicOr(e32, lane_bit);
            __syncwarp();
            const u64 e = myMH[d];
            __syncwarp();
            pm = valid ? (u32)e : 0u; old = (u32)(e >> 32);
            if (valid && (pm & lt_mask) == 0) myMH[d] = (u64)(old + __popc(pm)) << 32;
        } else {
            u32 differ = 0;
#pragma unroll
            for (int bit = 0; bit < 8; ++bit) {
                const bool b = (d >> bit) & 1u;
                differ |= __ballot_sync(0xFFFFFFFFu, b) ^ (b ? 0xFFFFFFFFu : 0u);
            }
            pm = ~differ;
            if (!full) { pm &= __ballot_sync(0xFFFFFFFFu, valid); if (!valid) pm = 0; }
            old = e32[1];
            __syncwarp();
            if (pm != 0 && (pm & lt_mask) == 0) e32[1] = old + __popc(pm);
        }
        __syncwarp();
        const u32 sl = old + __popc(pm & lt_mask);
        if (valid) ST_TILE(sl, key[k]);
        if (HAS_VAL) slot[k] = sl;
    }
    if (HAS_VAL) { // the values follow through the same buffer; their loads overlap the key stores
        const u64* src = vin + tile_base + wbase + lane;
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k) key[k] = (full || wbase + k * 32 + lane < tile_n) ? src[k * 32] : 0ull;
    }
    if (!RS_LB_FIRST && tid < 256) look_back();
    __syncthreads();

    // every digit's records leave as one contiguous burst
    u32 dst[HAS_VAL ? RS_IPT : 1];
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const u32 s = tid + j * RS_NT;
        if (full || s < tile_n) {
            const u64 kk = LD_TILE(s);
            const u32 d = DIGIT(kk);
            const u32 o = sBase[d] + s;
            if (USE_LUT) sPeer[d][o] = kk; // the bins may live in other GPUs' memory (NVLink peer stores)
            else kout[o] = kk;
            if (HAS_VAL) dst[j] = o;
        }
    }
    if (HAS_VAL) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k)
            if (full || wbase + k * 32 + lane < tile_n) ST_TILE(slot[k], key[k]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RS_IPT; ++j) {
            const u32 s = tid + j * RS_NT;
            if (full || s < tile_n) vout[dst[j]] = LD_TILE(s);
        }
    }
}

template <bool HAS_VAL, bool USE_LUT>
__global__ void __launch_bounds__(RS_NT, RS_MINB) k_onesweep(const u64* __restrict__ kin, u64* __restrict__ kout,
                                                    const u64* __restrict__ vin, u64* __restrict__ vout, u32 n,
                                                    const u32* __restrict__ digit_base /*[256] exclusive*/,
                                                    u64* lookback /*[tiles][256]*/, u32* ticket, int shift, u32 dmask, const u8* __restrict__ lut,
                                                    u64* const* __restrict__ peers /*[256] or null: output array of every bin */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* sKeys = reinterpret_cast<u64*>(smem_raw);                    // RS_TILE
    u64* sMH = sKeys + RS_TILE;                                        // RS_NW * 256 entries
    u32* sBase = reinterpret_cast<u32*>(sMH + RS_NW * 256);            // 256: global index of tile slot 0 of the digit
    __shared__ u32 sTile;
    __shared__ u32 sWarpSums[8];
    __shared__ u8 sLut[256];
    __shared__ u64* sPeer[USE_LUT ? 256 : 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    if (USE_LUT && tid < 256) { sLut[tid] = lut[tid]; sPeer[tid] = peers ? peers[tid] : kout; }
    for (int i = tid; i < RS_NW * 256; i += RS_NT) sMH[i] = 0;
    __syncthreads();
    const u32 tile = sTile;
    const u64 tile_base = (u64)tile * RS_TILE;
    const u32 tile_n = (u32)min((u64)RS_TILE, (u64)n - tile_base);
    RsShared sh{sKeys, sMH, sBase, sWarpSums, sLut, sPeer};
    if (tile_n == RS_TILE) onesweep_tile<HAS_VAL, USE_LUT, true>(sh, kin, kout, vin, vout, digit_base, lookback, shift, dmask, tile, tile_base, tile_n);
    else onesweep_tile<HAS_VAL, USE_LUT, false>(sh, kin, kout, vin, vout, digit_base, lookback, shift, dmask, tile, tile_base, tile_n);
}

// exclusive scan of each pass's 256-bin histogram (one block per pass)
__global__ void __launch_bounds__(256) k_scan_hist(const u32* __restrict__ hist, u32* __restrict__ base) {
    __shared__ u32 s[256];
    const int t = threadIdx.x;
    const u32* h = hist + blockIdx.x * 256;
    s[t] = h[t];
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        u32 v = t >= o ? s[t - o] : 0;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    base[blockIdx.x * 256 + t] = s[t] - h[t];
}

size_t radix_smem_bytes(bool) { return (size_t)RS_TILE * 8 + RS_NW * 256 * 8 + 256 * 4; }
u32 radix_tile_size() { return RS_TILE; }

void launch_scan_hist(const u32* d_hist, u32* d_base, int npass, cudaStream_t st) {
    k_scan_hist<<<npass, 256, 0, st>>>(d_hist, d_base);
}

// the dynamic shared memory opt-in is per function AND per device (a process may hold contexts on several devices)
#define RS_MAX_DEV 64
static bool g_attr_set[RS_MAX_DEV][4] = {};

cudaError_t launch_onesweep(const u64* kin, u64* kout, const u64* vin, u64* vout, u32 n, const u32* d_digit_base,
                            u64* d_lookback, u32* d_ticket, int shift, int bits, cudaStream_t st, const u8* lut, u64* const* peers) {
    if (n == 0) return cudaSuccess;
    bool hv = vin != nullptr;
    if (lut && hv && peers) return cudaErrorInvalidValue; // 16-byte records are partitioned into local send buffers only
    size_t smem = radix_smem_bytes(hv);
    int variant = lut ? (hv ? 3 : 2) : (hv ? 1 : 0);
    int dev = 0;
    cudaError_t de = cudaGetDevice(&dev);
    if (de != cudaSuccess) return de;
    const bool tracked = dev >= 0 && dev < RS_MAX_DEV;
    if (!tracked || !g_attr_set[dev][variant]) {
        cudaError_t e = variant == 3 ? cudaFuncSetAttribute(k_onesweep<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : variant == 2 ? cudaFuncSetAttribute(k_onesweep<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : variant == 1 ? cudaFuncSetAttribute(k_onesweep<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                     : cudaFuncSetAttribute(k_onesweep<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (tracked) g_attr_set[dev][variant] = true;
    }
    u32 tiles = div_up(n, RS_TILE);
    u32 dmask = (1u << bits) - 1;
    if (variant == 3) k_onesweep<true, true><<<tiles, RS_NT, smem, st>>>(kin, kout, vin, vout, n, d_digit_base, d_lookback, d_ticket, shift, dmask, lut, peers);
    else if (variant == 2) k_onesweep<false, true><<<tiles, RS_NT, smem, st>>>(kin, kout, vin, vout, n, d_digit_base, d_lookback, d_ticket, shift, dmask, lut, peers);
    else if (variant == 1) k_onesweep<true, false><<<tiles, RS_NT, smem, st>>>(kin, kout, vin, vout, n, d_digit_base, d_lookback, d_ticket, shift, dmask, lut, peers);
    else k_onesweep<false, false><<<tiles, RS_NT, smem, st>>>(kin, kout, vin, vout, n, d_digit_base, d_lookback, d_ticket, shift, dmask, lut, peers);
    return cudaGetLastError();
}

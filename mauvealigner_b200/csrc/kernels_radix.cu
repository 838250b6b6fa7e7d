// kernels_radix.cu — a4 (sorted mer list construction) + a6 (N-way merge) of SURVEY.md §8a as ONE
// stable LSD radix sort of all genomes' seed records.
//
// Replaces libMems DNAMemorySML::Create / FileSML::Create (per-genome sort) and the streaming merge
// of MatchFinder::FindMatchSeeds (call sites /root/reference/src/SeedMatchEnumerator.h:61,
// src/mauveAligner.cpp:465,585).  Records are emitted in (genome, position) order and every pass
// is stable, so equal seeds end up ordered by (genome, position) = Appendix A D6.
//
// One pass = one kernel ("onesweep"): tiles are taken in order through an atomic ticket, each tile
// ranks its keys per 8-bit digit (warp __match_any_sync ranking), publishes its digit counts,
// resolves its global digit offsets by decoupled look-back over earlier tiles, reorders the tile
// in shared memory and writes each digit's keys as one contiguous burst.
// Traffic per pass: one read + one write of every record (2R bytes/record).
#include "common.cuh"
#include "kernels.h"

#ifndef RS_NT
#define RS_NT 512
#endif
#define RS_NW (RS_NT / 32)
#ifndef RS_IPT
#define RS_IPT 12
#endif
#define RS_TILE (RS_NT * RS_IPT)

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

#define DIGIT(x) (USE_LUT ? (u32)sLut[(u32)((x) >> shift) & dmask] : ((u32)((x) >> shift) & dmask))
#ifndef RS_MINB
#define RS_MINB 2
#endif
template <bool HAS_VAL, bool USE_LUT>
__global__ void __launch_bounds__(RS_NT, RS_MINB) k_onesweep(const u64* __restrict__ kin, u64* __restrict__ kout,
                                                    const u64* __restrict__ vin, u64* __restrict__ vout, u32 n,
                                                    const u32* __restrict__ digit_base /*[256] exclusive*/,
                                                    u64* lookback /*[tiles][256]*/, u32* ticket, int shift, u32 dmask, const u8* __restrict__ lut,
                                                    u64* const* __restrict__ peers /*[256] or null: output array of every bin */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* sKeys = reinterpret_cast<u64*>(smem_raw);                    // RS_TILE
    u64* sVals = sKeys + RS_TILE;                                      // RS_TILE when HAS_VAL
    u32* sWarpHist = reinterpret_cast<u32*>(sKeys + (HAS_VAL ? 2 : 1) * RS_TILE); // RS_NW*256
    u32* sTilePrefix = sWarpHist + RS_NW * 256;                        // 256 exclusive digit offsets in tile
    u32* sGlobBase = sTilePrefix + 256;                                // 256: global index of slot 0 of digit (mod 2^32)
    __shared__ u32 sTile;
    __shared__ u32 sWarpSums[8];
    __shared__ u8 sLut[256];
    __shared__ u64* sPeer[USE_LUT ? 256 : 1];
#ifdef RS_CLAIM_RANK
    __shared__ u8 sClaim[RS_NW][256];
#endif

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    if (USE_LUT && tid < 256) { sLut[tid] = lut[tid]; sPeer[tid] = peers ? peers[tid] : kout; }
    for (int i = tid; i < RS_NW * 256; i += RS_NT) sWarpHist[i] = 0;
    __syncthreads();
    const u32 tile = sTile;
    const u64 tile_base = (u64)tile * RS_TILE;
    const u32 tile_n = (u32)min((u64)RS_TILE, (u64)n - tile_base);

    // warp-striped load; rk[k] = digit << 16 | rank (the rank is filled in below; a tile has < 65536 records)
    u64 key[RS_IPT];
    u32 rk[RS_IPT];
    const u32 wbase = warp * (RS_IPT * 32);
    const bool full = tile_n == RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        u32 o = wbase + k * 32 + lane;
        key[k] = (o < tile_n) ? kin[tile_base + o] : ~0ull;
        rk[k] = DIGIT(key[k]) << 16;
    }
    // early counts: the tile's digit histogram by shared atomics, published before the (slower) ranking so that
    // the look-back of later tiles never waits for this tile's ranking
    {
        u32* sEarly = reinterpret_cast<u32*>(sGlobBase); // 256 x u32, overwritten by sGlobBase only after the look-back
        if (tid < 256) sEarly[tid] = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k)
            if (full || wbase + k * 32 + lane < tile_n) atomicAdd(&sEarly[rk[k] >> 16], 1u);
        __syncthreads();
        if (tid < 256) st_volatile_u64(lookback + (u64)tile * 256 + tid, (tile == 0 ? LB_FLAG_INC : LB_FLAG_AGG) | (u64)sEarly[tid]);
    }
    // rank within the warp, per digit, in (k, lane) order.  Peers of a record = lanes holding the same digit:
    // 8 ballots (one per digit bit; lanes whose bit equals mine = ballot ^ (bit ? 0 : ~0)) are much cheaper
    // than MATCH.ANY here (measured: 0.44 -> 0.33 ms per pass over 40 M records).  Every peer reads the warp's
    // running digit counter, the lowest peer lane advances it.
    u32* myHist = sWarpHist + warp * 256;
    const u32 lt_mask = (1u << lane) - 1;
#ifdef RS_CLAIM_RANK
    // EXPERIMENTAL (build with EXTRA=-DRS_CLAIM_RANK; not the default: written at the end of r01 without GPU time left to
    // measure it, logic checked lane by lane in a host simulation).  Per-warp claim table: every lane stores its lane id at
    // its digit and reads the entry back.  A lane that reads itself back either holds its digit alone in this round — the
    // usual case with 256 bins and 32 lanes — or is the one winner of a colliding group; the lanes that read someone
    // else's id know they collide.  Only the colliding groups (about two per round on uniform digits) cost a ballot each.
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        const u32 d = rk[k] >> 16;
        const bool valid = full || wbase + k * 32 + lane < tile_n;
        if (valid) sClaim[warp][d] = (u8)lane;
        __syncwarp();
        const u32 w = valid ? (u32)sClaim[warp][d] : (u32)lane;
        u32 lost = __ballot_sync(0xFFFFFFFFu, w != (u32)lane); // also orders this round's loads before the next round's stores
        u32 below = 0, size = 1;
        while (lost) { // warp-uniform
            const int leader = __ffs(lost) - 1;
            const u32 dl = __shfl_sync(0xFFFFFFFFu, d, leader);
            const bool in = valid && d == dl;
            const u32 grp = __ballot_sync(0xFFFFFFFFu, in);
            if (in) { below = __popc(grp & lt_mask); size = __popc(grp); }
            lost &= ~grp;
        }
        const u32 old = valid ? myHist[d] : 0u;
        __syncwarp();
        if (valid && below == 0) myHist[d] = old + size;
        __syncwarp();
        rk[k] |= old + below;
    }
#else
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        const u32 d = rk[k] >> 16;
        u32 pm = 0xFFFFFFFFu;
#pragma unroll
        for (int bit = 0; bit < 8; ++bit) {
            const bool b = (d >> bit) & 1u;
            pm &= __ballot_sync(0xFFFFFFFFu, b) ^ (b ? 0u : 0xFFFFFFFFu);
        }
        if (!full) {
            const bool valid = wbase + k * 32 + lane < tile_n;
            pm &= __ballot_sync(0xFFFFFFFFu, valid);
            if (!valid) pm = 0;
        }
        const u32 old = myHist[d];
        __syncwarp();
        const u32 below = pm & lt_mask;
        if (below == 0 && pm != 0) myHist[d] = old + __popc(pm);
        __syncwarp();
        rk[k] |= old + __popc(below);
    }
#endif
    __syncthreads();

    // per digit: exclusive scan across warps, tile total, look-back
    u32 tile_count = 0;
    if (tid < 256) {
        u32 acc = 0;
#pragma unroll
        for (int w = 0; w < RS_NW; ++w) {
            u32 t = sWarpHist[w * 256 + tid];
            sWarpHist[w * 256 + tid] = acc;
            acc += t;
        }
        tile_count = acc;
    }
    // exclusive scan of tile_count over the 256 digits (threads 0..255 = 8 warps)
    {
        u32 v = tile_count, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (tid < 256 && lane == 31) sWarpSums[warp] = x;
        __syncthreads();
        if (tid < 256) {
            u32 pre = 0;
            for (int w = 0; w < warp; ++w) pre += sWarpSums[w];
            sTilePrefix[tid] = pre + x - v;
        }
    }
    if (tid < 256) {
        u64 excl = 0;
        if (tile > 0) {
            i64 t = (i64)tile - 1;
            while (true) {
                u64 s = ld_volatile_u64(lookback + (u64)t * 256 + tid);
                if ((s >> 62) == 0) continue; // not published yet
                excl += s & LB_MASK;
                if (s & LB_FLAG_INC) break;
                --t;
            }
            st_volatile_u64(lookback + (u64)tile * 256 + tid, LB_FLAG_INC | (excl + tile_count));
        }
        sGlobBase[tid] = digit_base[tid] + (u32)excl - sTilePrefix[tid];
    }
    __syncthreads();

    // reorder in shared memory
    u32 slot[RS_IPT];
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        u32 o = wbase + k * 32 + lane;
        if (full || o < tile_n) {
            u32 d = rk[k] >> 16;
            slot[k] = sTilePrefix[d] + myHist[d] + (rk[k] & 0xFFFFu);
            sKeys[slot[k]] = key[k];
        }
    }
    if (HAS_VAL) {
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k) {
            u32 o = wbase + k * 32 + lane;
            if (o < tile_n) sVals[slot[k]] = vin[tile_base + o];
        }
    }
    __syncthreads();
    for (u32 s = tid; s < tile_n; s += RS_NT) {
        u64 kk = sKeys[s];
        u32 d = DIGIT(kk);
        u32 dst = sGlobBase[d] + s; // n < 2^31: 32-bit wrap-around arithmetic is exact
        if (USE_LUT) sPeer[d][dst] = kk; // the bins may live in other GPUs' memory (NVLink peer stores)
        else kout[dst] = kk;
        if (HAS_VAL) vout[dst] = sVals[s];
    }
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

size_t radix_smem_bytes(bool has_val) {
    return (size_t)(has_val ? 2 : 1) * RS_TILE * 8 + RS_NW * 256 * 4 + 256 * 4 + 256 * 8;
}
u32 radix_tile_size() { return RS_TILE; }

void launch_scan_hist(const u32* d_hist, u32* d_base, int npass, cudaStream_t st) {
    k_scan_hist<<<npass, 256, 0, st>>>(d_hist, d_base);
}

// the dynamic shared memory opt-in is per function AND per device (a process may hold contexts on several devices)
#define RS_MAX_DEV 64
static bool g_attr_set[RS_MAX_DEV][3] = {};

cudaError_t launch_onesweep(const u64* kin, u64* kout, const u64* vin, u64* vout, u32 n, const u32* d_digit_base,
                            u64* d_lookback, u32* d_ticket, int shift, int bits, cudaStream_t st, const u8* lut, u64* const* peers) {
    if (n == 0) return cudaSuccess;
    bool hv = vin != nullptr;
    if (lut && hv) return cudaErrorInvalidValue;
    size_t smem = radix_smem_bytes(hv);
    int variant = lut ? 2 : (hv ? 1 : 0);
    int dev = 0;
    cudaError_t de = cudaGetDevice(&dev);
    if (de != cudaSuccess) return de;
    const bool tracked = dev >= 0 && dev < RS_MAX_DEV;
    if (!tracked || !g_attr_set[dev][variant]) {
        cudaError_t e = variant == 2 ? cudaFuncSetAttribute(k_onesweep<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : variant == 1 ? cudaFuncSetAttribute(k_onesweep<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                     : cudaFuncSetAttribute(k_onesweep<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (tracked) g_attr_set[dev][variant] = true;
    }
    u32 tiles = div_up(n, RS_TILE);
    u32 dmask = (1u << bits) - 1;
    if (variant == 2) k_onesweep<false, true><<<tiles, RS_NT, smem, st>>>(kin, kout, vin, vout, n, d_digit_base, d_lookback, d_ticket, shift, dmask, lut, peers);
    else if (variant == 1) k_onesweep<true, false><<<tiles, RS_NT, smem, st>>>(kin, kout, vin, vout, n, d_digit_base, d_lookback, d_ticket, shift, dmask, lut, peers);
    else k_onesweep<false, false><<<tiles, RS_NT, smem, st>>>(kin, kout, vin, vout, n, d_digit_base, d_lookback, d_ticket, shift, dmask, lut, peers);
    return cudaGetLastError();
}

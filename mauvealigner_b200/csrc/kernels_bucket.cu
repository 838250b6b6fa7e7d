// kernels_bucket.cu — a5, a6 (bucket formation), a7, a8 of SURVEY.md §8a.
//
//  k_find_runs      equal-seed runs of the sorted records = the IdmerList buckets that libMems'
//                   MatchFinder::FindMatchSeeds hands to EnumerateMatches (call site
//                   /root/reference/src/SeedMatchEnumerator.h:61); counts distinct seeds =
//                   SortedMerList::UniqueMerCount (/root/reference/src/uniqueMerCount.cpp:39).
//  k_select         per-bucket policy:
//                   MODE_UNIQUE   /root/reference/src/UniqueMatchFinder.cpp:36-60 (keep genomes that
//                                 occur exactly once; need >= 2), MaskedMemHash mask test
//                                 (/root/reference/src/mauveAligner.cpp:525-531)
//                   MODE_SEED_ENUM /root/reference/src/SeedMatchEnumerator.h:71-123 (multiplicity
//                                 window on the un-projected bucket, forward-only projection)
//  k_emit_unique    MemHash::HashMatch + SetDirection (same body as SeedMatchEnumerator.h:127-141)
//  k_emit_enum      SeedMatchEnumerator::HashMatch component list (position+1, sign vs first)
#include "common.cuh"
#include "kernels.h"
#include "lookback.cuh"

// ------------------------------------------------------------------------------------ find runs
// Thread = FR_IPT consecutive records (16-byte loads), so all neighbour comparisons but the two at the thread's
// ends are in registers; per-thread head / unique counts are scanned over the warp, the block and (decoupled
// look-back) the tiles, and every thread writes its run heads at consecutive ranks.
#ifndef FR_NT
#define FR_NT 256
#endif
#define FR_IPT 8
#define FR_TILE (FR_NT * FR_IPT)

struct Rec { u64 key; u32 g; };

__device__ __forceinline__ Rec load_rec(const RecFmt& f, const u64* __restrict__ k, const u64* __restrict__ v, u64 i) {
    Rec r;
    u64 kk = k[i];
    if (f.wide) { r.key = kk; r.g = (u32)(v[i] >> 33); }
    else { r.key = kk >> f.kshift; r.g = (u32)((kk >> (f.pbits + 1)) & ((1u << f.gbits) - 1)); }
    return r;
}

// Two streaming kernels and a one-block scan between them instead of one kernel with a decoupled look-back: the ncu
// source page of the single-pass version had 62 % of its stall samples at the barrier behind the warp that waits for its
// predecessors (tiles of 2048 records are over before the look-back answers).  Pass A reads the records once and leaves
// two bytes per thread (head / unique masks of its 8 records) and one (heads, uniques) pair per tile; pass B turns the
// masks into the compact run arrays with the scanned tile offsets — no dependency between tiles.
__global__ void __launch_bounds__(FR_NT) k_find_runs_a(const u64* __restrict__ keys, const u64* __restrict__ vals, u32 n, RecFmt fmt,
                                                       unsigned short* __restrict__ masks, u64* __restrict__ tile_counts,
                                                       u64* __restrict__ per_seq_count /*[nseq] or null*/) {
    __shared__ u32 sWarp[FR_NT / 32];
    __shared__ u32 sSeq[MB_MAX_SEQ];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (per_seq_count) {
        if (tid < MB_MAX_SEQ) sSeq[tid] = 0;
        __syncthreads();
    }
    const u32 tile = blockIdx.x;
    const u64 base = (u64)tile * FR_TILE;
    const u64 i0 = base + (u64)tid * FR_IPT;
    // records i0 .. i0+7 (keys beyond n read as a key no record has)
    u64 key[FR_IPT];
    u32 g[FR_IPT];
    if (i0 + FR_IPT <= n) {
        u64 raw[FR_IPT], rv[FR_IPT];
#pragma unroll
        for (int k = 0; k < FR_IPT; k += 2) {
            ulonglong2 q = *reinterpret_cast<const ulonglong2*>(keys + i0 + k);
            raw[k] = q.x; raw[k + 1] = q.y;
            if (fmt.wide) {
                ulonglong2 w = *reinterpret_cast<const ulonglong2*>(vals + i0 + k);
                rv[k] = w.x; rv[k + 1] = w.y;
            }
        }
#pragma unroll
        for (int k = 0; k < FR_IPT; ++k) {
            if (fmt.wide) { key[k] = raw[k]; g[k] = (u32)(rv[k] >> 33); }
            else { key[k] = raw[k] >> fmt.kshift; g[k] = (u32)((raw[k] >> (fmt.pbits + 1)) & ((1u << fmt.gbits) - 1)); }
        }
    } else {
#pragma unroll
        for (int k = 0; k < FR_IPT; ++k) {
            key[k] = ~0ull; g[k] = 0xFFFFFFFFu;
            if (i0 + k < n) { Rec r = load_rec(fmt, keys, vals, i0 + k); key[k] = r.key; g[k] = r.g; }
        }
    }
    // neighbours across the thread ends: shuffles inside the warp, global loads at the warp edges
    u64 pkey = __shfl_up_sync(0xFFFFFFFFu, key[FR_IPT - 1], 1);
    u32 pg = __shfl_up_sync(0xFFFFFFFFu, g[FR_IPT - 1], 1);
    u64 nkey = __shfl_down_sync(0xFFFFFFFFu, key[0], 1);
    u32 ng = __shfl_down_sync(0xFFFFFFFFu, g[0], 1);
    if (lane == 0) { pkey = ~0ull; pg = 0xFFFFFFFFu; if (i0 > 0 && i0 <= n) { Rec r = load_rec(fmt, keys, vals, i0 - 1); pkey = r.key; pg = r.g; } }
    if (lane == 31) { nkey = ~0ull; ng = 0xFFFFFFFFu; if (i0 + FR_IPT < n) { Rec r = load_rec(fmt, keys, vals, i0 + FR_IPT); nkey = r.key; ng = r.g; } }
    u32 headm = 0, uniqm = 0;
    u32 run_g = 0xFFFFFFFFu, run_c = 0; // per-genome count of "first record of its genome in its bucket", flushed on change
#pragma unroll
    for (int k = 0; k < FR_IPT; ++k) {
        const u64 i = i0 + k;
        const bool valid = i < n;
        const u64 pk = k ? key[k - 1] : pkey, nk = k + 1 < FR_IPT ? key[k + 1] : nkey;
        const u32 pgk = k ? g[k - 1] : pg, ngk = k + 1 < FR_IPT ? g[k + 1] : ng;
        const bool head = valid && (i == 0 || pk != key[k]);
        const bool newg = valid && (head || pgk != g[k]);
        const bool last = valid && (i + 1 >= n || nk != key[k] || ngk != g[k]);
        headm |= (head ? 1u : 0u) << k;
        uniqm |= (newg && last ? 1u : 0u) << k;
        if (per_seq_count && newg) {
            if (g[k] != run_g) { if (run_c) atomicAdd(&sSeq[run_g], run_c); run_g = g[k]; run_c = 0; }
            ++run_c;
        }
    }
    if (per_seq_count && run_c) atomicAdd(&sSeq[run_g], run_c);
    masks[(u64)tile * FR_NT + tid] = (unsigned short)(headm | (uniqm << 8));
    // tile totals (heads | uniques << 16)
    u32 x = __reduce_add_sync(0xFFFFFFFFu, (u32)__popc(headm) | ((u32)__popc(uniqm) << 16));
    if (lane == 0) sWarp[warp] = x;
    __syncthreads();
    if (tid == 0) {
        u32 tot = 0;
#pragma unroll
        for (int w = 0; w < FR_NT / 32; ++w) tot += sWarp[w];
        tile_counts[tile] = (u64)(tot & 0xFFFF) | ((u64)(tot >> 16) << 32);
    }
    if (per_seq_count && tid < MB_MAX_SEQ && sSeq[tid]) atomicAdd((unsigned long long*)&per_seq_count[tid], (unsigned long long)sSeq[tid]);
}

// exclusive scan of the tile pairs (one block), totals and sentinels
__global__ void __launch_bounds__(1024) k_find_runs_scan(u64* __restrict__ tile_counts, u32 n_tiles, u32 n, u32* __restrict__ run_start,
                                                         u32* __restrict__ run_u, u32* __restrict__ totals) {
    __shared__ u64 scratch[1024 / 32 + 1];
    u64 carry = 0;
    for (u32 base = 0; base < n_tiles; base += 1024 * 8) { // 8 consecutive tiles per thread and step
        const u32 t0 = base + threadIdx.x * 8;
        u64 v[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] = t0 + j < n_tiles ? tile_counts[t0 + j] : 0; sum += v[j]; }
        u64 total;
        u64 ex = carry + block_excl_scan_u64<1024>(sum, scratch, total);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (t0 + j < n_tiles) tile_counts[t0 + j] = ex;
            ex += v[j];
        }
        carry += total;
    }
    if (threadIdx.x == 0) {
        const u32 nr = (u32)carry, nu = (u32)(carry >> 32);
        totals[0] = nr; totals[1] = nu;
        run_start[nr] = n;
        if (run_u) run_u[nr] = nu;
    }
}

__global__ void __launch_bounds__(FR_NT) k_find_runs_b(const unsigned short* __restrict__ masks, const u64* __restrict__ tile_excl, u32 n,
                                                       u32* __restrict__ run_start, u32* __restrict__ run_u) {
    __shared__ u32 sWarp[FR_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 tile = blockIdx.x;
    const u64 i0 = (u64)tile * FR_TILE + (u64)tid * FR_IPT;
    const u32 mk = masks[(u64)tile * FR_NT + tid];
    const u32 headm = mk & 0xFFu, uniqm = mk >> 8;
    const u32 mine = (u32)__popc(headm) | ((u32)__popc(uniqm) << 16);
    u32 x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sWarp[warp] = x;
    __syncthreads();
    u32 wpre = 0;
#pragma unroll
    for (int w = 0; w < FR_NT / 32; ++w) wpre += (w < warp) ? sWarp[w] : 0u;
    const u32 pre = wpre + x - mine;
    const u64 te = tile_excl[tile];
    u32 rh = (u32)te + (pre & 0xFFFF), ru = (u32)(te >> 32) + (pre >> 16);
#pragma unroll
    for (int k = 0; k < FR_IPT; ++k) {
        if ((headm >> k) & 1) {
            run_start[rh] = (u32)(i0 + k);
            if (run_u) run_u[rh] = ru;
            ++rh;
        }
        ru += (uniqm >> k) & 1;
    }
    (void)n;
}

// status: workspace of at least 2 * tiles + FR_NT/4 * tiles ... words (see find_runs_workspace_words)
const unsigned short* launch_find_runs(const u64* keys, const u64* vals, u32 n, const RecFmt& fmt, u32* run_start, u32* run_u, u64* status,
                                       u32* ticket, u64* per_seq_count, u32* totals, cudaStream_t st) {
    (void)ticket;
    if (n == 0) return nullptr;
    const u32 tiles = div_up(n, FR_TILE);
    u64* tile_counts = status;                                                   // [tiles]
    unsigned short* masks = reinterpret_cast<unsigned short*>(status + tiles + 1); // [tiles * FR_NT]
    k_find_runs_a<<<tiles, FR_NT, 0, st>>>(keys, vals, n, fmt, masks, tile_counts, per_seq_count);
    k_find_runs_scan<<<1, 1024, 0, st>>>(tile_counts, tiles, n, run_start, run_u, totals);
    k_find_runs_b<<<tiles, FR_NT, 0, st>>>(masks, tile_counts, n, run_start, run_u);
    return masks; // one 16-bit word per 8 records: bit k = record 8 j + k heads a run, bit 8 + k = it is the only record of its genome in its run
}
// 8-byte words of workspace launch_find_runs needs behind `status`
size_t find_runs_workspace_words(u32 n) {
    const size_t tiles = div_up(n, FR_TILE);
    return tiles + 1 + (tiles * FR_NT * 2 + 7) / 8 + 1;
}
u32 find_runs_tile() { return FR_TILE; }

// --------------------------------------------------------------------------------------- select
#ifndef SL_NT
#define SL_NT 256
#endif
#ifndef SL_IPT
#define SL_IPT 8 // measured: 4 -> 8 runs per thread takes 8 % off find_runs + select (C2 0.49 -> 0.45 ms, C5 2.69 -> 2.47 ms)
#endif
#define SL_TILE (SL_NT * SL_IPT)

// value packing for the (candidates, components) pair: 30 + 32 bits
#define SL_PACK(c, m) ((u64)(c) | ((u64)(m) << 30))

__global__ void __launch_bounds__(SL_NT) k_select(SelectArgs a, RecFmt fmt) {
    __shared__ u64 scratch[SL_NT / 32 + 1];
    __shared__ u32 sTile;
    __shared__ u64 sExcl;
    const int tid = threadIdx.x;
    if (tid == 0) sTile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const u32 tile = sTile;
    const u32 n_runs = *a.n_runs_ptr;
    if ((u64)tile * SL_TILE >= n_runs) return; // grid is sized for the upper bound n_runs <= n
    const u64 r0 = (u64)tile * SL_TILE + (u64)tid * SL_IPT;
    u32 ncand[SL_IPT], ncomp[SL_IPT];
    u32 nb = 0;
    u64 mine = 0;
#pragma unroll
    for (int j = 0; j < SL_IPT; ++j) {
        u64 r = r0 + j;
        ncand[j] = 0; ncomp[j] = 0;
        if (r >= n_runs) continue;
        u32 s = a.run_start[r], e = a.run_start[r + 1];
        u32 len = e - s;
        if (len < 2) continue;
        ++nb;
        if (a.mode == MB_MODE_SEED_ENUM_) {
            if ((u64)len > a.max_multi || (u64)len < a.min_multi) continue;
            u32 m = len;
            if (a.direct_only) {
                // forward = same strand bit as the first (lowest position) occurrence
                u64 v0 = fmt.wide ? a.vals[s] : a.keys[s];
                u32 s0 = rec_strand(v0), fw = 0;
                for (u32 i = s; i < e; ++i) {
                    u64 v = fmt.wide ? a.vals[i] : a.keys[i];
                    fw += rec_strand(v) == s0;
                }
                if (fw != len) { // found_reverse: forward-only projection, needs > 1 component
                    if (fw < 2) continue;
                    m = fw;
                }
            }
            ncand[j] = 1; ncomp[j] = m;
        } else if (a.mode == MB_MODE_REPEAT_) {
            // RepeatHash: every occurrence is a component (one sequence); the candidate formats hold up to 255 components
            if ((u64)len > a.max_multi || (u64)len < a.min_multi || len > 255) continue;
            ncand[j] = 1; ncomp[j] = len;
        } else {
            u32 u = a.run_u[r + 1] - a.run_u[r];
            if (u < 2) continue;
            if (a.mode == MB_MODE_PAIRWISE_) {
                ncand[j] = u * (u - 1) / 2; ncomp[j] = u * (u - 1);
            } else {
                if (a.nway_mask) {
                    if (u != (u32)__popcll(a.nway_mask)) continue;
                    u64 present = 0;
                    for (u32 i = s; i < e; ++i) {
                        u32 g = rec_genome(fmt, fmt.wide ? a.vals[i] : a.keys[i]);
                        bool first = i == s || rec_genome(fmt, fmt.wide ? a.vals[i - 1] : a.keys[i - 1]) != g;
                        bool lastg = i + 1 == e || rec_genome(fmt, fmt.wide ? a.vals[i + 1] : a.keys[i + 1]) != g;
                        if (first && lastg) present |= 1ull << g;
                    }
                    if (present != a.nway_mask) continue;
                }
                ncand[j] = 1; ncomp[j] = u;
            }
        }
        mine += SL_PACK(ncand[j], ncomp[j]);
    }
    u64 total;
    u64 ex = block_excl_scan_u64<SL_NT>(mine, scratch, total);
    // bucket statistic
    {
        u32 wnb = __reduce_add_sync(0xFFFFFFFFu, nb);
        if ((tid & 31) == 0 && wnb) atomicAdd((unsigned long long*)a.n_buckets, (unsigned long long)wnb);
    }
    if (tid < 32) {
        u64 excl = lookback_exclusive(a.status, tile, total);
        if (tid == 0) {
            sExcl = excl;
            if ((u64)(tile + 1) * SL_TILE >= n_runs) {
                u64 tot = excl + total;
                u32 nc = (u32)(tot & ((1u << 30) - 1));
                u32 nm = (u32)(tot >> 30);
                a.totals[0] = nc; a.totals[1] = nm;
                a.cand_off[nc] = nm;
            }
        }
    }
    __syncthreads();
    u64 pre = sExcl + ex;
#pragma unroll
    for (int j = 0; j < SL_IPT; ++j) {
        if (!ncand[j]) continue;
        u32 c = (u32)(pre & ((1u << 30) - 1));
        u32 off = (u32)(pre >> 30);
        u32 per = ncomp[j] / ncand[j];
        for (u32 t = 0; t < ncand[j]; ++t) {
            a.cand_run[c + t] = (u32)(r0 + j);
            a.cand_off[c + t] = off + t * per;
            if (a.cand_aux) a.cand_aux[c + t] = t;
        }
        pre += SL_PACK(ncand[j], ncomp[j]);
    }
}

void launch_select(const SelectArgs& a, const RecFmt& fmt, u32 n_runs_upper, cudaStream_t st) {
    if (n_runs_upper == 0) return;
    k_select<<<div_up(n_runs_upper, SL_TILE), SL_NT, 0, st>>>(a, fmt);
}
u32 select_tile() { return SL_TILE; }

// ---------------------------------------------------------------------------------- emit unique
// One thread per candidate: gather the unique-genome members of its bucket into the candidate CSR
// (comp_pos, comp_gs = genome | reverse<<7) and set the candidate's bit in the (first genome,
// position) bitmap used by the de-dup stage.
__global__ void __launch_bounds__(256) k_emit_unique(EmitUniqueArgs a, RecFmt fmt, GenomeTable gt) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    u32 nc = a.totals[0];
    if (c >= nc) return;
    u32 r = a.cand_run[c];
    u32 s = a.run_start[r], e = a.run_start[r + 1];
    u32 off = a.cand_off[c];
    u32 k = 0, strand0 = 0;
    int pa = -1, pb = -1;
    if (a.mode == MB_MODE_PAIRWISE_) {
        // pair index t -> (pa, pb) with pa < pb in lexicographic order over the unique list
        u32 u = a.run_u[r + 1] - a.run_u[r];
        u32 t = a.cand_aux[c];
        u32 x = 0;
        while (t >= u - 1 - x) { t -= u - 1 - x; ++x; }
        pa = (int)x; pb = (int)(x + 1 + t);
    }
    u32 ui = 0;
    u32 x0 = 0, g_first = 0, g_second = 0;
    u64 h = 0x9E3779B97F4A7C15ull, h2 = 0xC2B2AE3D27D4EB4Full;
    for (u32 i = s; i < e; ++i) {
        // "the only record of its genome in its bucket" was worked out by k_find_runs_a: bit 8 + (i & 7) of the mask word of
        // the 8 records around i (RepeatHash takes every occurrence, in position order)
        if (a.mode != MB_MODE_REPEAT_ && !((a.masks[i >> 3] >> (8 + (i & 7))) & 1u)) continue;
        u64 v = fmt.wide ? a.vals[i] : a.keys[i];
        u32 g = rec_genome(fmt, v);
        bool take = pa < 0 || (int)ui == pa || (int)ui == pb;
        ++ui;
        if (!take) continue;
        u32 p = rec_pos(fmt, v), sb = rec_strand(v);
        if (k == 0) { strand0 = sb; x0 = p; g_first = g; }
        if (k == 1) g_second = g;
        bool rev = sb != strand0;
        a.comp_pos[off + k] = p;
        a.comp_gs[off + k] = (u8)(g | (rev ? 0x80u : 0u));
        // hash of the D16 group key: genome, strand and diagonal of every component
        u64 dg = rev ? (u64)p + x0 + a.seedL : (u64)(u32)(p - x0);
        h ^= (dg << 8) | (u64)(g | (rev ? 0x80u : 0u));
        h *= 0xFF51AFD7ED558CCDull; h ^= h >> 29;
        h2 = (h2 + ((dg << 8) | (u64)(g | (rev ? 0x80u : 0u)))) * 0x9FB21C651E98DF25ull; h2 ^= h2 >> 32;
        ++k;
    }
    a.ghash[c] = h;
    a.ghash2[c] = h2 & ~0xFFull;
    if (a.bitmap) {
        u64 gp = gt.vbase[vgenome(gt, g_first, g_second)] + x0;
        atomicOr((unsigned long long*)&a.bitmap[gp >> 6], 1ull << (gp & 63));
    }
}

void launch_emit_unique(const EmitUniqueArgs& a, const RecFmt& fmt, const GenomeTable& gt, u32 n_cand_upper, cudaStream_t st) {
    if (n_cand_upper == 0) return;
    k_emit_unique<<<div_up(n_cand_upper, 256), 256, 0, st>>>(a, fmt, gt);
}

// ------------------------------------------------------------------------------------ emit enum
// MODE_SEED_ENUM: sort key of a match = its first emitted position (unique per bucket).
__global__ void __launch_bounds__(256) k_enum_keys(EmitEnumArgs a, RecFmt fmt) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    u32 nc = a.totals[0];
    if (c >= nc) return;
    u32 r = a.cand_run[c];
    u32 s = a.run_start[r];
    u64 v = fmt.wide ? a.vals[s] : a.keys[s];
    a.sort_key[c] = (u64)rec_pos(fmt, v);
    a.sort_val[c] = c;
    a.ncomp[c] = a.cand_off[c + 1] - a.cand_off[c];
}

// after sorting: out slot j <- candidate perm[j]; comps written at out_off[j]
__global__ void __launch_bounds__(256) k_enum_gather(EmitEnumArgs a, RecFmt fmt, u32 seedL) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    u32 nc = a.totals[0];
    if (j >= nc) return;
    u32 c = (u32)a.sorted_val[j];
    u32 r = a.cand_run[c];
    u32 s = a.run_start[r], e = a.run_start[r + 1];
    u32 m = a.cand_off[c + 1] - a.cand_off[c];
    bool project = m != e - s;
    u64 o = a.out_off[j];
    a.out_len[j] = seedL;
    u32 s0 = rec_strand(fmt.wide ? a.vals[s] : a.keys[s]);
    for (u32 i = s; i < e; ++i) {
        u64 v = fmt.wide ? a.vals[i] : a.keys[i];
        bool rev = rec_strand(v) != s0;
        if (project && rev) continue;
        const int32_t st = (int32_t)rec_pos(fmt, v) + 1;
        a.out_seq[o] = 0;
        a.out_start[o] = rev ? -st : st;
        ++o;
    }
}

void launch_enum_keys(const EmitEnumArgs& a, const RecFmt& fmt, u32 n_upper, cudaStream_t st) {
    if (n_upper) k_enum_keys<<<div_up(n_upper, 256), 256, 0, st>>>(a, fmt);
}
void launch_enum_gather(const EmitEnumArgs& a, const RecFmt& fmt, u32 seedL, u32 n_upper, cudaStream_t st) {
    if (n_upper) k_enum_gather<<<div_up(n_upper, 256), 256, 0, st>>>(a, fmt, seedL);
}

// ---------------------------------------------------------------------------------- generic scan
// out[i] = sum_{j<i} f(in, j); total -> *total_out.  Single pass, decoupled look-back.
#define SC_NT 256
#define SC_IPT 8
#define SC_TILE (SC_NT * SC_IPT)
template <int KIND> // 0: u32 values, 1: popcount of u64 words
__global__ void __launch_bounds__(SC_NT) k_scan(const void* __restrict__ in, u64 n, u32* __restrict__ out32, u64* __restrict__ out64,
                                                u64* status, u32* ticket, u64* total_out) {
    __shared__ u64 scratch[SC_NT / 32 + 1];
    __shared__ u32 sTile;
    __shared__ u64 sExcl;
    const int tid = threadIdx.x;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = sTile;
    const u64 i0 = (u64)tile * SC_TILE + (u64)tid * SC_IPT;
    u32 v[SC_IPT];
    u64 mine = 0;
#pragma unroll
    for (int j = 0; j < SC_IPT; ++j) {
        u64 i = i0 + j;
        v[j] = 0;
        if (i < n) v[j] = KIND == 0 ? reinterpret_cast<const u32*>(in)[i] : (u32)__popcll(reinterpret_cast<const u64*>(in)[i]);
        mine += v[j];
    }
    u64 total;
    u64 ex = block_excl_scan_u64<SC_NT>(mine, scratch, total);
    if (tid < 32) {
        u64 excl = lookback_exclusive(status, tile, total);
        if (tid == 0) {
            sExcl = excl;
            if ((u64)(tile + 1) * SC_TILE >= n) {
                if (total_out) *total_out = excl + total;
                if (out64) out64[n] = excl + total;
                if (out32) out32[n] = (u32)(excl + total);
            }
        }
    }
    __syncthreads();
    u64 pre = sExcl + ex;
#pragma unroll
    for (int j = 0; j < SC_IPT; ++j) {
        u64 i = i0 + j;
        if (i < n) {
            if (out64) out64[i] = pre;
            if (out32) out32[i] = (u32)pre;
        }
        pre += v[j];
    }
}

u32 scan_tile() { return SC_TILE; }
void launch_scan_u32(const u32* in, u64 n, u32* out32, u64* out64, u64* status, u32* ticket, u64* total_out, cudaStream_t st) {
    if (n == 0) return;
    k_scan<0><<<div_up(n, SC_TILE), SC_NT, 0, st>>>(in, n, out32, out64, status, ticket, total_out);
}
void launch_scan_popc(const u64* in, u64 n, u32* out32, u64* status, u32* ticket, u64* total_out, cudaStream_t st) {
    if (n == 0) return;
    k_scan<1><<<div_up(n, SC_TILE), SC_NT, 0, st>>>(in, n, out32, nullptr, status, ticket, total_out);
}

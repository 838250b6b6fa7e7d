// kernels_seed.cu — a2 (2-bit packing) and a3 (per-position canonical seed mer) of SURVEY.md §8a.
//
// Replaces libMems SortedMerList::SetSequence/translate and FillDnaSeedSML/GetSeedMer (absent from
// /root/reference; semantics = SURVEY.md Appendix A D1-D4).  The strand bit emitted here is the one
// /root/reference/src/SeedMatchEnumerator.h:133,139 reads through GetMer(pos) & 1.
#include "common.cuh"
#include "kernels.h"

// ------------------------------------------------------------------------------------------ pack
// One thread = 16 bases = one 32-bit half of a D2 word.  16-byte vector load of ASCII.
__device__ __forceinline__ u32 code_of(u32 c) {
    u32 u = c & 0xDFu; // upper-case
    u32 x = (c >> 1) & 3u;
    x ^= x >> 1; // A->0 C->1 G->2 T->3
    bool ok = (u == 0x41u) | (u == 0x43u) | (u == 0x47u) | (u == 0x54u);
    return ok ? x : 0u;
}

__global__ void __launch_bounds__(256) k_pack(const u8* __restrict__ ascii, u64 len, u32* __restrict__ out32, u64 n_half) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_half) return;
    u64 b0 = t * 16;
    u32 v = 0;
    if (b0 + 16 <= len && ((((uintptr_t)ascii) & 15) == 0)) {
        uint4 q = *reinterpret_cast<const uint4*>(ascii + b0);
        u32 ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) v = (v << 2) | code_of((ws[i] >> (8 * j)) & 0xFFu);
    } else {
        for (int j = 0; j < 16; ++j) {
            u64 b = b0 + j;
            v = (v << 2) | (b < len ? code_of(ascii[b]) : 0u);
        }
    }
    // D2 word k = (first 16 bases) << 32 | (next 16 bases); little-endian memory: low half first
    out32[t ^ 1] = v;
}

void launch_pack(const u8* d_ascii, u64 len, u64* d_words, u64 n_words, cudaStream_t st) {
    u64 n_half = n_words * 2;
    if (n_half == 0) return;
    k_pack<<<div_up(n_half, 256), 256, 0, st>>>(d_ascii, len, reinterpret_cast<u32*>(d_words), n_half);
}

// --------------------------------------------------------------------------------------- extract
// Tile = EX_TILE consecutive seed positions of one genome, its packed words (plus halo) staged in shared memory; the
// forward window of a position is one unaligned 128-bit funnel extraction.  The pattern is palindromic, so the cared
// bases of the reverse-complement window are the reverse complement of the cared bases of the forward window: the
// reverse key is rc(forward key) — one bit reversal instead of a second window (no reverse-complemented copy of the
// tile, no second extraction, mask and 128-bit compare).  Thread t handles positions t, t+NT, ... so record stores
// are fully coalesced.
#define EX_NT 256
#define EX_IPT 16
#define EX_TILE (EX_NT * EX_IPT)
#define EX_WORDS (EX_TILE / 32 + 4) // tile words + halo for L <= 64 (+2) + 2 words read-ahead

struct ExtractArgs {
    const u64* packed;
    u64* keys;      // packed records or keys
    u64* vals;      // wide only
    u64* mers;      // MODE 2 only: per-position mers of one genome
    u32* hist;      // [npass][256] or null
    int npass;
    int only_genome; // MODE 2
    u32 tile0;       // first tile of this launch (a rank's slice of the tiles in the multi-GPU path)
    u32 out_base;    // record index of the slice's first seed
};

__device__ __forceinline__ u64 gather_key(const SeedDev& sd, u64 hi, u64 lo) {
    u64 k = 0;
    for (int r = 0; r < sd.nrun; ++r) {
        u64 a = hi & sd.runmask_hi[r];
        int s = sd.lshift[r];
        if (sd.wide) {
            u64 b = lo & sd.runmask_lo[r];
            k |= (s >= 64) ? (b << (s - 64)) : shl128_hi(a, b, s);
        } else {
            k |= a << s;
        }
    }
    return k; // left aligned: 2w bits at the top
}

template <int MODE> // 0: packed records, 1: wide records, 2: mers dump
__global__ void __launch_bounds__(EX_NT) k_extract(ExtractArgs a, GenomeTable gt, SeedDev sd, RecFmt fmt, const u32* __restrict__ tile_first /*[nseq+1]*/) {
    __shared__ u64 sF[EX_WORDS + 2];
    __shared__ u32 sHist[8 * 256];
    __shared__ u32 sG;

    const int tid = threadIdx.x;
    if (tid == 0) {
        u32 g = 0;
        if (MODE == 2) g = a.only_genome;
        else
            while (g + 1 < gt.nseq && blockIdx.x + a.tile0 >= tile_first[g + 1]) ++g;
        sG = g;
    }
    if (a.hist)
        for (int i = tid; i < a.npass * 256; i += EX_NT) sHist[i] = 0;
    __syncthreads();
    const u32 g = sG;
    const u32 tile_in_g = (MODE == 2) ? blockIdx.x : blockIdx.x + a.tile0 - tile_first[g];
    const u32 len = gt.len[g];
    const u32 nseeds = len >= (u32)sd.L ? len - sd.L + 1 : 0;
    const u32 p0 = tile_in_g * EX_TILE;
    const u64* gw = a.packed + gt.word_base[g];
    const int m = EX_TILE / 32 + 2; // staged span in words (covers L-1 <= 63 halo bases)

    for (int i = tid; i < m + 2; i += EX_NT) sF[i] = gw[(p0 >> 5) + i]; // padding words after each genome keep this in bounds
    __syncthreads();

    const u32 kshift_la = 64 - 2 * sd.w; // (fmt.kbits = 2w, plus the problem index bits of a segmented search)
#pragma unroll 4
    for (int it = 0; it < EX_IPT; ++it) {
        u32 o = it * EX_NT + tid;
        u32 p = p0 + o;
        if (p >= nseeds) continue;
        u64 fh, fl;
        load_window(sF, o, sd.wide, fh, fl);
        const u64 kf_la = gather_key(sd, fh & sd.mask_hi, fl & sd.mask_lo);     // forward key, left-aligned (2w bits at the top)
        const u64 kf = kf_la >> kshift_la;
        const u64 kr = rc_word(kf_la) & (~0ull >> kshift_la);                   // rc of the 32-base word: the key's rc in the low 2w bits
        const bool strand = kr < kf;                                            // D4: the smaller of the two, ties forward
        u64 key = strand ? kr : kf;
        if (MODE != 2 && gt.n_seg) { // segmented search: the problem index leads the key; a window across a boundary gets a key of its own
            u32 lo, hi;
            const u32 si = seg_range(gt, g, p, lo, hi);
            key = (p + sd.L <= hi) ? (((u64)si << gt.seg_field) | key) : (((u64)gt.n_seg << gt.seg_field) | ((u64)g << fmt.pbits) | p);
        }
        if (MODE == 2) {
            a.mers[p] = (key << kshift_la) | (u64)strand;
        } else {
            u32 idx = gt.seed_base[g] + p - a.out_base;
            if (MODE == 0) {
                a.keys[idx] = (((((key << fmt.gbits) | g) << fmt.pbits) | p) << 1) | (u64)strand;
            } else {
                a.keys[idx] = key;
                a.vals[idx] = ((u64)g << 33) | ((u64)p << 1) | (u64)strand;
            }
            if (a.hist) {
                for (int ps = 0; ps < a.npass; ++ps) atomicAdd(&sHist[ps * 256 + (u32)((key >> (8 * ps)) & 255u)], 1u);
            }
        }
    }
    if (MODE != 2 && a.hist) {
        __syncthreads();
        for (int i = tid; i < a.npass * 256; i += EX_NT) {
            u32 c = sHist[i];
            if (c) atomicAdd(&a.hist[i], c);
        }
    }
}

u32 extract_tile_size() { return EX_TILE; }

void launch_extract_records(const u64* d_packed, u64* d_keys, u64* d_vals, u32* d_hist, int npass, const GenomeTable& gt,
                            const SeedDev& sd, const RecFmt& fmt, const u32* d_tile_first, u32 n_tiles, cudaStream_t st, u32 tile0,
                            u32 out_base) {
    if (n_tiles == 0) return;
    ExtractArgs a{};
    a.packed = d_packed; a.keys = d_keys; a.vals = d_vals; a.hist = d_hist; a.npass = npass; a.tile0 = tile0; a.out_base = out_base;
    if (fmt.wide) k_extract<1><<<n_tiles, EX_NT, 0, st>>>(a, gt, sd, fmt, d_tile_first);
    else k_extract<0><<<n_tiles, EX_NT, 0, st>>>(a, gt, sd, fmt, d_tile_first);
}

void launch_mers(const u64* d_packed, const GenomeTable& gt, const SeedDev& sd, const RecFmt& fmt, int genome, u64* d_mers,
                 cudaStream_t st) {
    u32 len = gt.len[genome];
    if (len < (u32)sd.L) return;
    u32 n = len - sd.L + 1;
    ExtractArgs a{};
    a.packed = d_packed; a.mers = d_mers; a.only_genome = genome;
    k_extract<2><<<div_up(n, EX_TILE), EX_NT, 0, st>>>(a, gt, sd, fmt, nullptr);
}

// ------------------------------------------------------------------------------------- histogram
// Stand-alone digit histogram for sorts whose keys are not produced by k_extract.
__global__ void __launch_bounds__(256) k_hist(const u64* __restrict__ keys, u32 n, int shift0, int kbits, int npass, u32* __restrict__ hist) {
    __shared__ u32 sHist[8 * 256];
    for (int i = threadIdx.x; i < npass * 256; i += 256) sHist[i] = 0;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
        u64 k = keys[i] >> shift0;
        if (kbits < 64) k &= (1ull << kbits) - 1;
        for (int ps = 0; ps < npass; ++ps) atomicAdd(&sHist[ps * 256 + (u32)((k >> (8 * ps)) & 255u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += 256) {
        u32 c = sHist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

void launch_hist(const u64* d_keys, u32 n, int shift0, int kbits, int npass, u32* d_hist, cudaStream_t st) {
    if (n == 0) return;
    u32 blocks = div_up(n, 256 * 16);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_hist<<<blocks, 256, 0, st>>>(d_keys, n, shift0, kbits, npass, d_hist);
}

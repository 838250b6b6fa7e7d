// common.cuh — shared types of the sm_100a seed-match pipeline.
//
// Data layout in HBM:
//   packed genomes   uint64 words, 32 bases/word, first base in the top bits (D2); every genome
//                    starts on a 16-byte boundary and is followed by >= 4 zero words of padding
//   seed records     R=8 : one uint64  [ key:2w | genome:gbits | pos:pbits | strand:1 ]  (right aligned)
//                    R=16: keys[i] = key (2w bits), vals[i] = genome<<33 | pos<<1 | strand
//   runs             u32 start index per maximal equal-key run of the sorted records
//   candidates       CSR (cand_off, comp_pos u32, comp_gs u8 = genome | rev<<7), in ascending key order
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;
typedef uint8_t u8;

#define MB_MAX_SEQ 64
#define MB_PAD_WORDS 4

// Seed descriptor (host-built, passed by value to kernels). Windows are handled LEFT-ALIGNED in
// 128 bits (hi:lo), first base in the top two bits of hi; mask2_* is the care mask in that frame.
struct SeedDev {
    int L, w;
    int wide;        // L > 32: windows need the lo word too
    int nrun;        // care runs, most significant first
    u64 mask_hi, mask_lo;
    // gather step r: out |= ((win & runmask[r]) << lshift[r]); operates on the left-aligned window,
    // result left-aligned in 64 bits (2w <= 62 significant bits)
    u64 runmask_hi[32], runmask_lo[32];
    int lshift[32];
    u8 care_off[32];  // the w cared window offsets, ascending
};

// Record format
struct RecFmt {
    int wide;            // 0: packed u64, 1: key u64 + val u64
    int kbits;           // 2w
    int gbits, pbits;    // packed only
    int kshift;          // bit position of the key's LSB inside the sorted word (0 when wide)
};

__host__ __device__ __forceinline__ u64 rec_key(const RecFmt& f, u64 k) { return f.wide ? k : (k >> f.kshift); }
__host__ __device__ __forceinline__ u32 rec_genome(const RecFmt& f, u64 v) {
    return f.wide ? (u32)(v >> 33) : (u32)((v >> (f.pbits + 1)) & ((1u << f.gbits) - 1));
}
__host__ __device__ __forceinline__ u32 rec_pos(const RecFmt& f, u64 v) {
    return f.wide ? (u32)(v >> 1) : (u32)((v >> 1) & ((1ull << f.pbits) - 1));
}
__host__ __device__ __forceinline__ u32 rec_strand(u64 v) { return (u32)(v & 1); }

struct GenomeTable {
    u64 word_base[MB_MAX_SEQ];   // first packed word of genome g
    u64 base_base[MB_MAX_SEQ];   // first global base index of genome g (for the candidate bitmap)
    u32 len[MB_MAX_SEQ];
    u32 seed_base[MB_MAX_SEQ];   // first record index of genome g in extraction order
    u32 nseq;
    // candidate bitmap axis: one "virtual genome" per first genome (MODE_UNIQUE) or per genome pair (MODE_PAIRWISE,
    // where several candidates of one bucket share their first component)
    u32 pairwise;
    u64 vbase[MB_MAX_SEQ];       // first bitmap index of virtual genome v
    // Segmented search (many small problems in one pass, mb_find_batch): every genome is a concatenation of n_seg
    // segments, segment i of every genome forming problem i.  seg[g * (n_seg + 1) + i] = first base of segment i of
    // genome g (ascending, the last entry = the genome's length).  Seeds only meet seeds of the same problem, a window
    // never crosses a segment boundary and extension stops at it.  n_seg == 0: plain search.
    const u32* seg;
    u32 n_seg;
    u32 seg_field;               // key bits below the problem index: max(2w, genome bits + position bits)
};
// segment [lo, hi) of genome g that holds base p (the whole genome without segments)
__device__ __forceinline__ u32 seg_range(const GenomeTable& gt, u32 g, u32 p, u32& lo, u32& hi) {
    if (!gt.n_seg) { lo = 0; hi = gt.len[g]; return 0; }
    const u32* b = gt.seg + (size_t)g * (gt.n_seg + 1);
    u32 a = 0, z = gt.n_seg; // largest i with b[i] <= p
    while (z - a > 1) { u32 mid = (a + z) >> 1; if (b[mid] <= p) a = mid; else z = mid; }
    lo = b[a]; hi = b[a + 1];
    return a;
}
__host__ __device__ __forceinline__ u32 vgenome(const GenomeTable& gt, u32 g0, u32 g1) {
    return gt.pairwise ? g0 * gt.nseq - g0 * (g0 + 1) / 2 + (g1 - g0 - 1) : g0;
}

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ u64 shl128_hi(u64 hi, u64 lo, int s) { // top 64 bits of (hi:lo) << s, 0 <= s < 64
    return s ? ((hi << s) | (lo >> (64 - s))) : hi;
}
// reverse-complement all 32 bases of a word
__device__ __forceinline__ u64 rc_word(u64 x) {
    x = __brevll(x);
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    return ~x;
}

// Load the left-aligned 128-bit window starting at base `p` of the packed stream `w` (global or
// shared).  Reads words p/32 .. p/32+2; callers guarantee the padding makes that safe.
__device__ __forceinline__ void load_window(const u64* __restrict__ w, u64 p, bool wide, u64& hi, u64& lo) {
    u64 i = p >> 5;
    int s = (int)(p & 31) * 2;
    u64 a = w[i], b = w[i + 1];
    hi = shl128_hi(a, b, s);
    lo = 0;
    if (wide) {
        u64 c = w[i + 2];
        lo = shl128_hi(b, c, s);
    }
}
// reverse complement of a left-aligned L-base window; result left-aligned
__device__ __forceinline__ void rc_window(int L, u64 hi, u64 lo, u64& ohi, u64& olo) {
    // 128-bit value V = hi:lo holds the window in its top 2L bits. rc(V as 64 bases) puts the
    // reversed window in the LOW 2L bits; shift left by 128-2L to re-align.
    u64 rhi = rc_word(lo), rlo = rc_word(hi); // rc of 64 bases
    int s = 128 - 2 * L;                        // 0 .. 122
    if (s >= 64) { rhi = rlo; rlo = 0; s -= 64; }
    ohi = shl128_hi(rhi, rlo, s);
    olo = s ? (rlo << s) : rlo;
}

#define CUDA_TRY(ctx, expr)                                                     \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) { (ctx)->set_cuda_error(_e, #expr, __LINE__); return MB_E_CUDA; } \
    } while (0)

static inline u32 div_up(u64 a, u64 b) { return (u32)((a + b - 1) / b); }

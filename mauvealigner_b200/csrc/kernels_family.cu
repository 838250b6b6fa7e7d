// kernels_family.cu — the persistent MemHash table across FindMatches calls (seed-family search).
//
// The reference's seed-family path (/root/reference/src/progressiveMauve.cpp:503-548) runs ONE UniqueMatchFinder over
// three seed patterns: FindMatches(cur_list), ClearSequences(), ... and one GetMatchList at the end.  The MemHash table
// persists until Clear(), so a candidate of a later pattern is dropped when an accepted match of an EARLIER call
// contains its seed (same D16 group: genome set, strands, diagonal), and is extended with its own pattern otherwise.
// On the device the accepted matches of the earlier calls are kept as
//   (group hash, second group hash, start on the first genome, end on the first genome)
// sorted by (hash, start) with the running maximum of the ends inside a group, so that "is my seed inside an earlier
// match of my group" is two binary searches.  The group hashes of k_emit_unique do not depend on the seed length (a
// reverse component enters with position + length + first position, which extension leaves unchanged), so matches and
// candidates of different patterns meet in the same group.
#include "common.cuh"
#include "kernels.h"

#define FAM_POISON 0x5A5A5A5A5A5A5A5Aull

// accepted reps of the call that just finished -> table entries (any order; the table is sorted before it is used)
__global__ void __launch_bounds__(256) k_family_append(FamilyArgs f, const u8* __restrict__ rstate, const u32* __restrict__ s_cand, u32 n_rep,
                                                       const u64* __restrict__ ghash, const u64* __restrict__ ghash2, const u32* __restrict__ cand_off,
                                                       const u32* __restrict__ comp_pos, const u32* __restrict__ ext_l, const u32* __restrict__ ext_r, u32 L,
                                                       u32* __restrict__ counter) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rep || (rstate[i] & 0x0Fu) != 1) return;
    const u32 c = s_cand[i];
    const u32 x0 = comp_pos[cand_off[c]];
    const u32 k = atomicAdd(counter, 1u);
    f.t_h1[k] = ghash[c]; f.t_h2[k] = ghash2[c];
    f.t_x[k] = x0 - ext_l[c];
    f.t_end[k] = x0 + L + ext_r[c];
}

// sort keys: first by start (key = start, value = entry), then stably by hash (key = hash of the permuted entry)
__global__ void __launch_bounds__(256) k_family_key_x(FamilyArgs f, u32 n, u64* __restrict__ key, u64* __restrict__ val) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { key[i] = f.t_x[i]; val[i] = i; }
}
__global__ void __launch_bounds__(256) k_family_key_h(FamilyArgs f, u32 n, const u64* __restrict__ val, u64* __restrict__ key) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = f.t_h1[val[i]];
}
__global__ void __launch_bounds__(256) k_family_gather(FamilyArgs f, u32 n, const u64* __restrict__ perm) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 j = (u32)perm[i];
    f.s_h1[i] = f.t_h1[j]; f.s_h2[i] = f.t_h2[j]; f.s_x[i] = f.t_x[j]; f.s_end[i] = f.t_end[j];
}
// running maximum of the ends inside every run of equal (hash, second hash), and the run's first index: the thread of
// a run's first entry walks it
__global__ void __launch_bounds__(256) k_family_pmax(FamilyArgs f, u32 n) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 h1 = f.s_h1[i], h2 = f.s_h2[i];
    if (i > 0 && f.s_h1[i - 1] == h1 && f.s_h2[i - 1] == h2) return;
    u32 m = 0;
    for (u32 j = i; j < n && f.s_h1[j] == h1 && f.s_h2[j] == h2; ++j) { m = max(m, f.s_end[j]); f.s_pmax[j] = m; f.s_run0[j] = i; }
}

// candidate c is contained in an earlier match of its group <=> some entry of the group starts at or before the seed and
// ends at or after it.  Entries of one 64-bit hash are contiguous and ordered by start.  Usual case: they all carry the
// candidate's second hash too (one run) and the running maximum answers; entries of another group under the same first
// hash (a 2^-64 event) make it an exact walk.
__global__ void __launch_bounds__(256) k_family_filter(FamilyArgs f, u32 n_tab, u32 n_cand, u64* __restrict__ ghash, const u64* __restrict__ ghash2,
                                                       const u32* __restrict__ cand_off, const u32* __restrict__ comp_pos, u32 L, u8* __restrict__ pre_drop,
                                                       u32* __restrict__ n_dropped) {
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cand) return;
    const u64 h1 = ghash[c], h2 = ghash2[c];
    const u32 x = comp_pos[cand_off[c]];
    u32 lo = 0, hi = n_tab; // first entry with hash >= h1
    while (lo < hi) { u32 mid = (lo + hi) >> 1; if (f.s_h1[mid] < h1) lo = mid + 1; else hi = mid; }
    const u32 g0 = lo;
    hi = n_tab;             // first entry after g0 with another hash or a start beyond x
    while (lo < hi) { u32 mid = (lo + hi) >> 1; if (f.s_h1[mid] == h1 && f.s_x[mid] <= x) lo = mid + 1; else hi = mid; }
    bool drop = false;
    if (lo > g0) {
        const u32 last = lo - 1;
        if (f.s_h2[last] == h2 && f.s_run0[last] == g0) drop = f.s_pmax[last] >= x + L;
        else
            for (u32 j = g0; j < lo && !drop; ++j) drop = f.s_h2[j] == h2 && f.s_end[j] >= x + L;
    }
    pre_drop[c] = drop ? 1 : 0;
    if (drop) { ghash[c] = h1 ^ FAM_POISON; atomicAdd(n_dropped, 1u); } // its slot never links into a chain
}

static inline u32 fam_div_up(u32 a, u32 b) { return (a + b - 1) / b; }
void launch_family_append(const FamilyArgs& f, const u8* rstate, const u32* s_cand, u32 n_rep, const u64* ghash, const u64* ghash2, const u32* cand_off,
                          const u32* comp_pos, const u32* ext_l, const u32* ext_r, u32 L, u32* counter, cudaStream_t st) {
    if (n_rep) k_family_append<<<fam_div_up(n_rep, 256), 256, 0, st>>>(f, rstate, s_cand, n_rep, ghash, ghash2, cand_off, comp_pos, ext_l, ext_r, L, counter);
}
void launch_family_key_x(const FamilyArgs& f, u32 n, u64* key, u64* val, cudaStream_t st) { if (n) k_family_key_x<<<fam_div_up(n, 256), 256, 0, st>>>(f, n, key, val); }
void launch_family_key_h(const FamilyArgs& f, u32 n, const u64* val, u64* key, cudaStream_t st) { if (n) k_family_key_h<<<fam_div_up(n, 256), 256, 0, st>>>(f, n, val, key); }
void launch_family_gather(const FamilyArgs& f, u32 n, const u64* perm, cudaStream_t st) {
    if (!n) return;
    k_family_gather<<<fam_div_up(n, 256), 256, 0, st>>>(f, n, perm);
    k_family_pmax<<<fam_div_up(n, 256), 256, 0, st>>>(f, n);
}
void launch_family_filter(const FamilyArgs& f, u32 n_tab, u32 n_cand, u64* ghash, const u64* ghash2, const u32* cand_off, const u32* comp_pos, u32 L,
                          u8* pre_drop, u32* n_dropped, cudaStream_t st) {
    if (n_cand) k_family_filter<<<fam_div_up(n_cand, 256), 256, 0, st>>>(f, n_tab, n_cand, ghash, ghash2, cand_off, comp_pos, L, pre_drop, n_dropped);
}

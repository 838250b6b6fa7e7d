// kernels_dist.cu — pack / unpack kernels of the multi-GPU path (api_dist.cu drives them).
//
// Exchange formats (all 8-byte granular so every exchange is one all-to-all of u64 words):
//   seed records    the u64 records of the single-GPU path, partitioned by destination = seed-key range
//   candidate rows  4 x u64 (group hash, second hash, first component, extents) — the candidates are extended at
//                   their source and their component lists never travel; partitioned by owner = f(group hash),
//                   rows keep ascending seed order inside a partition
//   verdicts        one byte per candidate row, back from the owner to the source, in row order
//   match rows      header 2 x u64 (ext_l | ext_r << 32, component count) + one u64 per component (pos | gs << 32)
#include "common.cuh"
#include "kernels.h"

// destination of every value of the top `tb` key bits: analytic splitters of the canonical-seed
// distribution F(x) = 1 - (1 - x)^2 (minimum of the two strands' keys), identical on every rank.
// Folds the 2^tb-bin histogram of the slice into per-destination counts and exclusive offsets.
__global__ void __launch_bounds__(256) k_fold_lut(const u32* __restrict__ hist, const u8* __restrict__ lut, u32 nbins, u32 world,
                                                  u32* __restrict__ digit_base /*[256]*/, u64* __restrict__ counts /*[world]*/) {
    __shared__ u32 sCnt[256];
    const u32 t = threadIdx.x;
    sCnt[t] = 0;
    __syncthreads();
    if (t < nbins && hist[t]) atomicAdd(&sCnt[lut[t]], hist[t]);
    __syncthreads();
    if (t == 0) {
        u32 acc = 0;
        for (u32 d = 0; d < 256; ++d) {
            u32 cnt = sCnt[d];
            digit_base[d] = acc;
            if (d < world) counts[d] = cnt;
            acc += cnt;
        }
    }
}
void launch_fold_lut(const u32* hist, const u8* lut, u32 nbins, u32 world, u32* digit_base, u64* counts, cudaStream_t st) {
    k_fold_lut<<<1, 256, 0, st>>>(hist, lut, nbins, world, digit_base, counts);
}

__device__ __forceinline__ u32 owner_of(u64 h, u32 world) { return (u32)((((h >> 16) & 0xFFFFFFFFull) * world) >> 32); }

__global__ void __launch_bounds__(256) k_owner_keys(const u64* __restrict__ ghash, u32 n, u32 world, u64* __restrict__ skey, u64* __restrict__ sval) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    skey[c] = owner_of(ghash[c], world);
    sval[c] = c;
}
void launch_owner_keys(const u64* ghash, u32 n, u32 world, u64* skey, u64* sval, cudaStream_t st) {
    if (n) k_owner_keys<<<div_up(n, 256), 256, 0, st>>>(ghash, n, world, skey, sval);
}

__global__ void __launch_bounds__(256) k_hdr_m(const u64* __restrict__ hdr, u32 n, u32* __restrict__ m_out) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) m_out[j] = (u32)(hdr[2 * (u64)j + 1] & 0xFFu);
}
void launch_hdr_m(const u64* hdr, u32 n, u32* m_out, cudaStream_t st) {
    if (n) k_hdr_m<<<div_up(n, 256), 256, 0, st>>>(hdr, n, m_out);
}

// received match rows -> "all accepted" candidate arrays (input of the output stage).  The component words of
// row j sit at cand_off[j] in the receive buffer AND in the candidate arrays, so the components are one flat
// elementwise pass; the headers are a second one.
__global__ void __launch_bounds__(256) k_unpack_match(const ulonglong2* __restrict__ hdr, const u64* __restrict__ comps, u32 n, u32 n_comp,
                                                      u32* __restrict__ comp_pos, u8* __restrict__ comp_gs, u32* __restrict__ ext_l,
                                                      u32* __restrict__ ext_r, u8* __restrict__ state, u32* __restrict__ item_cand) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_comp) {
        const u64 w = comps[i];
        comp_pos[i] = (u32)w;
        comp_gs[i] = (u8)(w >> 32);
    }
    if (i < n) {
        const u64 e = hdr[i].x;
        ext_l[i] = (u32)e; ext_r[i] = (u32)(e >> 32);
        state[i] = 1;
        item_cand[i] = i;
    }
}
void launch_unpack_match(const u64* hdr, const u64* comps, const u32* cand_off, u32 n, u32 n_comp, u32* comp_pos, u8* comp_gs, u32* ext_l, u32* ext_r,
                         u8* state, u32* item_cand, cudaStream_t st) {
    (void)cand_off;
    const u32 items = n > n_comp ? n : n_comp;
    if (items) k_unpack_match<<<div_up(items, 256), 256, 0, st>>>(reinterpret_cast<const ulonglong2*>(hdr), comps, n, n_comp, comp_pos, comp_gs, ext_l, ext_r,
                                                                 state, item_cand);
}

// destination rank of row j: the last d with bound[d] <= j (world + 1 entries, the last one is the sentinel)
__device__ __forceinline__ u32 dest_of(const PeerDst* __restrict__ tab, u32 world, u32 j) {
    u32 lo = 0, hi = world; // invariant: bound[lo] <= j < bound[hi]
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (tab[mid].bound <= j) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- extend-at-source protocol: 4-word candidate rows (the extents travel, the component lists stay at the source)
__global__ void __launch_bounds__(256) k_pack_rows(const u64* __restrict__ perm, u32 n, const u64* __restrict__ ghash, const u64* __restrict__ ghash2,
                                                   const u32* __restrict__ cand_off, const u32* __restrict__ comp_pos, const u8* __restrict__ comp_gs,
                                                   GenomeTable gt, const u32* __restrict__ ext_l, const u32* __restrict__ ext_r,
                                                   const PeerDst* __restrict__ tab, u32 world, u32* __restrict__ perm_out) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const PeerDst dst = tab[dest_of(tab, world, j)];
    ulonglong2* rows = reinterpret_cast<ulonglong2*>(dst.base); // this rank's send buffer or, fused with exchange 2, a peer's receive buffer
    const size_t rj = (size_t)dst.off + (j - dst.bound);
    const u32 c = (u32)perm[j];
    const u32 off = cand_off[c], m = cand_off[c + 1] - off;
    const u32 g0 = comp_gs[off] & 0x7F;
    const u32 vg = vgenome(gt, g0, comp_gs[off + 1] & 0x7F);
    rows[2 * rj] = make_ulonglong2(ghash[c], ghash2[c]);
    rows[2 * rj + 1] = make_ulonglong2((u64)g0 | ((u64)vg << 8) | ((u64)m << 24) | ((u64)comp_pos[off] << 32), (u64)ext_l[c] | ((u64)ext_r[c] << 32));
    perm_out[j] = c;
}
void launch_pack_rows(const u64* perm, u32 n, const u64* ghash, const u64* ghash2, const u32* cand_off, const u32* comp_pos, const u8* comp_gs,
                      const GenomeTable& gt, const u32* ext_l, const u32* ext_r, const PeerDst* tab, u32 world, u32* perm_out, cudaStream_t st) {
    if (n) k_pack_rows<<<div_up(n, 256), 256, 0, st>>>(perm, n, ghash, ghash2, cand_off, comp_pos, comp_gs, gt, ext_l, ext_r, tab, world, perm_out);
}

// owner: one bit per received candidate at (virtual genome, position) — the slot axis of the chains
__global__ void __launch_bounds__(256) k_rows_bitmap(const u64* __restrict__ rows, u32 n, GenomeTable gt, u64* __restrict__ bitmap) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const u64 w = rows[4 * (size_t)j + 2];
    const u64 gp = gt.vbase[(u32)(w >> 8) & 0xFFFFu] + (u32)(w >> 32);
    atomicOr((unsigned long long*)&bitmap[gp >> 6], 1ull << (gp & 63));
}
void launch_rows_bitmap(const u64* rows, u32 n, const GenomeTable& gt, u64* bitmap, cudaStream_t st) {
    if (n) k_rows_bitmap<<<div_up(n, 256), 256, 0, st>>>(rows, n, gt, bitmap);
}

// owner: acc[row] = 1 for every accepted rep (acc zeroed before)
__global__ void __launch_bounds__(256) k_accept_mark(const u8* __restrict__ rstate, const u32* __restrict__ s_cand, u32 n_rep, u8* __restrict__ acc) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rep && (rstate[i] & 15u) == 1u) acc[s_cand[i]] = 1;
}
void launch_accept_mark(const u8* rstate, const u32* s_cand, u32 n_rep, u8* acc, cudaStream_t st) {
    if (n_rep) k_accept_mark<<<div_up(n_rep, 256), 256, 0, st>>>(rstate, s_cand, n_rep, acc);
}

// source: verdicts come back in row order; candidate c = perm[row]
__global__ void __launch_bounds__(256) k_apply_accept(const u8* __restrict__ acc, const u32* __restrict__ perm, u32 n, u8* __restrict__ state,
                                                      u32* __restrict__ item_cand) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const u32 c = perm[j];
    state[c] = acc[j] ? 1 : 2;
    item_cand[c] = c;
}
void launch_apply_accept(const u8* acc, const u32* perm, u32 n, u8* state, u32* item_cand, cudaStream_t st) {
    if (n) k_apply_accept<<<div_up(n, 256), 256, 0, st>>>(acc, perm, n, state, item_cand);
}

// ---- distributed output: canonical key of every accepted match, and its key histogram (top 12 key bits).
// The matches of one run crowd into few bins (most share their first genome), so the histogram is built per
// block in shared memory and flushed once.
__global__ void __launch_bounds__(256) k_match_keys(const u8* __restrict__ state, const u32* __restrict__ item_cand, const u32* __restrict__ match_idx,
                                                    const u32* __restrict__ cand_off, const u8* __restrict__ comp_gs, const u32* __restrict__ comp_pos,
                                                    const u32* __restrict__ ext_l, u32 n_items, int sbits, int binshift, u64* __restrict__ key,
                                                    u32* __restrict__ item_of, u64* __restrict__ hist) {
    __shared__ u32 sHist[4096];
    for (u32 b = threadIdx.x; b < 4096; b += blockDim.x) sHist[b] = 0;
    __syncthreads();
    for (u32 it = blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += gridDim.x * blockDim.x) {
        if ((state[it] & 15u) != 1u) continue;
        u32 c = item_cand[it], j = match_idx[it];
        u32 off = cand_off[c];
        u32 f = comp_gs[off] & 0x7F;
        u64 st = (u64)comp_pos[off] - ext_l[c] + 1;
        u64 k = ((u64)(MB_MAX_SEQ - 1 - f) << sbits) | st; // same key as k_uniq_keys: the ranks' ranges concatenate to D18
        key[j] = k;
        item_of[j] = it;
        atomicAdd(&sHist[(k >> binshift) & 4095u], 1u);
    }
    __syncthreads();
    for (u32 b = threadIdx.x; b < 4096; b += blockDim.x) {
        u32 v = sHist[b];
        if (v) atomicAdd((unsigned long long*)&hist[b], (unsigned long long)v);
    }
}
void launch_match_keys(const u8* state, const u32* item_cand, const u32* match_idx, const u32* cand_off, const u8* comp_gs, const u32* comp_pos,
                       const u32* ext_l, u32 n_items, int sbits, int binshift, u64* key, u32* item_of, u64* hist, cudaStream_t st) {
    if (n_items) k_match_keys<<<std::min<u32>(div_up(n_items, 256), 148 * 8), 256, 0, st>>>(state, item_cand, match_idx, cand_off, comp_gs, comp_pos, ext_l, n_items, sbits, binshift, key, item_of, hist);
}

__global__ void __launch_bounds__(256) k_dest_keys(const u64* __restrict__ key, u32 n, int binshift, const u8* __restrict__ lut, u64* __restrict__ skey,
                                                   u64* __restrict__ sval) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    skey[j] = lut[(key[j] >> binshift) & 4095u];
    sval[j] = j;
}
void launch_dest_keys(const u64* key, u32 n, int binshift, const u8* lut, u64* skey, u64* sval, cudaStream_t st) {
    if (n) k_dest_keys<<<div_up(n, 256), 256, 0, st>>>(key, n, binshift, lut, skey, sval);
}

__global__ void __launch_bounds__(256) k_match_perm_m(const u64* __restrict__ perm, const u32* __restrict__ item_of, const u32* __restrict__ item_cand,
                                                      const u32* __restrict__ cand_off, u32 n, u32* __restrict__ m_out) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    u32 c = item_cand[item_of[(u32)perm[t]]];
    m_out[t] = cand_off[c + 1] - cand_off[c];
}
void launch_match_perm_m(const u64* perm, const u32* item_of, const u32* item_cand, const u32* cand_off, u32 n, u32* m_out, cudaStream_t st) {
    if (n) k_match_perm_m<<<div_up(n, 256), 256, 0, st>>>(perm, item_of, item_cand, cand_off, n, m_out);
}

// lane = row for the header; the component words of the warp's 32 rows are copied by all lanes together
// (coalesced on both sides).  Destinations through the PeerDst table: the local send buffers or, fused with
// exchange 3, the destination ranks' receive buffers over NVLink.
__global__ void __launch_bounds__(256) k_pack_match_perm(const u64* __restrict__ perm, const u32* __restrict__ item_of, const u32* __restrict__ item_cand,
                                                         const u64* __restrict__ poff, const u32* __restrict__ cand_off, const u32* __restrict__ comp_pos,
                                                         const u8* __restrict__ comp_gs, const u32* __restrict__ ext_l, const u32* __restrict__ ext_r, u32 n,
                                                         const PeerDst* __restrict__ tab, u32 world) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 lane = threadIdx.x & 31;
    u32 off = 0, m = 0;
    u64* o = nullptr;
    if (t < n) {
        const u32 c = item_cand[item_of[(u32)perm[t]]];
        off = cand_off[c]; m = cand_off[c + 1] - off;
        const PeerDst dst = tab[dest_of(tab, world, t)];
        reinterpret_cast<ulonglong2*>(dst.base)[dst.off + t - dst.bound] = make_ulonglong2((u64)ext_l[c] | ((u64)ext_r[c] << 32), m);
        o = dst.base2 + dst.off2 + (poff[t] - dst.cbound);
    }
    for (int j = 0; j < 32; ++j) {
        const u32 mj = __shfl_sync(0xFFFFFFFFu, m, j);
        if (mj == 0) continue; // uniform
        const u32 offj = __shfl_sync(0xFFFFFFFFu, off, j);
        u64* oj = reinterpret_cast<u64*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(o), j));
        for (u32 k = lane; k < mj; k += 32) oj[k] = (u64)comp_pos[offj + k] | ((u64)comp_gs[offj + k] << 32);
    }
}
void launch_pack_match_perm(const u64* perm, const u32* item_of, const u32* item_cand, const u64* poff, const u32* cand_off, const u32* comp_pos,
                            const u8* comp_gs, const u32* ext_l, const u32* ext_r, u32 n, const PeerDst* tab, u32 world, cudaStream_t st) {
    if (n) k_pack_match_perm<<<div_up(n, 256), 256, 0, st>>>(perm, item_of, item_cand, poff, cand_off, comp_pos, comp_gs, ext_l, ext_r, n, tab, world);
}

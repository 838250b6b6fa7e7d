// api_dist.cu — the multi-GPU entry points of include/mauve_b200.h.
//
// One context per rank (one process per GPU); the packed genomes are replicated on every rank.  The
// library runs the per-rank stages and owns every exchange buffer; the exchanges themselves (three
// variable all-to-alls of u64 words over NCCL / NVLink) are issued by the caller between the stages
// (mauvealigner_b200/dist.py over torch.distributed), so the C ABI stays free of communicator types.
//
//   stage 1  mb_dist_extract   seeds of this rank's slice of the (genome, position) space, stably
//                              partitioned by destination = seed-key range        --> exchange 1
//   stage 2  mb_dist_local     sort / runs / policy over the received key range -> candidates in
//                              ascending seed order, stably partitioned by owner = f(group hash)
//                                                                                 --> exchange 2
//   stage 3  mb_dist_dedup     chains / extension / resolve over the owned groups (ranks keep the
//                              global seed order: rows arrive in source-rank order) -> accepted
//                              matches + histogram of their canonical keys        --> all-reduce (sum)
//            mb_dist_match_partition  match rows by destination = key range       --> exchange 3
//   stage 4  mb_dist_output    every rank: canonical order (D18) + CSR of its key range; the pieces in
//                              rank order are the result (mb_fetch_result per rank)
// MODE_UNIQUE with 8-byte records only (every BASELINE config that names several GPUs).
#include "ctx.h"

extern "C" {

// stage 1
int mb_dist_extract(mb_ctx* c, int rank, int world, void** d_send, uint64_t* h_counts) {
    if (!c || !d_send || !h_counts || world < 1 || world > 256 || rank < 0 || rank >= world) return MB_E_ARG;
    MbiRun run;
    TRY(mbi_setup_run(c, run));
    if (c->fmt.wide) return MB_E_ARG;
    cudaStream_t st = c->stream;
    const RecFmt& fmt = c->fmt;
    c->d_rank = rank; c->d_world = world; c->d_bases = run.bases; c->d_maxlen = run.maxlen;
    for (int i = 0; i < 8; ++i) if (!c->ev_d[i]) cudaEventCreate(&c->ev_d[i]);
    cudaEventRecord(c->ev_d[0], st);
    // this rank's tiles and the record index of their first seed
    const u32 ET = extract_tile_size();
    const u32 t0 = (u32)((u64)run.n_tiles * rank / world), t1 = (u32)((u64)run.n_tiles * (rank + 1) / world);
    auto first_record = [&](u32 t) -> u64 {
        if (t >= run.n_tiles) return run.n;
        u32 g = 0;
        while (g + 1 < c->gt.nseq && t >= run.tile_first[g + 1]) ++g;
        return (u64)c->gt.seed_base[g] + (u64)(t - run.tile_first[g]) * ET;
    };
    const u64 r0 = first_record(t0), r1 = first_record(t1);
    const u32 ns = (u32)(r1 - r0);
    TRY(c->reserve(c->keysA, ((size_t)ns + 8) * 8));
    TRY(c->reserve(c->keysB, ((size_t)ns + 8) * 8));
    launch_extract_records(c->packed.as<u64>(), c->keysA.as<u64>(), nullptr, nullptr, 0, c->gt, c->sd, fmt, c->tile_first.as<u32>(), t1 - t0, st, t0,
                           (u32)r0);
    if (t1 > t0) { LAUNCHED(c); CHECK_LAUNCH(c); }
    // destination of every value of the top tb key bits: splitters of F(x) = 1 - (1 - x)^2
    const int tb = std::min(8, fmt.kbits);
    const u32 T = 1u << tb;
    uint8_t lut[256];
    for (u32 t = 0; t < 256; ++t) {
        u64 q = 2ull * T - 1 - 2ull * std::min(t, T - 1);
        u64 d = (u64)world * (4ull * T * T - q * q) / (4ull * T * T);
        lut[t] = (uint8_t)std::min<u64>(d, (u64)world - 1);
    }
    TRY(c->reserve(c->x_lut, 256));
    TRY(c->reserve(c->x_counts, 256 * 8));
    CUDA_TRY(c, cudaMemcpyAsync(c->x_lut.p, lut, 256, cudaMemcpyHostToDevice, st));
    const int shift = fmt.kshift + fmt.kbits - tb;
    if (ns) {
        launch_hist(c->keysA.as<u64>(), ns, shift, tb, 1, c->hist.as<u32>(), st); LAUNCHED(c);
    }
    launch_fold_lut(c->hist.as<u32>(), c->x_lut.as<u8>(), T, (u32)world, c->digit_base.as<u32>(), c->x_counts.as<u64>(), st); LAUNCHED(c);
    if (ns) {
        u32 tiles = div_up(ns, radix_tile_size());
        CUDA_TRY(c, cudaMemsetAsync(c->lookback.p, 0, (size_t)tiles * 256 * 8, st));
        cudaError_t e = launch_onesweep(c->keysA.as<u64>(), c->keysB.as<u64>(), nullptr, nullptr, ns, c->digit_base.as<u32>(), c->lookback.as<u64>(),
                                        c->ticket(), shift, tb, st, c->x_lut.as<u8>());
        LAUNCHED(c);
        if (e != cudaSuccess) { c->set_cuda_error(e, "onesweep(partition)", __LINE__); return MB_E_CUDA; }
    }
    cudaEventRecord(c->ev_d[1], st);
    CUDA_TRY(c, cudaMemcpyAsync(h_counts, c->x_counts.p, (size_t)world * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    *d_send = c->keysB.p;
    return MB_OK;
}

// receive buffers (library-owned device memory the caller's all-to-all writes into)
//   which: 0 seed records, 1 candidate headers, 2 candidate components, 3 match headers, 4 match components
int mb_dist_recv_buffer(mb_ctx* c, int which, uint64_t n_words, void** d_ptr) {
    if (!c || !d_ptr || which < 0 || which > 4) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    DBuf* b = which == 0 ? &c->keysA : (which == 1 || which == 3) ? &c->x_hdr_r : &c->x_comp_r;
    TRY(c->reserve(*b, ((size_t)n_words + 16) * 8));
    *d_ptr = b->p;
    return MB_OK;
}

// stage 2: n_recv records sit in receive buffer 0, concatenated in source-rank order
int mb_dist_local(mb_ctx* c, const mb_params* prm, uint64_t n_recv, uint64_t* h_cand_counts, uint64_t* h_comp_counts, void** d_hdr, void** d_comps) {
    if (!c || !prm || !h_cand_counts || !h_comp_counts || !d_hdr || !d_comps) return MB_E_ARG;
    if (prm->mode != MB_MODE_UNIQUE) return MB_E_ARG;
    if (n_recv >= (1ull << 31)) return MB_E_TOOLONG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const RecFmt& fmt = c->fmt;
    const int world = c->d_world;
    const u32 n = (u32)n_recv;
    u64* scal = c->scalars.as<u64>();
    for (int r = 0; r < world; ++r) { h_cand_counts[r] = 0; h_comp_counts[r] = 0; }
    *d_hdr = nullptr; *d_comps = nullptr;
    c->d_ncand = 0; c->d_nccomp = 0;
    cudaEventRecord(c->ev_d[2], st);
    TRY(c->reserve(c->keysB, ((size_t)n + 8) * 8));
    u64 *kA = c->keysA.as<u64>(), *kB = c->keysB.as<u64>();
    TRY(mbi_sort_records(c, &kA, &kB, nullptr, nullptr, n, fmt.kshift, fmt.kbits, false, true));
    c->sorted_keys = kA; c->sorted_vals = nullptr;
    cudaEventRecord(c->ev_d[3], st);
    if (n == 0) { cudaEventRecord(c->ev_d[4], st); return MB_OK; }
    u32* run_start = reinterpret_cast<u32*>(kB);
    u32* run_u = run_start + (n + 2);
    launch_find_runs(kA, nullptr, n, fmt, run_start, run_u, c->status_slice(div_up(n, find_runs_tile())), c->ticket(), nullptr,
                     reinterpret_cast<u32*>(scal + SC_RUNS), st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    const u32 cand_cap = n / 2 + 2;
    TRY(c->reserve(c->cand_run, (size_t)cand_cap * 4));
    TRY(c->reserve(c->cand_off, (size_t)(cand_cap + 1) * 4));
    SelectArgs sa{};
    sa.keys = kA; sa.vals = nullptr; sa.run_start = run_start; sa.run_u = run_u;
    sa.n_runs_ptr = reinterpret_cast<u32*>(scal + SC_RUNS);
    sa.mode = prm->mode; sa.direct_only = prm->direct_only;
    sa.min_multi = prm->min_multi; sa.max_multi = prm->max_multi; sa.nway_mask = prm->nway_mask;
    sa.status = c->status_slice(div_up(n, select_tile())); sa.ticket = c->ticket();
    sa.n_buckets = scal + SC_NBUCKETS;
    sa.totals = reinterpret_cast<u32*>(scal + SC_CAND);
    sa.cand_run = c->cand_run.as<u32>(); sa.cand_off = c->cand_off.as<u32>(); sa.cand_aux = nullptr;
    launch_select(sa, fmt, n, st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    TRY(mbi_read_scalars(c));
    const u32* hs32 = reinterpret_cast<const u32*>(c->h_scal);
    const u64* hs64 = reinterpret_cast<const u64*>(c->h_scal);
    const u32 n_cand = hs32[2 * SC_CAND], n_ccomp = hs32[2 * SC_CAND + 1];
    c->stats.n_runs = hs32[2 * SC_RUNS]; c->stats.n_buckets = hs64[SC_NBUCKETS]; c->stats.n_candidates = n_cand;
    c->r_unique = hs32[2 * SC_RUNS];
    c->d_ncand = n_cand; c->d_nccomp = n_ccomp;
    if (n_cand == 0) { cudaEventRecord(c->ev_d[4], st); return MB_OK; }
    const size_t nc = (size_t)n_cand + 8;
    TRY(c->reserve(c->comp_pos, ((size_t)n_ccomp + 8) * 4));
    TRY(c->reserve(c->comp_gs, (size_t)n_ccomp + 8));
    TRY(c->reserve(c->ghash, nc * 8));
    TRY(c->reserve(c->ghash2, nc * 8));
    TRY(c->reserve(c->sort_kA, nc * 8)); TRY(c->reserve(c->sort_kB, nc * 8));
    TRY(c->reserve(c->sort_vA, nc * 8)); TRY(c->reserve(c->sort_vB, nc * 8));
    TRY(c->reserve(c->x_m, nc * 4));
    TRY(c->reserve(c->out_off, nc * 8));
    TRY(c->reserve(c->x_hdr_s, nc * 16));
    TRY(c->reserve(c->x_comp_s, ((size_t)n_ccomp + 8) * 8));
    EmitUniqueArgs eu{};
    eu.keys = kA; eu.vals = nullptr; eu.run_start = run_start; eu.run_u = run_u;
    eu.cand_run = c->cand_run.as<u32>(); eu.cand_off = c->cand_off.as<u32>(); eu.cand_aux = nullptr;
    eu.totals = reinterpret_cast<u32*>(scal + SC_CAND);
    eu.mode = prm->mode; eu.comp_pos = c->comp_pos.as<u32>(); eu.comp_gs = c->comp_gs.as<u8>(); eu.bitmap = nullptr; eu.ghash = c->ghash.as<u64>(); eu.ghash2 = c->ghash2.as<u64>();
    launch_emit_unique(eu, fmt, c->gt, n_cand, st); LAUNCHED(c); CHECK_LAUNCH(c);
    // stable partition of the candidate rows by owner
    u64 *skA = c->sort_kA.as<u64>(), *skB = c->sort_kB.as<u64>(), *svA = c->sort_vA.as<u64>(), *svB = c->sort_vB.as<u64>();
    launch_owner_keys(c->ghash.as<u64>(), n_cand, (u32)world, skA, svA, st); LAUNCHED(c);
    const int obits = std::max(1, mbi_bits_for((u64)world - 1));
    TRY(mbi_sort_records(c, &skA, &skB, &svA, &svB, n_cand, 0, obits, false)); // leaves the owner histogram in c->hist
    launch_perm_m(svA, c->cand_off.as<u32>(), n_cand, c->x_m.as<u32>(), st); LAUNCHED(c);
    launch_scan_u32(c->x_m.as<u32>(), n_cand, nullptr, c->out_off.as<u64>(), c->status_slice(div_up(n_cand, scan_tile())), c->ticket(), nullptr, st);
    LAUNCHED(c);
    launch_pack_cand(svA, c->out_off.as<u64>(), c->cand_off.as<u32>(), c->comp_pos.as<u32>(), c->comp_gs.as<u8>(), c->ghash.as<u64>(), c->ghash2.as<u64>(),
                     n_cand, c->x_hdr_s.as<u64>(), c->x_comp_s.as<u64>(), st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    cudaEventRecord(c->ev_d[4], st);
    // per-owner row and component counts
    std::vector<u32> oh(256);
    CUDA_TRY(c, cudaMemcpyAsync(oh.data(), c->hist.p, 256 * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    std::vector<u64> bound(world + 1, 0);
    u64 acc = 0;
    for (int r = 0; r < world; ++r) { h_cand_counts[r] = oh[r]; acc += oh[r]; bound[r + 1] = acc; }
    if (acc != n_cand) return MB_E_STATE;
    std::vector<u64> cb(world + 1, 0);
    for (int r = 1; r <= world; ++r) CUDA_TRY(c, cudaMemcpyAsync(&cb[r], c->out_off.as<u64>() + bound[r], 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    for (int r = 0; r < world; ++r) h_comp_counts[r] = cb[r + 1] - cb[r];
    *d_hdr = c->x_hdr_s.p; *d_comps = c->x_comp_s.p;
    return MB_OK;
}

// stage 3a: n_cand rows (headers in receive buffer 1, n_comp component words in buffer 2), source-rank order.
// De-dup of the owned groups; leaves a 4096-bin histogram of the accepted matches' canonical sort keys at
// *d_hist (uint64 counts, library-owned) for the caller to sum over all ranks IN PLACE before stage 3b.
int mb_dist_dedup(mb_ctx* c, uint64_t n_cand64, uint64_t n_comp64, void** d_hist) {
    if (!c || !d_hist) return MB_E_ARG;
    if (n_cand64 >= (1ull << 31) || n_comp64 >= (1ull << 32)) return MB_E_TOOLONG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    u64* scal = c->scalars.as<u64>();
    const u32 n_cand = (u32)n_cand64, n_ccomp = (u32)n_comp64;
    c->d_nmatch = 0; c->d_nmcomp = 0; c->n_rep = 0;
    TRY(c->reserve(c->x_counts, 4096 * 8));
    CUDA_TRY(c, cudaMemsetAsync(c->x_counts.p, 0, 4096 * 8, st));
    *d_hist = c->x_counts.p;
    cudaEventRecord(c->ev_d[5], st);
    if (n_cand == 0) { cudaEventRecord(c->ev_d[6], st); CUDA_TRY(c, cudaStreamSynchronize(st)); return MB_OK; }
    TRY(mbi_reserve_candidates(c, n_cand, n_ccomp, c->d_bases));
    TRY(c->reserve(c->x_m, ((size_t)n_cand + 8) * 4));
    const u64 bm_words = c->d_bases / 64 + 2;
    CUDA_TRY(c, cudaMemsetAsync(c->bitmap.p, 0, bm_words * 8, st));
    const u64* hdr = c->x_hdr_r.as<u64>();
    const u64* comps = c->x_comp_r.as<u64>();
    launch_hdr_m(hdr, n_cand, c->x_m.as<u32>(), st); LAUNCHED(c);
    launch_scan_u32(c->x_m.as<u32>(), n_cand, c->cand_off.as<u32>(), nullptr, c->status_slice(div_up(n_cand, scan_tile())), c->ticket(), nullptr, st);
    LAUNCHED(c);
    launch_unpack_cand(hdr, comps, c->cand_off.as<u32>(), n_cand, c->gt, c->comp_pos.as<u32>(), c->comp_gs.as<u8>(), c->ghash.as<u64>(),
                       c->ghash2.as<u64>(), c->bitmap.as<u64>(), st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    TRY(mbi_dedup(c, n_cand, c->d_bases));
    // accepted matches (among the reps): index, canonical key, key histogram
    const u32 n_rep = c->n_rep;
    TRY(c->reserve(c->flags, ((size_t)n_rep + 8) * 4));
    TRY(c->reserve(c->match_idx, ((size_t)n_rep + 8) * 4));
    OutputArgs oa{};
    oa.n_items = n_rep; oa.state = c->s_cand.as<u8>(); oa.item_cand = c->rep_cand.as<u32>(); oa.flags = c->flags.as<u32>();
    launch_uniq_flags(oa, st); LAUNCHED(c);
    launch_scan_u32(c->flags.as<u32>(), n_rep, c->match_idx.as<u32>(), nullptr, c->status_slice(div_up(n_rep, scan_tile())), c->ticket(),
                    scal + SC_NMATCH, st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    TRY(mbi_read_scalars(c));
    const u64* hs64 = reinterpret_cast<const u64*>(c->h_scal);
    const u32* hs32 = reinterpret_cast<const u32*>(c->h_scal);
    const u32 n_match = (u32)hs64[SC_NMATCH];
    c->stats.dedup_iters = hs32[2 * SC_DDCTR + 9];
    c->d_nmatch = n_match;
    const size_t nm = (size_t)n_match + 8;
    TRY(c->reserve(c->sort_kA, nm * 8)); TRY(c->reserve(c->sort_kB, nm * 8));
    TRY(c->reserve(c->sort_vA, nm * 8)); TRY(c->reserve(c->sort_vB, nm * 8));
    TRY(c->reserve(c->x_key, nm * 8));
    TRY(c->reserve(c->x_item, nm * 4));
    const int sbits = mbi_bits_for(c->d_maxlen);
    launch_match_keys(c->s_cand.as<u8>(), c->rep_cand.as<u32>(), c->match_idx.as<u32>(), c->cand_off.as<u32>(), c->comp_gs.as<u8>(),
                      c->comp_pos.as<u32>(), c->ext_l.as<u32>(), n_rep, sbits, std::max(0, sbits + 6 - 12), c->x_key.as<u64>(), c->x_item.as<u32>(),
                      c->x_counts.as<u64>(), st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    cudaEventRecord(c->ev_d[6], st);
    CUDA_TRY(c, cudaStreamSynchronize(st));
    return MB_OK;
}

// stage 3b: *d_hist now holds the histogram summed over all ranks.  Partition the accepted matches by
// destination = range of the canonical sort key (ranks hold ascending ranges of the final order) and pack
// their rows; per-destination row / component-word counts go to the host arrays.
int mb_dist_match_partition(mb_ctx* c, uint64_t* h_match_counts, uint64_t* h_comp_counts, void** d_hdr, void** d_comps) {
    if (!c || !h_match_counts || !h_comp_counts || !d_hdr || !d_comps) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int world = c->d_world;
    const u32 n_match = c->d_nmatch;
    for (int r = 0; r < world; ++r) { h_match_counts[r] = 0; h_comp_counts[r] = 0; }
    *d_hdr = nullptr; *d_comps = nullptr;
    // destination of every key bin: equal shares of the global match count, whole bins
    std::vector<u64> gh(4096);
    CUDA_TRY(c, cudaMemcpyAsync(gh.data(), c->x_counts.p, 4096 * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    u64 total = 0;
    for (u64 v : gh) total += v;
    std::vector<uint8_t> lut(4096, 0);
    u64 before = 0;
    for (int b = 0; b < 4096; ++b) {
        u64 d = total ? (before * (u64)world) / total : 0;
        lut[b] = (uint8_t)std::min<u64>(d, (u64)world - 1);
        before += gh[b];
    }
    if (n_match == 0) return MB_OK;
    TRY(c->reserve(c->x_lut, 4096));
    CUDA_TRY(c, cudaMemcpyAsync(c->x_lut.p, lut.data(), 4096, cudaMemcpyHostToDevice, st));
    const int sbits = mbi_bits_for(c->d_maxlen);
    u64 *skA = c->sort_kA.as<u64>(), *skB = c->sort_kB.as<u64>(), *svA = c->sort_vA.as<u64>(), *svB = c->sort_vB.as<u64>();
    launch_dest_keys(c->x_key.as<u64>(), n_match, std::max(0, sbits + 6 - 12), c->x_lut.as<u8>(), skA, svA, st); LAUNCHED(c);
    const int obits = std::max(1, mbi_bits_for((u64)world - 1));
    TRY(mbi_sort_records(c, &skA, &skB, &svA, &svB, n_match, 0, obits, false)); // leaves the destination histogram in c->hist
    TRY(c->reserve(c->x_m, ((size_t)n_match + 8) * 4));
    TRY(c->reserve(c->out_off, ((size_t)n_match + 8) * 8));
    launch_match_perm_m(svA, c->x_item.as<u32>(), c->rep_cand.as<u32>(), c->cand_off.as<u32>(), n_match, c->x_m.as<u32>(), st); LAUNCHED(c);
    launch_scan_u32(c->x_m.as<u32>(), n_match, nullptr, c->out_off.as<u64>(), c->status_slice(div_up(n_match, scan_tile())), c->ticket(), nullptr, st);
    LAUNCHED(c);
    std::vector<u32> oh(256);
    CUDA_TRY(c, cudaMemcpyAsync(oh.data(), c->hist.p, 256 * 4, cudaMemcpyDeviceToHost, st));
    u64 n_mcomp = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&n_mcomp, c->out_off.as<u64>() + n_match, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    std::vector<u64> bound(world + 1, 0);
    u64 acc = 0;
    for (int r = 0; r < world; ++r) { h_match_counts[r] = oh[r]; acc += oh[r]; bound[r + 1] = acc; }
    if (acc != n_match) return MB_E_STATE;
    std::vector<u64> cb(world + 1, 0);
    for (int r = 1; r <= world; ++r) CUDA_TRY(c, cudaMemcpyAsync(&cb[r], c->out_off.as<u64>() + bound[r], 8, cudaMemcpyDeviceToHost, st));
    TRY(c->reserve(c->x_hdr_s, ((size_t)n_match + 8) * 16));
    TRY(c->reserve(c->x_comp_s, ((size_t)n_mcomp + 8) * 8));
    launch_pack_match_perm(svA, c->x_item.as<u32>(), c->rep_cand.as<u32>(), c->out_off.as<u64>(), c->cand_off.as<u32>(), c->comp_pos.as<u32>(),
                           c->comp_gs.as<u8>(), c->ext_l.as<u32>(), c->ext_r.as<u32>(), n_match, c->x_hdr_s.as<u64>(), c->x_comp_s.as<u64>(), st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    CUDA_TRY(c, cudaStreamSynchronize(st));
    for (int r = 0; r < world; ++r) h_comp_counts[r] = cb[r + 1] - cb[r];
    c->d_nmcomp = (u32)n_mcomp;
    *d_hdr = c->x_hdr_s.p; *d_comps = c->x_comp_s.p;
    return MB_OK;
}

// stage 4 (every rank): the n_match rows of this rank's key range (receive buffers 3 / 4) -> canonical match CSR
// on the device; the ranks' pieces, in rank order, are the whole result
int mb_dist_output(mb_ctx* c, uint64_t n_match64, uint64_t n_comp64) {
    if (!c) return MB_E_ARG;
    if (n_match64 >= (1ull << 31) || n_comp64 >= (1ull << 32)) return MB_E_TOOLONG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const u32 n_match = (u32)n_match64, n_mcomp = (u32)n_comp64;
    c->last_mode = MB_MODE_UNIQUE;
    c->r_matches = 0; c->r_comps = 0;
    if (n_match) {
        TRY(mbi_reserve_candidates(c, n_match, n_mcomp, 0));
        TRY(c->reserve(c->x_m, ((size_t)n_match + 8) * 4));
        const u64* hdr = c->x_hdr_r.as<u64>();
        const u64* comps = c->x_comp_r.as<u64>();
        launch_hdr_m(hdr, n_match, c->x_m.as<u32>(), st); LAUNCHED(c);
        launch_scan_u32(c->x_m.as<u32>(), n_match, c->cand_off.as<u32>(), nullptr, c->status_slice(div_up(n_match, scan_tile())), c->ticket(), nullptr,
                        st);
        LAUNCHED(c);
        launch_unpack_match(hdr, comps, c->cand_off.as<u32>(), n_match, c->comp_pos.as<u32>(), c->comp_gs.as<u8>(), c->ext_l.as<u32>(),
                            c->ext_r.as<u32>(), c->s_cand.as<u8>(), c->rep_cand.as<u32>(), st);
        LAUNCHED(c); CHECK_LAUNCH(c);
        TRY(mbi_output_unique(c, n_match, c->d_maxlen));
    }
    cudaEventRecord(c->ev_d[7], st);
    c->stats.n_matches = c->r_matches; c->stats.n_comps = c->r_comps;
    // the stage events of the single-GPU driver are not recorded on this path
    for (int i = 0; i < EV_COUNT; ++i) cudaEventRecord(c->ev[i], st);
    c->have_result = true;
    return MB_OK;
}

// device milliseconds of this rank's stages of the last distributed run: [0] extract+partition,
// [1] sort, [2] runs/policy/candidate rows, [3] de-dup + match rows
int mb_dist_stage_ms(mb_ctx* c, float* out4) {
    if (!c || !out4) return MB_E_ARG;
    auto ms = [&](int a, int b) { float t = 0; if (!c->ev_d[a] || !c->ev_d[b] || cudaEventElapsedTime(&t, c->ev_d[a], c->ev_d[b]) != cudaSuccess) { cudaGetLastError(); t = 0; } return t; };
    out4[0] = ms(0, 1); out4[1] = ms(2, 3); out4[2] = ms(3, 4); out4[3] = ms(5, 6);
    return MB_OK;
}

} // extern "C"

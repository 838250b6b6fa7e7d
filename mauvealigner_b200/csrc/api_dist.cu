// api_dist.cu — the multi-GPU entry points of include/mauve_b200.h.
//
// One context per rank (one process per GPU); the packed genomes are replicated on every rank.  The
// library runs the per-rank stages and owns every exchange buffer; the exchanges themselves are issued by
// the caller between the stages (mauvealigner_b200/dist.py over torch.distributed), so the C ABI stays free
// of communicator types — except exchange 1, which mb_dist_partition can fuse into the partition kernel as
// NVLink peer stores into the destination ranks' receive arrays (CUDA IPC handles travel through the caller).
//
//   stage 1  mb_dist_extract[_count] / mb_dist_partition   seeds of this rank's slice of the (genome, position)
//                              space, stably partitioned by destination = seed-key range     --> exchange 1
//   stage 2  mb_dist_local     sort / runs / policy over the received key range -> candidates in ascending
//                              seed order; every candidate is EXTENDED HERE (pure function of the candidate
//                              and the replicated genomes); 4-word rows (group hashes, first component,
//                              extents) stably partitioned by owner = f(group hash)         --> exchange 2
//   stage 3a mb_dist_resolve   owner: chains / resolve over the owned groups (rows arrive in source-rank order,
//                              which keeps the global seed order inside every group) -> one verdict byte per row
//                                                                                     --> exchange 2b (back)
//   stage 3b mb_dist_accept    source: accepted candidates = matches; histogram of their canonical keys
//                                                                                     --> all-reduce (sum)
//   stage 3c mb_dist_match_partition  match rows (with their component lists, which never left the source)
//                              by destination = range of the canonical order           --> exchange 3
//   stage 4  mb_dist_output    every rank: canonical order (D18) + CSR of its key range; the pieces in
//                              rank order are the result (mb_fetch_result per rank)
// The stages above serve MODE_UNIQUE with 8-byte records (C5, the BASELINE config that names several GPUs).
// MODE_UNIQUE_COUNT / MODE_SEED_ENUM, with 8- or 16-byte records, need stage 1 and exchange 1 only:
//   mb_dist_extract_records -> exchange 1 (one all-to-all per record word) -> mb_dist_enum_local (sort + the mode's tail
//   over the key range); counts add up over the ranks, match lists are disjoint.
#include "ctx.h"

extern "C" {

// stage 1a: extract this rank's slice of seeds (into keysA) and count them per destination rank
int mb_dist_extract_count(mb_ctx* c, int rank, int world, uint64_t* h_counts) {
    if (!c || !h_counts || world < 1 || world > 256 || rank < 0 || rank >= world) return MB_E_ARG;
    MbiRun run;
    if (c->n_seg) return MB_E_STATE; // segmented searches are single-GPU
    TRY(mbi_setup_run(c, run));
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
    c->d_nslice = ns;
    TRY(c->reserve(c->keysA, ((size_t)ns + 8) * 8));
    if (fmt.wide) TRY(c->reserve(c->valsA, ((size_t)ns + 8) * 8)); // 16-byte records: (seed, genome | position | strand)
    launch_extract_records(c->packed.as<u64>(), c->keysA.as<u64>(), fmt.wide ? c->valsA.as<u64>() : nullptr, nullptr, 0, c->gt, c->sd, fmt,
                           c->tile_first.as<u32>(), t1 - t0, st, t0, (u32)r0);
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
    TRY(c->reserve(c->x_lut, 4096));
    TRY(c->reserve(c->x_counts, 4096 * 8));
    CUDA_TRY(c, cudaMemcpyAsync(c->x_lut.p, lut, 256, cudaMemcpyHostToDevice, st));
    const int shift = fmt.kshift + fmt.kbits - tb;
    if (ns) { launch_hist(c->keysA.as<u64>(), ns, shift, tb, 1, c->hist.as<u32>(), st); LAUNCHED(c); }
    launch_fold_lut(c->hist.as<u32>(), c->x_lut.as<u8>(), T, (u32)world, c->digit_base.as<u32>(), c->x_counts.as<u64>(), st); LAUNCHED(c);
    CUDA_TRY(c, cudaMemcpyAsync(h_counts, c->x_counts.p, (size_t)world * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    return MB_OK;
}

// stage 1b: stable partition of the slice by destination.  peer_bases == NULL: into this rank's send buffer
// (*d_send; the caller's all-to-all moves it).  Otherwise the partition pass writes every record straight into
// its destination rank's receive array over NVLink (peer_bases[d] = device pointer of rank d's array, mapped into
// this process; peer_offsets[d] = record index where this rank's block starts there): the exchange is fused
// into the kernel.  The caller synchronises all ranks before anyone reads its receive array.
int mb_dist_partition(mb_ctx* c, void* const* peer_bases, const uint64_t* peer_offsets, void** d_send) {
    if (!c || (peer_bases && !peer_offsets) || (!peer_bases && !d_send)) return MB_E_ARG;
    if (peer_bases && c->fmt.wide) return MB_E_ARG; // 16-byte records travel through local send buffers (mb_dist_extract_records)
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const RecFmt& fmt = c->fmt;
    const int world = c->d_world;
    const u32 ns = c->d_nslice;
    const int tb = std::min(8, fmt.kbits);
    const int shift = fmt.kshift + fmt.kbits - tb;
    u64* const* d_peers = nullptr;
    if (peer_bases) {
        TRY(c->reserve(c->x_peers, 256 * 8));
        std::vector<u64*> pp(256, nullptr);
        std::vector<u32> base(256, 0);
        for (int d = 0; d < world; ++d) {
            if (peer_offsets[d] >= (1ull << 31)) return MB_E_TOOLONG;
            pp[d] = (u64*)peer_bases[d]; base[d] = (u32)peer_offsets[d];
        }
        CUDA_TRY(c, cudaMemcpyAsync(c->x_peers.p, pp.data(), 256 * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(c, cudaMemcpyAsync(c->digit_base.p, base.data(), 256 * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(c, cudaStreamSynchronize(st)); // the host vectors die at return
        d_peers = c->x_peers.as<u64*>();
    } else {
        TRY(c->reserve(c->keysB, ((size_t)ns + 8) * 8));
        if (fmt.wide) TRY(c->reserve(c->valsB, ((size_t)ns + 8) * 8));
    }
    if (ns) {
        u32 tiles = div_up(ns, radix_tile_size());
        CUDA_TRY(c, cudaMemsetAsync(c->lookback.p, 0, (size_t)tiles * 256 * 8, st));
        cudaError_t e = launch_onesweep(c->keysA.as<u64>(), c->keysB.as<u64>(), fmt.wide ? c->valsA.as<u64>() : nullptr, fmt.wide ? c->valsB.as<u64>() : nullptr, ns,
                                        c->digit_base.as<u32>(), c->lookback.as<u64>(), c->ticket(), shift, tb, st, c->x_lut.as<u8>(), d_peers);
        LAUNCHED(c);
        if (e != cudaSuccess) { c->set_cuda_error(e, "onesweep(partition)", __LINE__); return MB_E_CUDA; }
    }
    cudaEventRecord(c->ev_d[1], st);
    if (d_send) *d_send = c->keysB.p;
    return MB_OK;
}

// stage 1 = 1a + 1b into the local send buffer
int mb_dist_extract(mb_ctx* c, int rank, int world, void** d_send, uint64_t* h_counts) {
    if (!d_send) return MB_E_ARG;
    TRY(mb_dist_extract_count(c, rank, world, h_counts));
    if (c->fmt.wide) return MB_E_ARG; // use mb_dist_extract_records
    TRY(mb_dist_partition(c, nullptr, nullptr, d_send));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MB_OK;
}
// the same for either record format: *d_send_vals = the second words of 16-byte records in the same order, or NULL
int mb_dist_extract_records(mb_ctx* c, int rank, int world, void** d_send_keys, void** d_send_vals, uint64_t* h_counts) {
    if (!d_send_keys || !d_send_vals) return MB_E_ARG;
    TRY(mb_dist_extract_count(c, rank, world, h_counts));
    TRY(mb_dist_partition(c, nullptr, nullptr, d_send_keys));
    *d_send_vals = c->fmt.wide ? c->valsB.p : nullptr;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MB_OK;
}

// ---- peer memory (fused exchange #1): a fixed receive array other ranks' partition kernels write into
int mb_dist_p2p_recv_array(mb_ctx* c, uint64_t capacity_records, void** d_ptr) {
    if (!c || !d_ptr) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (!c->x_recv.p || c->x_recv.cap < (capacity_records + 16) * 8) {
        if (c->x_recv.p) { cudaFree(c->x_recv.p); c->x_recv.p = nullptr; c->x_recv.cap = 0; }
        size_t bytes = (capacity_records + 16) * 8;
        cudaError_t e = cudaMalloc(&c->x_recv.p, bytes); // exact, never regrown behind the peers' backs
        if (e != cudaSuccess) { c->set_cuda_error(e, "cudaMalloc(p2p receive array)", __LINE__); return MB_E_NOMEM; }
        c->x_recv.cap = bytes;
    }
    *d_ptr = c->x_recv.p;
    return MB_OK;
}
int mb_ipc_export(mb_ctx* c, void* d_ptr, uint8_t* handle64) {
    if (!c || !d_ptr || !handle64) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(c, cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle64, &h, 64);
    return MB_OK;
}
int mb_ipc_import(mb_ctx* c, const uint8_t* handle64, void** d_ptr) {
    if (!c || !d_ptr || !handle64) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CUDA_TRY(c, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MB_OK;
}
int mb_ipc_close(mb_ctx* c, void* d_ptr) {
    if (!c || !d_ptr) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaIpcCloseMemHandle(d_ptr));
    return MB_OK;
}
// stage 2 reads the seed records from the p2p receive array instead of receive buffer 0
int mb_dist_use_p2p_recv(mb_ctx* c, int on) {
    if (!c) return MB_E_ARG;
    c->d_use_p2p = on != 0;
    return MB_OK;
}

// receive buffers (library-owned device memory the caller's all-to-all writes into)
//   which: 0 seed records (first words), 2 second words of 16-byte seed records, 1 candidate rows, 3 match headers,
//          4 match components (n_words 8-byte words each), 5 verdict bytes (n_words = BYTES)
int mb_dist_recv_buffer(mb_ctx* c, int which, uint64_t n_words, void** d_ptr) {
    if (!c || !d_ptr || which < 0 || which > 5) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    DBuf* b = which == 0 ? &c->keysA : which == 2 ? &c->valsA : which == 5 ? &c->x_acc_r : (which == 1 || which == 3) ? &c->x_hdr_r : &c->x_comp_r;
    TRY(c->reserve(*b, which == 5 ? (size_t)n_words + 64 : ((size_t)n_words + 16) * 8));
    *d_ptr = b->p;
    return MB_OK;
}

// stage 2 (source side): n_recv seed records of this rank's key range, concatenated in source-rank order (receive
// buffer 0 or the p2p receive array).  Sort / runs / policy -> candidates in ascending seed order; EVERY candidate
// is extended here (a pure function of the candidate and the replicated genomes), so that only 4-word rows
// (two group hashes, first component, extents) travel to the owner of the de-dup group — the component lists stay.
int mb_dist_local(mb_ctx* c, const mb_params* prm, uint64_t n_recv, uint64_t* h_row_counts) {
    if (!c || !prm || !h_row_counts) return MB_E_ARG;
    if (prm->mode != MB_MODE_UNIQUE) return MB_E_ARG;
    if (n_recv >= (1ull << 31)) return MB_E_TOOLONG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const RecFmt& fmt = c->fmt;
    const int world = c->d_world;
    const u32 n = (u32)n_recv;
    u64* scal = c->scalars.as<u64>();
    for (int r = 0; r < world; ++r) h_row_counts[r] = 0;
    c->d_ncand = 0; c->d_nccomp = 0;
    c->d_bound.assign(world + 1, 0); c->d_cbound.assign(world + 1, 0);
    cudaEventRecord(c->ev_d[2], st);
    TRY(c->reserve(c->keysB, ((size_t)n + 8) * 8));
    u64 *kA = c->d_use_p2p ? c->x_recv.as<u64>() : c->keysA.as<u64>(), *kB = c->keysB.as<u64>();
    if (c->d_use_p2p && (size_t)(n + 8) * 8 > c->x_recv.cap) return MB_E_STATE;
    TRY(mbi_sort_records(c, &kA, &kB, nullptr, nullptr, n, fmt.kshift, fmt.kbits, false, true));
    c->sorted_keys = kA; c->sorted_vals = nullptr;
    cudaEventRecord(c->ev_d[3], st);
    if (n == 0) { cudaEventRecord(c->ev_d[4], st); cudaEventRecord(c->ev_d[5], st); return MB_OK; }
    u32* run_start = reinterpret_cast<u32*>(kB);
    u32* run_u = run_start + (n + 2);
    const unsigned short* run_masks = launch_find_runs(kA, nullptr, n, fmt, run_start, run_u, c->status_slice(find_runs_workspace_words(n)), c->ticket(), nullptr,
                     reinterpret_cast<u32*>(scal + SC_RUNS), st);
    LAUNCHED(c); LAUNCHED(c); LAUNCHED(c); CHECK_LAUNCH(c);
    const u32 cand_cap = n / 2 + 2;
    TRY(c->reserve(c->cand_run, (size_t)cand_cap * 4));
    TRY(c->reserve(c->q_off, (size_t)(cand_cap + 1) * 4));
    SelectArgs sa{};
    sa.keys = kA; sa.vals = nullptr; sa.run_start = run_start; sa.run_u = run_u;
    sa.n_runs_ptr = reinterpret_cast<u32*>(scal + SC_RUNS);
    sa.mode = prm->mode; sa.direct_only = prm->direct_only;
    sa.min_multi = prm->min_multi; sa.max_multi = prm->max_multi; sa.nway_mask = prm->nway_mask;
    sa.status = c->status_slice(div_up(n, select_tile())); sa.ticket = c->ticket();
    sa.n_buckets = scal + SC_NBUCKETS;
    sa.totals = reinterpret_cast<u32*>(scal + SC_CAND);
    sa.cand_run = c->cand_run.as<u32>(); sa.cand_off = c->q_off.as<u32>(); sa.cand_aux = nullptr;
    launch_select(sa, fmt, n, st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    TRY(mbi_read_scalars(c));
    const u32* hs32 = reinterpret_cast<const u32*>(c->h_scal);
    const u64* hs64 = reinterpret_cast<const u64*>(c->h_scal);
    const u32 n_cand = hs32[2 * SC_CAND], n_ccomp = hs32[2 * SC_CAND + 1];
    c->stats.n_runs = hs32[2 * SC_RUNS]; c->stats.n_buckets = hs64[SC_NBUCKETS]; c->stats.n_candidates = n_cand;
    c->r_unique = hs32[2 * SC_RUNS];
    c->d_ncand = n_cand; c->d_nccomp = n_ccomp;
    if (n_cand == 0) { cudaEventRecord(c->ev_d[4], st); cudaEventRecord(c->ev_d[5], st); return MB_OK; }
    const size_t nc = (size_t)n_cand + 8;
    TRY(c->reserve(c->q_pos, ((size_t)n_ccomp + 8) * 4));
    TRY(c->reserve(c->q_gs, (size_t)n_ccomp + 8));
    TRY(c->reserve(c->q_el, nc * 4)); TRY(c->reserve(c->q_er, nc * 4));
    TRY(c->reserve(c->q_perm, nc * 4));
    TRY(c->reserve(c->ghash, nc * 8));
    TRY(c->reserve(c->ghash2, nc * 8));
    TRY(c->reserve(c->xrec, nc * 16)); TRY(c->reserve(c->xstate, nc * 16));
    TRY(c->reserve(c->wd_a, nc * 4)); TRY(c->reserve(c->wd_b, nc * 4)); TRY(c->reserve(c->wl_long, nc * 4));
    TRY(c->reserve(c->sort_kA, nc * 8)); TRY(c->reserve(c->sort_kB, nc * 8));
    TRY(c->reserve(c->sort_vA, nc * 8)); TRY(c->reserve(c->sort_vB, nc * 8));
    TRY(c->reserve(c->x_hdr_s, nc * 32));
    EmitUniqueArgs eu{};
    eu.keys = kA; eu.vals = nullptr; eu.run_start = run_start; eu.run_u = run_u;
    eu.cand_run = c->cand_run.as<u32>(); eu.cand_off = c->q_off.as<u32>(); eu.cand_aux = nullptr;
    eu.totals = reinterpret_cast<u32*>(scal + SC_CAND);
    eu.mode = prm->mode; eu.comp_pos = c->q_pos.as<u32>(); eu.comp_gs = c->q_gs.as<u8>(); eu.bitmap = nullptr; eu.ghash = c->ghash.as<u64>(); eu.ghash2 = c->ghash2.as<u64>();
    eu.seedL = (u32)c->sd.L; eu.masks = run_masks;
    launch_emit_unique(eu, fmt, c->gt, n_cand, st); LAUNCHED(c); CHECK_LAUNCH(c);
    cudaEventRecord(c->ev_d[4], st);
    // extension of every candidate (rep index = candidate; no slot axis here: the owner derives the slot ranges)
    CUDA_TRY(c, cudaMemsetAsync(scal + SC_DDCTR, 0, 8 * 8, st));
    DedupArgs da{};
    da.packed = c->packed.as<u64>(); da.n_cand = n_cand; da.n_rep = n_cand;
    da.cand_off = c->q_off.as<u32>(); da.comp_pos = c->q_pos.as<u32>(); da.comp_gs = c->q_gs.as<u8>();
    da.xrec = c->xrec.as<uint4>(); da.xstate = c->xstate.as<uint4>();
    da.ext_l = c->q_el.as<u32>(); da.ext_r = c->q_er.as<u32>();
    da.wd0 = c->wd_a.as<u32>(); da.wd1 = c->wd_b.as<u32>(); da.wl_long = c->wl_long.as<u32>();
    da.ctr = reinterpret_cast<u32*>(scal + SC_DDCTR);
    launch_cand_xrec(da, c->gt, st);
    da.ghash = c->ghash.as<u64>(); da.ghash2 = c->ghash2.as<u64>();
    launch_extend(da, c->gt, c->sd, st);
    c->stats.kernel_launches += extend_launches();
    if (n_cand) TRY(mbi_extend_long(c, da));
    CHECK_LAUNCH(c);
    c->stats.n_extended = n_cand;
    cudaEventRecord(c->ev_d[5], st);
    // stable partition of the rows by owner of the de-dup group
    u64 *skA = c->sort_kA.as<u64>(), *skB = c->sort_kB.as<u64>(), *svA = c->sort_vA.as<u64>(), *svB = c->sort_vB.as<u64>();
    launch_owner_keys(c->ghash.as<u64>(), n_cand, (u32)world, skA, svA, st); LAUNCHED(c);
    const int obits = std::max(1, mbi_bits_for((u64)world - 1));
    TRY(mbi_sort_records(c, &skA, &skB, &svA, &svB, n_cand, 0, obits, false)); // leaves the owner histogram in c->hist
    c->d_sperm = svA;
    std::vector<u32> oh(256);
    CUDA_TRY(c, cudaMemcpyAsync(oh.data(), c->hist.p, 256 * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    u64 acc = 0;
    for (int r = 0; r < world; ++r) { h_row_counts[r] = oh[r]; acc += oh[r]; c->d_bound[r + 1] = acc; }
    if (acc != n_cand) return MB_E_STATE;
    return MB_OK;
}

// stage 2 of MB_MODE_UNIQUE_COUNT / MB_MODE_SEED_ENUM: n_recv seed records of this rank's key range (receive buffers 0 and,
// for 16-byte records, 2).  Equal seeds all live on one rank, so the counts of the ranks ADD UP (the caller's all-reduce of
// mb_fetch_result's unique_mers / unique_mers_per_seq) and the ranks' match lists are disjoint: every rank's piece is in
// canonical order (by first position); the union of the pieces is the result ("sum / concat only").
int mb_dist_enum_local(mb_ctx* c, const mb_params* prm, uint64_t n_recv) {
    if (!c || !prm) return MB_E_ARG;
    if (prm->mode != MB_MODE_UNIQUE_COUNT && prm->mode != MB_MODE_SEED_ENUM) return MB_E_ARG;
    if (prm->mode == MB_MODE_SEED_ENUM && c->seq_len.size() > 1) return MB_E_SEQCOUNT;
    if (n_recv >= (1ull << 31)) return MB_E_TOOLONG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const RecFmt& fmt = c->fmt;
    const u32 n = (u32)n_recv;
    for (int i = EV_START; i <= EV_EXTRACT; ++i) cudaEventRecord(c->ev[i], st);
    cudaEventRecord(c->ev_d[2], st);
    TRY(c->reserve(c->keysB, ((size_t)n + 8) * 8));
    if (fmt.wide) TRY(c->reserve(c->valsB, ((size_t)n + 8) * 8));
    if ((size_t)(n + 8) * 8 > c->keysA.cap || (fmt.wide && (size_t)(n + 8) * 8 > c->valsA.cap)) return MB_E_STATE; // receive buffers not set up
    u64 *kA = c->keysA.as<u64>(), *kB = c->keysB.as<u64>(), *vA = fmt.wide ? c->valsA.as<u64>() : nullptr, *vB = fmt.wide ? c->valsB.as<u64>() : nullptr;
    TRY(mbi_sort_records(c, &kA, &kB, fmt.wide ? &vA : nullptr, fmt.wide ? &vB : nullptr, n, fmt.kshift, fmt.kbits, false, true));
    c->sorted_keys = kA; c->sorted_vals = vA;
    c->n_seeds = n;
    cudaEventRecord(c->ev[EV_SORT], st);
    cudaEventRecord(c->ev_d[3], st);
    c->last_mode = prm->mode;
    c->r_matches = 0; c->r_comps = 0; c->r_unique = 0;
    if (n == 0) {
        for (int i = EV_BUCKET; i < EV_COUNT; ++i) cudaEventRecord(c->ev[i], st);
        memset(c->h_perseq, 0, MB_MAX_SEQ * 8);
        c->have_result = true;
        return MB_OK;
    }
    TRY(mbi_count_or_enum(c, prm, kA, kB, vA, vB, n));
    cudaEventRecord(c->ev_d[4], st); // "buckets" of mb_dist_stage_ms: runs / policy / output of the key range
    cudaEventRecord(c->ev_d[5], st);
    return MB_OK;
}

// destination table of a pack kernel -> device (x_peers).  bases == NULL: everything into this rank's own send
// buffers (the caller's all-to-all moves it); otherwise bases[d] / offsets[d] = rank d's receive buffer mapped
// into this process and the row (word) index where this rank's block starts in it.
static int upload_peer_table(mb_ctx* c, void* const* bases, const uint64_t* offs, void* local, void* const* bases2, const uint64_t* offs2, void* local2,
                             u64 n_rows) {
    const int world = c->d_world;
    std::vector<PeerDst> tab(world + 1);
    for (int d = 0; d <= world; ++d) {
        PeerDst& t = tab[d];
        const int e = std::min(d, world - 1);
        t.bound = (u32)(d < world ? c->d_bound[d] : n_rows);
        t.cbound = d < world ? c->d_cbound[d] : 0;
        t.base = (u64*)(bases ? bases[e] : local); t.off = bases ? offs[e] : c->d_bound[e];
        t.base2 = (u64*)(bases2 ? bases2[e] : local2); t.off2 = bases2 ? offs2[e] : c->d_cbound[e];
        t.pad = 0;
    }
    TRY(c->reserve(c->x_peers, 257 * sizeof(PeerDst)));
    CUDA_TRY(c, cudaMemcpyAsync(c->x_peers.p, tab.data(), tab.size() * sizeof(PeerDst), cudaMemcpyHostToDevice, c->stream)); // pageable: staged before return
    return MB_OK;
}

// stage 2, last step: write the candidate rows in owner order — into this rank's send buffer (*d_rows, for the
// caller's all-to-all; peer_bases = NULL) or, fused with exchange 2, straight into the owners' receive buffers over
// NVLink (peer_bases[d] = rank d's receive buffer 1 mapped into this process, peer_row_offsets[d] = row index of this
// rank's block there).  Asynchronous; with peer stores all ranks must synchronise before stage 3a.
int mb_dist_rows_pack(mb_ctx* c, void* const* peer_bases, const uint64_t* peer_row_offsets, void** d_rows) {
    if (!c || (peer_bases && !peer_row_offsets) || (!peer_bases && !d_rows)) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const u32 n_cand = c->d_ncand;
    if (d_rows) *d_rows = c->x_hdr_s.p;
    if (n_cand == 0) return MB_OK;
    TRY(upload_peer_table(c, peer_bases, peer_row_offsets, c->x_hdr_s.p, nullptr, nullptr, nullptr, n_cand));
    launch_pack_rows(c->d_sperm, n_cand, c->ghash.as<u64>(), c->ghash2.as<u64>(), c->q_off.as<u32>(), c->q_pos.as<u32>(), c->q_gs.as<u8>(), c->gt,
                     c->q_el.as<u32>(), c->q_er.as<u32>(), c->x_peers.as<PeerDst>(), (u32)c->d_world, c->q_perm.as<u32>(), st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    return MB_OK;
}

// Push the blocks of a local send buffer (as laid out by mb_dist_rows_pack / mb_dist_match_pack with bases = NULL:
// destination blocks back to back, counts[d] units of `unit_words` 8-byte words each) into the destination ranks'
// receive buffers with device-to-device copies on the context stream: the copy engines move the bytes over NVLink
// while the SMs are free.  peer_bases[d] = rank d's receive buffer mapped into this process, dst_offsets[d] = unit
// index of this rank's block there.  Asynchronous; all ranks must synchronise before the receivers read.
int mb_dist_push(mb_ctx* c, const void* d_src, const uint64_t* counts, uint32_t unit_words, void* const* peer_bases, const uint64_t* dst_offsets) {
    if (!c || !counts || !peer_bases || !dst_offsets || unit_words == 0) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t unit = (size_t)unit_words * 8;
    const int world = c->d_world;
    std::vector<size_t> so(world + 1, 0); // byte offset of every destination's block in the send buffer
    for (int d = 0; d < world; ++d) so[d + 1] = so[d] + (size_t)counts[d] * unit;
    // start with the next rank so that the ranks do not all push to the same destination at the same time
    for (int k = 0; k < world; ++k) {
        const int d = (c->d_rank + 1 + k) % world;
        if (counts[d] == 0) continue;
        if (!d_src || !peer_bases[d]) return MB_E_ARG;
        CUDA_TRY(c, cudaMemcpyAsync((char*)peer_bases[d] + (size_t)dst_offsets[d] * unit, (const char*)d_src + so[d], (size_t)counts[d] * unit,
                                    cudaMemcpyDeviceToDevice, c->stream));
    }
    return MB_OK;
}

// stage 3a (owner side): n_rows candidate rows in receive buffer 1, in source-rank order (= ascending seed order
// inside every group).  Chains / reps / resolve over the owned groups (no extension: the extents came with the rows);
// *d_verdict = one byte per row, in row order (1 accepted, 0 dropped), to be returned to the rows' sources.
int mb_dist_resolve(mb_ctx* c, uint64_t n_rows64, void** d_verdict) {
    if (!c || !d_verdict) return MB_E_ARG;
    if (n_rows64 >= (1ull << 31)) return MB_E_TOOLONG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const u32 n = (u32)n_rows64;
    c->n_rep = 0;
    TRY(c->reserve(c->x_acc_s, (size_t)n + 64));
    CUDA_TRY(c, cudaMemsetAsync(c->x_acc_s.p, 0, (size_t)n + 8, st));
    *d_verdict = c->x_acc_s.p;
    cudaEventRecord(c->ev_d[6], st);
    if (n == 0) { cudaEventRecord(c->ev_d[7], st); return MB_OK; }
    TRY(mbi_reserve_candidates(c, n, 0, c->d_bases));
    const u64 bm_words = c->d_bases / 64 + 2;
    CUDA_TRY(c, cudaMemsetAsync(c->bitmap.p, 0, bm_words * 8, st));
    const u64* rows = c->x_hdr_r.as<u64>();
    launch_rows_bitmap(rows, n, c->gt, c->bitmap.as<u64>(), st); LAUNCHED(c);
    TRY(mbi_dedup(c, n, c->d_bases, rows));
    launch_accept_mark(c->s_cand.as<u8>(), c->rep_cand.as<u32>(), c->n_rep, c->x_acc_s.as<u8>(), st); LAUNCHED(c); CHECK_LAUNCH(c);
    cudaEventRecord(c->ev_d[7], st);
    return MB_OK;
}

// stage 3b (source side): the verdicts of this rank's rows are in receive buffer 5, in the row order of stage 2.
// Accepted candidates = matches; leaves a 4096-bin histogram of their canonical sort keys at *d_hist (uint64
// counts, library-owned) for the caller to sum over all ranks IN PLACE before stage 3c.
int mb_dist_accept(mb_ctx* c, void** d_hist) {
    if (!c || !d_hist) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    u64* scal = c->scalars.as<u64>();
    const u32 n_cand = c->d_ncand;
    c->d_nmatch = 0; c->d_nmcomp = 0;
    TRY(c->reserve(c->x_counts, 4096 * 8));
    CUDA_TRY(c, cudaMemsetAsync(c->x_counts.p, 0, 4096 * 8, st));
    *d_hist = c->x_counts.p;
    if (n_cand == 0) { CUDA_TRY(c, cudaStreamSynchronize(st)); return MB_OK; }
    if (!c->x_acc_r.p || c->x_acc_r.cap < n_cand) return MB_E_STATE;
    const size_t nc = (size_t)n_cand + 8;
    TRY(c->reserve(c->q_state, nc));
    TRY(c->reserve(c->q_item, nc * 4));
    TRY(c->reserve(c->flags, nc * 4));
    TRY(c->reserve(c->match_idx, nc * 4));
    launch_apply_accept(c->x_acc_r.as<u8>(), c->q_perm.as<u32>(), n_cand, c->q_state.as<u8>(), c->q_item.as<u32>(), st); LAUNCHED(c);
    OutputArgs oa{};
    oa.n_items = n_cand; oa.state = c->q_state.as<u8>(); oa.item_cand = c->q_item.as<u32>(); oa.flags = c->flags.as<u32>();
    launch_uniq_flags(oa, st); LAUNCHED(c);
    launch_scan_u32(c->flags.as<u32>(), n_cand, c->match_idx.as<u32>(), nullptr, c->status_slice(div_up(n_cand, scan_tile())), c->ticket(),
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
    launch_match_keys(c->q_state.as<u8>(), c->q_item.as<u32>(), c->match_idx.as<u32>(), c->q_off.as<u32>(), c->q_gs.as<u8>(), c->q_pos.as<u32>(),
                      c->q_el.as<u32>(), n_cand, sbits, std::max(0, sbits + 6 - 12), c->x_key.as<u64>(), c->x_item.as<u32>(), c->x_counts.as<u64>(), st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    return MB_OK;
}

// stage 3c: *d_hist now holds the histogram summed over all ranks.  Partition the accepted matches by
// destination = range of the canonical sort key (ranks hold ascending ranges of the final order) and pack
// their rows; per-destination row / component-word counts go to the host arrays.
int mb_dist_match_partition(mb_ctx* c, uint64_t* h_match_counts, uint64_t* h_comp_counts) {
    if (!c || !h_match_counts || !h_comp_counts) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int world = c->d_world;
    const u32 n_match = c->d_nmatch;
    for (int r = 0; r < world; ++r) { h_match_counts[r] = 0; h_comp_counts[r] = 0; }
    c->d_bound.assign(world + 1, 0); c->d_cbound.assign(world + 1, 0);
    c->d_nmcomp = 0;
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
    launch_match_perm_m(svA, c->x_item.as<u32>(), c->q_item.as<u32>(), c->q_off.as<u32>(), n_match, c->x_m.as<u32>(), st); LAUNCHED(c);
    launch_scan_u32(c->x_m.as<u32>(), n_match, nullptr, c->out_off.as<u64>(), c->status_slice(div_up(n_match, scan_tile())), c->ticket(), nullptr, st);
    LAUNCHED(c);
    std::vector<u32> oh(256);
    CUDA_TRY(c, cudaMemcpyAsync(oh.data(), c->hist.p, 256 * 4, cudaMemcpyDeviceToHost, st));
    u64 n_mcomp = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&n_mcomp, c->out_off.as<u64>() + n_match, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    u64 acc = 0;
    for (int r = 0; r < world; ++r) { h_match_counts[r] = oh[r]; acc += oh[r]; c->d_bound[r + 1] = acc; }
    if (acc != n_match) return MB_E_STATE;
    for (int r = 1; r <= world; ++r) CUDA_TRY(c, cudaMemcpyAsync(&c->d_cbound[r], c->out_off.as<u64>() + c->d_bound[r], 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    for (int r = 0; r < world; ++r) h_comp_counts[r] = c->d_cbound[r + 1] - c->d_cbound[r];
    c->d_nmcomp = (u32)n_mcomp;
    c->d_sperm = svA;
    return MB_OK;
}

// stage 3c, last step: write the match rows in destination order — into this rank's send buffers (*d_hdr, *d_comps;
// bases = NULL) or, fused with exchange 3, into the destination ranks' receive buffers 3 / 4 over NVLink
// (hdr_row_offsets[d] / comp_word_offsets[d] = where this rank's block starts there).  Asynchronous.
int mb_dist_match_pack(mb_ctx* c, void* const* hdr_bases, const uint64_t* hdr_row_offsets, void* const* comp_bases, const uint64_t* comp_word_offsets,
                       void** d_hdr, void** d_comps) {
    const bool peers = hdr_bases != nullptr;
    if (!c || (peers && (!hdr_row_offsets || !comp_bases || !comp_word_offsets)) || (!peers && (!d_hdr || !d_comps))) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const u32 n_match = c->d_nmatch;
    TRY(c->reserve(c->x_hdr_s, ((size_t)n_match + 8) * 16));
    TRY(c->reserve(c->x_comp_s, ((size_t)c->d_nmcomp + 8) * 8));
    if (d_hdr) *d_hdr = c->x_hdr_s.p;
    if (d_comps) *d_comps = c->x_comp_s.p;
    if (n_match == 0) return MB_OK;
    TRY(upload_peer_table(c, hdr_bases, hdr_row_offsets, c->x_hdr_s.p, comp_bases, comp_word_offsets, c->x_comp_s.p, n_match));
    launch_pack_match_perm(c->d_sperm, c->x_item.as<u32>(), c->q_item.as<u32>(), c->out_off.as<u64>(), c->q_off.as<u32>(), c->q_pos.as<u32>(),
                           c->q_gs.as<u8>(), c->q_el.as<u32>(), c->q_er.as<u32>(), n_match, c->x_peers.as<PeerDst>(), (u32)c->d_world, st);
    LAUNCHED(c); CHECK_LAUNCH(c);
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
        launch_unpack_match(hdr, comps, c->cand_off.as<u32>(), n_match, n_mcomp, c->comp_pos.as<u32>(), c->comp_gs.as<u8>(), c->ext_l.as<u32>(),
                            c->ext_r.as<u32>(), c->s_cand.as<u8>(), c->rep_cand.as<u32>(), st);
        LAUNCHED(c); CHECK_LAUNCH(c);
        TRY(mbi_output_unique(c, n_match, c->d_maxlen));
    }
    c->stats.n_matches = c->r_matches; c->stats.n_comps = c->r_comps;
    // the stage events of the single-GPU driver are not recorded on this path
    for (int i = 0; i < EV_COUNT; ++i) cudaEventRecord(c->ev[i], st);
    c->have_result = true;
    return MB_OK;
}

// device milliseconds of this rank's stages of the last distributed run: [0] extract+partition, [1] sort,
// [2] runs/policy/candidates, [3] extension of all candidates, [4] chains + resolve of the owned groups; [5..7] zero
int mb_dist_stage_ms(mb_ctx* c, float* out8) {
    if (!c || !out8) return MB_E_ARG;
    auto ms = [&](int a, int b) { float t = 0; if (!c->ev_d[a] || !c->ev_d[b] || cudaEventElapsedTime(&t, c->ev_d[a], c->ev_d[b]) != cudaSuccess) { cudaGetLastError(); t = 0; } return t; };
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    out8[0] = ms(0, 1); out8[1] = ms(2, 3); out8[2] = ms(3, 4); out8[3] = ms(4, 5); out8[4] = ms(6, 7);
    return MB_OK;
}

} // extern "C"

// api.cu — the C ABI of include/mauve_b200.h: context, inputs, the seed-to-multi-MUM pipeline driver,
// result transfer.  No CPU fallback: every compute entry point needs a CUDA device.
#include "ctx.h"

static void free_buf(DBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

static int build_seed(uint64_t pattern, SeedDev& sd) {
    int L = 0; { u64 p = pattern; while (p) { ++L; p >>= 1; } }
    int w = __builtin_popcountll(pattern);
    if (L == 0 || !(w & 1) || w < 3 || w > 31) return MB_E_SEED;
    for (int j = 0; j < L; ++j)
        if (((pattern >> j) & 1) != ((pattern >> (L - 1 - j)) & 1)) return MB_E_SEED;
    memset(&sd, 0, sizeof(sd));
    sd.L = L; sd.w = w; sd.wide = L > 32;
    // window offset j (0 = first base) sits at left-aligned bits [126-2j, 127-2j] of hi:lo
    auto set2 = [](u64& hi, u64& lo, int j) {
        int b = 126 - 2 * j;
        if (b >= 64) hi |= 3ull << (b - 64); else lo |= 3ull << b;
    };
    int nrun = 0, j = 0, gaps = 0;
    while (j < L) {
        bool care = (pattern >> (L - 1 - j)) & 1;
        if (!care) { ++gaps; ++j; continue; }
        u64 hi = 0, lo = 0;
        while (j < L && ((pattern >> (L - 1 - j)) & 1)) { set2(hi, lo, j); set2(sd.mask_hi, sd.mask_lo, j); ++j; }
        if (nrun >= 32) return MB_E_SEED;
        sd.runmask_hi[nrun] = hi; sd.runmask_lo[nrun] = lo; sd.lshift[nrun] = 2 * gaps;
        ++nrun;
    }
    sd.nrun = nrun;
    int t = 0;
    for (int o = 0; o < L; ++o)
        if ((pattern >> (L - 1 - o)) & 1) sd.care_off[t++] = (u8)o;
    return MB_OK;
}

extern "C" {

const char* mb_version(void) { return "mauvealigner_b200 0.1 (sm_100a)"; }

const char* mb_strerror(int code) {
    switch (code) {
    case MB_OK: return "ok";
    case MB_E_ARG: return "bad argument";
    case MB_E_SEED: return "seed pattern must be palindromic with odd weight in [3,31]";
    case MB_E_NOSEQ: return "no sequence added";
    case MB_E_SEQCOUNT: return "sequence count not supported by this mode";
    case MB_E_TOOLONG: return "sequence too long";
    case MB_E_CUDA: return "CUDA error";
    case MB_E_NOMEM: return "out of memory";
    case MB_E_STATE: return "call out of order";
    default: return "unknown error";
    }
}

int mb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* mb_last_cuda_error(mb_ctx* ctx) { return ctx ? ctx->err : ""; }

int mb_ctx_create(mb_ctx** out, int device) {
    if (!out) return MB_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) { cudaGetLastError(); return MB_E_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MB_E_CUDA;
    if (prop.major != 10) return MB_E_CUDA; // sm_100a cubins only; no other code path exists
    mb_ctx* c = new (std::nothrow) mb_ctx();
    if (!c) return MB_E_NOMEM;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return MB_E_CUDA; }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return MB_E_CUDA; }
    c->own_stream = true;
    for (int i = 0; i < EV_COUNT; ++i) cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < 4; ++i) cudaEventCreate(&c->ev_x[i]);
    for (int i = 0; i < 16; ++i) cudaEventCreate(&c->ev_r[i]);
    cudaMallocHost(&c->h_perseq, MB_MAX_SEQ * sizeof(u64));
    cudaMallocHost(&c->h_scal, SC_COUNT * sizeof(u64));
    *out = c;
    return MB_OK;
}

int mb_ctx_destroy(mb_ctx* c) {
    if (!c) return MB_E_ARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DBuf* bufs[] = {&c->packed, &c->ascii_stage, &c->keysA, &c->keysB, &c->valsA, &c->valsB, &c->hist, &c->digit_base, &c->lookback, &c->tickets,
                    &c->status, &c->scalars, &c->per_seq, &c->tile_first, &c->cand_run, &c->cand_off, &c->cand_aux, &c->comp_pos, &c->comp_gs,
                    &c->bitmap, &c->bmrank, &c->cand_at, &c->cstate, &c->covered, &c->ghash2, &c->rep_cand, &c->s_h2, &c->reach, &c->xstate, &c->xrec, &c->x_lut, &c->x_counts, &c->x_hdr_s, &c->x_comp_s, &c->x_hdr_r, &c->x_comp_r, &c->x_m, &c->x_key, &c->x_item, &c->x_peers, &c->x_recv, &c->q_off, &c->q_pos, &c->q_gs, &c->q_el, &c->q_er, &c->q_perm, &c->q_state, &c->q_item, &c->x_acc_s, &c->x_acc_r, &c->minrank, &c->ext_l, &c->ext_r, &c->wl_a, &c->wl_b, &c->wl_c, &c->wl_long, &c->wd_a, &c->wd_b, &c->wd_c, &c->live_bits, &c->trace, &c->ghash, &c->slot_gp, &c->slot_hash, &c->link_bits, &c->chain_min, &c->rep_bits, &c->rep_rank, &c->s_hash, &c->s_cand, &c->rng_lo, &c->rng_hi, &c->flags,
                    &c->match_idx, &c->sort_kA, &c->sort_kB, &c->sort_vA, &c->sort_vB, &c->ncomp, &c->mers_tmp, &c->out_len, &c->out_off,
                    &c->out_seq, &c->out_start, &c->fam_th1, &c->fam_th2, &c->fam_tx, &c->fam_tend, &c->fam_sh1, &c->fam_sh2, &c->fam_sx, &c->fam_send,
                    &c->fam_spmax, &c->fam_srun0, &c->fam_drop, &c->seg, &c->pos_match, &c->pos_comp, &c->wide_seq, &c->wide_start, &c->shadow, &c->cell_r, &c->cell_new};
    for (DBuf* b : bufs) free_buf(*b);
    void* hs[] = {c->h_len, c->h_off, c->h_seq, c->h_start, c->h_perseq, c->h_scal, c->h_posm, c->h_posc};
    for (void* h : hs) if (h) cudaFreeHost(h);
    for (int i = 0; i < EV_COUNT; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 4; ++i) if (c->ev_x[i]) cudaEventDestroy(c->ev_x[i]);
    for (int i = 0; i < 16; ++i) if (c->ev_r[i]) cudaEventDestroy(c->ev_r[i]);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return MB_OK;
}

int mb_set_stream(mb_ctx* c, void* s) {
    if (!c) return MB_E_ARG;
    cudaSetDevice(c->device);
    if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (s) { c->stream = (cudaStream_t)s; c->own_stream = false; }
    else {
        CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    return MB_OK;
}

/* The MemHash table of libMems persists across FindMatches calls until Clear() (the seed-family search of
 * src/progressiveMauve.cpp:503-548 relies on it).  on != 0: every MB_MODE_UNIQUE search from now on drops the candidates
 * whose seed lies inside a match of an earlier search of the same group and adds its own matches to the table;
 * on == 0: the table is forgotten (MemHash::Clear()). */
int mb_accumulate(mb_ctx* c, int on) {
    if (!c) return MB_E_ARG;
    if (on) { c->fam_on = true; return MB_OK; }
    c->fam_on = false; c->fam_n = 0; c->fam_dirty = false; c->fam_pre = false;
    return MB_OK;
}

int mb_clear_sequences(mb_ctx* c) {
    if (!c) return MB_E_ARG;
    c->seq_len.clear(); c->seq_word_base.clear(); c->words_used = 0; c->have_result = false;
    c->n_seg = 0; c->h_seg.clear();
    c->stats.h2d_bytes = 0;
    return MB_OK;
}

int mb_set_seed(mb_ctx* c, uint64_t pattern) {
    if (!c) return MB_E_ARG;
    SeedDev sd;
    TRY(build_seed(pattern, sd));
    c->sd = sd; c->pattern = pattern; c->seed_set = true;
    return MB_OK;
}

// grow the packed-genome buffer, keeping its contents
static int grow_packed(mb_ctx* c, u64 need_words) {
    const u64 slack = 512; // read-ahead of the extraction tiles past the last genome
    size_t bytes = (need_words + slack) * 8;
    if (bytes <= c->packed.cap) return MB_OK;
    DBuf nb;
    size_t want = bytes * 2;
    cudaError_t e = cudaMalloc(&nb.p, want);
    if (e != cudaSuccess) { c->set_cuda_error(e, "cudaMalloc(packed)", __LINE__); return MB_E_NOMEM; }
    nb.cap = want;
    CUDA_TRY(c, cudaMemsetAsync(nb.p, 0, want, c->stream));
    if (c->packed.p && c->words_used) CUDA_TRY(c, cudaMemcpyAsync(nb.p, c->packed.p, c->words_used * 8, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    free_buf(c->packed);
    c->packed = nb;
    return MB_OK;
}

static int add_sequence_common(mb_ctx* c, const void* data, uint64_t len, int kind /*0 host ascii, 1 host packed, 2 device ascii, 3 device packed*/, int* out_id) {
    if (!c || (!data && len)) return MB_E_ARG;
    if (len >= (1ull << 32)) return MB_E_TOOLONG;
    if (c->seq_len.size() >= MB_MAX_SEQ) return MB_E_SEQCOUNT;
    CUDA_TRY(c, cudaSetDevice(c->device));
    u64 n_words = (len + 31) / 32;
    // every genome has >= MB_PAD_WORDS zero words on both sides (the extension reads a little past either end)
    u64 base = (std::max<u64>(c->words_used, MB_PAD_WORDS) + 1) & ~1ull; // 16-byte aligned
    u64 end = base + n_words + MB_PAD_WORDS;
    TRY(grow_packed(c, end));
    u64* dst = c->packed.as<u64>() + base;
    // padding after the genome must read as zero
    CUDA_TRY(c, cudaMemsetAsync(dst, 0, (n_words + MB_PAD_WORDS) * 8, c->stream));
    if (len) {
        if (kind == 3) {
            CUDA_TRY(c, cudaMemcpyAsync(dst, data, n_words * 8, cudaMemcpyDeviceToDevice, c->stream));
        } else if (kind == 1) {
            CUDA_TRY(c, cudaMemcpyAsync(dst, data, n_words * 8, cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream)); // the caller keeps ownership of `data` (a pinned buffer would still be in flight)
            c->stats.h2d_bytes += n_words * 8;
        } else {
            const u8* d_ascii = (const u8*)data;
            if (kind == 0) {
                TRY(c->reserve(c->ascii_stage, len + 16));
                CUDA_TRY(c, cudaMemcpyAsync(c->ascii_stage.p, data, len, cudaMemcpyHostToDevice, c->stream));
                c->stats.h2d_bytes += len;
                d_ascii = c->ascii_stage.as<u8>();
            }
            launch_pack(d_ascii, len, dst, n_words, c->stream);
            CHECK_LAUNCH(c);
            if (kind == 0) CUDA_TRY(c, cudaStreamSynchronize(c->stream)); // the staging buffer is reused by the next call
        }
    }
    c->seq_len.push_back(len);
    c->seq_word_base.push_back(base);
    c->words_used = end;
    c->have_result = false;
    if (out_id) *out_id = (int)c->seq_len.size() - 1;
    return MB_OK;
}

int mb_add_sequence(mb_ctx* c, const uint8_t* data, uint64_t len, int is_packed, int* out_id) {
    return add_sequence_common(c, data, len, is_packed ? 1 : 0, out_id);
}
int mb_add_sequence_device(mb_ctx* c, const void* dev_ascii, uint64_t len, int* out_id) {
    return add_sequence_common(c, dev_ascii, len, 2, out_id);
}
int mb_add_sequence_device_packed(mb_ctx* c, const void* dev_words, uint64_t len, int* out_id) {
    return add_sequence_common(c, dev_words, len, 3, out_id);
}
int mb_copy_packed_device(mb_ctx* c, int seq, void* dst_dev) {
    if (!c || seq < 0 || (size_t)seq >= c->seq_len.size() || !dst_dev) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const u64 nw = (c->seq_len[seq] + 31) / 32;
    if (nw) CUDA_TRY(c, cudaMemcpyAsync(dst_dev, c->packed.as<u64>() + c->seq_word_base[seq], nw * 8, cudaMemcpyDeviceToDevice, c->stream));
    return MB_OK;
}
int mb_get_packed_device(mb_ctx* c, int seq, const void** dev_words, uint64_t* n_words) {
    if (!c || seq < 0 || (size_t)seq >= c->seq_len.size() || !dev_words || !n_words) return MB_E_ARG;
    *dev_words = c->packed.as<u64>() + c->seq_word_base[seq];
    *n_words = (c->seq_len[seq] + 31) / 32;
    return MB_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------- sort driver
// LSD radix sort of n (key[, val]) records on key bits [shift, shift+kbits).  On return *kA/*vA hold
// the sorted data (the buffers are swapped as needed).
int mbi_sort_records(mb_ctx* c, u64** kA, u64** kB, u64** vA, u64** vB, u32 n, int shift, int kbits, bool hist_ready, bool time_passes) {
    int npass = (kbits + 7) / 8;
    if (n == 0 || npass == 0) return MB_OK;
    if (!hist_ready) {
        CUDA_TRY(c, cudaMemsetAsync(c->hist.p, 0, 8 * 256 * 4, c->stream));
        launch_hist(*kA, n, shift, kbits, npass, c->hist.as<u32>(), c->stream); LAUNCHED(c);
    }
    launch_scan_hist(c->hist.as<u32>(), c->digit_base.as<u32>(), npass, c->stream); LAUNCHED(c);
    u32 tiles = div_up(n, radix_tile_size());
    for (int ps = 0; ps < npass; ++ps) {
        CUDA_TRY(c, cudaMemsetAsync(c->lookback.p, 0, (size_t)tiles * 256 * 8, c->stream));
        int bits = std::min(8, kbits - 8 * ps);
        if (time_passes && ps < 8) cudaEventRecord(c->ev_r[2 * ps], c->stream);
        cudaError_t e = launch_onesweep(*kA, *kB, vA ? *vA : nullptr, vB ? *vB : nullptr, n, c->digit_base.as<u32>() + ps * 256,
                                        c->lookback.as<u64>(), c->ticket(), shift + 8 * ps, bits, c->stream);
        LAUNCHED(c);
        if (time_passes && ps < 8) { cudaEventRecord(c->ev_r[2 * ps + 1], c->stream); c->n_timed_passes = ps + 1; }
        if (e != cudaSuccess) { c->set_cuda_error(e, "onesweep", __LINE__); return MB_E_CUDA; }
        std::swap(*kA, *kB);
        if (vA) std::swap(*vA, *vB);
    }
    c->stats.radix_passes += npass;
    return MB_OK;
}

int mbi_bits_for(u64 maxval) { int b = 0; while (maxval) { ++b; maxval >>= 1; } return b; }

int mbi_read_scalars(mb_ctx* c) {
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->scalars.p, SC_COUNT * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MB_OK;
}

extern "C" {

} // extern "C"

// Genome table, record format, small workspaces and counters of one run (shared with api_dist.cu).
int mbi_setup_run(mb_ctx* c, MbiRun& r) {
    if (!c->seed_set) return MB_E_SEED;
    const u32 nseq = (u32)c->seq_len.size();
    if (nseq == 0) return MB_E_NOSEQ;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const SeedDev& sd = c->sd;
    const u32 L = sd.L;
    u64 h2d = c->stats.h2d_bytes;
    memset(&c->stats, 0, sizeof(c->stats));
    c->stats.h2d_bytes = h2d;
    c->have_result = false;
    c->ticket_next = 0; c->status_next = 0; c->n_timed_passes = 0;
    GenomeTable& gt = c->gt;
    memset(&gt, 0, sizeof(gt));
    gt.nseq = nseq;
    u64 n64 = 0, bases = 0, maxlen = 0;
    r.tile_first.assign(nseq + 1, 0);
    const u32 ET = extract_tile_size();
    u32 n_tiles = 0;
    for (u32 g = 0; g < nseq; ++g) {
        u64 len = c->seq_len[g];
        gt.word_base[g] = c->seq_word_base[g];
        gt.base_base[g] = bases;
        gt.len[g] = (u32)len;
        gt.seed_base[g] = (u32)n64;
        u64 ns = len >= L ? len - L + 1 : 0;
        r.tile_first[g] = n_tiles;
        n_tiles += div_up(ns, ET);
        n64 += ns; bases += len; maxlen = std::max(maxlen, len);
    }
    r.tile_first[nseq] = n_tiles;
    gt.pairwise = 0;
    for (u32 g = 0; g < nseq; ++g) gt.vbase[g] = gt.base_base[g];
    if (c->n_seg && c->h_seg.size() != (size_t)nseq * (c->n_seg + 1)) return MB_E_STATE; // segments set for another sequence count
    if (n64 >= (1ull << 31)) return MB_E_TOOLONG;
    const u32 n = (u32)n64;
    r.n = n; r.n_tiles = n_tiles; r.bases = bases; r.maxlen = maxlen;
    c->n_seeds = n;
    c->stats.n_seeds = n;
    RecFmt& fmt = c->fmt;
    fmt.kbits = 2 * sd.w;
    fmt.gbits = mbi_bits_for(nseq - 1);
    fmt.pbits = mbi_bits_for(maxlen ? maxlen - 1 : 0);
    if (fmt.pbits == 0) fmt.pbits = 1;
    if (c->n_seg) { // segmented search: problem index | max(seed, genome + position) bits
        gt.n_seg = c->n_seg;
        gt.seg_field = (u32)std::max(fmt.kbits, fmt.gbits + fmt.pbits);
        fmt.kbits = (int)gt.seg_field + mbi_bits_for(c->n_seg);
        if (fmt.kbits > 64) return MB_E_TOOLONG;
        TRY(c->reserve(c->seg, c->h_seg.size() * 4));
        CUDA_TRY(c, cudaMemcpyAsync(c->seg.p, c->h_seg.data(), c->h_seg.size() * 4, cudaMemcpyHostToDevice, st));
        gt.seg = c->seg.as<u32>();
    }
    fmt.wide = (fmt.kbits + fmt.gbits + fmt.pbits + 1) > 64;
    fmt.kshift = fmt.wide ? 0 : fmt.gbits + fmt.pbits + 1;
    c->stats.record_bytes = fmt.wide ? 16 : 8;
    TRY(c->reserve(c->hist, 8 * 256 * 4));
    TRY(c->reserve(c->digit_base, 8 * 256 * 4));
    // the scans and sorts of the later stages run over candidates, reps and matches: at most n / 2 of them, except for
    // MB_MODE_PAIRWISE, where a bucket of u unique genomes yields u (u - 1) / 2 candidates (up to n (nseq - 1) / 2)
    const u64 items = r.mode == MB_MODE_PAIRWISE ? std::max<u64>(n, (u64)n * (nseq - 1) / 2 + 2) : n;
    TRY(c->reserve(c->lookback, (size_t)(div_up(items, radix_tile_size()) + 1) * 256 * 8));
    TRY(c->reserve(c->tickets, 256 * 4));
    size_t status_words = find_runs_workspace_words(n) + div_up(n, select_tile()) + 6 * (size_t)div_up(items, scan_tile()) +
                          div_up(bases * std::min<u64>(nseq, 8) / 64 + 2, scan_tile()) + 2 * (size_t)div_up(items, chain_tile()) + 64;
    TRY(c->reserve(c->status, status_words * 8));
    TRY(c->reserve(c->scalars, SC_COUNT * 8));
    TRY(c->reserve(c->per_seq, MB_MAX_SEQ * 8));
    TRY(c->reserve(c->tile_first, (nseq + 1) * 4));
    CUDA_TRY(c, cudaMemsetAsync(c->tickets.p, 0, 256 * 4, st));
    CUDA_TRY(c, cudaMemsetAsync(c->status.p, 0, status_words * 8, st));
    CUDA_TRY(c, cudaMemsetAsync(c->scalars.p, 0, SC_COUNT * 8, st));
    CUDA_TRY(c, cudaMemsetAsync(c->per_seq.p, 0, MB_MAX_SEQ * 8, st));
    CUDA_TRY(c, cudaMemsetAsync(c->hist.p, 0, 8 * 256 * 4, st));
    CUDA_TRY(c, cudaMemcpyAsync(c->tile_first.p, r.tile_first.data(), (nseq + 1) * 4, cudaMemcpyHostToDevice, st));
    c->last_mode = -1;
    c->r_matches = 0; c->r_comps = 0; c->r_unique = 0;
    return MB_OK;
}

// MB_MODE_UNIQUE_COUNT / MB_MODE_SEED_ENUM over n sorted seed records (kA / vA; kB / vB = the free ping-pong buffers): runs of
// equal seed -> counts, or bucket policy -> matches in canonical order (by first position) + CSR.  Shared by mb_find_device
// and by the multi-GPU path (mb_dist_enum_local, where the records are one key range of the whole set).
int mbi_count_or_enum(mb_ctx* c, const mb_params* prm, u64* kA, u64* kB, u64* vA, u64* vB, u32 n) {
    cudaStream_t st = c->stream;
    const RecFmt& fmt = c->fmt;
    const int mode = prm->mode;
    const u32 L = c->sd.L;
    u64* scal = c->scalars.as<u64>();
    u32* run_start = reinterpret_cast<u32*>(kB);
    u32* run_u = fmt.wide ? reinterpret_cast<u32*>(vB) : run_start + (n + 2);
    const bool need_counts = mode == MB_MODE_UNIQUE_COUNT;
    launch_find_runs(kA, vA, n, fmt, run_start, nullptr, c->status_slice(find_runs_workspace_words(n)), c->ticket(),
                     need_counts ? c->per_seq.as<u64>() : nullptr, reinterpret_cast<u32*>(scal + SC_RUNS), st);
    LAUNCHED(c); LAUNCHED(c); LAUNCHED(c); CHECK_LAUNCH(c); // masks, one-block scan, compaction

    if (mode == MB_MODE_UNIQUE_COUNT) {
        for (int i = EV_BUCKET; i < EV_COUNT; ++i) cudaEventRecord(c->ev[i], st);
        CUDA_TRY(c, cudaMemcpyAsync(c->h_perseq, c->per_seq.p, MB_MAX_SEQ * 8, cudaMemcpyDeviceToHost, st));
        TRY(mbi_read_scalars(c));
        c->r_unique = reinterpret_cast<u32*>((u64*)c->h_scal + SC_RUNS)[0];
        c->stats.n_runs = c->r_unique;
        c->have_result = true;
        return MB_OK;
    }


    const size_t cand_cap = (size_t)n / 2 + 2; // one candidate per selected bucket
    if (cand_cap >= (1ull << 30)) return MB_E_TOOLONG;
    TRY(c->reserve(c->cand_run, cand_cap * 4));
    TRY(c->reserve(c->cand_off, (cand_cap + 1) * 4));
    SelectArgs sa{};
    sa.keys = kA; sa.vals = vA; sa.run_start = run_start; sa.run_u = run_u;
    sa.n_runs_ptr = reinterpret_cast<u32*>(scal + SC_RUNS);
    sa.mode = mode; sa.direct_only = prm->direct_only;
    sa.min_multi = prm->min_multi; sa.max_multi = prm->max_multi; sa.nway_mask = prm->nway_mask;
    sa.status = c->status_slice(div_up(n, select_tile())); sa.ticket = c->ticket();
    sa.n_buckets = scal + SC_NBUCKETS;
    sa.totals = reinterpret_cast<u32*>(scal + SC_CAND);
    sa.cand_run = c->cand_run.as<u32>(); sa.cand_off = c->cand_off.as<u32>(); sa.cand_aux = nullptr;
    launch_select(sa, fmt, n, st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    memset(c->h_perseq, 0, MB_MAX_SEQ * 8); // per-sequence counts are a MODE_UNIQUE_COUNT product
    TRY(mbi_read_scalars(c));
    const u32* hs32 = reinterpret_cast<const u32*>(c->h_scal);
    const u64* hs64 = reinterpret_cast<const u64*>(c->h_scal);
    const u32 n_runs = hs32[2 * SC_RUNS], n_cand = hs32[2 * SC_CAND], n_ccomp = hs32[2 * SC_CAND + 1];
    c->stats.n_runs = n_runs; c->stats.n_buckets = hs64[SC_NBUCKETS]; c->stats.n_candidates = n_cand;
    c->r_unique = n_runs;
    cudaEventRecord(c->ev[EV_BUCKET], st);

    {
        cudaEventRecord(c->ev[EV_DEDUP], st);
        // matches = selected buckets; canonical order (D18) = by first position
        TRY(c->reserve(c->sort_kA, (size_t)(n_cand + 8) * 8)); TRY(c->reserve(c->sort_kB, (size_t)(n_cand + 8) * 8));
        TRY(c->reserve(c->sort_vA, (size_t)(n_cand + 8) * 8)); TRY(c->reserve(c->sort_vB, (size_t)(n_cand + 8) * 8));
        TRY(c->reserve(c->ncomp, (size_t)(n_cand + 8) * 4));
        TRY(c->reserve(c->out_len, (size_t)(n_cand + 8) * 4));
        TRY(c->reserve(c->out_off, (size_t)(n_cand + 8) * 8));
        TRY(c->reserve(c->out_seq, (size_t)(n_ccomp + 8)));
        TRY(c->reserve(c->out_start, (size_t)(n_ccomp + 8) * 4));
        if (n_cand) {
            EmitEnumArgs ea{};
            ea.keys = kA; ea.vals = vA; ea.run_start = run_start; ea.cand_run = c->cand_run.as<u32>(); ea.cand_off = c->cand_off.as<u32>();
            ea.totals = reinterpret_cast<u32*>(scal + SC_CAND);
            u64 *skA = c->sort_kA.as<u64>(), *skB = c->sort_kB.as<u64>(), *svA = c->sort_vA.as<u64>(), *svB = c->sort_vB.as<u64>();
            ea.sort_key = skA; ea.sort_val = svA; ea.ncomp = c->ncomp.as<u32>();
            launch_enum_keys(ea, fmt, n_cand, st); LAUNCHED(c); CHECK_LAUNCH(c);
            TRY(mbi_sort_records(c, &skA, &skB, &svA, &svB, n_cand, 0, fmt.pbits, false));
            // component counts in sorted order -> offsets
            // (ncomp was written in candidate order; permute through the sorted values inside the gather scan input)
            OutputArgs oa{};
            oa.n_items = n_cand; oa.cand_off = c->cand_off.as<u32>(); oa.ncomp = c->ncomp.as<u32>();
            oa.n_matches_ptr = scal + SC_NMATCH;
            c->tmp_u64 = n_cand;
            CUDA_TRY(c, cudaMemcpyAsync(scal + SC_NMATCH, &c->tmp_u64, 8, cudaMemcpyHostToDevice, st));
            launch_uniq_ncomp(oa, svA, n_cand, st); LAUNCHED(c); CHECK_LAUNCH(c);
            launch_scan_u32(c->ncomp.as<u32>(), n_cand, nullptr, c->out_off.as<u64>(), c->status_slice(div_up(n_cand, scan_tile())), c->ticket(),
                            scal + SC_NCOMP, st);
            LAUNCHED(c); CHECK_LAUNCH(c);
            ea.sorted_val = svA; ea.out_off = c->out_off.as<u64>();
            ea.out_len = c->out_len.as<u32>(); ea.out_seq = c->out_seq.as<u8>(); ea.out_start = c->out_start.as<int32_t>();
            launch_enum_gather(ea, fmt, L, n_cand, st); LAUNCHED(c); CHECK_LAUNCH(c);
        }
        c->r_matches = n_cand; c->r_comps = n_ccomp;
        cudaEventRecord(c->ev[EV_OUTPUT], st);
        c->stats.n_matches = n_cand; c->stats.n_comps = n_ccomp;
        c->have_result = true;
        return MB_OK;
    }

}

extern "C" {

int mb_find_device(mb_ctx* c, const mb_params* prm) {
    if (!c || !prm) return MB_E_ARG;
    const int mode = prm->mode;
    if (mode < MB_MODE_UNIQUE || mode > MB_MODE_REPEAT) return MB_E_ARG;
    if (mode == MB_MODE_PAIRWISE && c->seq_len.size() > 8) return MB_E_SEQCOUNT; // the reference uses it for <= 4 genomes
    if ((mode == MB_MODE_SEED_ENUM || mode == MB_MODE_REPEAT) && c->seq_len.size() > 1) return MB_E_SEQCOUNT;
    MbiRun run;
    run.mode = mode;
    TRY(mbi_setup_run(c, run));
    cudaStream_t st = c->stream;
    const SeedDev& sd = c->sd;
    const u32 L = sd.L;
    GenomeTable& gt = c->gt;
    RecFmt& fmt = c->fmt;
    const u32 n = run.n, n_tiles = run.n_tiles;
    u64 bases = run.bases; // length of the candidate bitmap axis
    const u64 maxlen = run.maxlen;
    if (mode == MB_MODE_PAIRWISE) {
        gt.pairwise = 1;
        bases = 0;
        for (u32 g0 = 0; g0 < gt.nseq; ++g0)
            for (u32 g1 = g0 + 1; g1 < gt.nseq; ++g1) { gt.vbase[vgenome(gt, g0, g1)] = bases; bases += gt.len[g0]; }
    }
    const int npass = (fmt.kbits + 7) / 8;
    const size_t nrec = (size_t)n + 8;
    TRY(c->reserve(c->keysA, nrec * 8));
    TRY(c->reserve(c->keysB, nrec * 8));
    if (fmt.wide) { TRY(c->reserve(c->valsA, nrec * 8)); TRY(c->reserve(c->valsB, nrec * 8)); }
    u64* scal = c->scalars.as<u64>();

    cudaEventRecord(c->ev[EV_START], st);
    // ---- a3: extract seed records (+ all digit histograms)
    u64 *kA = c->keysA.as<u64>(), *kB = c->keysB.as<u64>(), *vA = fmt.wide ? c->valsA.as<u64>() : nullptr, *vB = fmt.wide ? c->valsB.as<u64>() : nullptr;
    launch_extract_records(c->packed.as<u64>(), kA, vA, c->hist.as<u32>(), npass, gt, sd, fmt, c->tile_first.as<u32>(), n_tiles, st);
    if (n_tiles) { LAUNCHED(c); CHECK_LAUNCH(c); }
    cudaEventRecord(c->ev[EV_EXTRACT], st);
    // ---- a4 + a6: one stable radix sort over the seed bits
    TRY(mbi_sort_records(c, &kA, &kB, fmt.wide ? &vA : nullptr, fmt.wide ? &vB : nullptr, n, fmt.kshift, fmt.kbits, true, true));
    c->sorted_keys = kA; c->sorted_vals = vA;
    cudaEventRecord(c->ev[EV_SORT], st);

    c->last_mode = mode;
    if (n == 0) {
        for (int i = EV_BUCKET; i < EV_COUNT; ++i) cudaEventRecord(c->ev[i], st);
        memset(c->h_perseq, 0, MB_MAX_SEQ * 8);
        c->have_result = true;
        return MB_OK;
    }

    if (mode == MB_MODE_UNIQUE_COUNT || mode == MB_MODE_SEED_ENUM) return mbi_count_or_enum(c, prm, kA, kB, vA, vB, n);

    // ---- a5/a6: runs of equal seed.  run arrays live in the now-free ping-pong buffer.
    u32* run_start = reinterpret_cast<u32*>(kB);
    u32* run_u = fmt.wide ? reinterpret_cast<u32*>(vB) : run_start + (n + 2);
    bool need_u = mode == MB_MODE_UNIQUE || mode == MB_MODE_PAIRWISE;
    bool need_counts = false;
    const unsigned short* run_masks = launch_find_runs(kA, vA, n, fmt, run_start, need_u ? run_u : nullptr, c->status_slice(find_runs_workspace_words(n)), c->ticket(),
                     need_counts ? c->per_seq.as<u64>() : nullptr, reinterpret_cast<u32*>(scal + SC_RUNS), st);
    LAUNCHED(c); LAUNCHED(c); LAUNCHED(c); CHECK_LAUNCH(c); // masks, one-block scan, compaction

    // ---- a7/a8: per-bucket policy -> candidates
    // one candidate per bucket; PAIRWISE: one per pair of unique genomes of a bucket (u (u-1) / 2 <= records * (nseq-1) / 2)
    const size_t cand_cap = mode == MB_MODE_PAIRWISE ? (size_t)n * (c->seq_len.size() - 1) / 2 + 2 : (size_t)n / 2 + 2;
    if (cand_cap >= (1ull << 30)) return MB_E_TOOLONG;
    TRY(c->reserve(c->cand_run, cand_cap * 4));
    TRY(c->reserve(c->cand_off, (cand_cap + 1) * 4));
    if (mode == MB_MODE_PAIRWISE) TRY(c->reserve(c->cand_aux, cand_cap * 4));
    SelectArgs sa{};
    sa.keys = kA; sa.vals = vA; sa.run_start = run_start; sa.run_u = run_u;
    sa.n_runs_ptr = reinterpret_cast<u32*>(scal + SC_RUNS);
    sa.mode = mode; sa.direct_only = prm->direct_only;
    sa.min_multi = prm->min_multi; sa.max_multi = prm->max_multi; sa.nway_mask = prm->nway_mask;
    sa.status = c->status_slice(div_up(n, select_tile())); sa.ticket = c->ticket();
    sa.n_buckets = scal + SC_NBUCKETS;
    sa.totals = reinterpret_cast<u32*>(scal + SC_CAND);
    sa.cand_run = c->cand_run.as<u32>(); sa.cand_off = c->cand_off.as<u32>(); sa.cand_aux = mode == MB_MODE_PAIRWISE ? c->cand_aux.as<u32>() : nullptr;
    launch_select(sa, fmt, n, st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    memset(c->h_perseq, 0, MB_MAX_SEQ * 8); // per-sequence counts are a MODE_UNIQUE_COUNT product
    TRY(mbi_read_scalars(c));
    const u32* hs32 = reinterpret_cast<const u32*>(c->h_scal);
    const u64* hs64 = reinterpret_cast<const u64*>(c->h_scal);
    const u32 n_runs = hs32[2 * SC_RUNS], n_cand = hs32[2 * SC_CAND], n_ccomp = hs32[2 * SC_CAND + 1];
    c->stats.n_runs = n_runs; c->stats.n_buckets = hs64[SC_NBUCKETS]; c->stats.n_candidates = n_cand;
    c->r_unique = n_runs;
    cudaEventRecord(c->ev[EV_BUCKET], st);

    // ---- MODE_UNIQUE: a9 candidates (HashMatch + SetDirection)
    TRY(mbi_reserve_candidates(c, n_cand, n_ccomp, bases));
    u32 n_matches = 0;
    u64 n_ocomp = 0;
    if (n_cand) {
        const u64 bm_words = bases / 64 + 2;
        CUDA_TRY(c, cudaMemsetAsync(c->bitmap.p, 0, bm_words * 8, st));
        EmitUniqueArgs eu{};
        eu.keys = kA; eu.vals = vA; eu.run_start = run_start; eu.run_u = run_u;
        eu.cand_run = c->cand_run.as<u32>(); eu.cand_off = c->cand_off.as<u32>(); eu.cand_aux = mode == MB_MODE_PAIRWISE ? c->cand_aux.as<u32>() : nullptr;
        eu.totals = reinterpret_cast<u32*>(scal + SC_CAND);
        eu.mode = mode; eu.comp_pos = c->comp_pos.as<u32>(); eu.comp_gs = c->comp_gs.as<u8>(); eu.bitmap = c->bitmap.as<u64>(); eu.ghash = c->ghash.as<u64>(); eu.ghash2 = c->ghash2.as<u64>();
        eu.seedL = L; eu.masks = run_masks;
        launch_emit_unique(eu, fmt, gt, n_cand, st); LAUNCHED(c); CHECK_LAUNCH(c);
        // ---- matches of earlier searches (mb_accumulate: the MemHash table persists until Clear())
        c->fam_pre = false;
        if (c->fam_on && mode == MB_MODE_UNIQUE) TRY(mbi_family_filter(c, n_cand));
        // ---- a10 + a11
        TRY(mbi_dedup(c, n_cand, bases));
        cudaEventRecord(c->ev[EV_DEDUP], st);
        // ---- a12
        TRY(mbi_output_unique(c, c->n_rep, maxlen));
        n_matches = (u32)c->r_matches; n_ocomp = c->r_comps;
        if (c->fam_on && mode == MB_MODE_UNIQUE) TRY(mbi_family_append(c));
    } else {
        cudaEventRecord(c->ev_x[0], st);
        cudaEventRecord(c->ev[EV_DEDUP], st);
    }
    cudaEventRecord(c->ev[EV_OUTPUT], st);
    c->r_matches = n_matches; c->r_comps = n_ocomp;
    c->stats.n_matches = n_matches; c->stats.n_comps = n_ocomp;
    c->have_result = true;
    return MB_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------- MODE_UNIQUE tail
// Shared by the single-GPU driver above and the distributed driver (api_dist.cu): the candidate
// CSR (cand_off, comp_pos, comp_gs, ghash; ascending seed order) is in place on this device.
int mbi_reserve_candidates(mb_ctx* c, u32 n_cand, u32 n_ccomp, u64 bases) {
    const u64 bm_words = bases / 64 + 2;
    const size_t nc = (size_t)n_cand + 8;
    TRY(c->reserve(c->cand_off, (nc + 1) * 4));
    TRY(c->reserve(c->comp_pos, (size_t)(n_ccomp + 8) * 4));
    TRY(c->reserve(c->comp_gs, (size_t)(n_ccomp + 8)));
    TRY(c->reserve(c->ghash, nc * 8));
    TRY(c->reserve(c->ghash2, nc * 8));
    TRY(c->reserve(c->bitmap, bm_words * 8));
    TRY(c->reserve(c->bmrank, (bm_words + 1) * 4));
    TRY(c->reserve(c->rep_bits, (nc / 64 + 2) * 8));
    TRY(c->reserve(c->rep_rank, (nc / 64 + 4) * 4));
    TRY(c->reserve(c->s_hash, nc * 16));
    TRY(c->reserve(c->s_cand, nc));
    TRY(c->reserve(c->rep_cand, nc * 4));
    TRY(c->reserve(c->s_h2, nc * 8));
    TRY(c->reserve(c->reach, nc * 4));
    TRY(c->reserve(c->xstate, nc * 16));
    TRY(c->reserve(c->xrec, nc * 16));
    TRY(c->reserve(c->sort_kA, nc * 8));
    TRY(c->reserve(c->sort_kB, nc * 8));
    TRY(c->reserve(c->slot_gp, nc * 32));
    TRY(c->reserve(c->link_bits, nc / 8 + 16));
    TRY(c->reserve(c->chain_min, nc * 4));
    TRY(c->reserve(c->minrank, nc * 16));
    TRY(c->reserve(c->ext_l, nc * 4));
    TRY(c->reserve(c->ext_r, nc * 4));
    TRY(c->reserve(c->rng_lo, nc * 4));
    TRY(c->reserve(c->rng_hi, nc * 4));
    TRY(c->reserve(c->wl_a, nc * 4));
    TRY(c->reserve(c->wl_b, nc * 4));
    TRY(c->reserve(c->wl_c, nc * 4));
    TRY(c->reserve(c->wl_long, nc * 4));
    TRY(c->reserve(c->wd_a, nc * 4));
    TRY(c->reserve(c->wd_b, nc * 4));
    TRY(c->reserve(c->wd_c, nc * 4));
    return MB_OK;
}

// a10 + a11: chains -> extension of the chain reps -> resolve (kernels_dedup.cu).  The candidate
// bitmap must hold one bit per candidate at (first genome, position).
// rows != null (multi-GPU owner side): the candidates are 4-word rows carrying their extents; no component lists,
// no extension here.
int mbi_dedup(mb_ctx* c, u32 n_cand, u64 bases, const u64* rows) {
    cudaStream_t st = c->stream;
    u64* scal = c->scalars.as<u64>();
    const u64 bm_words = bases / 64 + 2;
    const size_t cw = (size_t)n_cand / 64 + 2; // 64-bit words of a per-slot / per-rep bitmap
    CUDA_TRY(c, cudaMemsetAsync(c->rep_bits.p, 0, cw * 8, st));
    CUDA_TRY(c, cudaMemsetAsync(scal + SC_DDCTR, 0, 8 * 8, st));
    launch_scan_popc(c->bitmap.as<u64>(), bm_words, c->bmrank.as<u32>(), c->status_slice(div_up(bm_words, scan_tile())), c->ticket(),
                     scal + SC_BMTOTAL, st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    DedupArgs da{};
    da.packed = c->packed.as<u64>(); da.n_cand = n_cand;
    da.cand_off = c->cand_off.as<u32>(); da.comp_pos = c->comp_pos.as<u32>(); da.comp_gs = c->comp_gs.as<u8>();
    da.bitmap = c->bitmap.as<u64>(); da.bmrank = c->bmrank.as<u32>(); da.ghash = c->ghash.as<u64>(); da.ghash2 = c->ghash2.as<u64>();
    da.slot_rec = c->slot_gp.as<ulonglong2>();
    da.link_bits = c->link_bits.as<u8>(); da.chain_min = c->chain_min.as<u32>();
    da.rep_bits = c->rep_bits.as<u64>(); da.rep_rank = c->rep_rank.as<u32>();
    da.s_rec = c->s_hash.as<ulonglong2>(); da.rstate = c->s_cand.as<u8>(); da.s_cand = c->rep_cand.as<u32>(); da.s_h2 = c->s_h2.as<u64>();
    da.reach = c->reach.as<u32>(); da.xstate = c->xstate.as<uint4>(); da.xrec = c->xrec.as<uint4>();
    da.rng_lo = c->rng_lo.as<u32>(); da.rng_hi = c->rng_hi.as<u32>(); da.minrank = c->minrank.as<u64>();
    da.ext_l = c->ext_l.as<u32>(); da.ext_r = c->ext_r.as<u32>();
    da.wl0 = c->wl_a.as<u32>(); da.wl1 = c->wl_b.as<u32>(); da.wl2 = c->wl_c.as<u32>();
    da.wl_long = c->wl_long.as<u32>();
    da.wd0 = c->wd_a.as<u32>(); da.wd1 = c->wd_b.as<u32>(); da.wd2 = c->wd_c.as<u32>();
    da.ctr = reinterpret_cast<u32*>(scal + SC_DDCTR);
    da.rows = rows;
    da.pre_drop = (!rows && c->fam_pre) ? c->fam_drop.as<u8>() : nullptr;
    // ---- chains: slots, links, reps
    launch_slot_scatter(da, c->gt, st); LAUNCHED(c);
    {
        u32 tiles = div_up(n_cand, chain_tile());
        u64* sf = c->status_slice(tiles);
        u64* sb = c->status_slice(tiles);
        launch_chains(da, sf, c->ticket(), sb, c->ticket(), st);
        LAUNCHED(c); LAUNCHED(c); CHECK_LAUNCH(c);
    }
    launch_scan_popc(c->rep_bits.as<u64>(), cw, c->rep_rank.as<u32>(), c->status_slice(div_up(cw, scan_tile())), c->ticket(), scal + SC_UNDECIDED, st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    TRY(mbi_read_scalars(c));
    const u32 n_rep = (u32)reinterpret_cast<const u64*>(c->h_scal)[SC_UNDECIDED];
    c->stats.n_extended = n_rep;
    c->n_rep = n_rep;
    da.n_rep = n_rep;
    // ---- sort keys of the reps (group colour, slot) and, single-GPU path, their extension records in slot order
    u64 *skA = c->sort_kA.as<u64>(), *skB = c->sort_kB.as<u64>();
#ifndef DD_XORDER
#define DD_XORDER 1
#endif
    const u32 nvirt = c->gt.pairwise ? c->gt.nseq * (c->gt.nseq - 1) / 2 : c->gt.nseq;
    if (DD_XORDER && !rows && n_rep && nvirt > 1 && nvirt <= 256) {
        // extension records by (block of positions, first genome, slot): see k_cell_bounds
        u64 maxlen = 0;
        for (u32 g = 0; g < c->gt.nseq; ++g) maxlen = std::max<u64>(maxlen, c->gt.len[g]);
        const size_t cells = extension_cells(nvirt, maxlen);
        TRY(c->reserve(c->cell_r, (cells + 8) * 4)); TRY(c->reserve(c->cell_new, (cells + 8) * 4));
        launch_extension_cells(da, c->gt, nvirt, maxlen, bases, c->cell_r.as<u32>(), c->cell_new.as<u32>(), st); LAUNCHED(c); LAUNCHED(c);
        launch_rep_keys(da, skA, st, c->cell_r.as<u32>(), c->cell_new.as<u32>(), nvirt); LAUNCHED(c);
    } else {
        launch_rep_keys(da, skA, st); LAUNCHED(c);
    }
    cudaEventRecord(c->ev_x[0], st);
    // ---- extend every rep (in slot order: neighbouring reps share genome sectors)
    if (!rows) {
        DedupArgs dx = da;
        dx.bitmap = nullptr; // extents only; the slot ranges follow in (colour, slot) order below
        launch_extend(dx, c->gt, c->sd, st);
        if (n_rep) {
            c->stats.kernel_launches += extend_launches();
            TRY(mbi_extend_long(c, dx));
            if (c->shadow_on) da.shadow = c->shadow.as<u8>();
        }
    }
    // ---- reps in (group colour, slot) order, per-rep records, slot ranges of the extents
    TRY(mbi_sort_records(c, &skA, &skB, nullptr, nullptr, n_rep, 32, 16, false));
    da.s_key = skA;
    launch_rep_setup(da, c->gt, st);
    if (n_rep) LAUNCHED(c);
    if (rows) { launch_extent_ranges(da, c->gt, st); if (n_rep) LAUNCHED(c); }
    CHECK_LAUNCH(c);
    cudaEventRecord(c->ev_x[3], st);
    const bool want_trace = getenv("MB_DEDUP_TRACE") != nullptr;
    if (want_trace) {
        TRY(c->reserve(c->trace, 8200 * 8));
        CUDA_TRY(c, cudaMemsetAsync(c->trace.p, 0, 8200 * 8, st));
        da.trace = c->trace.as<u64>();
    }
    {
        cudaError_t e = launch_resolve(da, st);
        if (n_rep) LAUNCHED(c);
        if (e != cudaSuccess) { c->set_cuda_error(e, "launch_resolve", __LINE__); return MB_E_CUDA; }
    }
    CHECK_LAUNCH(c);
    if (want_trace) {
        std::vector<u64> tr(8200);
        CUDA_TRY(c, cudaStreamSynchronize(st));
        CUDA_TRY(c, cudaMemcpy(tr.data(), c->trace.p, 8200 * 8, cudaMemcpyDeviceToHost));
        static const char* names[] = {"", "enter", "", "extend", "long", "claim", "decide"};
        u64 prev = 0;
        for (u64 k = 0; k < tr[0] && k < 4000; ++k) {
            u64 tag = tr[2 + 2 * k], t = tr[3 + 2 * k];
            if (k) fprintf(stderr, "[dedup-trace] %-12s %8.1f us\n", names[tag], (double)(t - prev) / 1e3);
            prev = t;
        }
        float t_chain = 0, t_ext = 0;
        cudaEventElapsedTime(&t_chain, c->ev[EV_BUCKET], c->ev_x[0]);
        cudaEventElapsedTime(&t_ext, c->ev_x[0], c->ev_x[3]);
        fprintf(stderr, "[dedup-trace] emit+chains+sort %.1f us, extend %.1f us\n", t_chain * 1e3, t_ext * 1e3);
    }
    return MB_OK;
}

// ------------------------------------------------------------------ persistent MemHash table (seed family)
static FamilyArgs family_args(mb_ctx* c) {
    FamilyArgs f{};
    f.t_h1 = c->fam_th1.as<u64>(); f.t_h2 = c->fam_th2.as<u64>(); f.t_x = c->fam_tx.as<u32>(); f.t_end = c->fam_tend.as<u32>();
    f.s_h1 = c->fam_sh1.as<u64>(); f.s_h2 = c->fam_sh2.as<u64>(); f.s_x = c->fam_sx.as<u32>(); f.s_end = c->fam_send.as<u32>();
    f.s_pmax = c->fam_spmax.as<u32>(); f.s_run0 = c->fam_srun0.as<u32>();
    return f;
}

// Before the de-dup of a search: candidates whose seed lies inside a match of an earlier search (same group) are marked
// dropped (c->fam_drop) and taken out of the chains.  The table is sorted here when it has grown since its last use.
int mbi_family_filter(mb_ctx* c, u32 n_cand) {
    const u32 nt = c->fam_n;
    if (nt == 0 || n_cand == 0) return MB_OK;
    cudaStream_t st = c->stream;
    u64* scal = c->scalars.as<u64>();
    if (c->fam_dirty) {
        const size_t n8 = (size_t)nt + 8;
        TRY(c->reserve(c->fam_sh1, n8 * 8)); TRY(c->reserve(c->fam_sh2, n8 * 8)); TRY(c->reserve(c->fam_sx, n8 * 4));
        TRY(c->reserve(c->fam_send, n8 * 4)); TRY(c->reserve(c->fam_spmax, n8 * 4)); TRY(c->reserve(c->fam_srun0, n8 * 4));
        TRY(c->reserve(c->sort_kA, n8 * 8)); TRY(c->reserve(c->sort_kB, n8 * 8)); TRY(c->reserve(c->sort_vA, n8 * 8)); TRY(c->reserve(c->sort_vB, n8 * 8));
        TRY(c->reserve(c->lookback, (size_t)(div_up(nt, radix_tile_size()) + 1) * 256 * 8)); // the table may be larger than this search's seed list
        FamilyArgs f = family_args(c);
        u64 *kA = c->sort_kA.as<u64>(), *kB = c->sort_kB.as<u64>(), *vA = c->sort_vA.as<u64>(), *vB = c->sort_vB.as<u64>();
        launch_family_key_x(f, nt, kA, vA, st); LAUNCHED(c);
        TRY(mbi_sort_records(c, &kA, &kB, &vA, &vB, nt, 0, 32, false));
        launch_family_key_h(f, nt, vA, kA, st); LAUNCHED(c);
        TRY(mbi_sort_records(c, &kA, &kB, &vA, &vB, nt, 0, 64, false));
        launch_family_gather(f, nt, vA, st); LAUNCHED(c); LAUNCHED(c);
        CHECK_LAUNCH(c);
        c->fam_dirty = false;
    }
    TRY(c->reserve(c->fam_drop, (size_t)n_cand + 8));
    launch_family_filter(family_args(c), nt, n_cand, c->ghash.as<u64>(), c->ghash2.as<u64>(), c->cand_off.as<u32>(), c->comp_pos.as<u32>(), (u32)c->sd.L,
                         c->fam_drop.as<u8>(), reinterpret_cast<u32*>(scal + SC_FAM) + 1, st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    c->fam_pre = true;
    return MB_OK;
}

// After a search: its accepted matches join the table (c->r_matches of them: the output stage has counted them).
int mbi_family_append(mb_ctx* c) {
    const u32 nm = (u32)c->r_matches;
    if (nm == 0 || c->n_rep == 0) return MB_OK;
    cudaStream_t st = c->stream;
    u64* scal = c->scalars.as<u64>();
    const size_t total = (size_t)c->fam_n + nm + 8;
    TRY(c->reserve_keep(c->fam_th1, total * 8, (size_t)c->fam_n * 8)); TRY(c->reserve_keep(c->fam_th2, total * 8, (size_t)c->fam_n * 8));
    TRY(c->reserve_keep(c->fam_tx, total * 4, (size_t)c->fam_n * 4)); TRY(c->reserve_keep(c->fam_tend, total * 4, (size_t)c->fam_n * 4));
    c->fam_tmp = c->fam_n;
    CUDA_TRY(c, cudaMemcpyAsync(scal + SC_FAM, &c->fam_tmp, 4, cudaMemcpyHostToDevice, st));
    // (a dropped candidate's first hash is poisoned, but a dropped candidate is never accepted)
    launch_family_append(family_args(c), c->s_cand.as<u8>(), c->rep_cand.as<u32>(), c->n_rep, c->ghash.as<u64>(), c->ghash2.as<u64>(), c->cand_off.as<u32>(),
                         c->comp_pos.as<u32>(), c->ext_l.as<u32>(), c->ext_r.as<u32>(), (u32)c->sd.L, reinterpret_cast<u32*>(scal + SC_FAM), st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    c->fam_n += nm;
    c->fam_dirty = true;
    return MB_OK;
}

// The reps whose extension the bounded rounds did not finish.  Few: one warp per rep.  Many (long window-consistent runs
// full of reps of one group): sorted into classes that share their walks (kernels_dedup.cu, k_extend_long_classes).
int mbi_extend_long(mb_ctx* c, const DedupArgs& da_in) {
    cudaStream_t st = c->stream;
    DedupArgs da = da_in;
    u32 n_long = 0;
    c->shadow_on = false;
    CUDA_TRY(c, cudaMemcpyAsync(&c->tmp_u64, da.ctr + 6, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    n_long = (u32)c->tmp_u64;
    if (n_long == 0) return MB_OK;
    // short lists: one warp per rep, no sort.  MB_LONG_CLASSES_MIN (test knob) moves the threshold.
    const char* knob = getenv("MB_LONG_CLASSES_MIN");
    const u32 classes_min = knob ? (u32)atoi(knob) : 256u;
    if (n_long < classes_min) {
        launch_extend_long(da, c->gt, c->sd, st); LAUNCHED(c); CHECK_LAUNCH(c);
        return MB_OK;
    }
    TRY(c->reserve(c->sort_vA, (size_t)(n_long + 8) * 8)); TRY(c->reserve(c->sort_vB, (size_t)(n_long + 8) * 8));
    TRY(c->reserve(c->lookback, (size_t)(div_up(n_long, radix_tile_size()) + 1) * 256 * 8));
    u64 *kA = c->sort_vA.as<u64>(), *kB = c->sort_vB.as<u64>();
    TRY(c->reserve(c->shadow, (size_t)da.n_cand + 8));
    CUDA_TRY(c, cudaMemsetAsync(c->shadow.p, 0, (size_t)da.n_cand + 8, st));
    da.shadow = c->shadow.as<u8>();
    c->shadow_on = true;
    launch_long_keys(da, (u32)c->sd.L, n_long, kA, st); LAUNCHED(c);
    // by (class, rep index): inside a class the members of one run must stay together (one walk per run), and the work list is in push order
    TRY(mbi_sort_records(c, &kA, &kB, nullptr, nullptr, n_long, 0, 64, false));
    launch_extend_long_classes(da, c->gt, c->sd, kA, n_long, st); LAUNCHED(c); CHECK_LAUNCH(c);
    return MB_OK;
}

// a12: compact the accepted items (reps; state in s_cand, candidate ids in rep_cand), canonical order (D18), CSR.
// Sets r_matches / r_comps.
int mbi_output_unique(mb_ctx* c, u32 n_cand, u64 maxlen) {
    cudaStream_t st = c->stream;
    u64* scal = c->scalars.as<u64>();
    const u32* hs32 = reinterpret_cast<const u32*>(c->h_scal);
    const u64* hs64 = reinterpret_cast<const u64*>(c->h_scal);
    const u32 L = c->sd.L;
    TRY(c->reserve(c->flags, (size_t)(n_cand + 8) * 4));
    TRY(c->reserve(c->match_idx, (size_t)(n_cand + 8) * 4));
    OutputArgs oa{};
    oa.n_items = n_cand; oa.state = c->s_cand.as<u8>(); oa.item_cand = c->rep_cand.as<u32>(); oa.cand_off = c->cand_off.as<u32>();
    oa.comp_pos = c->comp_pos.as<u32>();
    oa.comp_gs = c->comp_gs.as<u8>();
    oa.ext_l = c->ext_l.as<u32>(); oa.ext_r = c->ext_r.as<u32>(); oa.flags = c->flags.as<u32>(); oa.match_idx = c->match_idx.as<u32>();
    oa.n_matches_ptr = scal + SC_NMATCH;
    oa.repeat = c->last_mode == MB_MODE_REPEAT;
    launch_uniq_flags(oa, st); LAUNCHED(c);
    launch_scan_u32(oa.flags, n_cand, c->match_idx.as<u32>(), nullptr, c->status_slice(div_up(n_cand, scan_tile())), c->ticket(),
                    scal + SC_NMATCH, st);
    LAUNCHED(c); CHECK_LAUNCH(c);
    TRY(mbi_read_scalars(c));
    const u32 n_matches = (u32)hs64[SC_NMATCH];
    u64 n_ocomp = 0;
    c->stats.dedup_batches = 1; c->stats.dedup_iters = hs32[2 * SC_DDCTR + 9];
    if (getenv("MB_DEDUP_TRACE"))
        fprintf(stderr, "[dedup] reps %u rounds %u wide %u long %u visited %u claims %u covers %u\n", (u32)c->stats.n_extended, hs32[2 * SC_DDCTR + 9],
                hs32[2 * SC_DDCTR + 10], hs32[2 * SC_DDCTR + 6], hs32[2 * SC_DDCTR + 12], hs32[2 * SC_DDCTR + 13], hs32[2 * SC_DDCTR + 14]);
    TRY(c->reserve(c->sort_kA, (size_t)(n_matches + 8) * 8)); TRY(c->reserve(c->sort_kB, (size_t)(n_matches + 8) * 8));
    TRY(c->reserve(c->sort_vA, (size_t)(n_matches + 8) * 8)); TRY(c->reserve(c->sort_vB, (size_t)(n_matches + 8) * 8));
    TRY(c->reserve(c->ncomp, (size_t)(n_matches + 8) * 4));
    TRY(c->reserve(c->out_len, (size_t)(n_matches + 8) * 4));
    TRY(c->reserve(c->out_off, (size_t)(n_matches + 8) * 8));
    if (n_matches) {
        u64 *skA = c->sort_kA.as<u64>(), *skB = c->sort_kB.as<u64>(), *svA = c->sort_vA.as<u64>(), *svB = c->sort_vB.as<u64>();
        oa.sort_key = skA; oa.sort_val = svA; oa.ncomp = c->ncomp.as<u32>();
        int sbits = mbi_bits_for(maxlen);
        launch_uniq_keys(oa, sbits, st); LAUNCHED(c); CHECK_LAUNCH(c);
        TRY(mbi_sort_records(c, &skA, &skB, &svA, &svB, n_matches, 0, sbits + 6, false));
        launch_uniq_tiefix(oa, skA, svA, L, n_matches, st); LAUNCHED(c);
        launch_uniq_ncomp(oa, svA, n_matches, st); LAUNCHED(c);
        launch_scan_u32(oa.ncomp, n_matches, nullptr, c->out_off.as<u64>(), c->status_slice(div_up(n_matches, scan_tile())), c->ticket(),
                        scal + SC_NCOMP, st);
        LAUNCHED(c); CHECK_LAUNCH(c);
        TRY(mbi_read_scalars(c));
        n_ocomp = hs64[SC_NCOMP];
        TRY(c->reserve(c->out_seq, (size_t)(n_ocomp + 8)));
        TRY(c->reserve(c->out_start, (size_t)(n_ocomp + 8) * 4));
        oa.out_off = c->out_off.as<u64>(); oa.out_len = c->out_len.as<u32>(); oa.out_seq = c->out_seq.as<u8>();
        oa.out_start = c->out_start.as<int32_t>();
        launch_uniq_gather(oa, svA, L, n_matches, st); LAUNCHED(c); CHECK_LAUNCH(c);
    }
    c->r_matches = n_matches; c->r_comps = n_ocomp;
    return MB_OK;
}

extern "C" {

// Device -> pinned host copy of the last result.  compact: the device arrays as they are (u8 sequence, i32 start);
// otherwise widened on the device to the u32 / i64 arrays of mb_result first.
static int fetch_common(mb_ctx* c, bool compact) {
    if (!c->have_result) return MB_E_STATE;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    u64 nm = c->r_matches, nc = c->r_comps;
    TRY(c->reserve_host(c->h_len, c->h_len_cap, (nm + 1) * 4));
    TRY(c->reserve_host(c->h_off, c->h_off_cap, (nm + 1) * 8));
    TRY(c->reserve_host(c->h_seq, c->h_seq_cap, (nc + 1) * (compact ? 1 : 4)));
    TRY(c->reserve_host(c->h_start, c->h_start_cap, (nc + 1) * (compact ? 4 : 8)));
    cudaEventRecord(c->ev_x[1], st);
    if (nm) {
        CUDA_TRY(c, cudaMemcpyAsync(c->h_len, c->out_len.p, nm * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaMemcpyAsync(c->h_off, c->out_off.p, (nm + 1) * 8, cudaMemcpyDeviceToHost, st));
        if (compact) {
            CUDA_TRY(c, cudaMemcpyAsync(c->h_seq, c->out_seq.p, nc, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(c, cudaMemcpyAsync(c->h_start, c->out_start.p, nc * 4, cudaMemcpyDeviceToHost, st));
        } else {
            TRY(c->reserve(c->wide_seq, (nc + 8) * 4)); TRY(c->reserve(c->wide_start, (nc + 8) * 8));
            launch_expand_result(c->out_seq.as<u8>(), c->out_start.as<int32_t>(), nc, c->wide_seq.as<u32>(), c->wide_start.as<i64>(), st);
            CHECK_LAUNCH(c);
            CUDA_TRY(c, cudaMemcpyAsync(c->h_seq, c->wide_seq.p, nc * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(c, cudaMemcpyAsync(c->h_start, c->wide_start.p, nc * 8, cudaMemcpyDeviceToHost, st));
        }
    } else {
        ((u64*)c->h_off)[0] = 0;
    }
    cudaEventRecord(c->ev_x[2], st);
    CUDA_TRY(c, cudaStreamSynchronize(st));
    c->stats.d2h_bytes = nm ? nm * 4 + (nm + 1) * 8 + nc * (compact ? 5 : 12) : 0;
    // stage timings
    auto ms = [&](cudaEvent_t a, cudaEvent_t b) { float t = 0; if (cudaEventElapsedTime(&t, a, b) != cudaSuccess) { cudaGetLastError(); t = 0; } return t; };
    c->stats.ms_extract = ms(c->ev[EV_START], c->ev[EV_EXTRACT]);
    c->stats.ms_sort = ms(c->ev[EV_EXTRACT], c->ev[EV_SORT]);
    c->stats.ms_bucket = ms(c->ev[EV_SORT], c->ev[EV_BUCKET]);
    c->stats.ms_dedup = ms(c->ev[EV_BUCKET], c->ev[EV_DEDUP]);
    c->stats.ms_output = ms(c->ev[EV_DEDUP], c->ev[EV_OUTPUT]);
    c->stats.ms_total_device = ms(c->ev[EV_START], c->ev[EV_OUTPUT]);
    c->stats.ms_d2h = ms(c->ev_x[1], c->ev_x[2]);
    c->stats.ms_radix_kernels = 0;
    for (int i = 0; i < c->n_timed_passes; ++i) c->stats.ms_radix_kernels += ms(c->ev_r[2 * i], c->ev_r[2 * i + 1]);
    c->stats.radix_launches = c->n_timed_passes;
    return MB_OK;
}

int mb_fetch_result(mb_ctx* c, const mb_result** out) {
    if (!c || !out) return MB_E_ARG;
    TRY(fetch_common(c, false));
    mb_result& r = c->res;
    r.n_matches = c->r_matches; r.n_comps = c->r_comps;
    r.length = (const u32*)c->h_len; r.comp_off = (const u64*)c->h_off; r.comp_seq = (const u32*)c->h_seq; r.comp_start = (const i64*)c->h_start;
    r.unique_mers = c->r_unique;
    r.unique_mers_per_seq = (const u64*)c->h_perseq;
    r.nseq = (u32)c->seq_len.size();
    *out = &c->res;
    return MB_OK;
}

int mb_fetch_result_compact(mb_ctx* c, const mb_result_compact** out) {
    if (!c || !out) return MB_E_ARG;
    TRY(fetch_common(c, true));
    mb_result_compact& r = c->cres;
    r.n_matches = c->r_matches; r.n_comps = c->r_comps;
    r.length = (const u32*)c->h_len; r.comp_off = (const u64*)c->h_off; r.comp_seq8 = (const u8*)c->h_seq; r.comp_start32 = (const int32_t*)c->h_start;
    r.unique_mers = c->r_unique;
    r.unique_mers_per_seq = (const u64*)c->h_perseq;
    r.nseq = (u32)c->seq_len.size();
    *out = &c->cres;
    return MB_OK;
}

int mb_find_compact(mb_ctx* c, const mb_params* prm, const mb_result_compact** out) {
    TRY(mb_find_device(c, prm));
    return mb_fetch_result_compact(c, out);
}

/* repeatoire's match position lookup table (src/repeatoire.cpp:1944-1966), built on the device from the last single-sequence
 * result (MB_MODE_SEED_ENUM / MB_MODE_REPEAT): entry p (1-based left end; n_pos = sequence length + 1 entries) = index of the
 * match and of its component that start at p, 0xFFFFFFFF where none does. */
int mb_position_table(mb_ctx* c, const uint32_t** match_of_pos, const uint32_t** comp_of_pos, uint64_t* n_pos) {
    if (!c || !match_of_pos || !comp_of_pos || !n_pos) return MB_E_ARG;
    if (!c->have_result || c->seq_len.size() != 1 || (c->last_mode != MB_MODE_SEED_ENUM && c->last_mode != MB_MODE_REPEAT)) return MB_E_STATE;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const u64 np = c->seq_len[0] + 1;
    TRY(c->reserve(c->pos_match, np * 4)); TRY(c->reserve(c->pos_comp, np * 4));
    TRY(c->reserve_host(c->h_posm, c->h_posm_cap, np * 4)); TRY(c->reserve_host(c->h_posc, c->h_posc_cap, np * 4));
    TRY(c->reserve(c->sort_kA, np * 8)); // scratch: (match + 1) << 32 | component per position
    CUDA_TRY(c, cudaMemsetAsync(c->sort_kA.p, 0, np * 8, st));
    launch_position_table(c->out_off.as<u64>(), c->out_start.as<int32_t>(), (u32)c->r_matches, c->sort_kA.as<u64>(), c->pos_match.as<u32>(), c->pos_comp.as<u32>(), np, st);
    CHECK_LAUNCH(c);
    CUDA_TRY(c, cudaMemcpyAsync(c->h_posm, c->pos_match.p, np * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_posc, c->pos_comp.p, np * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    *match_of_pos = (const uint32_t*)c->h_posm; *comp_of_pos = (const uint32_t*)c->h_posc; *n_pos = np;
    return MB_OK;
}

int mb_find(mb_ctx* c, const mb_params* prm, const mb_result** out) {
    TRY(mb_find_device(c, prm));
    return mb_fetch_result(c, out);
}

int mb_get_stats(mb_ctx* c, mb_stats* out) {
    if (!c || !out) return MB_E_ARG;
    if (c->n_timed_passes) { // the seed sort's per-pass events (complete once the stream has been synchronised)
        float tot = 0;
        bool ok = true;
        for (int i = 0; i < c->n_timed_passes; ++i) {
            float t = 0;
            if (cudaEventElapsedTime(&t, c->ev_r[2 * i], c->ev_r[2 * i + 1]) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
            tot += t;
        }
        if (ok) { c->stats.ms_radix_kernels = tot; c->stats.radix_launches = c->n_timed_passes; }
    }
    *out = c->stats;
    return MB_OK;
}

int mb_get_mers(mb_ctx* c, int seq, uint64_t* out_mers, uint64_t capacity, uint64_t* out_n) {
    if (!c || seq < 0 || (size_t)seq >= c->seq_len.size() || !out_n) return MB_E_ARG;
    if (!c->seed_set) return MB_E_SEED;
    CUDA_TRY(c, cudaSetDevice(c->device));
    u64 len = c->seq_len[seq];
    u64 n = len >= (u64)c->sd.L ? len - c->sd.L + 1 : 0;
    *out_n = n;
    if (n == 0) return MB_OK;
    if (!out_mers || capacity < n) return MB_E_ARG;
    GenomeTable gt{};
    gt.nseq = (u32)c->seq_len.size();
    for (u32 g = 0; g < gt.nseq; ++g) { gt.word_base[g] = c->seq_word_base[g]; gt.len[g] = (u32)c->seq_len[g]; }
    RecFmt fmt{}; fmt.kbits = 2 * c->sd.w;
    TRY(c->reserve(c->mers_tmp, n * 8));
    launch_mers(c->packed.as<u64>(), gt, c->sd, fmt, seq, c->mers_tmp.as<u64>(), c->stream);
    CHECK_LAUNCH(c);
    CUDA_TRY(c, cudaMemcpyAsync(out_mers, c->mers_tmp.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MB_OK;
}

int mb_get_sml(mb_ctx* c, int seq, uint32_t* out_pos, uint64_t capacity, uint64_t* out_n) {
    if (!c || seq < 0 || (size_t)seq >= c->seq_len.size() || !out_n) return MB_E_ARG;
    if (!c->have_result || !c->sorted_keys) return MB_E_STATE;
    CUDA_TRY(c, cudaSetDevice(c->device));
    u64 len = c->seq_len[seq];
    u64 n = len >= (u64)c->sd.L ? len - c->sd.L + 1 : 0;
    *out_n = n;
    if (n == 0) return MB_OK;
    if (!out_pos || capacity < n) return MB_E_ARG;
    // debugging / .sslist export path: filter the globally sorted records on the host
    u32 total = c->n_seeds;
    std::vector<u64> rec(total);
    const u64* src = c->fmt.wide ? c->sorted_vals : c->sorted_keys;
    CUDA_TRY(c, cudaMemcpyAsync(rec.data(), src, (size_t)total * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    u64 k = 0;
    for (u32 i = 0; i < total; ++i)
        if (rec_genome(c->fmt, rec[i]) == (u32)seq) out_pos[k++] = rec_pos(c->fmt, rec[i]);
    return k == n ? MB_OK : MB_E_STATE;
}

// Tuning aid (tools/bench_radix.py): sorts n pseudo-random records on key bits [shift, shift+kbits) `reps` times and
// returns the mean device time of one radix pass in ms_out[0] (CUDA events around each pass) and of one whole sort
// in ms_out[1].  Records look like the path's: min of two uniform keys in the key field, payload bits below.
int mb_debug_radix(mb_ctx* c, uint64_t n64, int shift, int kbits, int reps, float* ms_out) {
    if (!c || !ms_out || n64 == 0 || n64 >= (1ull << 31) || kbits < 1 || shift + kbits > 64) return MB_E_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const u32 n = (u32)n64;
    TRY(c->reserve(c->keysA, ((size_t)n + 8) * 8));
    TRY(c->reserve(c->keysB, ((size_t)n + 8) * 8));
    TRY(c->reserve(c->sort_kA, ((size_t)n + 8) * 8));
    TRY(c->reserve(c->hist, 8 * 256 * 4));
    TRY(c->reserve(c->digit_base, 8 * 256 * 4));
    TRY(c->reserve(c->lookback, (size_t)(div_up(n, radix_tile_size()) + 1) * 256 * 8));
    TRY(c->reserve(c->tickets, 256 * 4));
    std::vector<u64> h(n);
    u64 x = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    const u64 kmask = kbits >= 64 ? ~0ull : ((1ull << kbits) - 1);
    for (u32 i = 0; i < n; ++i) {
        u64 a = rnd() & kmask, b = rnd() & kmask;
        h[i] = (std::min(a, b) << shift) | (shift ? (rnd() & ((1ull << shift) - 1)) : 0);
    }
    CUDA_TRY(c, cudaMemcpy(c->sort_kA.p, h.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    float tot_pass = 0, tot_sort = 0;
    int npass_total = 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < reps + 1; ++r) {
        CUDA_TRY(c, cudaMemcpyAsync(c->keysA.p, c->sort_kA.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->tickets.p, 0, 256 * 4, c->stream));
        c->ticket_next = 0; c->n_timed_passes = 0;
        u64 *kA = c->keysA.as<u64>(), *kB = c->keysB.as<u64>();
        cudaEventRecord(e0, c->stream);
        TRY(mbi_sort_records(c, &kA, &kB, nullptr, nullptr, n, shift, kbits, false, true));
        cudaEventRecord(e1, c->stream);
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (r == 0) continue; // warm-up
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        tot_sort += t;
        for (int i = 0; i < c->n_timed_passes; ++i) { cudaEventElapsedTime(&t, c->ev_r[2 * i], c->ev_r[2 * i + 1]); tot_pass += t; ++npass_total; }
        if (r == reps) { // check order
            std::vector<u64> out(n);
            CUDA_TRY(c, cudaMemcpy(out.data(), kA, (size_t)n * 8, cudaMemcpyDeviceToHost));
            for (u32 i = 1; i < n; ++i)
                if (((out[i - 1] >> shift) & kmask) > ((out[i] >> shift) & kmask)) {
                    cudaEventDestroy(e0); cudaEventDestroy(e1);
                    ms_out[0] = npass_total ? tot_pass / npass_total : 0; ms_out[1] = reps ? tot_sort / reps : 0;
                    return MB_E_STATE;
                }
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ms_out[0] = npass_total ? tot_pass / npass_total : 0;
    ms_out[1] = reps ? tot_sort / reps : 0;
    return MB_OK;
}

} // extern "C"

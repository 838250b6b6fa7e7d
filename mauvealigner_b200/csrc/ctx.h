// ctx.h — the library context (device buffers, last-run state) shared by api.cu and api_dist.cu.
#pragma once
#include "../../include/mauve_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

enum { EV_START, EV_EXTRACT, EV_SORT, EV_BUCKET, EV_DEDUP, EV_OUTPUT, EV_COUNT };
enum { SC_RUNS = 0 /*u32[2]*/, SC_CAND = 1 /*u32[2]*/, SC_NBUCKETS = 2, SC_UNDECIDED = 3, SC_EXTENDED = 4, SC_NMATCH = 5, SC_NCOMP = 6,
       SC_BMTOTAL = 7, SC_DDCTR = 8 /* 16 x u32 */, SC_FAM = 16 /* u32[2]: table entries, candidates dropped by the table */, SC_COUNT = 18 };

struct mb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    char err[512] = {0};

    // inputs
    std::vector<u64> seq_len;
    std::vector<u64> seq_word_base;
    u64 words_used = 0;
    DBuf packed, ascii_stage;
    u64 pattern = 0;
    SeedDev sd{};
    bool seed_set = false;

    // workspace
    DBuf keysA, keysB, valsA, valsB, hist, digit_base, lookback, tickets, status, scalars, per_seq, tile_first;
    DBuf cand_run, cand_off, cand_aux, comp_pos, comp_gs, bitmap, bmrank, cand_at, cstate, covered, minrank, ext_l, ext_r;
    DBuf trace, ghash2, rep_cand, s_h2, reach, xstate, xrec, shadow, cell_r, cell_new;
    bool shadow_on = false;                 // this search's long extensions left shadow flags (mbi_extend_long)
    u32 n_rep = 0;
    DBuf x_lut, x_counts, x_hdr_s, x_comp_s, x_hdr_r, x_comp_r, x_m, x_key, x_item, x_peers, x_recv; // multi-GPU exchange buffers (api_dist.cu)
    // source-side candidate arrays of the distributed path: they must outlive the owner stage that runs in between
    DBuf q_off, q_pos, q_gs, q_el, q_er, q_perm, q_state, q_item, x_acc_s, x_acc_r;
    const u64* d_sperm = nullptr;           // partition order of the rows being packed (between the count and the pack call)
    std::vector<u64> d_bound, d_cbound;     // rows / component words before every destination's block
    int d_rank = 0, d_world = 1;
    u32 d_nslice = 0;
    bool d_use_p2p = false;
    u32 d_ncand = 0, d_nccomp = 0, d_nmatch = 0, d_nmcomp = 0;
    u64 d_bases = 0, d_maxlen = 0;
    cudaEvent_t ev_d[8] = {nullptr};
    DBuf wl_a, wl_b, wl_c, wl_long, wd_a, wd_b, wd_c, live_bits, ghash, slot_gp, slot_hash, link_bits, chain_min, rep_bits, rep_rank, s_hash, s_cand, rng_lo, rng_hi;
    DBuf flags, match_idx, sort_kA, sort_kB, sort_vA, sort_vB, ncomp, mers_tmp;
    DBuf out_len, out_off, out_seq, out_start;   // device result: out_seq u8[], out_start i32[] (compact)
    DBuf wide_seq, wide_start;                   // widened copies for mb_fetch_result
    mb_result_compact cres{};
    // segmented search (mb_set_segments / mb_find_batch): host copy of the bounds [nseq][n_seg + 1], device copy
    std::vector<u32> h_seg;
    u32 n_seg = 0;
    DBuf seg;
    // result of mb_find_batch (host, owned by the context)
    std::vector<u64> b_moff, b_coff;
    std::vector<u32> b_len, b_seq;
    std::vector<i64> b_start;
    mb_batch_result bres{};
    std::vector<u8> b_concat;
    DBuf pos_match, pos_comp;                 // mb_position_table
    void* h_posm = nullptr; void* h_posc = nullptr;
    size_t h_posm_cap = 0, h_posc_cap = 0;
    // persistent MemHash table across searches (mb_accumulate; kernels_family.cu)
    bool fam_on = false, fam_dirty = false, fam_pre = false;
    u32 fam_n = 0;
    u64 fam_tmp = 0;
    DBuf fam_th1, fam_th2, fam_tx, fam_tend, fam_sh1, fam_sh2, fam_sx, fam_send, fam_spmax, fam_srun0, fam_drop;
    u32 ticket_next = 0;
    u64 tmp_u64 = 0;
    size_t status_next = 0;

    // last run
    GenomeTable gt{};
    RecFmt fmt{};
    const u64* sorted_keys = nullptr;
    const u64* sorted_vals = nullptr;
    u32 n_seeds = 0;
    bool have_result = false;
    u64 r_matches = 0, r_comps = 0, r_unique = 0;
    int last_mode = 0;

    // host result (pinned)
    void* h_len = nullptr; void* h_off = nullptr; void* h_seq = nullptr; void* h_start = nullptr; void* h_perseq = nullptr; void* h_scal = nullptr;
    size_t h_len_cap = 0, h_off_cap = 0, h_seq_cap = 0, h_start_cap = 0;
    mb_result res{};
    mb_stats stats{};
    cudaEvent_t ev[EV_COUNT] = {nullptr};
    cudaEvent_t ev_x[4] = {nullptr};
    cudaEvent_t ev_r[16] = {nullptr};
    int n_timed_passes = 0;

    void set_cuda_error(cudaError_t e, const char* what, int line) {
        snprintf(err, sizeof(err), "%s (%s) at api.cu:%d: %s", cudaGetErrorName(e), cudaGetErrorString(e), line, what);
    }
    int reserve(DBuf& b, size_t bytes) {
        if (bytes <= b.cap) return MB_OK;
        if (b.p) cudaFree(b.p);
        b.p = nullptr; b.cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) { set_cuda_error(e, "cudaMalloc", __LINE__); return e == cudaErrorMemoryAllocation ? MB_E_NOMEM : MB_E_CUDA; }
        b.cap = want;
        return MB_OK;
    }
    // grow a buffer whose first `keep` bytes must survive
    int reserve_keep(DBuf& b, size_t bytes, size_t keep) {
        if (bytes <= b.cap) return MB_OK;
        DBuf nb;
        int rc = reserve(nb, bytes + bytes / 2);
        if (rc != MB_OK) return rc;
        if (b.p && keep) {
            cudaError_t e = cudaMemcpyAsync(nb.p, b.p, keep, cudaMemcpyDeviceToDevice, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) { set_cuda_error(e, "reserve_keep", __LINE__); cudaFree(nb.p); return MB_E_CUDA; }
        }
        if (b.p) cudaFree(b.p);
        b = nb;
        return MB_OK;
    }
    int reserve_host(void*& p, size_t& cap, size_t bytes) {
        if (bytes <= cap) return MB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { set_cuda_error(e, "cudaMallocHost", __LINE__); return MB_E_NOMEM; }
        cap = want;
        return MB_OK;
    }
    u32* ticket() { return tickets.as<u32>() + (ticket_next++); }
    u64* status_slice(size_t n_tiles) {
        u64* p = status.as<u64>() + status_next;
        status_next += n_tiles + 1;
        return p;
    }
};

#define TRY(expr) do { int _rc = (expr); if (_rc != MB_OK) return _rc; } while (0)
#define LAUNCHED(ctx) do { ++(ctx)->stats.kernel_launches; } while (0)
#define CHECK_LAUNCH(ctx) CUDA_TRY(ctx, cudaGetLastError())


// helpers defined in api.cu
struct MbiRun { std::vector<u32> tile_first; u32 n = 0, n_tiles = 0; u64 bases = 0, maxlen = 0; int mode = 0; };
int mbi_setup_run(mb_ctx* c, MbiRun& r);
int mbi_sort_records(mb_ctx* c, u64** kA, u64** kB, u64** vA, u64** vB, u32 n, int shift, int kbits, bool hist_ready, bool time_passes = false);
int mbi_read_scalars(mb_ctx* c);
int mbi_bits_for(u64 maxval);
int mbi_count_or_enum(mb_ctx* c, const mb_params* prm, u64* kA, u64* kB, u64* vA, u64* vB, u32 n);
// stages of the MODE_UNIQUE tail, shared by the single-GPU and the distributed drivers
int mbi_reserve_candidates(mb_ctx* c, u32 n_cand, u32 n_ccomp, u64 bases);
int mbi_dedup(mb_ctx* c, u32 n_cand, u64 bases, const u64* rows = nullptr);
int mbi_output_unique(mb_ctx* c, u32 n_cand, u64 maxlen);
// finishes the extensions launch_extend left on the long list (reads its length back; uses sort_vA / sort_vB as scratch)
int mbi_extend_long(mb_ctx* c, const DedupArgs& da);
// persistent MemHash table (mb_accumulate): drop the candidates contained in matches of earlier searches / add this search's matches
int mbi_family_filter(mb_ctx* c, u32 n_cand);
int mbi_family_append(mb_ctx* c);

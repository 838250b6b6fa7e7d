/*
 * mauve_b200.h — C ABI of the B200-native seed-match anchoring path.
 *
 * This is the drop-in boundary.  In the reference the path sits
 * behind a C++ virtual-method protocol, not an FFI; each entry point below names
 * the reference call it stands in for.  The host classes in include/mems_compat/
 * (MatchList, MatchFinder, MemHash, UniqueMatchFinder, SeedMatchEnumerator, ...)
 * are thin C++ over exactly these functions; INTEGRATION.md shows the binding.
 *
 * Conventions: plain C types only; 0 = MB_OK, negative = error (mb_strerror);
 * no exceptions cross the boundary; a context is driven by one host thread;
 * there is NO CPU fallback and no backend dispatch behind this ABI — every
 * compute entry point fails with MB_E_CUDA when no sm_100 device is usable.
 */
#ifndef MAUVE_B200_H
#define MAUVE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mb_ctx mb_ctx;

enum {
    MB_OK = 0,
    MB_E_ARG = -1,       /* null / out-of-range argument                                   */
    MB_E_SEED = -2,      /* pattern not palindromic / weight not odd in [3,31]              */
    MB_E_NOSEQ = -3,     /* no sequence added                                              */
    MB_E_SEQCOUNT = -4,  /* SEED_ENUM needs exactly 1 sequence (SeedMatchEnumerator.h:59-65);
                            UNIQUE / PAIRWISE take at most 64                              */
    MB_E_TOOLONG = -5,   /* a sequence >= 2^32 bases, or >= 2^31 seed positions in total    */
    MB_E_CUDA = -6,      /* CUDA runtime error or no usable device (see mb_last_cuda_error) */
    MB_E_NOMEM = -7,
    MB_E_STATE = -8      /* call order (e.g. fetch before find)                            */
};

/* The per-bucket policy.  The reference selects it by which MatchFinder subclass
 * overrides EnumerateMatches/HashMatch; a virtual call per bucket cannot cross to
 * the device, so it is an enum here. */
enum {
    MB_MODE_UNIQUE = 0,       /* UniqueMatchFinder::EnumerateMatches (src/UniqueMatchFinder.cpp:36-60)
                                 -> MemHash::HashMatch: extend + containment de-dup.  With nway_mask:
                                 MaskedMemHash (src/mauveAligner.cpp:523-531).                        */
    MB_MODE_SEED_ENUM = 1,    /* SeedMatchEnumerator::HashMatch (src/SeedMatchEnumerator.h:71-123)    */
    MB_MODE_UNIQUE_COUNT = 2, /* SortedMerList::UniqueMerCount (src/uniqueMerCount.cpp:39)            */
    MB_MODE_PAIRWISE = 3,     /* PairwiseMatchFinder (src/progressiveMauve.cpp:496-501); at most 8 sequences */
    MB_MODE_REPEAT = 4        /* RepeatHash (mauveAligner --repeats, src/mauveAligner.cpp:480-487): ONE sequence; every bucket of
                               * min_multi..max_multi (at most 255) occurrences is one match with a component per occurrence,
                               * extended and de-duplicated like a MemHash entry; result columns as MB_MODE_SEED_ENUM */
};

typedef struct mb_params {
    int32_t mode;
    int32_t direct_only;  /* SeedMatchEnumerator::FindMatches(..., direct_repeats_only)       */
    uint64_t min_multi;   /* SeedMatchEnumerator::FindMatches(..., min_multi = 2, ...)         */
    uint64_t max_multi;   /* SeedMatchEnumerator::FindMatches(..., max_multi = 1000, ...)      */
    uint64_t nway_mask;   /* MaskedMemHash::SetMask(uint64), 0 = off                           */
} mb_params;

/* Result of one search, host side (pinned memory owned by the context; valid until
 * the next mb_find / mb_fetch_result / mb_ctx_destroy on that context).
 * CSR over matches; canonical order = SURVEY.md Appendix A D18.
 * Stands in for MemHash::GetMatchList(MatchList&) (src/progressiveMauve.cpp:545)
 * and SeedMatchEnumerator's mlist (src/SeedMatchEnumerator.h:31-32). */
typedef struct mb_result {
    uint64_t n_matches;
    uint64_t n_comps;
    const uint32_t* length;     /* [n_matches]                                              */
    const uint64_t* comp_off;   /* [n_matches + 1]                                          */
    const uint32_t* comp_seq;   /* [n_comps] sequence index (0 for SEED_ENUM)               */
    const int64_t* comp_start;  /* [n_comps] signed 1-based left end; < 0 = reverse strand  */
    uint64_t unique_mers;                /* distinct seeds over all sequences               */
    const uint64_t* unique_mers_per_seq; /* [nseq] SortedMerList::UniqueMerCount per SML    */
    uint32_t nseq;
} mb_result;

typedef struct mb_stats {
    uint64_t n_seeds, n_runs, n_buckets, n_candidates, n_extended, n_matches, n_comps;
    uint32_t radix_passes, record_bytes, dedup_batches, dedup_iters;
    /* device milliseconds per stage of the last mb_find (CUDA events on the context stream) */
    float ms_h2d, ms_pack, ms_extract, ms_sort, ms_bucket, ms_dedup, ms_output, ms_d2h, ms_total_device;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t kernel_launches;   /* launches of this library's own kernels in the last mb_find */
    /* the dominant kernel (one radix pass over all seed records): summed device time and launch count
     * of the seed sort's k_onesweep launches, CUDA events on the context stream */
    float ms_radix_kernels;
    uint32_t radix_launches;
} mb_stats;

/* ---- context ----------------------------------------------------------------- */
int mb_ctx_create(mb_ctx** out, int device);
int mb_ctx_destroy(mb_ctx* ctx);
/* Use an existing CUDA stream (cudaStream_t / CUstream handle, e.g. torch's); NULL = own stream. */
int mb_set_stream(mb_ctx* ctx, void* cuda_stream);

/* ---- inputs ------------------------------------------------------------------ */
/* MatchFinder::AddSequence(SortedMerList*, gnSequence*) (src/SeedMatchEnumerator.h:25,
 * src/mauveAligner.cpp:577).  `data` is a HOST buffer: ASCII bases (is_packed = 0; A/C/G/T any
 * case, anything else reads as A) or 2-bit words (is_packed = 1; 32 bases per uint64, first base
 * in the top bits).  The library copies to the device; the caller keeps ownership. */
int mb_add_sequence(mb_ctx* ctx, const uint8_t* data, uint64_t len, int is_packed, int* out_id);
/* Same, but `dev_ascii` is already DEVICE memory (used by the resident-input benchmark leg).  The pack kernel is only
 * ENQUEUED on the context stream: the buffer must stay valid until that stream has passed it (stream order is enough:
 * free it with a stream-ordered free, or after any later synchronising call on the context). */
int mb_add_sequence_device(mb_ctx* ctx, const void* dev_ascii, uint64_t len, int* out_id);
/* The 2-bit words of a sequence as the context holds them on the device (32 bases per uint64, first base in the top
 * bits; stream-ordered after the mb_add_sequence* call that produced them), and the way back in: a sequence given as
 * packed words already in DEVICE memory.  With several GPUs every rank uploads and packs only its share of the genomes and
 * the ranks exchange the packed words over NVLink (80 MB at C5) instead of every rank pulling every genome over PCIe. */
int mb_get_packed_device(mb_ctx* ctx, int seq, const void** dev_words, uint64_t* n_words);
int mb_copy_packed_device(mb_ctx* ctx, int seq, void* dst_dev);   /* the same words copied to dst_dev on the context stream */
int mb_add_sequence_device_packed(mb_ctx* ctx, const void* dev_words, uint64_t len, int* out_id);
/* MatchFinder::ClearSequences() (src/progressiveMauve.cpp:542) */
int mb_clear_sequences(mb_ctx* ctx);
/* getSeed() result handed to LoadSMLs/CreateMemorySMLs (src/mauveAligner.cpp:456,465): the raw
 * 64-bit pattern, MSB-first from its highest set bit, 1 = care. */
int mb_set_seed(mb_ctx* ctx, uint64_t pattern);

/* ---- the path ---------------------------------------------------------------- */
/* Whole path, host to host: stands in for LoadSMLs + XFinder.FindMatches(MatchList&) +
 * GetMatchList (src/progressiveMauve.cpp:446-451,492-495; src/repeatoire.cpp:1850,1866-1867;
 * src/mauveAligner.cpp:465,585).  Equivalent to mb_find_device + mb_fetch_result. */
int mb_find(mb_ctx* ctx, const mb_params* params, const mb_result** out);
/* MemHash keeps its hash table across FindMatches calls until Clear() — the seed-family search relies on it
 * (src/progressiveMauve.cpp:503-548: FindMatches per seed pattern, ClearSequences() in between, ONE GetMatchList and
 * Clear() at the end).  on != 0: every MB_MODE_UNIQUE search of this context from now on drops the candidates whose
 * seed lies inside a match of an earlier search (same genome set, strands and diagonal) before its own de-dup, and adds
 * the matches it accepts to the table; each search still returns only its own matches (the host finder unions them,
 * include/mems_compat/mems_compat.h).  on == 0: the table is forgotten (MemHash::Clear()).  Single-GPU searches only. */
int mb_accumulate(mb_ctx* ctx, int on);
/* Device part only: packed sequences resident in HBM -> canonical match CSR resident in HBM.
 * Launches on the context stream; it reads a few counters back on the way (buffer sizes), so it returns with the stream
 * drained up to the output stage.  FindMatchesFromPosition's start offsets (src/mauveAligner.cpp:585) have no
 * counterpart: a search always covers the whole sequences. */
int mb_find_device(mb_ctx* ctx, const mb_params* params);
/* Device -> pinned host copy of the last mb_find_device result (synchronises the stream). */
int mb_fetch_result(mb_ctx* ctx, const mb_result** out);
/* The same result in the compact layout the device keeps it in: one byte per component for the sequence index (at most
 * 64 sequences) and a 4-byte signed start (every sequence is shorter than 2^31 bases) — 5 instead of 12 bytes per
 * component over PCIe (C5: 1.3 instead of 3.0 GB).  mems_compat.h builds its Match objects from this form.  The two
 * result forms share the context's pinned buffers: a result is valid until the next fetch / search on the context. */
typedef struct mb_result_compact {
    uint64_t n_matches;
    uint64_t n_comps;
    const uint32_t* length;          /* [n_matches]      */
    const uint64_t* comp_off;        /* [n_matches + 1]  */
    const uint8_t* comp_seq8;        /* [n_comps]        */
    const int32_t* comp_start32;     /* [n_comps] signed, 1-based; < 0 = reverse strand */
    uint64_t unique_mers;
    const uint64_t* unique_mers_per_seq;
    uint32_t nseq;
} mb_result_compact;
int mb_fetch_result_compact(mb_ctx* ctx, const mb_result_compact** out);
int mb_find_compact(mb_ctx* ctx, const mb_params* params, const mb_result_compact** out);

/* repeatoire's match position lookup table (src/repeatoire.cpp:1944-1966: every component of every seed match entered
 * at its left end into a table over the sequence), built on the device from the last single-sequence result
 * (MB_MODE_SEED_ENUM / MB_MODE_REPEAT).  n_pos = sequence length + 1; entry p (1-based left end) = index of the match — in
 * result order, which is ascending LeftEnd(0), the order of repeatoire's seed_sort_list (:1920-1935) — and of its
 * component that start at p; 0xFFFFFFFF where none does.  Pinned host arrays owned by the context. */
int mb_position_table(mb_ctx* ctx, const uint32_t** match_of_pos, const uint32_t** comp_of_pos, uint64_t* n_pos);

/* ---- many small problems in one pass (recursive anchoring: the aligners re-run the search inside every gap between
 * anchors — `recursive` flag src/mauveAligner.cpp:94,698; SetRecursive src/progressiveMauve.cpp:661-664) ------------------
 * mb_find_batch: n_problems independent searches, each over its own nseq sequences (seqs[i * nseq + g] / lens[i * nseq + g]:
 * HOST ASCII of sequence g of problem i; a length of 0 = the problem has no sequence g), with one seed and one policy.
 * The library lays the problems side by side, runs the pipeline ONCE (the problem index leads the sort key, windows and
 * extension stop at the piece boundaries) and cuts the result by problem.  Every problem's matches are exactly those
 * of its own mb_find, in its canonical order, with coordinates relative to its own sequences.
 * Single GPU; not MB_MODE_UNIQUE_COUNT.  The result is owned by the context until its next batch / destroy. */
typedef struct mb_batch_result {
    uint64_t n_problems;
    uint64_t n_matches, n_comps;
    const uint64_t* match_off;   /* [n_problems + 1]: matches of problem i = [match_off[i], match_off[i + 1])      */
    const uint32_t* length;      /* [n_matches]                                                                    */
    const uint64_t* comp_off;    /* [n_matches + 1]                                                                */
    const uint32_t* comp_seq;    /* [n_comps]                                                                      */
    const int64_t* comp_start;   /* [n_comps] signed, 1-based, relative to the problem's own sequence              */
} mb_batch_result;
int mb_find_batch(mb_ctx* ctx, const mb_params* params, uint32_t n_problems, uint32_t nseq, const uint8_t* const* seqs, const uint64_t* lens,
                  const mb_batch_result** out);
/* The mechanism underneath, for callers that hold the concatenated sequences already: after the mb_add_sequence calls,
 * bounds[g * (n_problems + 1) + i] = first base of piece i of sequence g (ascending, bounds[..][0] = 0, the last entry =
 * the sequence length).  The following searches are segmented (coordinates stay those of the concatenations);
 * n_problems = 0 or mb_clear_sequences ends it. */
int mb_set_segments(mb_ctx* ctx, uint32_t n_problems, const uint64_t* bounds);

/* ---- multi-GPU path ---------------------------------------------
 * One context per rank / GPU, the same sequences added to every context (replicated genomes).
 * The library runs the per-rank stages and owns the exchange buffers; the exchanges between the stages are
 * either the caller's (variable all-to-alls over NCCL/NVLink; mauvealigner_b200/dist.py drives them over
 * torch.distributed) or, for exchange 1, fused into the partition kernel as NVLink peer stores.
 * MB_MODE_UNIQUE, 8-byte records.  Stands in for the same reference calls as mb_find (the reference is
 * single-process).
 *   1 mb_dist_extract      -> *d_send: this rank's seed records grouped by destination rank (= seed-key range),
 *                             h_counts[world] records per destination               (exchange 1: 1 word / record)
 *     or mb_dist_extract_count + mb_dist_partition(peer arrays): the same, with the exchange fused into the kernel
 *   2 mb_dist_recv_buffer(0, n) is where exchange 1 must deliver (source-rank order);
 *     mb_dist_local        -> sort / runs / policy over the key range, EVERY candidate extended at its source;
 *                             h_row_counts[world] = candidates per owner of their de-dup group;
 *     mb_dist_rows_pack    -> 4-word candidate rows (two group hashes, first component, extents) grouped by owner:
 *                             at *d_rows (exchange 2: 4 words / row), or stored into the owners' buffers directly
 *   3 mb_dist_recv_buffer(1, 4 n) receives the rows (source-rank order);
 *     mb_dist_resolve      -> chains / resolve over the owned groups; *d_verdict = one byte per received row
 *                             (1 accepted), to go back to the sources in row order    (exchange 2b: 1 byte / row)
 *     mb_dist_recv_buffer(5, n_bytes) receives this rank's verdicts (the row order of stage 2);
 *     mb_dist_accept       -> *d_hist = 4096 uint64 counts (device) of the accepted matches' canonical keys:
 *                             sum it over all ranks IN PLACE (all-reduce)
 *     mb_dist_match_partition -> matches grouped by destination = range of the canonical order: counts per destination;
 *     mb_dist_match_pack   -> their rows (*d_hdr 2 words per row, *d_comps 1 word per component)   (exchange 3),
 *                             or stored into the destinations' buffers directly
 *   4 mb_dist_recv_buffer(3 / 4, n) receive them; mb_dist_output builds the canonical CSR of this rank's
 *     range, after which mb_fetch_result works as for mb_find_device.  The ranks' pieces, concatenated in
 *     rank order, are the result. */
int mb_dist_extract(mb_ctx* ctx, int rank, int world, void** d_send, uint64_t* h_counts);
/* MB_MODE_UNIQUE_COUNT and MB_MODE_SEED_ENUM over several GPUs (no exchange after the first — the ranks' counts
 * add up, their match lists are disjoint), for 8-byte and 16-byte seed records (long genomes / heavy seeds: the BASELINE
 * configs C3 and C4):
 *   1 mb_dist_extract_records = mb_dist_extract for either format: *d_send_keys (and, for 16-byte records, *d_send_vals;
 *     NULL otherwise) hold the slice's records grouped by destination key range, h_counts[world] records each   (exchange 1)
 *   2 mb_dist_recv_buffer(0, n) (and (2, n)) receive them in source-rank order; mb_dist_enum_local sorts them and runs the
 *     mode's tail over this key range, after which mb_fetch_result works as for mb_find_device:
 *       MB_MODE_UNIQUE_COUNT  unique_mers and unique_mers_per_seq of the key range — SUM them over the ranks;
 *       MB_MODE_SEED_ENUM     this rank's matches in canonical order (by first position).  The ranks' pieces are disjoint and
 *                             their union is the result; merging them by first position (unique per match) gives the
 *                             canonical list of the single-GPU search.
 * Replaces, per rank, the SortedMerList creation + UniqueMerCount / SeedMatchEnumerator::FindMatches of
 * /root/reference/src/uniqueMerCount.cpp:35-43 and /root/reference/src/SeedMatchEnumerator.h:19-33,59-68. */
int mb_dist_extract_records(mb_ctx* ctx, int rank, int world, void** d_send_keys, void** d_send_vals, uint64_t* h_counts);
int mb_dist_enum_local(mb_ctx* ctx, const mb_params* params, uint64_t n_recv);
/* Stage 1 in two steps, for the fused exchange: mb_dist_extract_count extracts the slice and counts it per
 * destination; after the ranks have shared their counts, mb_dist_partition either fills the local send buffer
 * (peer_bases = NULL, as mb_dist_extract does) or writes every record straight into its destination rank's
 * receive array over NVLink peer stores (peer_bases[d] = rank d's array mapped into this process, peer_offsets[d] =
 * record index of this rank's block in it).  All ranks must be synchronised before stage 2 reads the arrays
 * (mb_dist_use_p2p_recv(ctx, 1) makes mb_dist_local read from the array of mb_dist_p2p_recv_array). */
int mb_dist_extract_count(mb_ctx* ctx, int rank, int world, uint64_t* h_counts);
int mb_dist_partition(mb_ctx* ctx, void* const* peer_bases, const uint64_t* peer_offsets, void** d_send);
int mb_dist_p2p_recv_array(mb_ctx* ctx, uint64_t capacity_records, void** d_ptr);
int mb_dist_use_p2p_recv(mb_ctx* ctx, int on);
/* CUDA IPC plumbing for the peer arrays (64-byte handles; cudaIpcGetMemHandle / OpenMemHandle / CloseMemHandle) */
int mb_ipc_export(mb_ctx* ctx, void* d_ptr, uint8_t* handle64);
int mb_ipc_import(mb_ctx* ctx, const uint8_t* handle64, void** d_ptr);
int mb_ipc_close(mb_ctx* ctx, void* d_ptr);
/* which: 0 seed records (first words), 2 second words of 16-byte seed records, 1 candidate rows, 3 match headers,
 * 4 match components (n = 8-byte words), 5 verdicts (n = bytes) */
int mb_dist_recv_buffer(mb_ctx* ctx, int which, uint64_t n, void** d_ptr);
int mb_dist_local(mb_ctx* ctx, const mb_params* params, uint64_t n_recv, uint64_t* h_row_counts);
/* Last step of stage 2: write the rows in owner order into this rank's send buffer (*d_rows; peer_bases = NULL) or,
 * fused with exchange 2, straight into the owners' receive buffers 1 over NVLink (peer_bases[d] mapped with
 * mb_ipc_import, peer_row_offsets[d] = row index of this rank's block in rank d's buffer).  Asynchronous. */
int mb_dist_rows_pack(mb_ctx* ctx, void* const* peer_bases, const uint64_t* peer_row_offsets, void** d_rows);
/* Push the destination blocks of a local send buffer (counts[d] units of unit_words 8-byte words, back to back) into
 * the destination ranks' mapped receive buffers with device-to-device copies (copy engines over NVLink) on the
 * context stream; dst_offsets[d] = unit index of this rank's block in rank d's buffer.  Asynchronous. */
int mb_dist_push(mb_ctx* ctx, const void* d_src, const uint64_t* counts, uint32_t unit_words, void* const* peer_bases,
                 const uint64_t* dst_offsets);
int mb_dist_resolve(mb_ctx* ctx, uint64_t n_rows, void** d_verdict);
int mb_dist_accept(mb_ctx* ctx, void** d_hist);
int mb_dist_match_partition(mb_ctx* ctx, uint64_t* h_match_counts, uint64_t* h_comp_counts);
/* Last step of stage 3c: write the match rows in destination order into this rank's send buffers (*d_hdr, *d_comps;
 * bases = NULL) or, fused with exchange 3, into the destination ranks' receive buffers 3 / 4 over NVLink. */
int mb_dist_match_pack(mb_ctx* ctx, void* const* hdr_bases, const uint64_t* hdr_row_offsets, void* const* comp_bases,
                       const uint64_t* comp_word_offsets, void** d_hdr, void** d_comps);
int mb_dist_output(mb_ctx* ctx, uint64_t n_match, uint64_t n_comp);
/* device ms of this rank's stages of the last run: [0] extract+partition, [1] sort, [2] runs/policy/candidates,
 * [3] extension, [4] chains + resolve; [5..7] reserved (0) */
int mb_dist_stage_ms(mb_ctx* ctx, float* out8);

/* The same multi-GPU search driven from C++ inside ONE process: ctxs[r] = the context of rank r (normally one per GPU;
 * contexts may share a device), the same sequences and seed set on every context.  One host thread per context runs
 * the stages above; every exchange goes device to device over NVLink peer access (exchange 1 fused into the partition
 * kernel, the others pushed by the copy engines), no NCCL, no torch.  Blocks until every rank's piece of the result is
 * on its device: then mb_fetch_result(ctxs[r]) per rank, pieces concatenated in rank order = the result of mb_find
 * (same reference calls as mb_find; MB_MODE_UNIQUE only). */
int mb_find_multi(mb_ctx* const* ctxs, int world, const mb_params* params);

/* ---- sorted mer list access (SortedMerList façade; "next" row of SURVEY.md §8f) ---- */
/* Positions of sequence `seq` sorted by (seed, position): the .sslist position array (a4).
 * out_pos must hold len-L+1 entries (host). Valid after a mb_find* call in any mode. */
int mb_get_sml(mb_ctx* ctx, int seq, uint32_t* out_pos, uint64_t capacity, uint64_t* out_n);
/* SortedMerList::GetMer-style per-position mers (key << (64-2w) | strand) of one sequence, for
 * record-by-record parity checks of the extraction kernel. */
int mb_get_mers(mb_ctx* ctx, int seq, uint64_t* out_mers, uint64_t capacity, uint64_t* out_n);

/* ---- diagnostics ------------------------------------------------------------- */
int mb_get_stats(mb_ctx* ctx, mb_stats* out);
const char* mb_strerror(int code);
const char* mb_last_cuda_error(mb_ctx* ctx);
int mb_device_count(void);
/* tuning aid: mean device ms of one radix pass / one whole sort over n synthetic records (tools/bench_radix.py) */
int mb_debug_radix(mb_ctx* ctx, uint64_t n, int shift, int kbits, int reps, float* ms_out);
const char* mb_version(void);

#ifdef __cplusplus
}
#endif
#endif

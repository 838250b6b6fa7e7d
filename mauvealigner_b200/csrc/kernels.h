// kernels.h — host-side launchers of the pipeline kernels (one translation unit per stage).
#pragma once
#include "common.cuh"

// internal copies of the public mode constants (kernels do not include the public header)
#define MB_MODE_UNIQUE_ 0
#define MB_MODE_SEED_ENUM_ 1
#define MB_MODE_UNIQUE_COUNT_ 2
#define MB_MODE_PAIRWISE_ 3
#define MB_MODE_REPEAT_ 4

// ---- kernels_seed.cu
struct ExtractArgs;
void launch_pack(const u8* d_ascii, u64 len, u64* d_words, u64 n_words, cudaStream_t st);
u32 extract_tile_size();
void launch_extract_records(const u64* d_packed, u64* d_keys, u64* d_vals, u32* d_hist, int npass, const GenomeTable& gt,
                            const SeedDev& sd, const RecFmt& fmt, const u32* d_tile_first, u32 n_tiles, cudaStream_t st, u32 tile0 = 0,
                            u32 out_base = 0);
void launch_mers(const u64* d_packed, const GenomeTable& gt, const SeedDev& sd, const RecFmt& fmt, int genome, u64* d_mers, cudaStream_t st);
void launch_hist(const u64* d_keys, u32 n, int shift0, int kbits, int npass, u32* d_hist, cudaStream_t st);

// ---- kernels_radix.cu
size_t radix_smem_bytes(bool has_val);
u32 radix_tile_size();
void launch_scan_hist(const u32* d_hist, u32* d_base, int npass, cudaStream_t st);
// lut (optional, device, 256 bytes): the bin of a record is lut[digit] instead of the digit itself (monotone
// key-range partition of the multi-GPU path); peers (optional, device, 256 pointers): bin b is written to peers[b],
// which may be another GPU's memory — digit_base[b] is then the index of this rank's block inside that array
cudaError_t launch_onesweep(const u64* kin, u64* kout, const u64* vin, u64* vout, u32 n, const u32* d_digit_base, u64* d_lookback,
                            u32* d_ticket, int shift, int bits, cudaStream_t st, const u8* lut = nullptr, u64* const* peers = nullptr);

// ---- kernels_bucket.cu
u32 find_runs_tile();
size_t find_runs_workspace_words(u32 n);
const unsigned short* launch_find_runs(const u64* keys, const u64* vals, u32 n, const RecFmt& fmt, u32* run_start, u32* run_u, u64* status, u32* ticket,
                      u64* per_seq_count, u32* totals, cudaStream_t st);
struct SelectArgs {
    const u64* keys; const u64* vals;
    const u32* run_start; const u32* run_u;
    const u32* n_runs_ptr;
    int mode, direct_only;
    u64 min_multi, max_multi, nway_mask;
    u64* status; u32* ticket;
    u64* n_buckets;
    u32* totals;     // [0] candidates, [1] components
    u32* cand_run; u32* cand_off; u32* cand_aux;
};
u32 select_tile();
void launch_select(const SelectArgs& a, const RecFmt& fmt, u32 n_runs_upper, cudaStream_t st);
struct EmitUniqueArgs {
    const u64* keys; const u64* vals;
    const u32* run_start; const u32* run_u;
    const u32* cand_run; const u32* cand_off; const u32* cand_aux;
    const u32* totals;
    int mode;
    u32* comp_pos; u8* comp_gs;
    u64* bitmap;
    u64* ghash;      // per candidate: hash of (genome set, strands, diagonal)
    u64* ghash2;     // second, independent hash of the same key (low 8 bits cleared)
    const unsigned short* masks; // per 8 sorted records: head / unique-genome flags (launch_find_runs)
    u32 seedL;       // seed length: a reverse component enters the hashes with position + seedL + first position, which
                     // extension leaves unchanged (so groups of different seed patterns meet, kernels_family.cu)
};
void launch_emit_unique(const EmitUniqueArgs& a, const RecFmt& fmt, const GenomeTable& gt, u32 n_cand_upper, cudaStream_t st);
struct EmitEnumArgs {
    const u64* keys; const u64* vals;
    const u32* run_start; const u32* cand_run; const u32* cand_off; const u32* totals;
    u64* sort_key; u64* sort_val; u32* ncomp;
    const u64* sorted_val; const u64* out_off;
    u32* out_len; u8* out_seq; int32_t* out_start;   // the device result is compact: 1-byte sequence, 4-byte signed start per component
    int repeat;               // MB_MODE_REPEAT: ties of the first start are ordered as for MB_MODE_SEED_ENUM (multiplicity, signed starts, length)
};
void launch_enum_keys(const EmitEnumArgs& a, const RecFmt& fmt, u32 n_upper, cudaStream_t st);
void launch_enum_gather(const EmitEnumArgs& a, const RecFmt& fmt, u32 seedL, u32 n_upper, cudaStream_t st);
u32 scan_tile();
void launch_scan_u32(const u32* in, u64 n, u32* out32, u64* out64, u64* status, u32* ticket, u64* total_out, cudaStream_t st);
void launch_scan_popc(const u64* in, u64 n, u32* out32, u64* status, u32* ticket, u64* total_out, cudaStream_t st);

// ---- kernels_dedup.cu
struct DedupArgs {
    const u64* packed;
    u32 n_cand;
    const u32* cand_off; const u32* comp_pos; const u8* comp_gs;
    const u64* bitmap; const u32* bmrank;   // candidates by (first genome, position): bitmap over global bases + word ranks
    const u64* ghash;   // candidate -> hash of its D16 group (genome set, strands, diagonal)
    const u64* ghash2;  // candidate -> second independent hash of the group
    // per slot (candidates in (first genome, position) order)
    ulonglong2* slot_rec; // slot -> (group hash | adjacent-to-previous-slot bit, candidate)
    u8* link_bits;      // one bit per slot: continues the chain of the previous slot
    u32* chain_min;     // slot -> lowest rank among the earlier members of its chain
    u64* rep_bits;      // one bit per slot: lowest rank of its chain
    const u32* rep_rank; // per 64-slot word: reps before it
    // per rep, in (group colour, slot) order
    u32 n_rep;
    const u64* s_key;   // colour << 32 | slot
    ulonglong2* s_rec;  // colour:16 | slot:32 | hash[15:0], hash[47:16] | candidate:32
    u8* rstate;         // 0 undecided, 1 accepted, 2 dropped; flag bits: see k_resolve
    u32* s_cand;        // rep -> candidate
    u64* s_h2;          // rep -> second group hash
    u32* reach;         // largest index distance of any claimer of this rep
    uint4* xrec;        // rep -> (candidate, first component row, components | first genome << 8, first position)
    uint4* xstate;      // saved walk state of a rep whose extension is continued by k_extend_more
    u32* rng_lo; u32* rng_hi; // slot range of the extent
    u64* minrank;       // [2][n_rep] lowest undecided claimer (rank << 32 | index) of the even / odd rounds
    u32* ext_l; u32* ext_r;   // per candidate
    // device-resident work lists (three rotating lists of undecided reps, long extensions, wide extents)
    // and their counters (layout: see k_resolve)
    u32* wl0; u32* wl1; u32* wl2; u32* wd0; u32* wd1; u32* wd2; u32* wl_long; u32* ctr;
    u64* trace;         // optional phase trace (debug): [0] count, then (tag, ns) pairs
    u8* shadow;         // per candidate, or null: set by the shared long walks for a rep whose round-0 claims a lower-rank rep of the
                        // same group and extent makes anyway (k_extend_long_classes); k_resolve skips those claims
    const u8* pre_drop; // per candidate, or null: dropped before the de-dup (contained in a match of an earlier call, kernels_family.cu)
    // multi-GPU owner side: the candidates arrive as 4-word rows (kernels_dist.cu) with their extents already
    // known and without component lists; null on the single-GPU path
    const u64* rows;
};
// persistent MemHash table across calls (kernels_family.cu): t_* = entries in arrival order, s_* = sorted by (hash, start)
struct FamilyArgs {
    u64* t_h1; u64* t_h2; u32* t_x; u32* t_end;
    u64* s_h1; u64* s_h2; u32* s_x; u32* s_end; u32* s_pmax; u32* s_run0;
};
void launch_family_append(const FamilyArgs& f, const u8* rstate, const u32* s_cand, u32 n_rep, const u64* ghash, const u64* ghash2, const u32* cand_off,
                          const u32* comp_pos, const u32* ext_l, const u32* ext_r, u32 L, u32* counter, cudaStream_t st);
void launch_family_key_x(const FamilyArgs& f, u32 n, u64* key, u64* val, cudaStream_t st);
void launch_family_key_h(const FamilyArgs& f, u32 n, const u64* val, u64* key, cudaStream_t st);
void launch_family_gather(const FamilyArgs& f, u32 n, const u64* perm, cudaStream_t st);
void launch_family_filter(const FamilyArgs& f, u32 n_tab, u32 n_cand, u64* ghash, const u64* ghash2, const u32* cand_off, const u32* comp_pos, u32 L,
                          u8* pre_drop, u32* n_dropped, cudaStream_t st);
void launch_slot_scatter(const DedupArgs& a, const GenomeTable& gt, cudaStream_t st);
u32 chain_tile();
void launch_chains(const DedupArgs& a, u64* status_fwd, u32* ticket_fwd, u64* status_bwd, u32* ticket_bwd, cudaStream_t st);
void launch_rep_keys(const DedupArgs& a, u64* skey, cudaStream_t st, const u32* cellR = nullptr, const u32* cellNew = nullptr, u32 V = 0);
// layout of the extension records by (block of positions, first genome, slot): cell tables for launch_rep_keys (two launches)
u32 extension_cells(u32 V, u64 maxlen);
void launch_extension_cells(const DedupArgs& a, const GenomeTable& gt, u32 V, u64 maxlen, u64 axis_end, u32* cellR, u32* cellNew, cudaStream_t st);
int extend_launches();
// extends the a.n_rep items of a.xrec (filled by launch_rep_keys in slot order, or by launch_cand_xrec); a.bitmap == null:
// extents only, the slot ranges are derived later (launch_rep_setup / launch_extent_ranges)
void launch_extend(const DedupArgs& a, const GenomeTable& gt, const SeedDev& sd, cudaStream_t st);
// the reps the bounded rounds of launch_extend left unfinished (a.ctr[6] of them on a.wl_long): warp per rep, or classes
// of reps of one group that share a walk (keys from launch_long_keys, sorted by the caller)
void launch_extend_long(const DedupArgs& a, const GenomeTable& gt, const SeedDev& sd, cudaStream_t st);
void launch_long_keys(const DedupArgs& a, u32 L, u32 n_long, u64* keys, cudaStream_t st);
void launch_extend_long_classes(const DedupArgs& a, const GenomeTable& gt, const SeedDev& sd, const u64* sorted_keys, u32 n_long, cudaStream_t st);
void launch_rep_setup(const DedupArgs& a, const GenomeTable& gt, cudaStream_t st);
// multi-GPU source side: every candidate is its own "rep" (a.n_rep = a.n_cand, extension records straight from the CSR)
void launch_cand_xrec(const DedupArgs& a, const GenomeTable& gt, cudaStream_t st);
// multi-GPU owner side: slot ranges of the reps' extents from the extents that came with the rows
void launch_extent_ranges(const DedupArgs& a, const GenomeTable& gt, cudaStream_t st);
cudaError_t launch_resolve(const DedupArgs& a, cudaStream_t st); // cooperative

// ---- kernels_output.cu
struct OutputArgs {
    u32 n_items;              // reps (or, on rank 0 of the multi-GPU path, gathered matches)
    const u8* state;          // item accepted iff (state & 15) == 1
    const u32* item_cand;     // item -> candidate
    const u32* cand_off; const u32* comp_pos; const u8* comp_gs;
    const u32* ext_l; const u32* ext_r;
    u32* flags; const u32* match_idx;
    const u64* n_matches_ptr;
    u64* sort_key; u64* sort_val;
    u32* ncomp;
    const u64* out_off;
    u32* out_len; u8* out_seq; int32_t* out_start;   // the device result is compact: 1-byte sequence, 4-byte signed start per component
    int repeat;               // MB_MODE_REPEAT: ties of the first start are ordered as for MB_MODE_SEED_ENUM (multiplicity, signed starts, length)
};
void launch_position_table(const u64* out_off, const int32_t* out_start, u32 n_matches, u64* tab, u32* match_of, u32* comp_of, u64 n_pos, cudaStream_t st);
// compact device result -> the wide arrays of mb_result (u32 sequence, i64 start)
void launch_expand_result(const u8* seq8, const int32_t* start32, u64 n_comps, u32* seq, i64* start, cudaStream_t st);
void launch_uniq_flags(const OutputArgs& a, cudaStream_t st);
void launch_uniq_keys(const OutputArgs& a, int sbits, cudaStream_t st);
void launch_uniq_tiefix(const OutputArgs& a, const u64* skey, u64* sval, u32 L, u32 n_upper, cudaStream_t st);
void launch_uniq_ncomp(const OutputArgs& a, const u64* sval, u32 n_upper, cudaStream_t st);
void launch_uniq_gather(const OutputArgs& a, const u64* sval, u32 L, u32 n_upper, cudaStream_t st);

// ---- kernels_dist.cu (multi-GPU exchange formats)
void launch_fold_lut(const u32* hist, const u8* lut, u32 nbins, u32 world, u32* digit_base, u64* counts, cudaStream_t st);
void launch_owner_keys(const u64* ghash, u32 n, u32 world, u64* skey, u64* sval, cudaStream_t st);
void launch_hdr_m(const u64* hdr, u32 n, u32* m_out, cudaStream_t st);
void launch_unpack_match(const u64* hdr, const u64* comps, const u32* cand_off, u32 n, u32 n_comp, u32* comp_pos, u8* comp_gs, u32* ext_l, u32* ext_r,
                         u8* state, u32* item_cand, cudaStream_t st);
void launch_match_keys(const u8* state, const u32* item_cand, const u32* match_idx, const u32* cand_off, const u8* comp_gs, const u32* comp_pos,
                       const u32* ext_l, u32 n_items, int sbits, int binshift, u64* key, u32* item_of, u64* hist, cudaStream_t st);
// Destination table of a pack kernel (device memory, world + 1 entries, entry `world` is the sentinel with
// bound = number of rows): the rows [bound[d], bound[d+1]) of the partition order go to rank d — into `base`
// (this rank's own send buffer, or rank d's receive buffer mapped over NVLink) at row `off` + (row - bound);
// base2 / off2 / cbound: the same for the component words of the match rows.
struct PeerDst { u64* base; u64 off; u64* base2; u64 off2; u64 cbound; u32 bound; u32 pad; };
// 4-word candidate rows of the extend-at-source protocol: [group hash, second hash, g0 | vg << 8 | m << 24 | p0 << 32,
// ext_l | ext_r << 32], written in partition (owner) order; perm_out[j] = candidate of row j
void launch_pack_rows(const u64* perm, u32 n, const u64* ghash, const u64* ghash2, const u32* cand_off, const u32* comp_pos, const u8* comp_gs,
                      const GenomeTable& gt, const u32* ext_l, const u32* ext_r, const PeerDst* tab, u32 world, u32* perm_out, cudaStream_t st);
void launch_rows_bitmap(const u64* rows, u32 n, const GenomeTable& gt, u64* bitmap, cudaStream_t st);
void launch_accept_mark(const u8* rstate, const u32* s_cand, u32 n_rep, u8* acc, cudaStream_t st);
void launch_apply_accept(const u8* acc, const u32* perm, u32 n, u8* state, u32* item_cand, cudaStream_t st);
void launch_dest_keys(const u64* key, u32 n, int binshift, const u8* lut, u64* skey, u64* sval, cudaStream_t st);
void launch_match_perm_m(const u64* perm, const u32* item_of, const u32* item_cand, const u32* cand_off, u32 n, u32* m_out, cudaStream_t st);
void launch_pack_match_perm(const u64* perm, const u32* item_of, const u32* item_cand, const u64* poff, const u32* cand_off, const u32* comp_pos,
                            const u8* comp_gs, const u32* ext_l, const u32* ext_r, u32 n, const PeerDst* tab, u32 world, cudaStream_t st);

// api_batch.cu — many small problems in ONE pass of the pipeline (recursive anchoring re-runs the multi-MUM search inside
// every gap between anchors: `recursive` flag /root/reference/src/mauveAligner.cpp:94,698; SetRecursive
// /root/reference/src/progressiveMauve.cpp:661-664 — thousands of searches over a few hundred bases each).
//
// The problems are laid side by side: genome g of the batch is the concatenation of every problem's g-th sequence, and a
// segment table tells the kernels where each problem's piece lies (GenomeTable::seg, common.cuh).  The problem index
// leads the sort key, so seeds only meet seeds of their own problem; seed windows never cross a piece boundary and the
// extension stops at it.  One extraction, one sort, one de-dup for the whole batch instead of ~40 launches and five host
// read-backs per problem; the canonical order of the batch restricted to one problem is that problem's canonical order,
// so the host only has to cut the result by problem and shift the coordinates.
#include "ctx.h"

extern "C" {

int mb_set_segments(mb_ctx* c, uint32_t n_problems, const uint64_t* bounds) {
    if (!c) return MB_E_ARG;
    if (n_problems == 0) { c->n_seg = 0; c->h_seg.clear(); return MB_OK; }
    if (!bounds) return MB_E_ARG;
    const size_t nseq = c->seq_len.size();
    if (nseq == 0) return MB_E_NOSEQ;
    std::vector<u32> h(nseq * ((size_t)n_problems + 1));
    for (size_t g = 0; g < nseq; ++g) {
        const uint64_t* b = bounds + g * ((size_t)n_problems + 1);
        if (b[0] != 0 || b[n_problems] != c->seq_len[g]) return MB_E_ARG;
        for (uint32_t i = 0; i <= n_problems; ++i) {
            if (i && b[i] < b[i - 1]) return MB_E_ARG;
            h[g * ((size_t)n_problems + 1) + i] = (u32)b[i];
        }
    }
    c->h_seg.swap(h);
    c->n_seg = n_problems;
    c->have_result = false;
    return MB_OK;
}

int mb_find_batch(mb_ctx* c, const mb_params* prm, uint32_t n_problems, uint32_t nseq, const uint8_t* const* seqs, const uint64_t* lens,
                  const mb_batch_result** out) {
    if (!c || !prm || !out || !lens || (!seqs && n_problems)) return MB_E_ARG;
    if (nseq == 0 || nseq > MB_MAX_SEQ) return MB_E_SEQCOUNT;
    if (prm->mode == MB_MODE_UNIQUE_COUNT) return MB_E_ARG;
    c->b_moff.assign((size_t)n_problems + 1, 0);
    c->b_coff.assign(1, 0);
    c->b_len.clear(); c->b_seq.clear(); c->b_start.clear();
    mb_batch_result& br = c->bres;
    memset(&br, 0, sizeof(br));
    br.n_problems = n_problems;
    auto publish = [&]() {
        br.n_matches = c->b_len.size(); br.n_comps = c->b_seq.size();
        br.match_off = c->b_moff.data(); br.length = c->b_len.data(); br.comp_off = c->b_coff.data();
        br.comp_seq = c->b_seq.data(); br.comp_start = c->b_start.data();
        *out = &br;
    };
    if (n_problems == 0) { publish(); return MB_OK; }
    // ---- the batch genomes: piece i of genome g = sequence g of problem i
    std::vector<u64> bounds((size_t)nseq * (n_problems + 1), 0);
    TRY(mb_clear_sequences(c));
    for (uint32_t g = 0; g < nseq; ++g) {
        u64 total = 0;
        for (uint32_t i = 0; i < n_problems; ++i) total += lens[(size_t)i * nseq + g];
        if (total >= (1ull << 32)) return MB_E_TOOLONG;
        c->b_concat.resize(total ? total : 1);
        u64 at = 0;
        for (uint32_t i = 0; i < n_problems; ++i) {
            const u64 l = lens[(size_t)i * nseq + g];
            bounds[(size_t)g * (n_problems + 1) + i] = at;
            if (l) {
                if (!seqs[(size_t)i * nseq + g]) return MB_E_ARG;
                memcpy(c->b_concat.data() + at, seqs[(size_t)i * nseq + g], l);
            }
            at += l;
        }
        bounds[(size_t)g * (n_problems + 1) + n_problems] = at;
        TRY(mb_add_sequence(c, c->b_concat.data(), total, 0, nullptr));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream)); // the staging buffer is reused for the next genome
    }
    TRY(mb_set_segments(c, n_problems, bounds.data()));
    const mb_result* r = nullptr;
    const bool fam = c->fam_on; // a batch is a set of independent searches: the table of mb_accumulate stays out of it
    c->fam_on = false;
    int rc = mb_find(c, prm, &r);
    c->fam_on = fam;
    if (rc != MB_OK) { mb_set_segments(c, 0, nullptr); return rc; }
    // ---- cut by problem (stable: the batch order restricted to a problem is the problem's canonical order)
    const u64 nm = r->n_matches;
    std::vector<u32> prob(nm);
    const bool columns = prm->mode == MB_MODE_SEED_ENUM || prm->mode == MB_MODE_REPEAT; // components all in sequence 0
    for (u64 m = 0; m < nm; ++m) {
        const u64 k = r->comp_off[m];
        const u32 g = columns ? 0 : r->comp_seq[k];
        const u64 p = (u64)(r->comp_start[k] < 0 ? -r->comp_start[k] : r->comp_start[k]) - 1;
        const u64* b = bounds.data() + (size_t)g * (n_problems + 1);
        const u32 i = (u32)(std::upper_bound(b, b + n_problems + 1, p) - b) - 1;
        prob[m] = i;
        ++c->b_moff[i + 1];
    }
    for (uint32_t i = 0; i < n_problems; ++i) c->b_moff[i + 1] += c->b_moff[i];
    std::vector<u64> order(nm), next(c->b_moff.begin(), c->b_moff.end() - 1);
    for (u64 m = 0; m < nm; ++m) order[next[prob[m]]++] = m;
    c->b_len.resize(nm); c->b_coff.resize(nm + 1);
    c->b_seq.resize(r->n_comps); c->b_start.resize(r->n_comps);
    u64 o = 0;
    for (u64 j = 0; j < nm; ++j) {
        const u64 m = order[j];
        const u32 i = prob[m];
        c->b_len[j] = r->length[m];
        c->b_coff[j] = o;
        for (u64 k = r->comp_off[m]; k < r->comp_off[m + 1]; ++k) {
            const u32 g = columns ? 0 : r->comp_seq[k];
            const i64 shift = (i64)bounds[(size_t)g * (n_problems + 1) + i];
            const i64 s = r->comp_start[k];
            c->b_seq[o] = r->comp_seq[k];
            c->b_start[o] = s < 0 ? s + shift : s - shift;
            ++o;
        }
    }
    c->b_coff[nm] = o;
    mb_set_segments(c, 0, nullptr);
    publish();
    return MB_OK;
}

} // extern "C"

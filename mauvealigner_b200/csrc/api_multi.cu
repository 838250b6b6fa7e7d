// api_multi.cu — mb_find_multi: the multi-GPU path of SURVEY.md §8e driven entirely from C++, inside ONE process.
//
// The torchrun driver (mauvealigner_b200/dist.py) runs one process per GPU and leaves the exchanges to NCCL.  A C++
// host that owns several GPUs itself (the reference's applications are single-process) needs neither: this entry point
// takes one context per GPU, starts one host thread per context, runs the same mb_dist_* stages, and moves every
// exchange device to device over NVLink peer access —
//   exchange 1   fused into the partition kernel (peer stores into the destinations' receive arrays),
//   exchanges 2, 2b, 3   packed locally, then pushed into the peers' receive buffers by the copy engines
//                (cudaMemcpyAsync device to device),
//   the 4096-bin key histogram   summed on the host (32 KB per rank).
// The ranks meet at host barriers: a rank's stream is synchronised before the barrier that publishes its stores.
// Contexts may also share a device (that is how the parity tests run world sizes 2..8 on one GPU).
// The result is the same as mb_find_device on one context, bit for bit, split into the ranks' ascending ranges of the
// canonical order: mb_fetch_result per context, pieces concatenated in rank order.
#include "ctx.h"

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

// Barrier of the rank threads that also makes the "has any rank failed" decision COLLECTIVE: the last arriver samples the
// error word once, under the lock, and every rank leaves the barrier with that same value.  (Reading the word after the
// barrier is a race: a fast rank can run its next step, fail and set the word before a slow rank has looked — the slow
// rank then returns at barrier k while the fast ones wait at barrier k + 1 for ever.)  The sampled value of a generation
// cannot be overwritten before every rank has read it: the next sample is taken when all ranks have arrived again.
class HostBarrier {
public:
    explicit HostBarrier(int n) : n_(n) {}
    int wait(const std::atomic<int>& failed) {
        std::unique_lock<std::mutex> lk(m_);
        const unsigned gen = gen_;
        if (++count_ == n_) { count_ = 0; decision_ = failed.load(); ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen_ != gen; });
        return decision_;
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0, decision_ = MB_OK;
    unsigned gen_ = 0;
};

struct Shared {
    int world;
    const mb_params* prm;
    mb_ctx* const* ctxs;
    HostBarrier bar;
    std::atomic<int> failed{MB_OK};
    std::vector<uint64_t> M1, M2, M3h, M3c;          // count matrices [src][dst]
    std::vector<void*> seed_ptr, hdr_ptr, comp_ptr, acc_ptr; // every rank's receive buffers
    std::vector<uint64_t> hist;                       // [rank][4096]
    Shared(int w, const mb_params* p, mb_ctx* const* c)
        : world(w), prm(p), ctxs(c), bar(w), M1((size_t)w * w), M2((size_t)w * w), M3h((size_t)w * w), M3c((size_t)w * w), seed_ptr(w), hdr_ptr(w),
          comp_ptr(w), acc_ptr(w), hist((size_t)w * 4096) {}
};

// bytes-granular push of a local send buffer's destination blocks into the peers' buffers (copy engines)
int push_blocks(mb_ctx* c, const void* src, const uint64_t* counts, size_t unit_bytes, void* const* bases, const uint64_t* dst_offsets, int world, int rank) {
    std::vector<size_t> so(world + 1, 0);
    for (int d = 0; d < world; ++d) so[d + 1] = so[d] + (size_t)counts[d] * unit_bytes;
    for (int k = 0; k < world; ++k) {
        const int d = (rank + 1 + k) % world;
        if (counts[d] == 0) continue;
        CUDA_TRY(c, cudaMemcpyAsync((char*)bases[d] + (size_t)dst_offsets[d] * unit_bytes, (const char*)src + so[d], (size_t)counts[d] * unit_bytes,
                                    cudaMemcpyDeviceToDevice, c->stream));
    }
    return MB_OK;
}

// every step of a rank: run it unless some rank has already failed, record the first error, always reach the barrier
#define STEP(...) do { if (S.failed.load() == MB_OK) { int _rc = (__VA_ARGS__); if (_rc != MB_OK) { int ok = MB_OK; S.failed.compare_exchange_strong(ok, _rc); } } } while (0)
#define MEET() do { if (S.bar.wait(S.failed) != MB_OK) return; } while (0)

void rank_main(Shared& S, int r) {
    const int W = S.world;
    mb_ctx* c = S.ctxs[r];
    auto col = [&](const std::vector<uint64_t>& M, int d) { uint64_t t = 0; for (int s = 0; s < W; ++s) t += M[(size_t)s * W + d]; return t; };
    auto before = [&](const std::vector<uint64_t>& M, int d) { uint64_t t = 0; for (int s = 0; s < r; ++s) t += M[(size_t)s * W + d]; return t; };
    auto sync = [&]() -> int { CUDA_TRY(c, cudaSetDevice(c->device)); CUDA_TRY(c, cudaStreamSynchronize(c->stream)); return MB_OK; };
    std::vector<uint64_t> offs(W), offs2(W);

    // peer access towards every other device (contexts may share a device)
    STEP([&]() -> int {
        CUDA_TRY(c, cudaSetDevice(c->device));
        for (int d = 0; d < W; ++d) {
            const int pd = S.ctxs[d]->device;
            if (pd == c->device) continue;
            int can = 0;
            CUDA_TRY(c, cudaDeviceCanAccessPeer(&can, c->device, pd));
            if (!can) { snprintf(c->err, sizeof(c->err), "device %d cannot access device %d (no peer path)", c->device, pd); return MB_E_CUDA; }
            cudaError_t e = cudaDeviceEnablePeerAccess(pd, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { c->set_cuda_error(e, "cudaDeviceEnablePeerAccess", __LINE__); return MB_E_CUDA; }
        }
        return MB_OK;
    }());
    // ---- stage 1 + exchange 1 (fused)
    STEP(mb_dist_extract_count(c, r, W, &S.M1[(size_t)r * W]));
    MEET();
    const uint64_t n_recv = col(S.M1, r);
    STEP(mb_dist_p2p_recv_array(c, n_recv + n_recv / 8 + 4096, &S.seed_ptr[r]));
    MEET();
    for (int d = 0; d < W; ++d) offs[d] = before(S.M1, d);
    STEP(mb_dist_partition(c, S.seed_ptr.data(), offs.data(), nullptr));
    STEP(sync());
    MEET();
    // ---- stage 2 + exchange 2
    STEP(mb_dist_use_p2p_recv(c, 1));
    STEP(mb_dist_local(c, S.prm, n_recv, &S.M2[(size_t)r * W]));
    MEET();
    const uint64_t n_rows = col(S.M2, r);
    STEP(mb_dist_recv_buffer(c, 1, 4 * n_rows, &S.hdr_ptr[r]));
    uint64_t n_sent = 0;
    for (int d = 0; d < W; ++d) n_sent += S.M2[(size_t)r * W + d];
    STEP(mb_dist_recv_buffer(c, 5, n_sent, &S.acc_ptr[r]));
    MEET();
    {
        void* src = nullptr;
        STEP(mb_dist_rows_pack(c, nullptr, nullptr, &src));
        for (int d = 0; d < W; ++d) offs[d] = before(S.M2, d);
        STEP(push_blocks(c, src, &S.M2[(size_t)r * W], 32, S.hdr_ptr.data(), offs.data(), W, r));
        STEP(sync());
    }
    MEET();
    // ---- stage 3a + exchange 2b: verdict bytes back to the rows' sources, in row order
    {
        void* verdict = nullptr;
        STEP(mb_dist_resolve(c, n_rows, &verdict));
        std::vector<uint64_t> cnt(W);
        for (int s = 0; s < W; ++s) {
            cnt[s] = S.M2[(size_t)s * W + r];  // rows that came from source s
            uint64_t o = 0;                     // this owner's block in source s's verdict array: after the lower owners'
            for (int q = 0; q < r; ++q) o += S.M2[(size_t)s * W + q];
            offs[s] = o;
        }
        STEP(push_blocks(c, verdict, cnt.data(), 1, S.acc_ptr.data(), offs.data(), W, r));
        STEP(sync());
    }
    MEET();
    // ---- stage 3b: accepted candidates -> key histogram, summed over the ranks on the host
    {
        void* d_hist = nullptr;
        STEP(mb_dist_accept(c, &d_hist));
        STEP([&]() -> int {
            CUDA_TRY(c, cudaMemcpyAsync(&S.hist[(size_t)r * 4096], d_hist, 4096 * 8, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            return MB_OK;
        }());
        MEET();
        STEP([&]() -> int {
            std::vector<uint64_t> sum(4096, 0);
            for (int s = 0; s < W; ++s)
                for (int b = 0; b < 4096; ++b) sum[b] += S.hist[(size_t)s * 4096 + b];
            CUDA_TRY(c, cudaMemcpyAsync(d_hist, sum.data(), 4096 * 8, cudaMemcpyHostToDevice, c->stream)); // pageable: staged before return
            return MB_OK;
        }());
    }
    // ---- stage 3c + exchange 3
    STEP(mb_dist_match_partition(c, &S.M3h[(size_t)r * W], &S.M3c[(size_t)r * W]));
    MEET();
    const uint64_t n_match = col(S.M3h, r), n_mcomp = col(S.M3c, r);
    STEP(mb_dist_recv_buffer(c, 3, 2 * n_match, &S.hdr_ptr[r]));
    STEP(mb_dist_recv_buffer(c, 4, n_mcomp, &S.comp_ptr[r]));
    MEET();
    {
        void *hsrc = nullptr, *csrc = nullptr;
        STEP(mb_dist_match_pack(c, nullptr, nullptr, nullptr, nullptr, &hsrc, &csrc));
        for (int d = 0; d < W; ++d) { offs[d] = before(S.M3h, d); offs2[d] = before(S.M3c, d); }
        STEP(push_blocks(c, hsrc, &S.M3h[(size_t)r * W], 16, S.hdr_ptr.data(), offs.data(), W, r));
        STEP(push_blocks(c, csrc, &S.M3c[(size_t)r * W], 8, S.comp_ptr.data(), offs2.data(), W, r));
        STEP(sync());
    }
    MEET();
    // ---- stage 4
    STEP(mb_dist_output(c, n_match, n_mcomp));
    STEP(sync());
    S.bar.wait(S.failed);
}

} // namespace

extern "C" int mb_find_multi(mb_ctx* const* ctxs, int world, const mb_params* prm) {
    if (!ctxs || !prm || world < 1 || world > 256) return MB_E_ARG;
    if (prm->mode != MB_MODE_UNIQUE) return MB_E_ARG;
    for (int r = 0; r < world; ++r) {
        if (!ctxs[r]) return MB_E_ARG;
        for (int q = 0; q < r; ++q)
            if (ctxs[q] == ctxs[r]) return MB_E_ARG;
    }
    Shared S(world, prm, ctxs);
    std::vector<std::thread> threads;
    threads.reserve(world);
    for (int r = 1; r < world; ++r) threads.emplace_back(rank_main, std::ref(S), r);
    rank_main(S, 0);
    for (auto& t : threads) t.join();
    return S.failed.load();
}

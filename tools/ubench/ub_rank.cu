// ub_rank.cu — micro-benchmark of the warp-level digit-ranking primitives considered for k_onesweep (sm_100a).
// Not part of the product; run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ub_rank ub_rank.cu && ./ub_rank
// Every variant ranks ROUNDS pseudo-random 8-bit digits per lane (stable rank of the record among the warp's records of the
// same digit) and reports SM cycles per record (all SMs busy, 2 CTAs x 512 threads per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned int u32;
typedef unsigned long long u64;
#define NT 512
#define NW (NT / 32)
#define ROUNDS 12
#define ITERS 64

__device__ __forceinline__ u32 rnd(u32& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

template <int V>
__global__ void __launch_bounds__(NT, 2) k(u32* out, u32 dmask) {
    __shared__ u64 sMH[NW][256];                                   // 32 KB, viewed either as {mask,count} entries or as two u32 tables
    u32 (*sH)[256] = reinterpret_cast<u32 (*)[256]>(&sMH[0][0]);
    u32 (*sM)[256] = reinterpret_cast<u32 (*)[256]>(&sMH[NW / 2][0]);
    __shared__ unsigned char sC[NW][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NW * 256; i += NT) (&sMH[0][0])[i] = 0;
    __syncthreads();
    u32 s = (blockIdx.x * NT + tid) * 2654435761u + 12345u;
    u32 acc = 0;
    const u32 lt = (1u << lane) - 1;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < ROUNDS; ++k) {
            const u32 d = (rnd(s) >> 7) & dmask;
            if (V == 0) { // 8 ballots + running counter (the r01 kernel)
                u32 pm = 0xFFFFFFFFu;
#pragma unroll
                for (int bit = 0; bit < 8; ++bit) { const bool b = (d >> bit) & 1u; pm &= __ballot_sync(0xFFFFFFFFu, b) ^ (b ? 0u : 0xFFFFFFFFu); }
                const u32 old = sH[warp][d];
                __syncwarp();
                const u32 below = pm & lt;
                if (below == 0) sH[warp][d] = old + __popc(pm);
                __syncwarp();
                acc += old + __popc(below);
            } else if (V == 1) { // match.any
                const u32 pm = __match_any_sync(0xFFFFFFFFu, d);
                const u32 old = sH[warp][d];
                __syncwarp();
                const u32 below = pm & lt;
                if (below == 0) sH[warp][d] = old + __popc(pm);
                __syncwarp();
                acc += old + __popc(below);
            } else if (V == 2) { // atomicOr match, separate mask / counter tables
                atomicOr(&sM[warp][d], 1u << lane);
                __syncwarp();
                const u32 pm = sM[warp][d];
                const u32 old = sH[warp][d];
                __syncwarp();
                const u32 below = pm & lt;
                if (below == 0) { sH[warp][d] = old + __popc(pm); sM[warp][d] = 0; }
                __syncwarp();
                acc += old + __popc(below);
            } else if (V == 3) { // atomicOr match, one 64-bit entry {mask, count}
                atomicOr(reinterpret_cast<u32*>(&sMH[warp][d]), 1u << lane);
                __syncwarp();
                const u64 e = sMH[warp][d];
                const u32 pm = (u32)e, old = (u32)(e >> 32);
                __syncwarp();
                const u32 below = pm & lt;
                if (below == 0) sMH[warp][d] = (u64)(old + __popc(pm)) << 32;
                __syncwarp();
                acc += old + __popc(below);
            } else if (V == 4) { // claim table + ballots over the colliding groups
                sC[warp][d] = (unsigned char)lane;
                __syncwarp();
                const u32 w = sC[warp][d];
                u32 lost = __ballot_sync(0xFFFFFFFFu, w != (u32)lane);
                u32 below = 0, size = 1;
                while (lost) {
                    const int leader = __ffs(lost) - 1;
                    const u32 dl = __shfl_sync(0xFFFFFFFFu, d, leader);
                    const bool in = d == dl;
                    const u32 grp = __ballot_sync(0xFFFFFFFFu, in);
                    if (in) { below = __popc(grp & lt); size = __popc(grp); }
                    lost &= ~grp;
                }
                const u32 old = sH[warp][d];
                __syncwarp();
                if (below == 0) sH[warp][d] = old + size;
                __syncwarp();
                acc += old + below;
            } else if (V == 5) { // histogram only: shared atomicAdd without return
                atomicAdd(&sH[warp][d], 1u);
            } else if (V == 6) { // shared atomicAdd with return (unstable rank)
                acc += atomicAdd(&sH[warp][d], 1u);
            } else if (V == 7) { // plain LDS + STS at random digits (smem floor)
                const u32 old = sH[warp][d];
                sM[warp][d] = old + lane;
                acc += old;
            } else if (V == 8) { // atomicOr only
                atomicOr(&sM[warp][d], 1u << lane);
            } else if (V == 10) { // ballots, one LOP3 per bit: differ |= ballot ^ (bit ? ~0 : 0); peers = ~differ
                u32 differ = 0;
#pragma unroll
                for (int bit = 0; bit < 8; ++bit) {
                    const bool b = (d >> bit) & 1u;
                    const u32 x = __ballot_sync(0xFFFFFFFFu, b);
                    differ |= x ^ (b ? 0xFFFFFFFFu : 0u);
                }
                const u32 pm = ~differ;
                const u32 old = sH[warp][d];
                __syncwarp();
                const u32 below = pm & lt;
                if (below == 0) sH[warp][d] = old + __popc(pm);
                __syncwarp();
                acc += old + __popc(below);
            } else if (V == 11) { // ballots with sign-extended bit masks (bfe.s32)
                u32 differ = 0;
#pragma unroll
                for (int bit = 0; bit < 8; ++bit) {
                    int m;
                    asm("bfe.s32 %0, %1, %2, 1;" : "=r"(m) : "r"(d), "r"(bit));
                    const u32 x = __ballot_sync(0xFFFFFFFFu, m != 0);
                    differ |= x ^ (u32)m;
                }
                const u32 pm = ~differ;
                const u32 old = sH[warp][d];
                __syncwarp();
                const u32 below = pm & lt;
                if (below == 0) sH[warp][d] = old + __popc(pm);
                __syncwarp();
                acc += old + __popc(below);
            } else if (V == 12) { // two rounds of ballots, one of atomicOr ({mask,count} entries)
                if (k % 3 != 2) {
                    u32 differ = 0;
#pragma unroll
                    for (int bit = 0; bit < 8; ++bit) {
                        const bool b = (d >> bit) & 1u;
                        const u32 x = __ballot_sync(0xFFFFFFFFu, b);
                        differ |= x ^ (b ? 0xFFFFFFFFu : 0u);
                    }
                    const u32 pm = ~differ;
                    u32* e = reinterpret_cast<u32*>(&sMH[warp][d]);
                    const u32 old = e[1];
                    __syncwarp();
                    const u32 below = pm & lt;
                    if (below == 0) e[1] = old + __popc(pm);
                    __syncwarp();
                    acc += old + __popc(below);
                } else {
                    atomicOr(reinterpret_cast<u32*>(&sMH[warp][d]), 1u << lane);
                    __syncwarp();
                    const u64 e = sMH[warp][d];
                    const u32 pm = (u32)e, old = (u32)(e >> 32);
                    __syncwarp();
                    const u32 below = pm & lt;
                    if (below == 0) sMH[warp][d] = (u64)(old + __popc(pm)) << 32;
                    __syncwarp();
                    acc += old + __popc(below);
                }
            } else if (V == 9) { // block-wide table instead of per-warp (atomicAdd no return)
                atomicAdd(&sH[0][d], 1u);
            }
        }
    }
    if (V == 5 || V == 8 || V == 9) acc += sH[warp][lane] + sM[warp][lane];
    out[blockIdx.x * NT + tid] = acc + (u32)sMH[warp][lane] + sC[warp][lane];
}

template <int V>
void run(const char* name, u32* d_out, int sms, double mhz, u32 dmask) {
    const int grid = sms * 2;
    k<V><<<grid, NT>>>(d_out, dmask);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<V><<<grid, NT>>>(d_out, dmask);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double recs_per_sm = 2.0 * NT * ROUNDS * ITERS;
    const double cyc = ms * 1e-3 * mhz * 1e6;
    printf("%-64s dmask=%3u  %8.4f ms  %6.3f SM-cycles/record  %6.2f cycles/warp-round\n", name, dmask, ms, cyc / recs_per_sm, cyc / recs_per_sm * 32);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double mhz = khz / 1000.0;
    printf("%s: %d SMs, %.0f MHz (max)\n", p.name, sms, mhz);
    u32* d_out; cudaMalloc(&d_out, (size_t)sms * 2 * NT * 4);
    for (u32 dm : {255u, 15u, 0u}) {
        run<0>("0 ballot x8 + counter", d_out, sms, mhz, dm);
        run<1>("1 match.any + counter", d_out, sms, mhz, dm);
        run<2>("2 atomicOr mask table + counter table", d_out, sms, mhz, dm);
        run<3>("3 atomicOr on {mask,count} 64-bit entries", d_out, sms, mhz, dm);
        run<4>("4 claim table + ballots over colliding groups", d_out, sms, mhz, dm);
        run<5>("5 shared atomicAdd, no return (per-warp table)", d_out, sms, mhz, dm);
        run<6>("6 shared atomicAdd with return", d_out, sms, mhz, dm);
        run<7>("7 LDS + STS at the digit", d_out, sms, mhz, dm);
        run<8>("8 shared atomicOr, no return", d_out, sms, mhz, dm);
        run<9>("9 shared atomicAdd, no return (one table per CTA)", d_out, sms, mhz, dm);
        run<10>("10 ballot x8, differ |= x ^ m", d_out, sms, mhz, dm);
        run<11>("11 ballot x8, bfe.s32 masks", d_out, sms, mhz, dm);
        run<12>("12 two ballot rounds + one atomicOr round", d_out, sms, mhz, dm);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

// psa_sim.cu -- sequence_similarity for batches of pairs.
//
// Reference: sequence_similarity (test_functions/pull_data.cpp:97-127) counts the positions k below
// min(l1, l2) with s1[k] == s2[k] and divides by max(l1, l2).  (Its own chunking counts a few positions
// twice when min(l1,l2) is not a multiple of the thread count -- pull_data.cpp:106-113 -- that artefact
// is not reproduced; an empty pair gives 0 instead of 0/0.)
//
// A byte kernel bound by HBM: the threads of a pair read both sequences as aligned 32-bit words, adjacent
// lanes adjacent words, four words in flight per thread; B is brought to A's alignment by a funnel shift, equal
// bytes are counted with __vcmpeq4 + popc.  Algorithmic traffic: 2 bytes per compared position.
// Measured on B200 (tools/sim_bench.py): 1 M x 150 B pairs 0.10 ms = 2.9 TB/s; 64 x 4 MB pairs 0.11 ms =
// 4.7 TB/s = 0.73 of the measured HBM copy peak.
#include "psa_common.cuh"

namespace {

struct SimArgs {
    const uint8_t* bases_a; const int64_t* off_a; const int32_t* len_a;
    const uint8_t* bases_b; const int64_t* off_b; const int32_t* len_b;
    long long n_pairs;
    double* out;
};

// Equal bytes among positions [0, L) of one pair, the part of it given to `nthr` cooperating threads
// (this thread = tid) and to slice `split` of `nsplit` of the aligned body; slice 0 also takes the unaligned
// head and the tail.  A's words are aligned; B is realigned from two aligned words that both lie inside B's own bytes.
__device__ __forceinline__ int sim_count(const uint8_t* pa, const uint8_t* pb, int L, int tid, int nthr, int split, int nsplit) {
    int cnt = 0;
    const int head = min(L, (int)((4 - ((uintptr_t)pa & 3)) & 3));
    const int words = (L - head >= 8) ? (L - head - 4) / 4 : 0;
    if (split == 0) {
        if (tid < head) cnt += (pa[tid] == pb[tid]);
        for (int k = head + 4 * words + tid; k < L; k += nthr) cnt += (pa[k] == pb[k]);     // at most 7 bytes
    }
    if (words > 0) {
        const int w_lo = (int)((long long)words * split / nsplit), w_hi = (int)((long long)words * (split + 1) / nsplit);
        const uint32_t* wa = reinterpret_cast<const uint32_t*>(pa + head);
        const uintptr_t b0 = (uintptr_t)(pb + head);
        const uint32_t* wb = reinterpret_cast<const uint32_t*>(b0 & ~(uintptr_t)3);
        const unsigned sh = 8u * (unsigned)(b0 & 3);
        int w = w_lo + tid;
        for (; w + 3 * nthr < w_hi; w += 4 * nthr) {          // four independent words in flight per thread
            uint32_t x[4], y0[4], y1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { x[u] = wa[w + u * nthr]; y0[u] = wb[w + u * nthr]; y1[u] = wb[w + u * nthr + 1]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) cnt += __popc(__vcmpeq4(x[u], __funnelshift_r(y0[u], y1[u], sh))) >> 3;
        }
        for (; w < w_hi; w += nthr) cnt += __popc(__vcmpeq4(wa[w], __funnelshift_r(wb[w], wb[w + 1], sh))) >> 3;
    }
    return cnt;
}

// Short and medium pairs: TPP (8 or 32) lanes per pair, 256 / TPP pairs per CTA.
template <int TPP>
__global__ void __launch_bounds__(256) psa_similarity_kernel(SimArgs A) {
    constexpr int PPB = 256 / TPP;
    const int sub = threadIdx.x / TPP, tl = threadIdx.x % TPP;
    const long long rounds = (A.n_pairs + PPB - 1) / PPB;
    for (long long rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
        const long long p = rd * PPB + sub;
        const bool have = p < A.n_pairs;
        int la = 0, lb = 0, cnt = 0;
        if (have) {
            la = A.len_a[p]; lb = A.len_b[p];
            cnt = sim_count(A.bases_a + A.off_a[p], A.bases_b + A.off_b[p], min(la, lb), tl, TPP, 0, 1);
        }
#pragma unroll
        for (int o = TPP / 2; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (tl == 0 && have) A.out[p] = max(la, lb) > 0 ? (double)cnt / (double)max(la, lb) : 0.0;
    }
}

// Long pairs: every pair is cut into `nsplit` word ranges, one CTA each; the counts meet in the pair's
// output slot, used as a 64-bit counter (zeroed by the launcher) until psa_similarity_finish_kernel
// turns it into the ratio.
__global__ void __launch_bounds__(256) psa_similarity_split_kernel(SimArgs A, int nsplit) {
    __shared__ int warp_sum[8];
    const long long units = A.n_pairs * nsplit;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const long long p = u / nsplit;
        const int split = (int)(u % nsplit);
        const int L = min(A.len_a[p], A.len_b[p]);
        int cnt = sim_count(A.bases_a + A.off_a[p], A.bases_b + A.off_b[p], L, threadIdx.x, 256, split, nsplit);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) tot += warp_sum[k];
            if (tot) atomicAdd(reinterpret_cast<unsigned long long*>(A.out) + p, (unsigned long long)tot);
        }
        __syncthreads();
    }
}

__global__ void psa_similarity_finish_kernel(SimArgs A) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.n_pairs) return;
    const unsigned long long c = reinterpret_cast<const unsigned long long*>(A.out)[p];
    const int mx = max(A.len_a[p], A.len_b[p]);
    A.out[p] = mx > 0 ? (double)c / (double)mx : 0.0;
}

}  // namespace

int psa_launch_similarity(psa_ctx* ctx, const psa_batch_args& args, int max_len, double* d_out, cudaStream_t st) {
    SimArgs A{args.bases_a, args.off_a, args.len_a, args.bases_b, args.off_b, args.len_b, args.n_pairs, d_out};
    const long long cap = (long long)ctx->sm_count * 8;            // 2048 resident threads per SM
    if (max_len <= 8192) {
        if (max_len <= 512) {
            const long long blocks = (args.n_pairs + 31) / 32;
            psa_similarity_kernel<8><<<(int)std::max<long long>(1, std::min(blocks, cap)), 256, 0, st>>>(A);
        } else {
            const long long blocks = (args.n_pairs + 7) / 8;
            psa_similarity_kernel<32><<<(int)std::max<long long>(1, std::min(blocks, cap)), 256, 0, st>>>(A);
        }
        PSA_CUDA_OK(ctx, cudaGetLastError());
        ctx->launches += 1;
        return PSA_OK;
    }
    // at least 16 KB per CTA and split, and enough splits to fill the machine twice over
    long long nsplit = std::min<long long>((max_len + 16383) / 16384, std::max<long long>(1, 2 * cap / std::max<long long>(1, args.n_pairs)));
    if (nsplit < 1) nsplit = 1;
    PSA_CUDA_OK(ctx, cudaMemsetAsync(d_out, 0, (size_t)args.n_pairs * sizeof(double), st));
    psa_similarity_split_kernel<<<(int)std::min<long long>(args.n_pairs * nsplit, cap), 256, 0, st>>>(A, (int)nsplit);
    PSA_CUDA_OK(ctx, cudaGetLastError());
    psa_similarity_finish_kernel<<<(int)((args.n_pairs + 255) / 256), 256, 0, st>>>(A);
    PSA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 2;
    return PSA_OK;
}

// psa_long.cu -- intra-pair wavefront kernels for long pairs (BASELINE configs 3, 4, 5).
//
// The DP matrix is cut into row blocks of R rows and column strips of W = 32*K columns.  One warp
// owns a row block and sweeps its tiles left to right (psa_tile::sweep); the bottom boundary of
// every tile (H and F per column) goes to a global "hbuf", the right boundary stays in the warp's
// shared memory for the next tile.  Row block rb may start strip s once row block rb-1 has
// finished strip s: a per-row-block progress counter published with release/acquire semantics.
// Row blocks are handed out by an atomic ticket in increasing order, so every dependency points at
// a warp that already holds an earlier ticket -- forward progress never depends on co-residency.
// This replaces the reference's Subproblem::compute_tables (subproblem_alignment.cpp:329-355):
// per-row fork/join + ParallelPrefixMax become a two-level wavefront (lanes inside a tile, warps
// across tiles), and the three O(mn) double tables (subproblem_alignment.h:66-73) become O(m+n)
// boundaries.
//
// Traceback (config 3) never needs the full matrix: with checkpoints enabled the fill keeps every
// row-block bottom row and every strip right column (8*m*n*(1/R + 1/W) bytes); the traceback
// kernel then recomputes only the tiles the path crosses, with 4-bit codes in shared memory -- the
// role the reference intended for its partial-balanced-partition stage (sequence_alignment/
// partial.cpp:81-163, never wired up).
#include <climits>
#include <vector>

#include "psa_tile.cuh"

using namespace psa_tile;

namespace {

constexpr int K = 8;
constexpr int W = 32 * K;      // strip width
constexpr int R = 128;         // row-block height
constexpr int WPB = 4;         // warps per CTA

struct LongJob {
    const uint8_t* a;
    const uint8_t* b;
    int m, n;
    int g, h;
    int mul8;                   // = 8 at run time (keeps the local-mode key an IMAD)
    int* hbufH;                 // bottom boundaries; row-block stride hb_stride ints (0 = one recycled row)
    int* hbufF;
    long long hb_stride;
    int* ckvH;                  // right-boundary checkpoint columns [S][m+1] (null without checkpoints)
    int* ckvE;
    int* progress;              // [NB] strips finished per row block (null in batch mode: no waiting)
    int* ticket;
    unsigned long long* best;   // local: packed (score, end_i, end_j) key, atomicMax
    int* corner;                // global: T1,T2,T3 of (m,n)
    int col0, n_total;          // always 0 / n (kept so that the tile engine can address a column window)
    int start_type, end_type;   // Subproblem border variants (global mode); -1 / -1 = the live case
    int row_off;                // ROWOFF instantiation only: height of the FIRST row block (partition sweeps); else 0
};

__device__ __forceinline__ unsigned long long pack_best(int score, int i, int j) {
    return ((unsigned long long)(unsigned)score << 42) | ((unsigned long long)(0x1FFFFF - i) << 21) |
           (unsigned long long)(0x1FFFFF - j);
}

// Polling load: relaxed at gpu scope (served by L2, no L1 invalidate).  ld.acquire would emit a
// CCTL.IVALL per poll, and a few spinning warps then starve the shared-memory pipe of the working
// warps on the same SM (measured: 1.35e9 invalidates, 4x slower tiles).  One fence follows the
// successful poll instead.
__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int RR>
struct WarpSmem {
    int bH[2][RR];     // ping-pong boundary columns: [cur] = left boundary in, [cur^1] = right out
    int bE[2][RR];
    uint8_t sA[RR];
};

// One row block: rows rb*R+1 .. , all strips.  Warp-collective.
// RR rows per block, KK columns per lane (strip width 32*KK); checkpoints and the traceback grid need <128, 8>.
template <int MODE, int RR, int KK, bool ROWOFF = false>
__device__ void process_rowblock(const LongJob& J, int rb, WarpSmem<RR>& sm, Track& tr) {
    constexpr int WW = 32 * KK;
    const int lane = threadIdx.x & 31;
    const int m = J.m, n = J.n, g = J.g, h = J.h;
    // ROWOFF (partition sweeps only): the first row block is J.row_off rows high, the grid is shifted accordingly
    const int i0 = (!ROWOFF || rb == 0) ? rb * RR : J.row_off + (rb - 1) * RR;
    const int nrows = (ROWOFF && rb == 0) ? min(J.row_off, m) : min(RR, m - i0);
    const int S = (n + WW - 1) / WW;
    for (int r = lane; r < nrows; r += 32) sm.sA[r] = J.a[i0 + r];
    // left boundary of strip 0: column 0 of the matrix (subproblem_alignment.cpp:282-292)
    for (int r = lane; r < nrows; r += 32) { sm.bH[0][r] = border_col0_H<MODE>(i0 + 1 + r, g, h, J.start_type); sm.bE[0][r] = PSA_KNEG; }
    int corner = border_col0_H<MODE>(i0, g, h, J.start_type);           // H[i0][0]
    __syncwarp();
    int cur = 0;
    const int* topH = J.hbufH + (J.hb_stride ? (long long)(rb - 1) * J.hb_stride : 0);
    const int* topF = J.hbufF + (J.hb_stride ? (long long)(rb - 1) * J.hb_stride : 0);
    int* botH = J.hbufH + (J.hb_stride ? (long long)rb * J.hb_stride : 0);
    int* botF = J.hbufF + (J.hb_stride ? (long long)rb * J.hb_stride : 0);
    int cap1 = PSA_KNEG, cap2 = PSA_KNEG, cap3 = PSA_KNEG;
    bool captured = false;
    for (int s = 0; s < S; ++s) {
        if (rb > 0 && J.progress != nullptr) {
            // Every lane polls the same word (one broadcast request): a lane-0-only spin loop left
            // the warp split into two convergence groups for the rest of the row block -- each
            // step then ran twice and the shuffles took the slow collective path (3.3x slower).
            unsigned ns = 32;
            while (ld_relaxed(J.progress + rb - 1) <= s) { __nanosleep(ns); if (ns < 2048) ns <<= 1; }
            __threadfence();              // acquire side: order the boundary reads after the flag read
            __syncwarp();
        }
        const int c0 = s * WW + lane * KK;
        ColsS<KK> cs;
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            const int j = c0 + k + 1;
            cs.b[k] = (j <= n) ? (int)J.b[j - 1] : 256;
            int H0, F0 = PSA_KNEG;
            if (rb == 0) H0 = border_row0_H<MODE>(J.col0 + j, g, h, J.start_type);
            else if (j <= n) { H0 = __ldcg(topH + j); F0 = __ldcg(topF + j); }
            else H0 = (MODE == PSA_LOCAL) ? 0 : PSA_KNEG;
            cs.set_top(k, H0, F0, g + h);
            cs.set_key_addend(k, j <= n, g + h);
        }
        // H[i0][c0] for every lane: the top value of the previous lane's last column; lane 0: the corner
        const int hlast = cs.top_H(KK - 1, g + h);
        int hd = __shfl_up_sync(0xffffffffu, hlast, 1);
        if (lane == 0) hd = corner;
        const int next_corner = __shfl_sync(0xffffffffu, hlast, 31);   // H[i0][(s+1)*WW]
        const bool has_cell = (J.col0 + n == J.n_total) && (i0 + nrows == m) && (n > s * WW) && (n <= (s + 1) * WW);
        int bestkey = 0, besti = 0;
        sweep_score<KK, MODE>(cs, hd, sm.bH[cur], sm.bE[cur], sm.bH[cur ^ 1], sm.bE[cur ^ 1], sm.sA, nrows, i0, c0, m, n, g, h,
                             J.mul8, bestkey, besti, cap1, cap2, cap3);
        if (MODE == PSA_LOCAL) {          // fold this tile's best (T1, first row, first column) into the lane's tracker
            const int t1v = bestkey / key_mult(KK);
            const int bj = c0 + (key_mult(KK) - 1 - (bestkey % key_mult(KK))) + 1;
            if (t1v > 0 && bj <= n) {
                const bool better = t1v > tr.best || (t1v == tr.best && (besti < tr.bi || (besti == tr.bi && bj < tr.bj)));
                if (better) { tr.best = t1v; tr.bi = besti; tr.bj = bj; }     // bj is strip-local; col0 is added when packed
            }
        }
        if (has_cell) captured = true;
        // publish the bottom boundary
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            const int j = c0 + k + 1;
            if (j <= n) { botH[j] = cs.top_H(k, g + h); botF[j] = cs.ff[k]; }
        }
        __syncwarp();
        if (J.ckvH != nullptr) {          // checkpoint the right boundary column (col (s+1)*WW)
            int* vh = J.ckvH + (long long)s * (m + 1);
            int* ve = J.ckvE + (long long)s * (m + 1);
            for (int r = lane; r < nrows; r += 32) { vh[i0 + 1 + r] = sm.bH[cur ^ 1][r]; ve[i0 + 1 + r] = sm.bE[cur ^ 1][r]; }
        }
        if (J.progress != nullptr) {
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release(J.progress + rb, s + 1);
        }
        corner = next_corner;
        cur ^= 1;
        __syncwarp();
    }
    if (MODE == PSA_GLOBAL && captured) {
        // exactly one lane of one tile holds cell (m, n)
        const int src = ((n - 1) % WW) / KK;
        if (lane == src) { J.corner[0] = cap1; J.corner[1] = cap2; J.corner[2] = cap3; }
    }
}

template <int MODE>
__device__ void flush_track(const LongJob& J, Track& tr) {
    if (MODE != PSA_LOCAL) return;
    unsigned long long key = tr.best > 0 ? pack_best(tr.best, tr.bi, J.col0 + tr.bj) : 0ull;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
        key = o > key ? o : key;
    }
    if ((threadIdx.x & 31) == 0 && key != 0ull) atomicMax(J.best, key);
}

// Single long pair: every warp of the grid pulls row blocks of the same job.
template <int MODE, int RR, int KK, bool ROWOFF = false>
__global__ void __launch_bounds__(WPB * 32) psa_long_single_kernel(LongJob J) {
    __shared__ WarpSmem<RR> smem[WPB];
    WarpSmem<RR>& sm = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int NB = ROWOFF ? 1 + max(0, J.m - J.row_off + RR - 1) / RR : (J.m + RR - 1) / RR;
    Track tr{0, 0, 0};
    for (;;) {
        int rb = 0;
        if (lane == 0) rb = atomicAdd(J.ticket, 1);
        rb = __shfl_sync(0xffffffffu, rb, 0);
        if (rb >= NB) break;
        process_rowblock<MODE, RR, KK, ROWOFF>(J, rb, sm, tr);
    }
    flush_track<MODE>(J, tr);
}

// Batch of long pairs, score only: one warp per pair (pairs are independent -- no flags).
struct LongBatch {
    psa_batch_args P;
    int* hbuf;                 // per resident warp: 2*(max_n+1) ints
    long long hbuf_warp_stride;
    int* ticket;
    const uint8_t* only_flagged;   // optional: one byte per pair, process only non-zero entries
    int mul8;
};

template <int MODE, int KK>      // KK columns per lane: 16 from 1 kbp (pairs are independent, wide lanes only amortise overhead), else 8
__global__ void __launch_bounds__(WPB * 32) psa_long_batch_kernel(LongBatch Bt) {
    __shared__ WarpSmem<R> smem[WPB];
    __shared__ unsigned long long s_best[WPB];
    __shared__ int s_corner[WPB][4];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem<R>& sm = smem[w];
    const long long gw = (long long)blockIdx.x * WPB + w;
    for (;;) {
        long long p = 0;
        if (lane == 0) p = atomicAdd(Bt.ticket, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= Bt.P.n_pairs) break;
        if (Bt.only_flagged != nullptr && Bt.only_flagged[p] == 0) continue;
        LongJob J;
        J.a = Bt.P.bases_a + Bt.P.off_a[p]; J.b = Bt.P.bases_b + Bt.P.off_b[p];
        J.m = Bt.P.len_a[p]; J.n = Bt.P.len_b[p]; J.g = Bt.P.g; J.h = Bt.P.h; J.mul8 = Bt.mul8;
        J.hbufH = Bt.hbuf + gw * Bt.hbuf_warp_stride; J.hbufF = J.hbufH + Bt.hbuf_warp_stride / 2; J.hb_stride = 0;
        J.ckvH = nullptr; J.ckvE = nullptr; J.progress = nullptr; J.ticket = nullptr;
        J.best = &s_best[w]; J.corner = s_corner[w];
        J.col0 = 0; J.n_total = J.n; J.row_off = 0;
        J.start_type = -1; J.end_type = -1;
        if (lane == 0) { s_best[w] = 0ull; s_corner[w][0] = s_corner[w][1] = s_corner[w][2] = PSA_KNEG; }
        __syncwarp();
        psa_batch_item r;
        r.start_i = 0; r.start_j = 0; r.aln_len = 0;
        if (J.m > 0 && J.n > 0) {
            Track tr{0, 0, 0};
            const int NB = (J.m + R - 1) / R;
            for (int rb = 0; rb < NB; ++rb) process_rowblock<MODE, R, KK>(J, rb, sm, tr);
            flush_track<MODE>(J, tr);
            __syncwarp();
        }
        if (lane == 0) {
            const int m = J.m, n = J.n, g = J.g, h = J.h;
            if (MODE == PSA_LOCAL) {
                const unsigned long long key = s_best[w];
                const int sc = (int)(key >> 42);
                r.t1 = sc; r.t2 = PSA_NEG_INF; r.t3 = PSA_NEG_INF; r.score = sc; r.end_state = 1;
                r.end_i = sc > 0 ? 0x1FFFFF - (int)((key >> 21) & 0x1FFFFF) : 0;
                r.end_j = sc > 0 ? 0x1FFFFF - (int)(key & 0x1FFFFF) : 0;
            } else {
                int c1, c2, c3;
                if (m > 0 && n > 0) { c1 = s_corner[w][0]; c2 = s_corner[w][1]; c3 = s_corner[w][2]; }
                else {
                    c1 = (m == 0 && n == 0) ? 0 : PSA_NEG_INF;
                    c2 = (m == 0 && n > 0) ? -h - g * n : PSA_NEG_INF;
                    c3 = (n == 0 && m > 0) ? -h - g * m : PSA_NEG_INF;
                }
                r.t1 = c1; r.t2 = c2; r.t3 = c3; r.score = imax(c1, imax(c2, c3));
                r.end_state = (c1 >= c2 && c1 >= c3) ? 1 : ((c2 >= c1 && c2 >= c3) ? 2 : 3);
                r.end_i = m; r.end_j = n;
            }
            Bt.P.items[p] = r;
        }
        __syncwarp();
    }
}

// ---- checkpointed traceback: one warp recomputes only the tiles the path crosses ----------
struct TbArgs {
    LongJob J;
    psa_batch_item* item;      // device result
    uint32_t* ops;             // 2-bit ops, traceback order
    uint32_t* band;            // optional: direction codes of BAND_TILES tiles per row block, recomputed in parallel
};

constexpr int BAND_TILES = 3;

// End cell and end state of the traceback (find_alignment, subproblem_alignment.cpp:112-145).
template <int MODE>
__device__ __forceinline__ void tb_end(const LongJob& J, psa_batch_item& r, int& state) {
    if (MODE == PSA_LOCAL) {
        const unsigned long long key = *J.best;
        const int sc = (int)(key >> 42);
        r.t1 = sc; r.t2 = PSA_NEG_INF; r.t3 = PSA_NEG_INF; r.score = sc;
        r.end_i = sc > 0 ? 0x1FFFFF - (int)((key >> 21) & 0x1FFFFF) : 0;
        r.end_j = sc > 0 ? 0x1FFFFF - (int)(key & 0x1FFFFF) : 0;
        state = 1;
    } else {
        const int c1 = J.corner[0], c2 = J.corner[1], c3 = J.corner[2];
        auto outv = [](int v) { return v < PSA_KNEG / 2 ? PSA_NEG_INF : v; };      // sentinel drift -> the ABI's -inf
        r.t1 = outv(c1); r.t2 = outv(c2); r.t3 = outv(c3); r.score = outv(imax(c1, imax(c2, c3)));
        // a positive end type forces the state; -2 / -3 credit h to T2 / T3 in the pick (cpp:112-146, h_prime)
        if (J.end_type > 0) state = J.end_type;
        else {
            const int e2 = c2 + (J.end_type == -2 ? J.h : 0), e3 = c3 + (J.end_type == -3 ? J.h : 0);
            state = (c1 >= e2 && c1 >= e3) ? 1 : ((e2 >= c1 && e2 >= e3) ? 2 : 3);
        }
        r.end_i = J.m; r.end_j = J.n_total;
    }
    r.end_state = state;
    r.start_i = 0; r.start_j = 0; r.aln_len = 0;
}

// First strip of the band in row block rb: the tiles around the straight line the path is expected to
// follow -- the corner-to-corner diagonal for a global alignment, the slope-1 diagonal through the end
// cell for a local one.  A path that leaves the band only costs the on-demand recompute it always had.
template <int MODE>
__device__ __forceinline__ int band_first_strip(const LongJob& J, int end_i, int end_j, int rb) {
    const int S = (J.n + W - 1) / W;
    const long long mid_i = (long long)rb * R + R / 2;
    long long cj = (MODE == PSA_LOCAL) ? (long long)end_j - ((long long)end_i - mid_i)
                                       : (long long)J.n * mid_i / (J.m > 0 ? J.m : 1);
    cj = cj < 0 ? 0 : (cj > J.n - 1 ? J.n - 1 : cj);
    int first = (int)(cj / W) - 1;
    if (first > S - BAND_TILES) first = S - BAND_TILES;
    return first < 0 ? 0 : first;
}

// Direction codes of rows i0+1 .. i0+nrows of tile (rb, s), recomputed from the checkpointed boundaries
// (row-block bottom rows in hbuf, strip right columns in ckv) into `dirs` (32 words per row).
template <int MODE>
__device__ __forceinline__ void recompute_tile(const LongJob& J, int rb, int s, int nrows, uint32_t* dirs, int* lbH, int* lbE,
                                               uint8_t* sA) {
    const int lane = threadIdx.x & 31;
    const int m = J.m, n = J.n, g = J.g, h = J.h;
    const int i0 = rb * R, c0 = s * W + lane * K;
    __syncwarp();
    for (int q = lane; q < nrows; q += 32) {
        sA[q] = J.a[i0 + q];
        if (s == 0) { lbH[q] = border_col0_H<MODE>(i0 + 1 + q, g, h, J.start_type); lbE[q] = PSA_KNEG; }
        else { lbH[q] = J.ckvH[(long long)(s - 1) * (m + 1) + i0 + 1 + q]; lbE[q] = J.ckvE[(long long)(s - 1) * (m + 1) + i0 + 1 + q]; }
    }
    Cols<K> cs;
    const int* topH = J.hbufH + (long long)(rb - 1) * J.hb_stride;
    const int* topF = J.hbufF + (long long)(rb - 1) * J.hb_stride;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int jj = c0 + k + 1;
        cs.b[k] = (jj <= n) ? (int)J.b[jj - 1] : 256;
        if (rb == 0) { cs.H[k] = border_row0_H<MODE>(jj, g, h, J.start_type); cs.F[k] = PSA_KNEG; }
        else if (jj <= n) { cs.H[k] = topH[jj]; cs.F[k] = topF[jj]; }
        else { cs.H[k] = PSA_KNEG; cs.F[k] = PSA_KNEG; }
    }
    int hd = __shfl_up_sync(0xffffffffu, cs.H[K - 1], 1);
    if (lane == 0) hd = (s == 0) ? border_col0_H<MODE>(i0, g, h, J.start_type) : (rb == 0 ? border_row0_H<MODE>(s * W, g, h, J.start_type) : topH[s * W]);
    __syncwarp();
    Track tr{0, 0, 0};
    int d1 = 0, d2 = 0, d3 = 0;
    sweep<K, MODE, true>(cs, hd, lbH, lbE, nullptr, nullptr, sA, nrows, i0, c0, m, n, g, h, dirs, tr, d1, d2, d3);
    __syncwarp();
}

// Every tile is a function of its checkpointed boundaries alone, so the tiles the path will most likely
// cross are recomputed by independent warps all over the GPU before the (inherently serial) walk starts;
// the walker then only copies 16 KB per tile into its shared memory instead of sweeping it.
constexpr int BAND_WPB = 2;
template <int MODE>
__global__ void __launch_bounds__(BAND_WPB * 32) psa_long_band_kernel(TbArgs T) {
    __shared__ uint32_t dirs[BAND_WPB][R * 32];
    __shared__ int lbH[BAND_WPB][R], lbE[BAND_WPB][R];
    __shared__ uint8_t sA[BAND_WPB][R];
    const LongJob& J = T.J;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NB = (J.m + R - 1) / R, S = (J.n + W - 1) / W;
    const int u = blockIdx.x * BAND_WPB + w;
    if (u >= NB * BAND_TILES) return;
    const int rb = u / BAND_TILES, bt = u % BAND_TILES;
    psa_batch_item r;
    int state;
    tb_end<MODE>(J, r, state);
    if (rb * R >= r.end_i) return;                       // rows below the end cell are never visited
    const int s = band_first_strip<MODE>(J, r.end_i, r.end_j, rb) + bt;
    if (s >= S) return;
    const int nrows = min(R, J.m - rb * R);
    recompute_tile<MODE>(J, rb, s, nrows, dirs[w], lbH[w], lbE[w], sA[w]);
    uint4* out = reinterpret_cast<uint4*>(T.band + (long long)u * (R * 32));
    const uint4* in = reinterpret_cast<const uint4*>(dirs[w]);
    for (int q = lane; q < nrows * 8; q += 32) out[q] = in[q];
}

template <int MODE>
__global__ void __launch_bounds__(32) psa_long_tb_kernel(TbArgs T) {
    __shared__ uint32_t dirs[R * 32];
    __shared__ int lbH[R], lbE[R];
    __shared__ uint8_t sA[R];
    const LongJob& J = T.J;
    const int lane = threadIdx.x;
    const int m = J.m;
    psa_batch_item r;
    int i, j, state;
    tb_end<MODE>(J, r, state);
    i = r.end_i; j = r.end_j;
    if (T.ops == nullptr) { if (lane == 0) *T.item = r; return; }

    int trb = -1, ts = -1;      // tile currently held in `dirs`
    int len = 0;
    int hits = 0, misses = 0;
    uint32_t acc = 0;
    bool done = !(i > 0 && j > 0);
    while (!done) {
        // lane 0 walks as far as the loaded tile allows
        int need_rb = -1, need_s = -1;
        if (lane == 0) {
            while (i > 0 && j > 0) {
                const int si = (state == 2) ? i : i - 1;
                const int sj = (state == 3) ? j : j - 1;
                const bool border = (si == 0 || sj == 0);
                int code = 0;
                if (!border) {
                    const int rb = (si - 1) / R, s = (sj - 1) / W;
                    if (rb != trb || s != ts) { need_rb = rb; need_s = s; break; }
                    const int cj = sj - 1 - s * W;
                    code = (dirs[(si - 1 - rb * R) * 32 + cj / K] >> (4 * (cj % K))) & 15;
                }
                acc |= (uint32_t)state << (2 * (len & 15));
                if ((len & 15) == 15) { T.ops[len >> 4] = acc; acc = 0; }
                ++len;
                r.start_i = i; r.start_j = j;
                const int ns = next_state<MODE>(state, code, border);
                if (ns == 0) { i = 0; break; }     // local alignment starts here
                state = ns; i = si; j = sj;
            }
        }
        need_rb = __shfl_sync(0xffffffffu, need_rb, 0);
        need_s = __shfl_sync(0xffffffffu, need_s, 0);
        if (need_rb < 0) { done = true; break; }
        const int rb = need_rb, s = need_s;
        const int need_row = __shfl_sync(0xffffffffu, (lane == 0) ? ((state == 2) ? i : i - 1) : 0, 0);   // source row of the pending step
        const int i0 = rb * R, nrows = min(min(R, m - i0), need_row - i0);      // rows below the entry row are never consulted
        const int bt = (T.band != nullptr) ? s - band_first_strip<MODE>(J, r.end_i, r.end_j, rb) : -1;
        if (bt >= 0 && bt < BAND_TILES) {
            // ---- the tile was recomputed ahead of time by psa_long_band_kernel: copy its codes in ----
            __syncwarp();
            const uint4* in = reinterpret_cast<const uint4*>(T.band + ((long long)rb * BAND_TILES + bt) * (R * 32));
            uint4* out = reinterpret_cast<uint4*>(dirs);
            const int total = nrows * 8;                       // 16-byte pieces; eight loads in flight per lane
            for (int q0 = 0; q0 < total; q0 += 256) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int q = q0 + u * 32 + lane; if (q < total) v[u] = __ldcg(in + q); }
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int q = q0 + u * 32 + lane; if (q < total) out[q] = v[u]; }
            }
            __syncwarp();
            if (lane == 0) ++hits;
        } else {
            // ---- recompute tile (rb, s) from its checkpointed boundaries ----
            recompute_tile<MODE>(J, rb, s, nrows, dirs, lbH, lbE, sA);
            if (lane == 0) ++misses;
        }
        trb = rb; ts = s;
    }
#ifdef PSA_TB_DEBUG
    if (lane == 0) printf("psa_long_tb_kernel: %d band tiles, %d recomputed tiles, %d ops\n", hits, misses, len);
#endif
    if (lane == 0) {
        if (len & 15) T.ops[len >> 4] = acc;
        r.aln_len = len;
        *T.item = r;
    }
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------
namespace {
int ensure_work(psa_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->d_work_bytes) return PSA_OK;
    if (ctx->d_work) cudaFree(ctx->d_work);
    ctx->d_work = nullptr; ctx->d_work_bytes = 0;
    bytes = ((bytes + bytes / 8) + (1 << 20) - 1) / (1 << 20) * (size_t)(1 << 20);
    if (cudaMalloc(&ctx->d_work, bytes) != cudaSuccess) {
        cudaGetLastError();
        return psa_fail(ctx, PSA_ERR_NOMEM, "cudaMalloc of long-pair work buffers failed (" + std::to_string(bytes >> 20) + " MiB)");
    }
    ctx->d_work_bytes = bytes;
    return PSA_OK;
}
size_t up256(size_t x) { return (x + 255) / 256 * 256; }
}  // namespace

// One long pair, sequences already on the device.  Writes *d_item (and d_ops when traceback).
int psa_launch_long_single(psa_ctx* ctx, const uint8_t* d_a, const uint8_t* d_b, int m, int n, int mode, int g, int h,
                           bool traceback, psa_batch_item* d_item, uint32_t* d_ops, cudaStream_t st,
                           int start_type, int end_type) {
    if (m <= 0 || n <= 0) return psa_fail(ctx, PSA_ERR_ARG, "long path needs m, n >= 1");
    if (m >= 0x1FFFFF || n >= 0x1FFFFF) return psa_fail(ctx, PSA_ERR_RANGE, "long path: lengths must be < 2^21 - 1");
    // One pair that is wide enough to fill the GPU with strips, score only: the column-stationary systolic kernel
    // (psa_systolic.cu) -- no per-tile drains, every resident warp busy for all m rows.
    {
        const bool plain = (start_type == -1 && end_type == -1 && !traceback);
        // on ONE GPU the row-block tiles below are faster (wider lanes, 12 instead of ~25 instructions per cell); the
        // systolic kernel is what several GPUs share one pair with (psa_align_long_cyclic_device) and an option here
        const bool want = ctx->opt.long_systolic == 1;
        if (plain && want) return psa_launch_systolic(ctx, d_a, d_b, m, n, mode, g, h, 0, 1, 0, nullptr, nullptr, 0, d_item, st);
    }
    const int NB = (m + R - 1) / R, S = (n + W - 1) / W;
    const size_t row = up256((size_t)(n + 1) * 4);
    const size_t hb = traceback ? row * NB : row;
    const size_t ckv = traceback ? up256((size_t)S * (m + 1) * 4) : 0;
    size_t o = 0;
    const size_t o_hH = o; o += hb;
    const size_t o_hF = o; o += hb;
    const size_t o_vH = o; o += ckv;
    const size_t o_vE = o; o += ckv;
    const size_t o_pr = o; o += up256((size_t)NB * 4);
    const size_t o_misc = o; o += 256;      // ticket(4) | pad | best(8 @ +8) | corner(3*4 @ +16)
    // direction codes of the band tiles (16 KB each), recomputed in parallel before the walk
    const size_t band_bytes = (size_t)NB * BAND_TILES * R * 32 * 4;
    const bool use_band = traceback && band_bytes <= ((size_t)1 << 30) && ctx->opt.long_band;
    const size_t o_band = o; o += use_band ? band_bytes : 0;
    int rc = ensure_work(ctx, o);
    if (rc) return rc;
    uint8_t* d = (uint8_t*)ctx->d_work;
    PSA_CUDA_OK(ctx, cudaMemsetAsync(d + o_pr, 0, (o_misc + 256) - o_pr, st));
    LongJob J;
    J.a = d_a; J.b = d_b; J.m = m; J.n = n; J.g = g; J.h = h; J.mul8 = 8;
    J.hbufH = (int*)(d + o_hH); J.hbufF = (int*)(d + o_hF);
    J.hb_stride = traceback ? (long long)(row / 4) : 0;
    J.ckvH = traceback ? (int*)(d + o_vH) : nullptr; J.ckvE = traceback ? (int*)(d + o_vE) : nullptr;
    J.progress = (int*)(d + o_pr);
    J.ticket = (int*)(d + o_misc);
    J.best = (unsigned long long*)(d + o_misc + 8);
    J.corner = (int*)(d + o_misc + 16);
    J.col0 = 0; J.n_total = n; J.row_off = 0;
    J.start_type = start_type; J.end_type = end_type;
    const bool typed = (J.start_type != -1 || J.end_type != -1);
    if (typed && mode != PSA_GLOBAL) return psa_fail(ctx, PSA_ERR_ARG, "start/end types apply to global alignment only");
    // geometry: checkpoints (traceback) fix the 128 x 256 tile grid; score-only runs may use taller row
    // blocks (less skew drain) and wider lanes (less per-step overhead)
    // At most one tile per strip column is in flight, so the wavefront is n/(32*K) tiles wide: pick the
    // widest lanes that still keep most of the resident warps busy (measured, same box: 1 Mbp <128,8> 865 ms,
    // <128,12> 815, <128,16> 725, <128,20> 639, <128,24> 635, <128,28> 638, <128,32> 698; 300 kbp <128,24> 156 ms,
    // <128,8> 159, <128,4> 136; 100 kbp <128,8> 42 ms vs <128,4> 36 ms).  Wide lanes amortise the per-step
    // overhead and leave sleeping warps' issue slots to the busy ones; narrow lanes widen the wavefront.
    int geo = 0;
    if (!traceback) {
        const long long resident = (long long)ctx->sm_count * 4 * WPB;
        if ((long long)n / 256 * 5 >= resident * 4)                    // from ~485 kbp (600 kbp: 332 ms vs 357 with <128,8>):
            geo = ((long long)m / 256 >= resident) ? 7 : 6;            // <256,24> when there are row blocks to spare (1 Mbp: 616 vs 632 ms), else <128,24>
        else geo = 4;                                                  // <128,4>
    }
    if (ctx->opt.long_geometry >= 0 && !traceback) geo = ctx->opt.long_geometry;
    auto launch = [&](auto kern, int RRv, int KKv) -> int {
        int per_sm = 0;
        PSA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WPB * 32, 0));
        if (ctx->opt.long_ctas_per_sm > 0) per_sm = std::max(1, std::min(per_sm, ctx->opt.long_ctas_per_sm));
        else if (per_sm > 4) per_sm = 4;
        const int NBv = (m + RRv - 1) / RRv;
        int grid = std::min((NBv + WPB - 1) / WPB, per_sm * ctx->sm_count);
        if (grid < 1) grid = 1;
        J.mul8 = key_mult(KKv);
        kern<<<grid, WPB * 32, 0, st>>>(J);
        PSA_CUDA_OK(ctx, cudaGetLastError());
        ctx->launches += 1;
        return PSA_OK;
    };
    int lrc;
    if (mode == PSA_LOCAL) {
        if (geo == 1) lrc = launch(psa_long_single_kernel<PSA_LOCAL, 256, 8>, 256, 8);
        else if (geo == 2) lrc = launch(psa_long_single_kernel<PSA_LOCAL, 128, 16>, 128, 16);
        else if (geo == 3) lrc = launch(psa_long_single_kernel<PSA_LOCAL, 256, 16>, 256, 16);
        else if (geo == 4) lrc = launch(psa_long_single_kernel<PSA_LOCAL, 128, 4>, 128, 4);
        else if (geo == 6) lrc = launch(psa_long_single_kernel<PSA_LOCAL, 128, 24>, 128, 24);
        else if (geo == 7) lrc = launch(psa_long_single_kernel<PSA_LOCAL, 256, 24>, 256, 24);
        else lrc = launch(psa_long_single_kernel<PSA_LOCAL, 128, 8>, 128, 8);
    } else {
        if (geo == 1) lrc = launch(psa_long_single_kernel<PSA_GLOBAL, 256, 8>, 256, 8);
        else if (geo == 2) lrc = launch(psa_long_single_kernel<PSA_GLOBAL, 128, 16>, 128, 16);
        else if (geo == 3) lrc = launch(psa_long_single_kernel<PSA_GLOBAL, 256, 16>, 256, 16);
        else if (geo == 4) lrc = launch(psa_long_single_kernel<PSA_GLOBAL, 128, 4>, 128, 4);
        else if (geo == 6) lrc = launch(psa_long_single_kernel<PSA_GLOBAL, 128, 24>, 128, 24);
        else if (geo == 7) lrc = launch(psa_long_single_kernel<PSA_GLOBAL, 256, 24>, 256, 24);
        else lrc = launch(psa_long_single_kernel<PSA_GLOBAL, 128, 8>, 128, 8);
    }
    if (lrc) return lrc;
    TbArgs T{J, d_item, traceback ? d_ops : nullptr, use_band ? (uint32_t*)(d + o_band) : nullptr};
    if (use_band) {
        const int units = NB * BAND_TILES;
        if (mode == PSA_LOCAL) psa_long_band_kernel<PSA_LOCAL><<<(units + BAND_WPB - 1) / BAND_WPB, BAND_WPB * 32, 0, st>>>(T);
        else psa_long_band_kernel<PSA_GLOBAL><<<(units + BAND_WPB - 1) / BAND_WPB, BAND_WPB * 32, 0, st>>>(T);
        PSA_CUDA_OK(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    if (mode == PSA_LOCAL) psa_long_tb_kernel<PSA_LOCAL><<<1, 32, 0, st>>>(T);
    else psa_long_tb_kernel<PSA_GLOBAL><<<1, 32, 0, st>>>(T);
    PSA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return PSA_OK;
}

// Batch of long pairs, score (+ end cell) only.  Scratch layout: [per-warp hbuf rows | ticket].
static const void* long_batch_kernel_for(int mode, int max_n, int* lane_cols) {
    *lane_cols = max_n >= 1024 ? 16 : 8;
    if (mode == PSA_LOCAL) return *lane_cols == 16 ? (const void*)psa_long_batch_kernel<PSA_LOCAL, 16> : (const void*)psa_long_batch_kernel<PSA_LOCAL, 8>;
    return *lane_cols == 16 ? (const void*)psa_long_batch_kernel<PSA_GLOBAL, 16> : (const void*)psa_long_batch_kernel<PSA_GLOBAL, 8>;
}

static int long_batch_grid(psa_ctx* ctx, long long n_pairs, const void* kern, int* grid) {
    int per_sm = 0;
    PSA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WPB * 32, 0));
    if (per_sm > 4) per_sm = 4;
    const long long want = (n_pairs + WPB - 1) / WPB;
    *grid = (int)std::min<long long>(want, (long long)per_sm * ctx->sm_count);
    if (*grid < 1) *grid = 1;
    return PSA_OK;
}

size_t psa_long_batch_scratch_bytes(psa_ctx* ctx, long long n_pairs, int max_n) {
    const size_t warp_stride = up256((size_t)(max_n + 1) * 4) / 4 * 2;
    const size_t max_grid = (size_t)4 * ctx->sm_count;
    return max_grid * WPB * warp_stride * 4 + 256;
}

int psa_launch_long_batch_at(psa_ctx* ctx, const psa_batch_args& args, int max_m, int max_n, int mode,
                             const uint8_t* d_flags, uint8_t* scratch, cudaStream_t st) {
    if (max_m >= 0x1FFFFF || max_n >= 0x1FFFFF) return psa_fail(ctx, PSA_ERR_RANGE, "long path: lengths must be < 2^21 - 1");
    int grid = 1, lane_cols = 8;
    const void* kern = long_batch_kernel_for(mode, max_n, &lane_cols);
    int rc = long_batch_grid(ctx, args.n_pairs, kern, &grid);
    if (rc) return rc;
    const size_t warp_stride = up256((size_t)(max_n + 1) * 4) / 4 * 2;       // ints: H row + F row
    const size_t hb = (size_t)4 * ctx->sm_count * WPB * warp_stride * 4;
    PSA_CUDA_OK(ctx, cudaMemsetAsync(scratch + hb, 0, 256, st));
    LongBatch Bt{args, (int*)scratch, (long long)warp_stride, (int*)(scratch + hb), d_flags, key_mult(lane_cols)};
    void* kargs[] = {&Bt};
    PSA_CUDA_OK(ctx, cudaLaunchKernel(kern, dim3(grid), dim3(WPB * 32), kargs, 0, st));
    ctx->launches += 1;
    return PSA_OK;
}

int psa_launch_long_batch(psa_ctx* ctx, const psa_batch_args& args, int max_m, int max_n, int mode, cudaStream_t st) {
    int rc = ensure_work(ctx, psa_long_batch_scratch_bytes(ctx, args.n_pairs, max_n));
    if (rc) return rc;
    return psa_launch_long_batch_at(ctx, args, max_m, max_n, mode, nullptr, (uint8_t*)ctx->d_work, st);
}


// ---- partition finder (the role of sequence_alignment/partial.cpp:81-163) ---------------------------
// Forward sweep of (A, B) and reverse sweep of (reverse A, reverse B), both score-only fills that keep the bottom row
// (H = max(T1,T2,T3) and F = T3 per column) of every 128-row block; the reverse sweep's first block is m mod 128 rows
// high, so that its block boundaries fall on the same matrix rows as the forward sweep's.  For a special row i the
// best crossing is  max_j max( Hf[i][j] + Hr[i][j],  Ff[i][j] + Fr[i][j] + h )  -- the path passes the node (i, j)
// between two operations, or a vertical gap spans the row (its opening penalty is charged by both sweeps, hence + h;
// partial.cpp:101-105).  Smallest j wins, node before gap (partial.cpp:108 uses the same >= priority).
namespace {

struct SweepRows { int* H; int* F; long long stride; int nb; };

// One checkpointed score-only fill into the region of ctx->d_work starting at byte `base`; returns the rows.
int sweep_rows(psa_ctx* ctx, uint8_t* region, size_t region_bytes, const uint8_t* d_a, const uint8_t* d_b, int m, int n, int g, int h,
               int row_off, int* d_corner, cudaStream_t st, SweepRows* out) {
    const int NB = row_off ? 1 + std::max(0, m - row_off + R - 1) / R : (m + R - 1) / R;
    const size_t row = up256((size_t)(n + 1) * 4);
    size_t o = 0;
    const size_t o_hH = o; o += row * NB;
    const size_t o_hF = o; o += row * NB;
    const size_t o_pr = o; o += up256((size_t)NB * 4);
    const size_t o_misc = o; o += 256;
    if (o > region_bytes) return psa_fail(ctx, PSA_ERR_NOMEM, "partition sweep region too small");
    PSA_CUDA_OK(ctx, cudaMemsetAsync(region + o_pr, 0, (o_misc + 256) - o_pr, st));
    LongJob J;
    J.a = d_a; J.b = d_b; J.m = m; J.n = n; J.g = g; J.h = h; J.mul8 = key_mult(K);
    J.hbufH = (int*)(region + o_hH); J.hbufF = (int*)(region + o_hF); J.hb_stride = (long long)(row / 4);
    J.ckvH = nullptr; J.ckvE = nullptr;
    J.progress = (int*)(region + o_pr); J.ticket = (int*)(region + o_misc);
    J.best = (unsigned long long*)(region + o_misc + 8); J.corner = d_corner;
    J.col0 = 0; J.n_total = n; J.row_off = row_off; J.start_type = -1; J.end_type = -1;
    int per_sm = 0;
    PSA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, psa_long_single_kernel<PSA_GLOBAL, R, K, true>, WPB * 32, 0));
    if (per_sm > 4) per_sm = 4;
    int grid = std::min((NB + WPB - 1) / WPB, per_sm * ctx->sm_count);
    if (grid < 1) grid = 1;
    if (row_off) psa_long_single_kernel<PSA_GLOBAL, R, K, true><<<grid, WPB * 32, 0, st>>>(J);
    else psa_long_single_kernel<PSA_GLOBAL, R, K, false><<<grid, WPB * 32, 0, st>>>(J);
    PSA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 1;
    out->H = J.hbufH; out->F = J.hbufF; out->stride = J.hb_stride; out->nb = NB;
    return PSA_OK;
}

// one CTA per special row: (i, j, type) of the best crossing
__global__ void __launch_bounds__(256) psa_crossing_kernel(SweepRows fwd, SweepRows rev, int m, int n, int g, int h, int row_off_rev,
                                                           const int* rows, int n_rows, int* out /* [n_rows][4]: i, j, type, value */) {
    __shared__ long long s_best[256];
    const int q = blockIdx.x;
    if (q >= n_rows) return;
    const int i = rows[q];                              // 0 < i < m, a multiple of R
    const int* Hf = fwd.H + (long long)(i / R - 1) * fwd.stride;
    const int* Ff = fwd.F + (long long)(i / R - 1) * fwd.stride;
    const int ir = m - i;                               // the same matrix row as the reverse sweep numbers it
    const int rbr = row_off_rev ? (ir - row_off_rev) / R : ir / R - 1;
    const int* Hr = rev.H + (long long)rbr * rev.stride;
    const int* Fr = rev.F + (long long)rbr * rev.stride;
    long long best = LLONG_MIN;
    for (int j = threadIdx.x; j <= n; j += blockDim.x) {
        const int jr = n - j;
        // column 0 of either sweep: only the vertical border gap reaches it (T3[i][0] = -h - g*i, cpp:290-292)
        const int hf = j > 0 ? Hf[j] : -h - g * i, ff = j > 0 ? Ff[j] : -h - g * i;
        const int hr = jr > 0 ? Hr[jr] : -h - g * ir, fr = jr > 0 ? Fr[jr] : -h - g * ir;
        const long long v1 = (long long)hf + hr, v3 = (long long)ff + fr + h;
        const long long v = v1 >= v3 ? v1 : v3;
        const int type = v1 >= v3 ? 1 : 3;
        // order: value desc, then j asc, then node (1) before gap (3)
        const long long key = (v << 24) | ((long long)(0x3FFFFF - j) << 2) | (type == 1 ? 1 : 0);
        best = key > best ? key : best;
    }
    s_best[threadIdx.x] = best;
    __syncthreads();
    for (int off = 128; off >= 1; off >>= 1) {
        if ((int)threadIdx.x < off) s_best[threadIdx.x] = s_best[threadIdx.x] > s_best[threadIdx.x + off] ? s_best[threadIdx.x] : s_best[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long key = s_best[0];
        out[4 * q + 0] = i;
        out[4 * q + 1] = 0x3FFFFF - (int)((key >> 2) & 0x3FFFFF);
        out[4 * q + 2] = (key & 1) ? 1 : 3;
        out[4 * q + 3] = (int)(key >> 24);
    }
}

}  // namespace

// Crossing points of an optimal global alignment on up to max_rows special rows (multiples of 128, evenly spread).
// d_a / d_b and their reversals resident on the device.  h_points[k] = {i, j, type, optimal score}.
int psa_find_crossings(psa_ctx* ctx, const uint8_t* d_a, const uint8_t* d_b, const uint8_t* d_ar, const uint8_t* d_br, int m, int n,
                       int g, int h, int max_rows, int* h_points, int* n_points, cudaStream_t st) {
    *n_points = 0;
    const int avail = (m - 1) / R;                      // rows R, 2R, ... < m
    const int want = std::min(max_rows, avail);
    if (want <= 0) return PSA_OK;
    if (m >= 0x1FFFFF || n >= 0x1FFFFF) return psa_fail(ctx, PSA_ERR_RANGE, "long path: lengths must be < 2^21 - 1");
    const int row_off_rev = m % R;                      // 0: the regular grid already lines up
    const int NBf = (m + R - 1) / R, NBr = row_off_rev ? 1 + (m - row_off_rev + R - 1) / R : NBf;
    const size_t row = up256((size_t)(n + 1) * 4);
    const size_t reg_f = 2 * row * NBf + up256((size_t)NBf * 4) + 512, reg_r = 2 * row * NBr + up256((size_t)NBr * 4) + 512;
    const size_t o_f = 0, o_r = up256(reg_f), o_rows = o_r + up256(reg_r), o_out = o_rows + up256((size_t)want * 4);
    const size_t total = o_out + up256((size_t)want * 16) + 256;
    int rc = ensure_work(ctx, total);
    if (rc) return rc;
    uint8_t* d = (uint8_t*)ctx->d_work;
    SweepRows F, Rv;
    int* d_corner = (int*)(d + o_out + up256((size_t)want * 16));
    rc = sweep_rows(ctx, d + o_f, reg_f, d_a, d_b, m, n, g, h, 0, d_corner, st, &F);
    if (rc) return rc;
    rc = sweep_rows(ctx, d + o_r, reg_r, d_ar, d_br, m, n, g, h, row_off_rev, d_corner + 4, st, &Rv);
    if (rc) return rc;
    std::vector<int> rows(want);
    for (int k = 0; k < want; ++k) {
        int idx = (int)((long long)(k + 1) * (avail + 1) / (want + 1));     // 1 .. avail, evenly spread
        idx = std::max(1, std::min(avail, idx));
        rows[k] = idx * R;
    }
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    const int nr = (int)rows.size();
    PSA_CUDA_OK(ctx, cudaMemcpyAsync(d + o_rows, rows.data(), (size_t)nr * 4, cudaMemcpyHostToDevice, st));
    psa_crossing_kernel<<<nr, 256, 0, st>>>(F, Rv, m, n, g, h, row_off_rev, (const int*)(d + o_rows), nr, (int*)(d + o_out));
    PSA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 1;
    PSA_CUDA_OK(ctx, cudaMemcpyAsync(h_points, d + o_out, (size_t)nr * 16, cudaMemcpyDeviceToHost, st));
    PSA_CUDA_OK(ctx, cudaStreamSynchronize(st));
    *n_points = nr;
    return PSA_OK;
}

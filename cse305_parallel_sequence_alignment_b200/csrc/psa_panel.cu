// psa_panel.cu -- column-stationary wavefront for ONE long pair (BASELINE configs 3 and 4).
//
// A warp owns a strip of PW = 128 columns for ALL m rows (lane t owns 4 columns; skewed wavefront,
// (H, E) of the column to the left arrive from lane t-1 by shuffle).  Adjacent strips are chained
// through a small ring in L2: lane 31 of strip w stores (H, E) of its last column row by row and,
// every 32 rows, publishes a "produced" counter with release semantics; strip w+1 polls it (all
// lanes, relaxed load + fence), block-loads 32 rows and feeds them to its lane 0 by shuffle, then
// bumps a "consumed" counter that lets the producer reuse ring slots.  All strips of a panel are
// co-resident (grid = one warp per strip, sized by the occupancy API), so strip w+1 trails strip w
// by ~64 rows and the whole panel advances as ONE anti-diagonal wavefront: the critical path is
// m + 32 * strips lane-steps, with no per-tile drain (the row-block kernel in psa_long.cu paid
// (m/128 + n/256) * 159 steps).  A matrix wider than one panel is processed panel by panel; the
// last strip of a panel writes its whole right boundary column (8 B/row) for the next panel -- or,
// across GPUs, straight into the next rank's HBM over NVLink with a system-scope counter.
//
// Recurrence, borders, corner capture and local end-cell key exactly as psa_tile.cuh
// (subproblem_alignment.cpp:229-292).  With checkpoints enabled the kernel also writes the
// row-block bottom rows (every 128 rows) and strip right columns (every 256 columns) that
// psa_long_tb_kernel recomputes tiles from.
#include "psa_tile.cuh"

using namespace psa_tile;

namespace {

constexpr int PK = 4;
constexpr int PW = 32 * PK;          // strip width
constexpr int RING = 1024;           // ring rows per strip (power of two, multiple of 32)
constexpr int PWPB = 4;              // warps (strips) per CTA
constexpr int CK_R = 128, CK_W = 256;   // checkpoint grid of psa_long_tb_kernel

__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_sys(const int* p) {
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct PanelJob {
    const uint8_t* a;
    const uint8_t* b;            // this launch's columns: b[0] is global column col_begin+1
    int m, g, h, mul4;
    int col_begin;               // global index of the column left of the panel
    int n_cols;                  // columns in this panel
    int n_total;                 // columns of the whole pair (corner / key validity)
    int nstrips;
    // left edge of the panel: null -> matrix column 0 (border formula)
    const int2* pin;             // [m] (H, E) of global column col_begin, rows 1..m
    const int* pin_count;        // rows published so far (epoch-biased, see `count_base`); null -> all there
    int pin_sys;                 // 1: written by another GPU (system scope)
    // right edge of the panel: null -> nobody needs it
    int2* pout;
    int* pout_count;
    int pout_sys;
    int count_base;              // counters hold count_base + rows (lets buffers be reused without clearing)
    // rings between adjacent strips of the panel
    int2* rings;                 // [nstrips][RING]
    int* produced;               // [nstrips]
    int* consumed;               // [nstrips]
    // checkpoints for the traceback kernel (null: score only)
    int* hbufH; int* hbufF; long long hb_stride;     // row (i/128 - 1), column j (1-based, global)
    int* ckvH; int* ckvE;                            // column block (j/256 - 1), row i (1-based); stride m+1
    unsigned long long* best;
    int* corner;
};

template <bool LOCAL, bool CAP>
__device__ __forceinline__ void panel_step(int (&H)[PK], int (&G)[PK], int (&F)[PK], const int (&b)[PK], const int (&ka)[PK],
                                           int& hlgo, int& el, int diag, int a, int ng, int go, int mul4, int& rowkey,
                                           int kcap, int& c1, int& c2, int& c3) {
    int key_prev = 0;
#pragma unroll
    for (int k = 0; k < PK; ++k) {
        const int t1 = diag + (a == b[k] ? 1 : 0);
        const int e = __viaddmax_s32(el, ng, hlgo);
        const int f = __viaddmax_s32(F[k], ng, G[k]);
        const int Hn = __vimax3_s32(t1, e, f);
        if (LOCAL) {
            const int key = t1 * mul4 + ka[k];
            if (k & 1) rowkey = __vimax3_s32(rowkey, key_prev, key);
            key_prev = key;
        }
        if (CAP) { if (k == kcap) { c1 = t1; c2 = e; c3 = f; } }
        diag = H[k];
        const int hg = Hn - go;
        H[k] = Hn; G[k] = hg; F[k] = f; hlgo = hg; el = e;
    }
}

template <int MODE>
__global__ void __launch_bounds__(PWPB * 32) psa_panel_kernel(PanelJob J) {
    constexpr bool LOCAL = (MODE == PSA_LOCAL);
    const int lane = threadIdx.x & 31;
    const int sidx = blockIdx.x * PWPB + (threadIdx.x >> 5);
    if (sidx >= J.nstrips) return;
    const int m = J.m, g = J.g, h = J.h, go = J.g + J.h, ng = -J.g;
    const int c0 = sidx * PW + lane * PK;                 // panel-local 0-based first column of this lane
    const int cg = J.col_begin + c0;                      // global index of the column left of it
    const bool first = (sidx == 0), last = (sidx == J.nstrips - 1);

    int H[PK], G[PK], F[PK], b[PK], ka[PK];
#pragma unroll
    for (int k = 0; k < PK; ++k) {
        const int jl = c0 + k;                            // panel-local 0-based column
        const bool valid = jl < J.n_cols;
        b[k] = valid ? (int)J.b[jl] : 256;
        H[k] = valid ? border_row0_H<MODE>(cg + k + 1, g, h) : (LOCAL ? 0 : PSA_KNEG);   // row 0 (cpp:222-224)
        F[k] = PSA_KNEG;
        G[k] = H[k] - go;
        ka[k] = valid ? (PK - 1 - k) : -(1 << 30);
    }
    int hd = border_row0_H<MODE>(cg, g, h);               // H[0][cg]
    int recv_h = PSA_KNEG, recv_e = PSA_KNEG;
    int bestkey = 0, besti = 0;
    int c1 = PSA_KNEG, c2 = PSA_KNEG, c3 = PSA_KNEG;
    // which k of this lane owns global column n_total (only in the panel that contains it)
    const int kcap = (!LOCAL && J.n_total > cg && J.n_total <= cg + PK) ? (J.n_total - 1 - cg) : -1;

    const int2* in_rows = first ? J.pin : J.rings + (size_t)(sidx - 1) * RING;
    const int* in_count = first ? J.pin_count : J.produced + (sidx - 1);
    const bool in_sys = first && J.pin_sys;
    const bool in_ring = !first;
    int2* out_rows = last ? J.pout : J.rings + (size_t)sidx * RING;
    int* out_count = last ? J.pout_count : J.produced + sidx;
    const bool out_sys = last && J.pout_sys;
    const bool out_ring = !last;
    const bool have_out = (out_rows != nullptr);
    // vertical checkpoint: this strip's right edge is a 256-column boundary of the traceback grid
    const int redge = J.col_begin + (sidx + 1) * PW;      // global column of this strip's right edge
    const int ckcol = (J.ckvH != nullptr && (redge % CK_W) == 0 && redge <= J.n_total) ? redge / CK_W - 1 : -1;

    int2 blk = make_int2(PSA_KNEG, PSA_KNEG);
    const int steps = m + 31;
    for (int st = 0; st < steps; ++st) {
        // ---- every 32 steps: fetch the next 32 rows of the left boundary (lane L <- row st+L) ----
        if ((st & 31) == 0 && st < m) {
            const int need = min(m, st + 32);
            if (in_rows == nullptr) {                    // matrix column 0 (cpp:282-292)
                const int i = st + lane + 1;
                blk = make_int2(border_col0_H<MODE>(i, g, h), PSA_KNEG);
            } else {
                if (in_count != nullptr) {
                    unsigned ns = 32;
                    if (in_sys) { while (ld_relaxed_sys(in_count) - J.count_base < need) { __nanosleep(ns); if (ns < 1024) ns <<= 1; } __threadfence_system(); }
                    else { while (ld_relaxed_gpu(in_count) - J.count_base < need) { __nanosleep(ns); if (ns < 1024) ns <<= 1; } __threadfence(); }
                }
                const int rr = st + lane;
                if (rr < m) blk = __ldcg(in_rows + (in_ring ? (rr & (RING - 1)) : rr));
                if (in_ring) {
                    __syncwarp();
                    if (lane == 0) st_relaxed_gpu(J.consumed + (sidx - 1), J.count_base + need);
                }
            }
            // ring space for the rows this strip is about to produce (rows st-31 .. st)
            if (have_out && out_ring) {
                unsigned ns = 32;
                // a counter that still holds 0 (or a stale epoch) reads as "nothing consumed yet"
                while (st + 32 - max(ld_relaxed_gpu(J.consumed + sidx) - J.count_base, 0) > RING) { __nanosleep(ns); if (ns < 1024) ns <<= 1; }
            }
        }
        const int r = st - lane;
        int hlgo, el;
        {
            const int bh = __shfl_sync(0xffffffffu, blk.x, st & 31);
            const int be = __shfl_sync(0xffffffffu, blk.y, st & 31);
            if (lane == 0) { hlgo = bh - go; el = be; } else { hlgo = recv_h; el = recv_e; }
        }
        const bool active = (r >= 0 && r < m);
        const bool capstep = !LOCAL && active && kcap >= 0 && r == m - 1;
        const bool anycap = LOCAL ? false : __any_sync(0xffffffffu, capstep);
        if (active) {
            const int a = J.a[r];
            const int hin = hlgo;
            int rowkey = 0;
            if (!anycap) panel_step<LOCAL, false>(H, G, F, b, ka, hlgo, el, hd, a, ng, go, J.mul4, rowkey, -1, c1, c2, c3);
            else panel_step<LOCAL, true>(H, G, F, b, ka, hlgo, el, hd, a, ng, go, J.mul4, rowkey, capstep ? kcap : -1, c1, c2, c3);
            hd = hin + go;
            if (LOCAL) {
                const bool up = rowkey > (bestkey | (PK - 1));
                bestkey = up ? rowkey : bestkey;
                besti = up ? (r + 1) : besti;
            }
            if (lane == 31) {
                if (have_out) out_rows[out_ring ? (r & (RING - 1)) : r] = make_int2(hlgo + go, el);
                if (ckcol >= 0) { J.ckvH[(long long)ckcol * (m + 1) + r + 1] = hlgo + go; J.ckvE[(long long)ckcol * (m + 1) + r + 1] = el; }
            }
            if (J.hbufH != nullptr && ((r + 1) % CK_R) == 0) {      // horizontal checkpoint row
                long long base = (long long)((r + 1) / CK_R - 1) * J.hb_stride;
#pragma unroll
                for (int k = 0; k < PK; ++k)
                    if (c0 + k < J.n_cols) { J.hbufH[base + cg + k + 1] = H[k]; J.hbufF[base + cg + k + 1] = F[k]; }
            }
        }
        // ---- publish every 32 produced rows (lane 31 finished row st-31) ----
        if (have_out && out_count != nullptr) {
            const int done = st - 31 + 1;                 // rows 0 .. st-31 are stored
            if (done > 0 && ((done & 31) == 0 || done == m)) {
                if (out_sys) __threadfence_system(); else __threadfence();
                __syncwarp();
                if (lane == 31) { if (out_sys) st_release_sys(out_count, J.count_base + done); else st_release_gpu(out_count, J.count_base + done); }
            }
        }
        recv_h = __shfl_up_sync(0xffffffffu, hlgo, 1);
        recv_e = __shfl_up_sync(0xffffffffu, el, 1);
    }

    if (LOCAL) {
        const int t1v = bestkey / PK;
        const int j = cg + (PK - 1 - (bestkey % PK)) + 1;
        unsigned long long key = 0ull;
        if (t1v > 0 && j <= J.n_total)
            key = ((unsigned long long)(unsigned)t1v << 42) | ((unsigned long long)(0x1FFFFF - besti) << 21) | (unsigned long long)(0x1FFFFF - j);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
            key = o > key ? o : key;
        }
        if (lane == 0 && key != 0ull) atomicMax(J.best, key);
    } else if (kcap >= 0) {
        J.corner[0] = c1; J.corner[1] = c2; J.corner[2] = c3;
    }
}

}  // namespace

// Resident strips of one launch (= panel width / 128).
int psa_panel_capacity(psa_ctx* ctx, int mode, int* strips) {
    int per_sm = 0;
    if (mode == PSA_LOCAL) PSA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, psa_panel_kernel<PSA_LOCAL>, PWPB * 32, 0));
    else PSA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, psa_panel_kernel<PSA_GLOBAL>, PWPB * 32, 0));
    int cap = 6;                                          // 24 warps/SM saturate the integer pipes; keeps every CTA resident
    if (ctx->opt.systolic_warps_per_sm > 0) cap = std::max(1, ctx->opt.systolic_warps_per_sm / PWPB);
    per_sm = std::min(per_sm, cap);
    *strips = per_sm * ctx->sm_count * PWPB;
    return PSA_OK;
}

size_t psa_panel_ring_bytes(int strips) { return (size_t)strips * RING * sizeof(int2) + (size_t)strips * 2 * sizeof(int) + 512; }

// One panel.  `scratch` holds [rings | produced | consumed]; counters are used epoch-biased.
int psa_launch_panel(psa_ctx* ctx, const psa_panel_args& P, cudaStream_t st) {
    PanelJob J;
    J.a = P.d_a; J.b = P.d_b; J.m = P.m; J.g = P.g; J.h = P.h; J.mul4 = PK;
    J.col_begin = P.col_begin; J.n_cols = P.n_cols; J.n_total = P.n_total;
    J.nstrips = (P.n_cols + PW - 1) / PW;
    J.pin = (const int2*)P.pin; J.pin_count = P.pin_count; J.pin_sys = P.pin_sys;
    J.pout = (int2*)P.pout; J.pout_count = P.pout_count; J.pout_sys = P.pout_sys;
    J.count_base = P.count_base;
    uint8_t* s = (uint8_t*)P.scratch;
    J.rings = (int2*)s;
    J.produced = (int*)(s + (size_t)P.scratch_strips * RING * sizeof(int2));
    J.consumed = J.produced + P.scratch_strips;
    J.hbufH = P.hbufH; J.hbufF = P.hbufF; J.hb_stride = P.hb_stride; J.ckvH = P.ckvH; J.ckvE = P.ckvE;
    J.best = P.best; J.corner = P.corner;
    if (J.nstrips > P.scratch_strips) return psa_fail(ctx, PSA_ERR_RANGE, "panel wider than the resident strip capacity");
    const int grid = (J.nstrips + PWPB - 1) / PWPB;
    if (P.mode == PSA_LOCAL) psa_panel_kernel<PSA_LOCAL><<<grid, PWPB * 32, 0, st>>>(J);
    else psa_panel_kernel<PSA_GLOBAL><<<grid, PWPB * 32, 0, st>>>(J);
    PSA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return PSA_OK;
}

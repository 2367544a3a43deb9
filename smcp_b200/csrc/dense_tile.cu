// Single-launch dense kernels for matrices that live in L2 (m up to a few thousand): the m x m
// Schur complement H of the band configuration (m = 1000), and the frontal matrices of the
// top set of the clique tree (rand_SDP n = 2000: a 1186-column root and 27 fronts of ~1130
// rows).  They replace launch chains (lapack.potrf, src/python/solvers.py:501, 1931, was 52
// launches at m = 1000; the blocked triangular solves of front.cu were ~38 launches per
// solve at n = 1186):
//
//   potrf_tile_kernel  right-looking tile Cholesky (64 x 64 tiles) as ONE cooperative kernel:
//                      every tile of the lower triangle is owned by a CTA; a step factors the
//                      diagonal tile (redundantly in every CTA that owns a tile of that block
//                      column: no broadcast round trip), solves the panel tiles, grid-syncs,
//                      applies the panel to the trailing tiles with DMMA, grid-syncs.
//                      `npiv` < m gives the partial factorisation of a frontal matrix (trailing
//                      block = update matrix); `ntc` limits the updated tile columns (panel mode
//                      of the blocked factorisation in dense.cu for large m).
//   trsm_slab_kernel   X = L^{-1} B / L^{-T} B with a dense lower-triangular L: a CTA owns 8
//                      right-hand sides for the whole solve and keeps them in shared memory
//                      (no inter-CTA dependency at all); 64 x 64 diagonal blocks are solved by one
//                      warp per right-hand side (shuffle substitution), the rest of each block
//                      column is applied with m16n8k8 DMMA whose A fragments are read straight
//                      from L in L2 (every entry of L is used exactly once per CTA).
#include "internal.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <vector>

namespace cg = cooperative_groups;

#define TT 64          // tile size
#define DLD 66         // leading dimension of the diagonal tile in shared memory (even: double2 loads)
#define OLDT 68        // leading dimension of the [k][row] operand tiles: conflict-free DMMA fragment loads
#define PT_THREADS 256
#define PT_MAXOWN 64   // owned tiles per CTA (host checks)

__device__ __forceinline__ void dmma16x8x8_t(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

__device__ __forceinline__ void pt_cp_async8(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}

struct PotrfTileArgs {
    double *H;
    long long ld;
    int mm;        // order of the (sub)matrix
    int npiv;      // leading columns to factor
    int ntc;       // tile columns that exist for this call (P: full trailing update; ceil(npiv/64): panel only)
    int *info;     // dpotrf's info (first non-positive pivot, 1-based), written if still 0
    int col_off;   // added to the local column index in info
    unsigned *bar; // grid barrier counter (zeroed before the launch)
    long long *dbg; // SMCP_B200_PT_DEBUG: per-step phase timestamps of CTA 0 (ns), 6 per step
};

__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define PT_STAMP(slot) do { if (a.dbg && blockIdx.x == 0 && tid == 0) a.dbg[6 * k + (slot)] = gtimer(); } while (0)

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Barrier over the (co-resident: cooperative launch) grid: one arrival per CTA on a monotone counter.
// cooperative_groups' grid.sync() costs ~4-5 us on 148 CTAs, this one ~2.
__device__ __forceinline__ void grid_barrier(unsigned *ctr, unsigned &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire_u32(ctr) < target) { }
    }
    __syncthreads();
}

// Cholesky of the 64 x 64 tile in D (kb pivots; the trailing part of the tile receives the Schur
// complement), all 256 threads: thread (i, jq) keeps the entries (i, 16 jq .. 16 jq + 15) of the lower
// triangle in REGISTERS.  The tile is processed in four 16-column sub-steps; per sub-step
//   (a) the 16 x 16 diagonal block is factored by the 16 lanes that hold its rows: the pivot column
//       travels by width-16 shuffles, every lane computes 1/sqrt(pivot) itself -- no block barrier
//       per pivot (the previous version had one __syncthreads per pivot: 0.45 us x 64 = 29 us a tile),
//   (b) the rows below it solve against the block from shared memory, thread-local in registers,
//   (c) the later column blocks take the rank-16 update; L(:, 16 s ..) is exchanged through Ls
//       (two 16 x 64 buffers alternating with s: two barriers per sub-step in all).
// rs[c] = 1/sqrt(pivot c).  The 1-based index of the first non-positive pivot (0 = none) goes to *bad_s.
#define PT_FULL 0xffffffffu
__device__ __forceinline__ void tile_chol64(double *D, double *rs, double *Ls, int *bad_s, int kb, int tid) {
    const int i = tid & 63, jq = tid >> 6, lane = tid & 31;
    double a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int j = 16 * jq + q;
        a[q] = (j <= i) ? D[j * DLD + i] : 0.0;
    }
    if (tid == 0) *bad_s = 0;
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        const int nbs = min(16, kb - 16 * s);
        if (nbs <= 0) break;
        double *Lb = Ls + (s & 1) * (16 * TT);
        if ((tid >> 5) == ((80 * s) >> 5)) {
            // (a) this warp holds rows 16 s .. 16 s + 15 of column block s in one of its halves; the other
            // half runs along on its own (unused) copy
            const int t = lane & 15;
            const bool mine = (i >> 4) == s;
            double b[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) b[q] = a[q];
            int bad = 0;
            double d = __shfl_sync(PT_FULL, b[0], 0, 16);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (c < nbs) {
                    if (!(d > 0.0)) { if (!bad) bad = c + 1; d = 1.0; }
                    const double r = rsqrt(d);
                    const double l = b[c] * r;
                    b[c] = l;
                    if (mine && t == c) rs[16 * s + c] = r;
                    if (c + 1 < 16) {
                        // next pivot straight from its owner (on lane c + 1 the shuffled l1 below is its own l:
                        // same value, one shuffle less on the pivot-to-pivot dependency chain)
                        d = __shfl_sync(PT_FULL, fma(-l, l, b[c + 1]), c + 1, 16);
                        const double l1 = __shfl_sync(PT_FULL, l, c + 1, 16);
                        b[c + 1] = fma(-l, l1, b[c + 1]);
                    }
#pragma unroll
                    for (int j = c + 2; j < 16; ++j) {
                        const double lj = __shfl_sync(PT_FULL, l, j, 16);
                        b[j] = fma(-l, lj, b[j]);
                    }
                }
            }
            if (mine) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    a[q] = b[q];
                    Lb[q * TT + i] = (q <= t) ? b[q] : 0.0;
                }
                if (t == 0 && bad && !*bad_s) *bad_s = 16 * s + bad;
            }
        }
        __syncthreads();
        if (jq == s && i >= 16 * s + 16) {
            // (b) l_ic = (a_ic - sum_{c' < c} l_ic' l_cc') / l_cc, right-looking over the block's columns
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (c < nbs) {
                    const double l = a[c] * rs[16 * s + c];
                    a[c] = l;
#pragma unroll
                    for (int j = c + 1; j < 16; ++j) a[j] = fma(-l, Lb[c * TT + 16 * s + j], a[j]);
                }
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) Lb[q * TT + i] = a[q];
        }
        __syncthreads();
        if (jq > s && i >= 16 * jq) {
            // (c) a_ij -= sum_c l_ic l_jc for the columns j of block jq
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (c < nbs) {
                    const double li = Lb[c * TT + i];
                    const double2 *lc = reinterpret_cast<const double2 *>(Lb + c * TT + 16 * jq);
#pragma unroll
                    for (int q2 = 0; q2 < 8; ++q2) {
                        const double2 lv = lc[q2];
                        a[2 * q2] = fma(-li, lv.x, a[2 * q2]);
                        a[2 * q2 + 1] = fma(-li, lv.y, a[2 * q2 + 1]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int j = 16 * jq + q;
        if (j <= i) D[j * DLD + i] = a[q];
    }
    __syncthreads();
}

// Panel tile X L^T = B against the factored diagonal tile in D (kb pivots; the columns past kb take the
// Schur complement): thread (r, jq) keeps the entries (r, 16 jq .. 16 jq + 15) of the tile in registers,
// same sub-step scheme as tile_chol64 without its phase (a).
__device__ __forceinline__ void tile_panel64(const double *D, const double *rs, double *Ls, double (&x)[16], int kb, int tid) {
    const int r = tid & 63, jq = tid >> 6;
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        const int nbs = min(16, kb - 16 * s);
        if (nbs <= 0) break;
        double *Lb = Ls + (s & 1) * (16 * TT);
        if (jq == s) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (c < nbs) {
                    const double l = x[c] * rs[16 * s + c];
                    x[c] = l;
#pragma unroll
                    for (int j = c + 1; j < 16; ++j) x[j] = fma(-l, D[(16 * s + c) * DLD + 16 * s + j], x[j]);
                }
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) Lb[q * TT + r] = x[q];
        }
        __syncthreads();
        if (jq > s) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (c < nbs) {
                    const double li = Lb[c * TT + r];
                    const double2 *lc = reinterpret_cast<const double2 *>(D + (16 * s + c) * DLD + 16 * jq);
#pragma unroll
                    for (int q2 = 0; q2 < 8; ++q2) {
                        const double2 lv = lc[q2];
                        x[2 * q2] = fma(-li, lv.x, x[2 * q2]);
                        x[2 * q2 + 1] = fma(-li, lv.y, x[2 * q2 + 1]);
                    }
                }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PT_THREADS, 1) potrf_tile_kernel(PotrfTileArgs a) {
    extern __shared__ __align__(16) double ptsm[];
    double *D = ptsm;                    // TT x DLD, D[c*DLD + r] = tile(r, c)
    double *rs = D + TT * DLD;           // 1/sqrt(pivot)
    double *Ls = rs + TT;                // 2 x (16 x TT): L(:, 16 s ..) of the current sub-step (tile_chol64 / tile_panel64)
    double *W = Ls + 2 * 16 * TT;        // phase B: two buffers of As | Bs
    __shared__ short2 own[PT_MAXOWN];
    __shared__ int nown_s, bad_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *H = a.H;
    const long long ld = a.ld;
    const int mm = a.mm;
    const int P = (mm + TT - 1) / TT, Pc = (a.npiv + TT - 1) / TT;
    unsigned bar_target = 0;
    if (tid == 0) {
        const long long ntiles = (long long)a.ntc * P - (long long)a.ntc * (a.ntc - 1) / 2;
        int cnt = 0, j = 0;
        long long off = 0;
        for (long long idx = blockIdx.x; idx < ntiles && cnt < PT_MAXOWN; idx += gridDim.x) {
            while (idx >= off + (P - j)) { off += P - j; ++j; }
            own[cnt++] = make_short2((short)(j + (int)(idx - off)), (short)j);
        }
        nown_s = cnt;
    }
    __syncthreads();
    const int nown = nown_s;

    for (int k = 0; k < Pc; ++k) {
        const int kb = min(TT, a.npiv - TT * k);
        const long long kc = (long long)TT * k;
        bool workA = false, workB = false, diag_owner = false;
        for (int q = 0; q < nown; ++q) {
            const short2 t = own[q];
            if (t.y == k) { workA = true; if (t.x == k) diag_owner = true; }
            else if (t.y > k) workB = true;
        }
        PT_STAMP(0);
        if (workA) {
            // ---- diagonal tile: load (identity beyond the matrix), factor kb pivots
            for (int idx = tid; idx < TT * TT; idx += PT_THREADS) {
                const int r = idx & 63, c = idx >> 6;
                double v = 0.0;
                if (r >= c) {
                    if (kc + r < mm && kc + c < mm) v = H[(kc + r) + (kc + c) * ld];
                    else if (r == c) v = 1.0;
                }
                D[c * DLD + r] = v;
            }
            __syncthreads();
            tile_chol64(D, rs, Ls, &bad_s, kb, tid);
            if (diag_owner && bad_s && tid == 0 && *a.info == 0) *a.info = a.col_off + (int)kc + bad_s;
            PT_STAMP(1);
            // ---- panel tiles (i, k), i > k: X L^T = B
            for (int q = 0; q < nown; ++q) {
                if (own[q].y != k || own[q].x == k) continue;
                const int r = tid & 63, jq = tid >> 6;
                const long long R = (long long)TT * own[q].x + r;
                const bool live = R < mm;
                double *Pg = H + R + (kc + 16 * jq) * ld;
                double x[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) x[e] = (live && kc + 16 * jq + e < mm) ? Pg[(long long)e * ld] : 0.0;
                tile_panel64(D, rs, Ls, x, kb, tid);
                if (live) {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (kc + 16 * jq + e < mm) Pg[(long long)e * ld] = x[e];
                }
            }
        }
        PT_STAMP(2);
        grid_barrier(a.bar, bar_target);
        PT_STAMP(3);
        if (diag_owner) {
            // written back only now: during phase A the other CTAs of this block column were still
            // reading the unfactored tile; D is not touched by the trailing updates below
            for (int idx = tid; idx < TT * TT; idx += PT_THREADS) {
                const int r = idx & 63, c = idx >> 6;
                if (r >= c && kc + r < mm) H[(kc + r) + (kc + c) * ld] = D[c * DLD + r];
            }
        }
        if (workB) {
            // ---- trailing tiles (i, j), j > k: C -= L(i,k) L(j,k)^T on the FP64 tensor cores.  The operand tiles of the
            // NEXT owned tile are fetched with cp.async into the other half of W while this one is multiplied, and the C
            // fragment is loaded before the wait: a tile cost load + load + 8 DMMA steps in sequence before (5.9 us
            // against 2.2 us of DMMA time)
            const int kb8 = (kb + 7) & ~7;
            const int g = lane >> 2, t = lane & 3;
            const int r0 = (warp & 3) * 16, c0w = (warp >> 2) * 32;
            auto fetch = [&](int q, int buf) {
                const int ti = own[q].x, tj = own[q].y;
                const long long ri = (long long)TT * ti, rj = (long long)TT * tj;
                double *As = W + buf * (2 * TT * OLDT), *Bs = As + TT * OLDT;
                for (int idx = tid; idx < TT * kb8; idx += PT_THREADS) {
                    const int r = idx & 63, c = idx >> 6;
                    if (c < kb && ri + r < mm) pt_cp_async8(As + c * OLDT + r, H + (ri + r) + (kc + c) * ld);
                    else As[c * OLDT + r] = 0.0;
                    if (ti != tj) {
                        if (c < kb && rj + r < mm) pt_cp_async8(Bs + c * OLDT + r, H + (rj + r) + (kc + c) * ld);
                        else Bs[c * OLDT + r] = 0.0;
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            int q = 0;
            while (q < nown && own[q].y <= k) ++q;
            int buf = 0;
            if (q < nown) fetch(q, 0);
            while (q < nown) {
                int qn = q + 1;
                while (qn < nown && own[qn].y <= k) ++qn;
                if (qn < nown) fetch(qn, buf ^ 1);
                const int ti = own[q].x, tj = own[q].y;
                const long long ri = (long long)TT * ti, rj = (long long)TT * tj;
                const double *As = W + buf * (2 * TT * OLDT), *Bs = As + TT * OLDT;
                const double *Bt = (ti != tj) ? Bs : As;
                double acc[4][4];
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const long long row = ri + r0 + g + 8 * (v >> 1), col = rj + c0w + b * 8 + 2 * t + (v & 1);
                        acc[b][v] = (row < mm && col < mm && row >= col) ? H[row + col * ld] : 0.0;
                    }
                if (qn < nown) asm volatile("cp.async.wait_group 1;" ::: "memory");
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();
                for (int kk = 0; kk < kb8; kk += 8) {
                    double af[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) af[v] = -As[(kk + t + 4 * (v >> 1)) * OLDT + r0 + g + 8 * (v & 1)];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        double bf[2];
#pragma unroll
                        for (int v = 0; v < 2; ++v) bf[v] = Bt[(kk + t + 4 * v) * OLDT + c0w + b * 8 + g];
                        dmma16x8x8_t(acc[b], af, bf);
                    }
                }
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const long long row = ri + r0 + g + 8 * (v >> 1), col = rj + c0w + b * 8 + 2 * t + (v & 1);
                        if (row < mm && col < mm && row >= col) H[row + col * ld] = acc[b][v];
                    }
                __syncthreads();            // this half of W is refilled by the fetch of the tile after next
                q = qn;
                buf ^= 1;
            }
        }
        PT_STAMP(4);
        if (k + 1 < Pc) grid_barrier(a.bar, bar_target);
        PT_STAMP(5);
    }
}

static size_t potrf_tile_smem() { return (size_t)(TT * DLD + TT + 2 * 16 * TT + 4 * TT * OLDT) * sizeof(double); }

// Largest order handled by one launch (tiles per CTA bounded by PT_MAXOWN)
bool potrf_tile_fits(const smcp_ctx *ctx, int64_t mm, int64_t npiv, bool panel_only) {
    const int64_t P = (mm + TT - 1) / TT, Pc = (npiv + TT - 1) / TT, ntc = panel_only ? Pc : P;
    const int64_t ntiles = ntc * P - ntc * (ntc - 1) / 2;
    return P < 32000 && ntiles <= (int64_t)PT_MAXOWN * ctx->num_sms;
}

int potrf_tile(smcp_ctx *ctx, double *H, int64_t ld, int64_t mm, int64_t npiv, bool panel_only, int32_t *info_dev, int col_off) {
    if (mm <= 0 || npiv <= 0) return 0;
    if (npiv > mm) npiv = mm;
    static bool attr = false;
    const size_t smem = potrf_tile_smem();
    if (!attr) {
        CUDA_TRY(cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    if (!potrf_tile_fits(ctx, mm, npiv, panel_only)) { smcp_set_error("potrf_tile: matrix too large for one launch"); return -2; }
    PotrfTileArgs a;
    a.H = H; a.ld = ld; a.mm = (int)mm; a.npiv = (int)npiv;
    const int64_t P = (mm + TT - 1) / TT, Pc = (npiv + TT - 1) / TT;
    a.ntc = (int)(panel_only ? Pc : P);
    a.info = info_dev; a.col_off = col_off;
    a.dbg = nullptr;
    // grid-barrier counter: a ring of 64 counters so that launches in flight on different streams never share one
    if (!ctx->gridbar) {
        CUDA_TRY(cudaMalloc(&ctx->gridbar, 64 * 64));
        CUDA_TRY(cudaMemset(ctx->gridbar, 0, 64 * 64));
    }
    a.bar = ctx->gridbar + 16 * (ctx->gridbar_next++ & 63);
    CUDA_TRY(cudaMemsetAsync(a.bar, 0, sizeof(unsigned), ctx->stream));
    static const bool dbg_on = getenv("SMCP_B200_PT_DEBUG") != nullptr;
    long long *dbg_dev = nullptr;
    if (dbg_on) {
        CUDA_TRY(cudaMalloc(&dbg_dev, (size_t)6 * Pc * sizeof(long long)));
        CUDA_TRY(cudaMemsetAsync(dbg_dev, 0, (size_t)6 * Pc * sizeof(long long), ctx->stream));
        a.dbg = dbg_dev;
    }
    const int64_t ntiles = (int64_t)a.ntc * P - (int64_t)a.ntc * (a.ntc - 1) / 2;
    // between lanes_fork and lanes_join several factorisations run side by side: each takes its share of the SMs
    // (a CTA then owns more tiles; the arithmetic of a tile does not depend on who owns it)
    int64_t gcap = ctx->num_sms;
    if (ctx->lanes_active > 1) gcap = std::max<int64_t>(ctx->num_sms / ctx->lanes_active, (ntiles + PT_MAXOWN - 1) / PT_MAXOWN);
    if (ctx->potrf_grid_cap > 0) gcap = std::max<int64_t>(std::min<int64_t>(gcap, ctx->potrf_grid_cap), (ntiles + PT_MAXOWN - 1) / PT_MAXOWN);
    const unsigned grid = (unsigned)std::min<int64_t>(std::min<int64_t>(gcap, ctx->num_sms), ntiles);
    void *args[] = {&a};
    LaunchScope ls(ctx, "potrf_tile", 1, (double)npiv * npiv * npiv / 3.0);
    CUDA_TRY(cudaLaunchCooperativeKernel((void *)potrf_tile_kernel, dim3(grid), dim3(PT_THREADS), args, smem, ctx->stream));
    if (dbg_on) {
        std::vector<long long> h((size_t)6 * Pc);
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaMemcpy(h.data(), dbg_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(dbg_dev);
        double ph[5] = {0, 0, 0, 0, 0};
        for (int64_t k = 0; k < Pc; ++k) {
            if (!h[6 * k + 1]) h[6 * k + 1] = h[6 * k];          // CTA 0 had no tile in this block column
            for (int q = 0; q < 5; ++q) ph[q] += (double)(h[6 * k + q + 1] - h[6 * k + q]) * 1e-3;
        }
        fprintf(stderr, "[potrf_tile m=%lld npiv=%lld grid=%u] CTA0 us: diag %.1f panel %.1f sync1 %.1f update %.1f sync2 %.1f total %.1f (%lld steps)\n",
                (long long)mm, (long long)npiv, grid, ph[0], ph[1], ph[2], ph[3], ph[4], (double)(h[6 * (Pc - 1) + 5] - h[0]) * 1e-3, (long long)Pc);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// slab triangular solve
// ---------------------------------------------------------------------------------------
#define SL_NC 8
#define SL_THREADS 256
#define SL_DLD 65

template <bool TRANS>
__global__ void __launch_bounds__(SL_THREADS, 1)
trsm_slab_kernel(const double *__restrict__ L, long long ldl, int n, double *__restrict__ B, long long ldb, long long nrhs, int ldS) {
    extern __shared__ __align__(16) double slsm[];
    double *D = slsm;                    // 64 x 65: D[c*65 + r] = L11(r, c)
    double *rinv = D + TT * SL_DLD;      // 64
    double *S = rinv + TT;               // 8 x ldS: S[c*ldS + r]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const long long c0 = (long long)blockIdx.x * SL_NC;
    const int ncol = (int)min((long long)SL_NC, nrhs - c0);
    {
        double *s = S + warp * ldS;
        const double *b = B + (c0 + warp) * ldb;
        for (int r = lane; r < ldS; r += 32) s[r] = (warp < ncol && r < n) ? b[r] : 0.0;
    }
    const int nb = (n + TT - 1) / TT;
    for (int bq = 0; bq < nb; ++bq) {
        const int bi = TRANS ? nb - 1 - bq : bq;
        const int k0 = bi * TT, kb = min(TT, n - k0);
        __syncthreads();
        for (int idx = tid; idx < TT * TT; idx += SL_THREADS) {
            const int r = idx & 63, c = idx >> 6;
            D[c * SL_DLD + r] = (r >= c && r < kb) ? L[(k0 + r) + (long long)(k0 + c) * ldl] : 0.0;
        }
        __syncthreads();
        if (tid < TT) rinv[tid] = tid < kb ? 1.0 / D[tid * SL_DLD + tid] : 1.0;
        __syncthreads();
        {
            // one warp per right-hand side: substitution inside the 64 x 64 block
            double *x = S + warp * ldS + k0;
            double x0 = x[lane], x1 = x[lane + 32];
            if (!TRANS) {
                for (int j = 0; j < kb; ++j) {
                    const double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31) * rinv[j];
                    if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
                    if (lane > j) x0 = fma(-D[j * SL_DLD + lane], xj, x0);
                    if (lane + 32 > j) x1 = fma(-D[j * SL_DLD + lane + 32], xj, x1);
                }
            } else {
                for (int j = kb - 1; j >= 0; --j) {
                    const double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31) * rinv[j];
                    if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
                    if (lane < j) x0 = fma(-D[lane * SL_DLD + j], xj, x0);
                    if (lane + 32 < j) x1 = fma(-D[(lane + 32) * SL_DLD + j], xj, x1);
                }
            }
            x[lane] = x0;
            x[lane + 32] = x1;
        }
        __syncthreads();
        // rest of the block column (forward: rows below; transposed: rows above) with DMMA
        const int kb8 = (kb + 7) & ~7;
        if (!TRANS) {
            const int rbeg = k0 + TT;
            if (rbeg < n) {
                const int nstrips = (n - rbeg + 15) / 16;
                for (int s = warp; s < nstrips; s += SL_THREADS / 32) {
                    const int row0 = rbeg + 16 * s;
                    double af[8][4];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int row = row0 + g + 8 * (v & 1);
                            af[q][v] = row < n ? -L[row + (long long)(k0 + 8 * q + t + 4 * (v >> 1)) * ldl] : 0.0;
                        }
                    double c[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) c[v] = S[(2 * t + (v & 1)) * ldS + row0 + g + 8 * (v >> 1)];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        double bf[2];
#pragma unroll
                        for (int v = 0; v < 2; ++v) bf[v] = S[g * ldS + k0 + 8 * q + t + 4 * v];
                        dmma16x8x8_t(c, af[q], bf);
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v) S[(2 * t + (v & 1)) * ldS + row0 + g + 8 * (v >> 1)] = c[v];
                }
            }
        } else if (k0 > 0) {
            const int nstrips = k0 / 16;
            for (int s = warp; s < nstrips; s += SL_THREADS / 32) {
                const int i0 = 16 * s;
                double c[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) c[v] = S[(2 * t + (v & 1)) * ldS + i0 + g + 8 * (v >> 1)];
                for (int kk = 0; kk < kb8; kk += 8) {
                    double af[4], bf[2];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int kr = kk + t + 4 * (v >> 1);
                        af[v] = kr < kb ? -L[(k0 + kr) + (long long)(i0 + g + 8 * (v & 1)) * ldl] : 0.0;
                    }
#pragma unroll
                    for (int v = 0; v < 2; ++v) bf[v] = S[g * ldS + k0 + kk + t + 4 * v];
                    dmma16x8x8_t(c, af, bf);
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) S[(2 * t + (v & 1)) * ldS + i0 + g + 8 * (v >> 1)] = c[v];
            }
        }
    }
    __syncthreads();
    if (warp < ncol) {
        const double *s = S + warp * ldS;
        double *b = B + (c0 + warp) * ldb;
        for (int r = lane; r < n; r += 32) b[r] = s[r];
    }
}

#define SL_NMAX 2816      // 8 x (2816 + 4) doubles of slab + the diagonal block = 214 KB of shared memory

bool trsm_slab_fits(int64_t n) { return n <= SL_NMAX; }

int trsm_slab(smcp_ctx *ctx, bool trans, const double *L, int64_t ldl, int64_t n, double *B, int64_t ldb, int64_t nrhs) {
    if (n <= 0 || nrhs <= 0) return 0;
    const int ldS = (int)(((n + 15) / 16) * 16 + 64 + 4);      // = 4 (mod 16); 64 extra rows: the last block reads x[lane + 32]
    const size_t smem = (size_t)(TT * SL_DLD + TT + (size_t)SL_NC * ldS) * sizeof(double);
    static size_t attr[2] = {0, 0};
    if (smem > attr[trans]) {
        if (trans) CUDA_TRY(cudaFuncSetAttribute(trsm_slab_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else CUDA_TRY(cudaFuncSetAttribute(trsm_slab_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[trans] = smem;
    }
    const long long nct = (nrhs + SL_NC - 1) / SL_NC;
    LaunchScope ls(ctx, "trsm_slab", 1, (double)n * n * nrhs);
    for (long long b0 = 0; b0 < nct; b0 += 1 << 30) {
        const unsigned grid = (unsigned)std::min<long long>(1 << 30, nct - b0);
        if (trans) trsm_slab_kernel<true><<<grid, SL_THREADS, smem, ctx->stream>>>(L, ldl, (int)n, B + b0 * SL_NC * ldb, ldb, nrhs - b0 * SL_NC, ldS);
        else trsm_slab_kernel<false><<<grid, SL_THREADS, smem, ctx->stream>>>(L, ldl, (int)n, B + b0 * SL_NC * ldb, ldb, nrhs - b0 * SL_NC, ldS);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

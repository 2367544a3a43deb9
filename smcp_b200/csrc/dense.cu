// Dense FP64 kernels for the m x m Schur complement H: tensor-core GEMM (DMMA), blocked
// Cholesky (lapack.potrf, src/python/solvers.py:501, 1931) and the two triangular solves of
// lapack.potrs (solvers.py:526, 1954).
//
// FP64 has no tcgen05/UMMA kind on sm_100a; the FP64 tensor path is mma.sync.m8n8k4.f64
// (SASS: DMMA.8x8x4).  The GEMM below stages K-major operand tiles in shared memory with a
// 3-stage cp.async pipeline, pads the tile rows so the per-thread 8-byte fragment loads
// are bank-conflict free, and keeps a 32x32 accumulator tile per warp in registers (16 warps per CTA).
#include "internal.cuh"
#include <algorithm>

// ---------------------------------------------------------------------------------------
// DMMA GEMM:  C(i, j) = [C(i, j) +] alpha * sum_k A'(i, k) B'(j, k)
//   TA = true : A'(i, k) = A[k + i*lda] (K contiguous)     TA = false: A'(i, k) = A[i + k*lda]
//   TB = true : B'(j, k) = B[k + j*ldb] (K contiguous)     TB = false: B'(j, k) = B[j + k*ldb]
//   (true, true) = A^T B: Schur assembly; (false, false) = A B^T: SYRK-like updates;
//   (false, true) = A B and (true, false) = A^T B^T: frontal updates of large supernodes
// tri: only tiles that intersect {i + tri_off >= j} are computed (lower triangle).
// ---------------------------------------------------------------------------------------
#define BM 128
#define BN 128
#define BK 16
#define LDK (BK + 4)       // 20 doubles: rows shifted by 4 banks -> conflict-free fragments
#define STAGES 3
#define GEMM_THREADS 512       // 16 warps: 4 x 4 warp tiles of 32 x 32 (4 warps per scheduler keep the DMMA pipe fed)

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool TN>
__device__ __forceinline__ void load_tile(double *sm, const double *G, long long ld, long long row0,
                                          long long nrows, long long k0, long long K, int tid) {
    // fills sm[r*LDK + k] for r in [0,128), k in [0,BK) with op(G)(row0+r, k0+k) or 0
    if (TN) {
        // G(k, r) at G[k + r*ld]; consecutive threads walk k (contiguous)
        for (int idx = tid; idx < 128 * BK; idx += GEMM_THREADS) {
            int k = idx % BK, r = idx / BK;
            double *dst = sm + r * LDK + k;
            if (row0 + r < nrows && k0 + k < K) cp_async8(dst, G + (k0 + k) + (row0 + r) * ld);
            else *dst = 0.0;
        }
    } else {
        // G(r, k) at G[r + k*ld]; consecutive threads walk r (contiguous)
        for (int idx = tid; idx < 128 * BK; idx += GEMM_THREADS) {
            int r = idx % 128, k = idx / 128;
            double *dst = sm + r * LDK + k;
            if (row0 + r < nrows && k0 + k < K) cp_async8(dst, G + (row0 + r) + (k0 + k) * ld);
            else *dst = 0.0;
        }
    }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_dmma_kernel(const double *__restrict__ A, long long lda, const double *__restrict__ B, long long ldb,
                 double *__restrict__ C, long long ldc, long long M, long long N, long long K, double alpha,
                 int accumulate, int tri, long long tri_off, long long kchunk, long long split_stride) {
    extern __shared__ double smem[];
    double *As = smem;                              // STAGES x 128 x LDK
    double *Bs = smem + STAGES * 128 * LDK;
    const long long i0 = (long long)blockIdx.x * BM;
    const long long j0 = (long long)blockIdx.y * BN;
    if (tri && (i0 + BM - 1 + tri_off < j0)) return;        // tile entirely above the diagonal
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 3) * 32;       // 4 warps along M
    const int wn = (warp >> 2) * 32;      // 4 warps along N
    const int g = lane >> 2, t = lane & 3;

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    // split-K: slice blockIdx.z works on k in [kbeg, kend) and writes its own partial result
    const long long kbeg = (long long)blockIdx.z * kchunk;
    const long long kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;
    C += (long long)blockIdx.z * split_stride;
    const long long nk = (kend - kbeg + BK - 1) / BK;
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            load_tile<TA>(As + s * 128 * LDK, A, lda, i0, M, kbeg + (long long)s * BK, kend, tid);
            load_tile<TB>(Bs + s * 128 * LDK, B, ldb, j0, N, kbeg + (long long)s * BK, kend, tid);
        }
        cp_async_commit();
    }
    for (long long kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            long long nx = kt + STAGES - 1;
            if (nx < nk) {
                int s = (int)(nx % STAGES);
                load_tile<TA>(As + s * 128 * LDK, A, lda, i0, M, kbeg + nx * BK, kend, tid);
                load_tile<TB>(Bs + s * 128 * LDK, B, ldb, j0, N, kbeg + nx * BK, kend, tid);
            }
            cp_async_commit();
        }
        const double *as = As + (kt % STAGES) * 128 * LDK;
        const double *bs = Bs + (kt % STAGES) * 128 * LDK;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = as[(wm + a * 8 + g) * LDK + kk + t];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = bs[(wn + b * 8 + g) * LDK + kk + t];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
    cp_async_wait<0>();
    // epilogue: C(i, j), i = i0 + wm + a*8 + g, j = j0 + wn + b*8 + 2t + {0,1}
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        long long i = i0 + wm + a * 8 + g;
        if (i >= M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                long long j = j0 + wn + b * 8 + 2 * t + e;
                if (j >= N) continue;
                if (tri && i + tri_off < j) continue;
                double *c = C + i + j * ldc;
                double v = alpha * acc[a][b][e];
                *c = accumulate ? (*c + v) : v;
            }
        }
    }
}

// sum of the split-K partial results in a fixed order (deterministic), lower part only when tri
__global__ void splitk_reduce_kernel(const double *__restrict__ P, long long split_stride, int splits, double *__restrict__ C,
                                     long long ldc, long long M, long long N, int tri, long long tri_off) {
    const long long total = M * N;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx % M, j = idx / M;
        if (tri && i + tri_off < j) continue;
        double s = 0.0;
        for (int z = 0; z < splits; ++z) s += P[(long long)z * split_stride + idx];
        C[i + j * ldc] = s;
    }
}

int launch_gemm(smcp_ctx *ctx, bool ta, bool tb, const double *A, int64_t lda, const double *B, int64_t ldb,
                       double *C, int64_t ldc, int64_t M, int64_t N, int64_t K, double alpha, int accumulate,
                       int tri, int64_t tri_off, const char *name) {
    if (M <= 0 || N <= 0) return 0;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
    // split-K when the (lower-triangular) tile grid cannot fill the GPU and K is long: the Schur
    // assembly contracts over |blkval| ~ 10^4..10^6 rows into an m x m block
    long long ntiles = 0;
    for (unsigned bj = 0; bj < grid.y; ++bj)
        for (unsigned bi = 0; bi < grid.x; ++bi)
            if (!tri || (long long)bi * BM + BM - 1 + tri_off >= (long long)bj * BN) ++ntiles;
    int splits = 1;
    if (!accumulate && alpha == 1.0 && K >= 2048 && ntiles > 0 && ntiles < ctx->num_sms) {
        splits = (int)(ctx->num_sms / ntiles);
        if (splits > 16) splits = 16;
        if ((long long)splits * 512 > K) splits = (int)(K / 512);
        if (splits < 1) splits = 1;
    }
    long long kchunk = K, split_stride = 0;
    double *Cout = C;
    int64_t ldout = ldc;
    if (splits > 1) {
        kchunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
        split_stride = M * N;
        if (grow((void **)&ctx->gemm_ws, &ctx->gemm_ws_cap, (size_t)splits * M * N * sizeof(double))) return -1;
        Cout = ctx->gemm_ws;
        ldout = M;
        grid.z = splits;
    }
    size_t smem = (size_t)2 * STAGES * 128 * LDK * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    // algorithmic flops: 2*K per computed entry of the (lower-triangular) result
    double pairs = 0.0;
    if (!tri) pairs = (double)M * (double)N;
    else
        for (int64_t j = 0; j < N; ++j) {
            int64_t lo = j - tri_off;            // rows i >= lo
            if (lo < 0) lo = 0;
            if (lo < M) pairs += (double)(M - lo);
        }
    {
        LaunchScope ls(ctx, name, 1, 2.0 * (double)K * pairs);
#define GEMM_LAUNCH(TA_, TB_) gemm_dmma_kernel<TA_, TB_><<<grid, GEMM_THREADS, smem, ctx->stream>>>(A, lda, B, ldb, Cout, ldout, M, N, K, alpha, accumulate, tri, tri_off, kchunk, split_stride)
        if (ta && tb) GEMM_LAUNCH(true, true);
        else if (!ta && !tb) GEMM_LAUNCH(false, false);
        else if (ta) GEMM_LAUNCH(true, false);
        else GEMM_LAUNCH(false, true);
#undef GEMM_LAUNCH
        if (splits > 1) {
            ctx->launches += 1;
            long long g = (M * N + 255) / 256;
            if (g > (long long)ctx->num_sms * 8) g = (long long)ctx->num_sms * 8;
            splitk_reduce_kernel<<<(unsigned)g, 256, 0, ctx->stream>>>(ctx->gemm_ws, split_stride, splits, C, ldc, M, N, tri, tri_off);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// C(i, j) = sum_k A(k, i) B(k, j) for i + row_lo_of_col0 >= j  (lower part of a column block of H)
int d_gemm_tn(smcp_ctx *ctx, const double *A, int64_t lda, const double *B, int64_t ldb, double *C,
              int64_t ldc, int64_t M, int64_t N, int64_t K, int64_t row_lo_of_col0) {
    return launch_gemm(ctx, true, true, A, lda, B, ldb, C, ldc, M, N, K, 1.0, 0, 1, row_lo_of_col0, "schur_gemm_dmma");
}

// ---------------------------------------------------------------------------------------
// blocked right-looking Cholesky (lower), column-major, in place (lapack.potrf, solvers.py:501)
//
// Per 64-column panel: ONE kernel factors the diagonal block and solves the panel below it, then
// the DMMA kernel applies the trailing update.  At m ~ 10^3 the factorisation is a chain of tiny
// dependent steps, so the diagonal block is factored with one ROW PER THREAD IN REGISTERS (64
// threads, two barriers per column, fully unrolled: ~8 us instead of ~65 us through shared
// memory) and every CTA of the panel kernel redoes that factorisation itself instead of waiting
// for another launch to publish it.  // ---------------------------------------------------------------------------------------
#define NB 64
#define LDT 66          // row stride of the transposed factor in shared memory

#define PP_THREADS 256
#define PP_ROWS 128       // panel rows per CTA

// Diagonal block: 16 x 16 threads, each owns a 4 x 4 register block of the 64 x 64 matrix; per
// column one barrier to publish the pivot column, one after 64 threads scaled it, then 16
// predicated FMAs per thread.  Loops stay rolled (the first version unrolled everything and was
// bound by instruction fetch: ncu showed stall_no_inst on 150 us launches).
// Panel: X L^T = B, one row per thread, the row lives in shared memory (column-major, conflict
// free), left-looking over 8-column blocks held in registers; L is read as broadcast double2.
__global__ void __launch_bounds__(PP_THREADS) potrf_panel_kernel(double *H, long long ld, int kb, long long k0, long long m, int *info) {
    extern __shared__ __align__(16) double ppsm[];
    double *LT = ppsm;                       // NB x LDT: LT[c*LDT + r] = L(r, c), zero elsewhere
    double *colbuf = LT + NB * LDT;          // 2 x NB
    double *pivot = colbuf + 2 * NB;         // 2 (+ pad)
    double *xs = pivot + 2;                  // NB x PP_ROWS
    const int tid = threadIdx.x;
    double *D = H + k0 + k0 * ld;
    const int br = tid >> 4, bc = tid & 15;
    double a[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int l2 = 0; l2 < 4; ++l2) {
            const int r = 4 * br + k, c = 4 * bc + l2;
            a[k][l2] = (r < kb && c <= r) ? D[r + (long long)c * ld] : 0.0;
        }
    for (int idx = tid; idx < NB * LDT; idx += PP_THREADS) LT[idx] = 0.0;
    for (int j = 0; j < kb; ++j) {
        double *cb = colbuf + (j & 1) * NB;
        const int jb = j >> 2, jj = j & 3;
        if (bc == jb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double v = (jj == 0) ? a[k][0] : (jj == 1) ? a[k][1] : (jj == 2) ? a[k][2] : a[k][3];
                cb[4 * br + k] = v;
                if (4 * br + k == j) pivot[j & 1] = v;
            }
        }
        __syncthreads();
        if (tid < NB) {
            double d = pivot[j & 1];
            const bool bad = !(d > 0.0);
            if (bad) d = 1.0;
            if (bad && tid == 0 && blockIdx.x == 0 && *info == 0) *info = (int)(k0 + j + 1);   // dpotrf's info
            const double sq = sqrt(d);
            const double v = cb[tid];
            cb[tid] = (tid == j) ? sq : ((tid > j) ? v / sq : 0.0);
        }
        __syncthreads();
        double lr[4], lc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            lr[k] = cb[4 * br + k];
            lc[k] = cb[4 * bc + k];
        }
        if (bc == jb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (jj == 0) a[k][0] = lr[k];
                else if (jj == 1) a[k][1] = lr[k];
                else if (jj == 2) a[k][2] = lr[k];
                else a[k][3] = lr[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int l2 = 0; l2 < 4; ++l2)
                if (4 * bc + l2 > j) a[k][l2] = fma(-lr[k], lc[l2], a[k][l2]);     // rows <= j have lr = 0 or are finished columns
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int l2 = 0; l2 < 4; ++l2) {
            const int r = 4 * br + k, c = 4 * bc + l2;
            if (r < kb && c <= r) {
                LT[c * LDT + r] = a[k][l2];
                if (blockIdx.x == 0) D[r + (long long)c * ld] = a[k][l2];
            }
        }
    __syncthreads();
    // ---- panel rows
    const long long row = k0 + kb + (long long)blockIdx.x * PP_ROWS + tid;
    if (tid >= PP_ROWS || row >= m) return;
    double *P = H + row + k0 * ld;
    for (int c = 0; c < kb; ++c) xs[c * PP_ROWS + tid] = P[(long long)c * ld];
    for (int cb8 = 0; cb8 < kb; cb8 += 8) {
        double x8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x8[k] = (cb8 + k < kb) ? xs[(cb8 + k) * PP_ROWS + tid] : 0.0;
        for (int p = 0; p < cb8; ++p) {
            const double xp = xs[p * PP_ROWS + tid];
            const double2 *lp = reinterpret_cast<const double2 *>(LT + p * LDT + cb8);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 lv = lp[k];
                x8[2 * k] = fma(-xp, lv.x, x8[2 * k]);
                x8[2 * k + 1] = fma(-xp, lv.y, x8[2 * k + 1]);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (cb8 + k < kb) {
                const double xk = x8[k] / LT[(cb8 + k) * LDT + cb8 + k];
                x8[k] = xk;
#pragma unroll
                for (int k2 = k + 1; k2 < 8; ++k2) x8[k2] = fma(-xk, LT[(cb8 + k) * LDT + cb8 + k2], x8[k2]);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (cb8 + k < kb) {
                xs[(cb8 + k) * PP_ROWS + tid] = x8[k];
                P[(long long)(cb8 + k) * ld] = x8[k];
            }
    }
}

int d_potrf(smcp_ctx *ctx, double *H, int64_t m, int32_t *info_dev, double *Dinv) {
    CUDA_TRY(cudaMemsetAsync(info_dev, 0, sizeof(int), ctx->stream));
    const size_t pp_smem = (size_t)(NB * LDT + 2 * NB + 2 + NB * PP_ROWS) * sizeof(double);
    static bool pp_attr = false;
    if (!pp_attr) {
        CUDA_TRY(cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp_smem));
        pp_attr = true;
    }
    for (int64_t k0 = 0; k0 < m; k0 += NB) {
        int kb = (int)((m - k0 < NB) ? (m - k0) : NB);
        int64_t rem = m - k0 - kb;
        {
            LaunchScope ls(ctx, "potrf_panel", 1, 16.0 * (double)(m - k0) * kb);
            unsigned grid = (unsigned)std::max<int64_t>(1, (rem + PP_ROWS - 1) / PP_ROWS);
            potrf_panel_kernel<<<grid, PP_THREADS, pp_smem, ctx->stream>>>(H, m, kb, k0, m, info_dev);
        }
        if (rem > 0) {
            const double *P = H + (k0 + kb) + k0 * m;
            double *Ct = H + (k0 + kb) + (k0 + kb) * m;
            if (launch_gemm(ctx, false, false, P, m, P, m, Ct, m, rem, rem, kb, -1.0, 1, 1, 0, "potrf_syrk_dmma")) return -1;
        }
    }
    (void)Dinv;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------
// potrs: y <- L^{-T} L^{-1} y (lapack.potrs, solvers.py:526), single right-hand side.
// A triangular solve is m dependent steps whatever the hardware, so this is a latency kernel:
// one CTA of 1024 threads, the right-hand side lives in shared memory for the whole solve, the
// 64 x 64 diagonal block is staged in shared memory, substitution inside a block is done by one
// warp (row values in registers, pivot broadcast by shuffle, multiplication by the reciprocal
// pivots computed off the critical path), and the off-diagonal part of every block column is a
// GEMV spread over all 32 warps with independent partial sums.  Plain substitution, no explicit
// inverses: backward stable like the LAPACK routine it replaces.
// ---------------------------------------------------------------------------------------
#define PS_THREADS 1024
__global__ void __launch_bounds__(PS_THREADS) potrs_kernel(const double *__restrict__ H, long long m, double *__restrict__ y) {
    extern __shared__ double psm[];
    double *Lk = psm;                    // NB x (NB+1)
    double *rinv = Lk + NB * (NB + 1);   // NB
    double *tb = rinv + NB;              // NB
    double *ys = tb + NB;                // m
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long nblkc = (m + NB - 1) / NB;
    for (long long i = tid; i < m; i += PS_THREADS) ys[i] = y[i];
    // forward: L x = y
    for (long long bi = 0; bi < nblkc; ++bi) {
        const long long k0 = bi * NB;
        const int kb = (int)min((long long)NB, m - k0);
        for (int idx = tid; idx < kb * kb; idx += PS_THREADS) {
            const int i = idx % kb, j = idx / kb;
            Lk[i * (NB + 1) + j] = H[(k0 + i) + (k0 + j) * m];
        }
        __syncthreads();
        if (tid < kb) rinv[tid] = 1.0 / Lk[tid * (NB + 1) + tid];
        __syncthreads();
        if (warp == 0) {
            double x0 = (lane < kb) ? ys[k0 + lane] : 0.0;
            double x1 = (lane + 32 < kb) ? ys[k0 + lane + 32] : 0.0;
            for (int j = 0; j < kb; ++j) {
                const double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31) * rinv[j];
                if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
                if (lane > j) x0 = fma(-Lk[lane * (NB + 1) + j], xj, x0);
                if (lane + 32 > j && lane + 32 < kb) x1 = fma(-Lk[(lane + 32) * (NB + 1) + j], xj, x1);
            }
            if (lane < kb) ys[k0 + lane] = x0;
            if (lane + 32 < kb) ys[k0 + lane + 32] = x1;
        }
        __syncthreads();
        for (long long i = k0 + kb + tid; i < m; i += PS_THREADS) {
            const double *Lr = H + i + k0 * m;
            const double *xk = ys + k0;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
            for (int j = 0; j < NB; j += 4) {       // here kb == NB (rows remain below the block)
                s0 = fma(Lr[(long long)j * m], xk[j], s0);
                s1 = fma(Lr[(long long)(j + 1) * m], xk[j + 1], s1);
                s2 = fma(Lr[(long long)(j + 2) * m], xk[j + 2], s2);
                s3 = fma(Lr[(long long)(j + 3) * m], xk[j + 3], s3);
            }
            ys[i] -= (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
    }
    // backward: L^T x = y
    for (long long bi = nblkc - 1; bi >= 0; --bi) {
        const long long k0 = bi * NB;
        const int kb = (int)min((long long)NB, m - k0);
        for (int idx = tid; idx < kb * kb; idx += PS_THREADS) {
            const int i = idx % kb, j = idx / kb;
            Lk[i * (NB + 1) + j] = H[(k0 + i) + (k0 + j) * m];
        }
        // t_c = y[k0+c] - sum_{i >= k0+kb} L(i, k0+c) y[i]: warp w takes columns w and w + 32
        for (int c = warp; c < kb; c += PS_THREADS / 32) {
            const double *Lc = H + (k0 + c) * m;
            double s0 = 0.0, s1 = 0.0;
            long long i = k0 + kb + lane;
            for (; i + 32 < m; i += 64) {
                s0 = fma(Lc[i], ys[i], s0);
                s1 = fma(Lc[i + 32], ys[i + 32], s1);
            }
            if (i < m) s0 = fma(Lc[i], ys[i], s0);
            double sacc = s0 + s1;
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_down_sync(0xffffffffu, sacc, o);
            if (lane == 0) tb[c] = ys[k0 + c] - sacc;
        }
        __syncthreads();
        if (tid < kb) rinv[tid] = 1.0 / Lk[tid * (NB + 1) + tid];
        __syncthreads();
        if (warp == 0) {
            double x0 = (lane < kb) ? tb[lane] : 0.0;
            double x1 = (lane + 32 < kb) ? tb[lane + 32] : 0.0;
            for (int j = kb - 1; j >= 0; --j) {
                const double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31) * rinv[j];
                if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
                if (lane < j) x0 = fma(-Lk[j * (NB + 1) + lane], xj, x0);
                if (lane + 32 < j) x1 = fma(-Lk[j * (NB + 1) + lane + 32], xj, x1);
            }
            if (lane < kb) ys[k0 + lane] = x0;
            if (lane + 32 < kb) ys[k0 + lane + 32] = x1;
        }
        __syncthreads();
    }
    for (long long i = tid; i < m; i += PS_THREADS) y[i] = ys[i];
}

int d_potrs(smcp_ctx *ctx, const double *H, int64_t m, const double *Dinv, double *y_dev) {
    (void)Dinv;
    const size_t smem = (size_t)(NB * (NB + 1) + 2 * NB + m + 8) * sizeof(double);
    if (smem > 200 * 1024) { smcp_set_error("potrs: m = %lld exceeds the single-CTA kernel (right-hand side in shared memory)", (long long)m); return -2; }
    static size_t attr = 0;
    if (smem > attr) {
        CUDA_TRY(cudaFuncSetAttribute(potrs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    {
        LaunchScope ls(ctx, "potrs", 1, 8.0 * (double)m * (double)m);
        potrs_kernel<<<1, PS_THREADS, smem, ctx->stream>>>(H, m, y_dev);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

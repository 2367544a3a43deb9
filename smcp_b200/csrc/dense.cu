// Dense FP64 kernels for the m x m Schur complement H: tensor-core GEMM (DMMA), blocked
// Cholesky (lapack.potrf, src/python/solvers.py:501, 1931) and the two triangular solves of
// lapack.potrs (solvers.py:526, 1954).
//
// FP64 has no tcgen05/UMMA kind on sm_100a; the FP64 tensor path is mma.sync.m8n8k4.f64
// (SASS: DMMA.8x8x4).  The GEMM below stages K-major operand tiles in shared memory with a
// 3-stage cp.async pipeline, pads the tile rows so the per-thread 8-byte fragment loads
// are bank-conflict free, and keeps a 64x32 accumulator tile per warp in registers.
#include "internal.cuh"

// ---------------------------------------------------------------------------------------
// DMMA GEMM:  C(MxN) = beta*C + alpha * op(A)^T op(B)
//   TN = true : A is K x M (column-major, K contiguous), B is K x N       (Schur assembly)
//   TN = false: A is M x K (column-major, M contiguous), B is N x K       (SYRK-like update)
// tri: only tiles that intersect {i + tri_off >= j} are computed (lower triangle).
// ---------------------------------------------------------------------------------------
#define BM 128
#define BN 128
#define BK 16
#define LDK (BK + 4)       // 20 doubles: rows shifted by 4 banks -> conflict-free fragments
#define STAGES 3
#define GEMM_THREADS 256

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

template <bool TN>
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
    const int wm = (warp & 1) * 64;       // 2 warps along M
    const int wn = (warp >> 1) * 32;      // 4 warps along N
    const int g = lane >> 2, t = lane & 3;

    double acc[8][4][2];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    // split-K: slice blockIdx.z works on k in [kbeg, kend) and writes its own partial result
    const long long kbeg = (long long)blockIdx.z * kchunk;
    const long long kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;
    C += (long long)blockIdx.z * split_stride;
    const long long nk = (kend - kbeg + BK - 1) / BK;
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            load_tile<TN>(As + s * 128 * LDK, A, lda, i0, M, kbeg + (long long)s * BK, kend, tid);
            load_tile<TN>(Bs + s * 128 * LDK, B, ldb, j0, N, kbeg + (long long)s * BK, kend, tid);
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
                load_tile<TN>(As + s * 128 * LDK, A, lda, i0, M, kbeg + nx * BK, kend, tid);
                load_tile<TN>(Bs + s * 128 * LDK, B, ldb, j0, N, kbeg + nx * BK, kend, tid);
            }
            cp_async_commit();
        }
        const double *as = As + (kt % STAGES) * 128 * LDK;
        const double *bs = Bs + (kt % STAGES) * 128 * LDK;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int a = 0; a < 8; ++a) af[a] = as[(wm + a * 8 + g) * LDK + kk + t];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = bs[(wn + b * 8 + g) * LDK + kk + t];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
    cp_async_wait<0>();
    // epilogue: C(i, j), i = i0 + wm + a*8 + g, j = j0 + wn + b*8 + 2t + {0,1}
#pragma unroll
    for (int a = 0; a < 8; ++a) {
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

static int launch_gemm(smcp_ctx *ctx, bool tn, const double *A, int64_t lda, const double *B, int64_t ldb,
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
        CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
        if (tn)
            gemm_dmma_kernel<true><<<grid, GEMM_THREADS, smem, ctx->stream>>>(A, lda, B, ldb, Cout, ldout, M, N, K, alpha, accumulate, tri, tri_off, kchunk, split_stride);
        else
            gemm_dmma_kernel<false><<<grid, GEMM_THREADS, smem, ctx->stream>>>(A, lda, B, ldb, Cout, ldout, M, N, K, alpha, accumulate, tri, tri_off, kchunk, split_stride);
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
    return launch_gemm(ctx, true, A, lda, B, ldb, C, ldc, M, N, K, 1.0, 0, 1, row_lo_of_col0, "schur_gemm_dmma");
}

// ---------------------------------------------------------------------------------------
// blocked right-looking Cholesky (lower), column-major, in place
// ---------------------------------------------------------------------------------------
#define NB 64

// factor the kb x kb diagonal block (one CTA); info = k0 + j + 1 at the first bad pivot
__global__ void potrf_diag_kernel(double *H, long long ld, int kb, long long k0, int *info) {
    __shared__ double T[NB][NB + 1];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int idx = tid; idx < kb * kb; idx += nt) {
        int i = idx % kb, j = idx / kb;
        T[i][j] = (i >= j) ? H[i + j * ld] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < kb; ++j) {
        double d = T[j][j];
        bool bad = !(d > 0.0);
        if (bad) {
            if (tid == 0 && *info == 0) *info = (int)(k0 + j + 1);
            d = 1.0;
        }
        double s = sqrt(d);
        __syncthreads();
        for (int i = j + tid; i < kb; i += nt) T[i][j] = (i == j) ? s : T[i][j] / s;
        __syncthreads();
        int nr = kb - j - 1;
        for (int idx = tid; idx < nr * nr; idx += nt) {
            int i = j + 1 + idx % nr, c = j + 1 + idx / nr;
            if (i >= c) T[i][c] = fma(-T[i][j], T[c][j], T[i][c]);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < kb * kb; idx += nt) {
        int i = idx % kb, j = idx / kb;
        if (i >= j) H[i + j * ld] = T[i][j];
    }
}

// panel rows: X L_kk^T = B, one row per thread
__global__ void potrf_trsm_kernel(double *H, long long ld, int kb, long long k0, long long m) {
    __shared__ double Lk[NB][NB + 1];
    const int tid = threadIdx.x;
    const double *D = H + k0 + k0 * ld;
    for (int idx = tid; idx < kb * kb; idx += blockDim.x) {
        int i = idx % kb, j = idx / kb;
        Lk[i][j] = (i >= j) ? D[i + j * ld] : 0.0;
    }
    __syncthreads();
    long long row = k0 + kb + (long long)blockIdx.x * blockDim.x + tid;
    if (row >= m) return;
    double *P = H + row + k0 * ld;
    double x[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) x[c] = (c < kb) ? P[c * ld] : 0.0;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        if (c < kb) {
            double s = x[c];
#pragma unroll
            for (int r = 0; r < c; ++r) s = fma(-x[r], Lk[c][r], s);
            x[c] = s / Lk[c][c];
        }
    }
#pragma unroll
    for (int c = 0; c < NB; ++c)
        if (c < kb) P[c * ld] = x[c];
}

int d_potrf(smcp_ctx *ctx, double *H, int64_t m, int32_t *info_dev) {
    CUDA_TRY(cudaMemsetAsync(info_dev, 0, sizeof(int), ctx->stream));
    for (int64_t k0 = 0; k0 < m; k0 += NB) {
        int kb = (int)((m - k0 < NB) ? (m - k0) : NB);
        {
            LaunchScope ls(ctx, "potrf_diag");
            potrf_diag_kernel<<<1, 256, 0, ctx->stream>>>(H + k0 + k0 * m, m, kb, k0, info_dev);
        }
        int64_t rem = m - k0 - kb;
        if (rem > 0) {
            {
                LaunchScope ls(ctx, "potrf_trsm");
                potrf_trsm_kernel<<<(unsigned)((rem + 127) / 128), 128, 0, ctx->stream>>>(H, m, kb, k0, m);
            }
            const double *P = H + (k0 + kb) + k0 * m;
            double *Ct = H + (k0 + kb) + (k0 + kb) * m;
            if (launch_gemm(ctx, false, P, m, P, m, Ct, m, rem, rem, kb, -1.0, 1, 1, 0, "potrf_syrk_dmma")) return -1;
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------
// potrs: y <- L^{-T} L^{-1} y, single right-hand side, one persistent CTA
// ---------------------------------------------------------------------------------------
#define PS_THREADS 512
__global__ void __launch_bounds__(PS_THREADS) potrs_kernel(const double *__restrict__ H, long long m, double *__restrict__ y) {
    __shared__ double Lk[NB][NB + 1];
    __shared__ double xb[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = PS_THREADS / 32;
    // forward: L x = y
    for (long long k0 = 0; k0 < m; k0 += NB) {
        int kb = (int)min((long long)NB, m - k0);
        for (int idx = tid; idx < kb * kb; idx += PS_THREADS) {
            int i = idx % kb, j = idx / kb;
            Lk[i][j] = H[(k0 + i) + (k0 + j) * m];
        }
        if (tid < kb) xb[tid] = y[k0 + tid];
        __syncthreads();
        if (warp == 0) {
            // lanes own rows lane and lane+32
            double x0 = (lane < kb) ? xb[lane] : 0.0;
            double x1 = (lane + 32 < kb) ? xb[lane + 32] : 0.0;
            for (int j = 0; j < kb; ++j) {
                double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31);
                xj = xj / Lk[j][j];
                if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
                if (lane > j && lane < kb) x0 = fma(-Lk[lane][j], xj, x0);
                if (lane + 32 > j && lane + 32 < kb) x1 = fma(-Lk[lane + 32][j], xj, x1);
            }
            if (lane < kb) xb[lane] = x0;
            if (lane + 32 < kb) xb[lane + 32] = x1;
        }
        __syncthreads();
        if (tid < kb) y[k0 + tid] = xb[tid];
        // y[k0+kb:] -= L[k0+kb:, k0:k0+kb] * xb
        for (long long i = k0 + kb + tid; i < m; i += PS_THREADS) {
            double s = y[i];
            const double *Lr = H + i + k0 * m;
            for (int j = 0; j < kb; ++j) s = fma(-Lr[j * m], xb[j], s);
            y[i] = s;
        }
        __syncthreads();
    }
    // backward: L^T x = y
    long long nblk = (m + NB - 1) / NB;
    for (long long bi = nblk - 1; bi >= 0; --bi) {
        long long k0 = bi * NB;
        int kb = (int)min((long long)NB, m - k0);
        // xb = y[k0:k0+kb] - L[k0+kb:, k0:k0+kb]^T y[k0+kb:]
        for (int j = warp; j < kb; j += nwarps) {
            const double *Lc = H + (k0 + j) * m;
            double s = 0.0;
            for (long long i = k0 + kb + lane; i < m; i += 32) s = fma(Lc[i], y[i], s);
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if (lane == 0) xb[j] = y[k0 + j] - s;
        }
        for (int idx = tid; idx < kb * kb; idx += PS_THREADS) {
            int i = idx % kb, j = idx / kb;
            Lk[i][j] = H[(k0 + i) + (k0 + j) * m];
        }
        __syncthreads();
        if (warp == 0) {
            double x0 = (lane < kb) ? xb[lane] : 0.0;
            double x1 = (lane + 32 < kb) ? xb[lane + 32] : 0.0;
            for (int j = kb - 1; j >= 0; --j) {
                double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31);
                xj = xj / Lk[j][j];
                if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
                if (lane < j) x0 = fma(-Lk[j][lane], xj, x0);
                if (lane + 32 < j) x1 = fma(-Lk[j][lane + 32], xj, x1);
            }
            if (lane < kb) xb[lane] = x0;
            if (lane + 32 < kb) xb[lane + 32] = x1;
        }
        __syncthreads();
        if (tid < kb) y[k0 + tid] = xb[tid];
        __syncthreads();
    }
}

int d_potrs(smcp_ctx *ctx, const double *H, int64_t m, double *y_dev) {
    {
        LaunchScope ls(ctx, "potrs");
        potrs_kernel<<<1, PS_THREADS, 0, ctx->stream>>>(H, m, y_dev);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

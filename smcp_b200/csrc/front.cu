// Large ROOT supernodes as dense multi-CTA kernels on the FP64 tensor cores.
//
// On patterns with non-trivial fill (rand_SDP, max-cut embeddings, mtxnorm) ~90 % of the flops of
// every chordal recursion sit in the root supernode of the clique tree (no separator, a dense
// nn x nn block with nn = 200..450 in the BASELINE configs).  The persistent tree kernels give a
// supernode to ONE CTA, which for a single matrix leaves 147 SMs idle for tens of milliseconds.
// For single matrices (Newton solves, line-search factorisations) the root is therefore taken
// out of the tree kernel and processed here with dense building blocks that use the whole GPU:
//   * DMMA GEMM (dense.cu, mma.sync m8n8k4 f64) for L L^T, D M D, L^{-T} L^{-1} and the
//     off-diagonal updates of the blocked triangular solves,
//   * blocked left triangular solves (64 x 64 diagonal blocks in shared memory, one right-hand
//     side per thread, 8-column register blocks),
//   * the blocked Cholesky of dense.cu (lapack.potrf replacement) for the root pivot block.
// Reference call sites of the routines: src/python/solvers.py:884 (cholesky), 874 (completion),
// 891 (projected_inverse), 904 (llt), 483/524/531 (hessian), 405 (inverse hessian).
// Children's update matrices are added with a precomputed inverse relative-index map so that every
// entry of the root front is summed by one thread in child order (deterministic, no atomics).
#include "internal.cuh"
#include <algorithm>
#include <cstdlib>

#define FNB 64
#define FLDT 66

// ---------------------------------------------------------------------------------------
// blocked triangular solves  X = L^{-1} B  /  X = L^{-T} B   (L lower, n x n, B n x nrhs)
// ---------------------------------------------------------------------------------------
// diagonal block: one right-hand side per thread, its kb values in shared memory (column-major
// over threads, conflict free), 8-row register blocks
template <bool TRANS>
__global__ void __launch_bounds__(128) trsm_diag_kernel(const double *__restrict__ L, long long ldl, int kb, double *B, long long ldb, long long nrhs) {
    extern __shared__ __align__(16) double tsm[];
    double *LT = tsm;                 // LT[c*FLDT + r] = L(r, c)
    double *xs = tsm + FNB * FLDT;    // kb x 128
    const int tid = threadIdx.x;
    for (int idx = tid; idx < FNB * FLDT; idx += 128) LT[idx] = 0.0;
    __syncthreads();
    for (int idx = tid; idx < kb * kb; idx += 128) {
        const int r = idx % kb, c = idx / kb;
        // forward solve reads columns of L (LT[c][r]); the transposed solve reads rows (LT[r][c])
        if (r >= c) LT[TRANS ? r * FLDT + c : c * FLDT + r] = L[r + (long long)c * ldl];
    }
    __syncthreads();
    const long long col = (long long)blockIdx.x * 128 + tid;
    if (col >= nrhs) return;
    double *P = B + col * ldb;
    for (int r = 0; r < kb; ++r) xs[r * 128 + tid] = P[r];
    if (!TRANS) {
        for (int c8 = 0; c8 < kb; c8 += 8) {
            double x8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x8[k] = (c8 + k < kb) ? xs[(c8 + k) * 128 + tid] : 0.0;
            for (int p = 0; p < c8; ++p) {
                const double xp = xs[p * 128 + tid];
                const double2 *lp = reinterpret_cast<const double2 *>(LT + p * FLDT + c8);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 lv = lp[k];
                    x8[2 * k] = fma(-xp, lv.x, x8[2 * k]);
                    x8[2 * k + 1] = fma(-xp, lv.y, x8[2 * k + 1]);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (c8 + k < kb) {
                    const double xk = x8[k] / LT[(c8 + k) * FLDT + c8 + k];
                    x8[k] = xk;
#pragma unroll
                    for (int k2 = k + 1; k2 < 8; ++k2) x8[k2] = fma(-xk, LT[(c8 + k) * FLDT + c8 + k2], x8[k2]);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (c8 + k < kb) {
                    xs[(c8 + k) * 128 + tid] = x8[k];
                    P[c8 + k] = x8[k];
                }
        }
    } else {
        // L^T x = b: backward over 8-row register blocks; LT[p*FLDT + c] = L(p, c)
        const int nb8 = (kb + 7) / 8;
        for (int b8 = nb8 - 1; b8 >= 0; --b8) {
            const int r8 = b8 * 8;
            double x8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x8[k] = (r8 + k < kb) ? xs[(r8 + k) * 128 + tid] : 0.0;
            for (int p = r8 + 8; p < kb; ++p) {
                const double xp = xs[p * 128 + tid];
                const double2 *lp = reinterpret_cast<const double2 *>(LT + p * FLDT + r8);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 lv = lp[k];
                    x8[2 * k] = fma(-xp, lv.x, x8[2 * k]);
                    x8[2 * k + 1] = fma(-xp, lv.y, x8[2 * k + 1]);
                }
            }
#pragma unroll
            for (int k = 7; k >= 0; --k) {
                if (r8 + k < kb) {
                    const double xk = x8[k] / LT[(r8 + k) * FLDT + r8 + k];
                    x8[k] = xk;
#pragma unroll
                    for (int k2 = 0; k2 < k; ++k2) x8[k2] = fma(-xk, LT[(r8 + k) * FLDT + r8 + k2], x8[k2]);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (r8 + k < kb) {
                    xs[(r8 + k) * 128 + tid] = x8[k];
                    P[r8 + k] = x8[k];
                }
        }
    }
}

int d_trsm_left_lower(smcp_ctx *ctx, bool trans, const double *L, int64_t ldl, int64_t n, double *B, int64_t ldb, int64_t nrhs) {
    if (n <= 0 || nrhs <= 0) return 0;
    const size_t smem = (size_t)(FNB * FLDT + FNB * 128) * sizeof(double);
    static bool attr = false;
    if (!attr) {
        CUDA_TRY(cudaFuncSetAttribute(trsm_diag_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(trsm_diag_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    const unsigned grid = (unsigned)((nrhs + 127) / 128);
    const int64_t nb = (n + FNB - 1) / FNB;
    // right-looking: after a 64-row block is solved the whole remainder is updated by ONE GEMM
    // with K = 64 and (rows left) x nrhs outputs, which fills the GPU; the left-looking form
    // (K growing, 64 output rows) runs on nrhs/128 CTAs and was 10x slower at n = 1186
    if (!trans) {
        for (int64_t bi = 0; bi < nb; ++bi) {
            const int64_t i0 = bi * FNB;
            const int kb = (int)std::min<int64_t>(FNB, n - i0);
            {
                LaunchScope ls(ctx, "front_trsm_diag");
                trsm_diag_kernel<false><<<grid, 128, smem, ctx->stream>>>(L + i0 + i0 * ldl, ldl, kb, B + i0, ldb, nrhs);
            }
            const int64_t below = n - i0 - kb;
            if (below > 0) {
                // B(i0+kb:n, :) -= L(i0+kb:n, i0:i0+kb) X(i0:i0+kb, :)
                if (launch_gemm(ctx, false, true, L + (i0 + kb) + i0 * ldl, ldl, B + i0, ldb, B + i0 + kb, ldb, below, nrhs, kb, -1.0, 1, 0, 0, "front_gemm_dmma")) return -1;
            }
        }
    } else {
        for (int64_t bi = nb - 1; bi >= 0; --bi) {
            const int64_t i0 = bi * FNB;
            const int kb = (int)std::min<int64_t>(FNB, n - i0);
            {
                LaunchScope ls(ctx, "front_trsm_diag");
                trsm_diag_kernel<true><<<grid, 128, smem, ctx->stream>>>(L + i0 + i0 * ldl, ldl, kb, B + i0, ldb, nrhs);
            }
            if (i0 > 0) {
                // B(0:i0, :) -= L(i0:i0+kb, 0:i0)^T X(i0:i0+kb, :)
                if (launch_gemm(ctx, true, true, L + i0, ldl, B + i0, ldb, B, ldb, i0, nrhs, kb, -1.0, 1, 0, 0, "front_gemm_dmma")) return -1;
            }
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------
// elementwise kernels on the root front
// ---------------------------------------------------------------------------------------
struct RootArgs {
    int nn, nch;
    const int *ch;          // children of the root
    const int *inv;         // nch x nn: position in the child's separator of root row p, or -1
    const int *na;          // per supernode
    const long long *updptr;
    const double *ub;       // update matrices of this matrix
};

enum { RA_LOWER = 0, RA_FULL = 1 };

// sum over the children of Uc(a, b) at root position (i, j); lower-stored or full update matrices
__device__ __forceinline__ double root_children(const RootArgs &r, int i, int j, bool lower_stored) {
    double acc = 0.0;
    for (int q = 0; q < r.nch; ++q) {
        const int a = r.inv[q * r.nn + i], b = r.inv[q * r.nn + j];
        if (a >= 0 && b >= 0) {
            const int c = r.ch[q], nac = r.na[c];
            const double *Uc = r.ub + r.updptr[c];
            acc += lower_stored ? Uc[max(a, b) + (long long)min(a, b) * nac] : Uc[a + (long long)b * nac];
        }
    }
    return acc;
}

// MODE 0: blk(lower) += children (lower), upper <- 0                         (cholesky front)
// MODE 1: F(full) = sym(blk lower) + children (full)                          (hessian pass 1)
// MODE 2: blk(lower) = T(lower) + children (lower), upper <- 0                (llt)
// MODE 3: blk(lower) = 0.5 (T(i,j) + ch(i,j) + T(j,i) + ch(j,i)), upper <- 0  (inverse hessian; children optional)
// MODE 4: F(full) = sym(blk lower)                                            (inverse hessian input)
// MODE 5: blk(lower) = 0.5 (T(i,j) + T(j,i)), upper <- 0                      (hessian scaling result)
// MODE 6: blk(lower) = T(lower), upper <- 0                                   (projected inverse / completion)
template <int MODE>
__global__ void root_elem_kernel(RootArgs r, double *blk, double *T) {
    const long long total = (long long)r.nn * r.nn;
    const int nn = r.nn;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % nn), j = (int)(idx / nn);
        if (MODE == 0) {
            blk[idx] = (i >= j) ? blk[idx] + root_children(r, i, j, true) : 0.0;
        } else if (MODE == 1) {
            const double base = (i >= j) ? blk[i + (long long)j * nn] : blk[j + (long long)i * nn];
            T[idx] = base + root_children(r, i, j, false);
        } else if (MODE == 2) {
            blk[idx] = (i >= j) ? T[idx] + root_children(r, i, j, true) : 0.0;
        } else if (MODE == 3) {
            if (i >= j) {
                const double a = T[i + (long long)j * nn] + root_children(r, i, j, false);
                const double b = T[j + (long long)i * nn] + root_children(r, j, i, false);
                blk[idx] = 0.5 * (a + b);
            } else blk[idx] = 0.0;
        } else if (MODE == 4) {
            T[idx] = (i >= j) ? blk[i + (long long)j * nn] : blk[j + (long long)i * nn];
        } else if (MODE == 5) {
            blk[idx] = (i >= j) ? 0.5 * (T[i + (long long)j * nn] + T[j + (long long)i * nn]) : 0.0;
        } else {
            blk[idx] = (i >= j) ? T[idx] : 0.0;
        }
    }
}

__global__ void transpose_kernel(const double *__restrict__ A, double *__restrict__ Bt, int n) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = bx + threadIdx.x, j = by + r;
        tile[r][threadIdx.x] = (i < n && j < n) ? A[i + (long long)j * n] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = by + threadIdx.x, j = bx + r;
        if (i < n && j < n) Bt[i + (long long)j * n] = tile[threadIdx.x][r];
    }
}

__global__ void set_identity_n(double *A, int n) {
    const long long total = (long long)n * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        A[idx] = (idx % n == idx / n) ? 1.0 : 0.0;
}

// completion: reversed symmetric copy of the lower-stored block, and M = (P Lc P)^T
__global__ void root_reverse_kernel(const double *__restrict__ blk, double *__restrict__ T, int n, int mode) {
    const long long total = (long long)n * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        if (mode == 0) {
            const int ri = n - 1 - i, rj = n - 1 - j;
            T[idx] = (ri >= rj) ? blk[ri + (long long)rj * n] : blk[rj + (long long)ri * n];
        } else {
            // M(i, j) = Lc(n-1-j, n-1-i) for i >= j
            T[idx] = (i >= j) ? blk[(n - 1 - j) + (long long)(n - 1 - i) * n] : 0.0;
        }
    }
}

__global__ void root_flag_kernel(const int *info, int *fail) {
    if (*info) *fail = 1;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int root_setup(smcp_sym *s, const smcp_sym_desc *D) {
    s->big_root = -1;
    const int nsn = (int)D->nsn;
    if (nsn < 1) return 0;
    const char *env = getenv("SMCP_B200_BIG_NN");
    const int thr = env ? atoi(env) : 96;
    if (thr <= 0) return 0;
    const int r = nsn - 1;
    const int nn = (int)(D->snptr[r + 1] - D->snptr[r]);
    const int nj = (int)(D->rowptr[r + 1] - D->rowptr[r]);
    if (D->snpar[r] != -1 || nj != nn || nn < thr) return 0;
    std::vector<int> ch;
    for (int64_t q = D->chptr[r]; q < D->chptr[r + 1]; ++q) ch.push_back((int)D->chidx[q]);
    std::vector<int> inv(std::max<size_t>(1, ch.size() * (size_t)nn), -1);
    for (size_t q = 0; q < ch.size(); ++q) {
        const int c = ch[q];
        const int nac = (int)(D->relptr[c + 1] - D->relptr[c]);
        for (int a = 0; a < nac; ++a) inv[q * nn + D->relidx[D->relptr[c] + a]] = a;
    }
    if (ch.empty()) ch.push_back(0);
    void *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, ch.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpy(d, ch.data(), ch.size() * sizeof(int), cudaMemcpyHostToDevice));
    s->allocs.push_back(d);
    s->root_ch = (const int *)d;
    CUDA_TRY(cudaMalloc(&d, inv.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpy(d, inv.data(), inv.size() * sizeof(int), cudaMemcpyHostToDevice));
    s->allocs.push_back(d);
    s->root_inv = (const int *)d;
    s->root_nch = (int)(D->chptr[r + 1] - D->chptr[r]);
    CUDA_TRY(cudaMalloc(&d, (size_t)3 * nn * nn * sizeof(double) + 64));
    s->allocs.push_back(d);
    s->root_ws = (double *)d;
    CUDA_TRY(cudaMalloc(&d, 64));
    s->allocs.push_back(d);
    s->root_info = (int *)d;
    s->big_root = r;
    s->root_nn = nn;
    return 0;
}

static RootArgs root_args(smcp_sym *s, int64_t b) {
    RootArgs r;
    r.nn = s->root_nn;
    r.nch = s->root_nch;
    r.ch = s->root_ch;
    r.inv = s->root_inv;
    r.na = s->d.na;
    r.updptr = s->d.updptr;
    r.ub = s->upd + (size_t)b * s->d.nupd;
    return r;
}

static unsigned elem_grid(smcp_sym *s) {
    long long g = ((long long)s->root_nn * s->root_nn + 255) / 256;
    return (unsigned)std::min<long long>(g, (long long)s->ctx->num_sms * 8);
}

#define ROOT_BLK(X, b) ((X) + (size_t)(b) * s->d.nblk + s->h_root_boff)

int root_cholesky(smcp_sym *s, double *X, int64_t b) {
    smcp_ctx *ctx = s->ctx;
    const int nn = s->root_nn;
    double *blk = ROOT_BLK(X, b);
    {
        LaunchScope ls(ctx, "front_elem");
        root_elem_kernel<0><<<elem_grid(s), 256, 0, ctx->stream>>>(root_args(s, b), blk, nullptr);
    }
    if (d_potrf(ctx, blk, nn, nn, nn, s->root_info, 0, 1)) return -1;
    root_flag_kernel<<<1, 1, 0, ctx->stream>>>(s->root_info, s->fail + b);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int root_llt(smcp_sym *s, double *X, int64_t b) {
    smcp_ctx *ctx = s->ctx;
    const int nn = s->root_nn;
    double *blk = ROOT_BLK(X, b), *T = s->root_ws;
    if (launch_gemm(ctx, false, false, blk, nn, blk, nn, T, nn, nn, nn, nn, 1.0, 0, 1, 0, "front_gemm_dmma")) return -1;
    {
        LaunchScope ls(ctx, "front_elem");
        root_elem_kernel<2><<<elem_grid(s), 256, 0, ctx->stream>>>(root_args(s, b), blk, T);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int root_transpose(smcp_sym *s, const double *A, double *Bt) {
    smcp_ctx *ctx = s->ctx;
    const int nn = s->root_nn;
    dim3 grid((nn + 31) / 32, (nn + 31) / 32), block(32, 8);
    LaunchScope ls(ctx, "front_elem");
    transpose_kernel<<<grid, block, 0, ctx->stream>>>(A, Bt, nn);
    return 0;
}

// forward Hessian, root part of pass 1 + scaling: M = D^{-1} (U_root + children) D^{-1}
int root_hess_up(smcp_sym *s, const double *Lt, double *X, int64_t b) {
    smcp_ctx *ctx = s->ctx;
    const int nn = s->root_nn;
    double *blk = ROOT_BLK(X, b), *T0 = s->root_ws, *T1 = s->root_ws + (size_t)nn * nn;
    const double *L = Lt + s->h_root_boff;
    {
        LaunchScope ls(ctx, "front_elem");
        root_elem_kernel<1><<<elem_grid(s), 256, 0, ctx->stream>>>(root_args(s, b), blk, T0);
    }
    if (d_trsm_left_lower(ctx, false, L, nn, nn, T0, nn, nn)) return -1;       // L^-1 F
    root_transpose(s, T0, T1);
    if (d_trsm_left_lower(ctx, false, L, nn, nn, T1, nn, nn)) return -1;       // L^-1 F L^-T   (symmetric)
    if (d_trsm_left_lower(ctx, true, L, nn, nn, T1, nn, nn)) return -1;        // L^-T (.)
    root_transpose(s, T1, T0);
    if (d_trsm_left_lower(ctx, true, L, nn, nn, T0, nn, nn)) return -1;        // D^-1 F D^-1
    {
        LaunchScope ls(ctx, "front_elem");
        root_elem_kernel<5><<<elem_grid(s), 256, 0, ctx->stream>>>(root_args(s, b), blk, T0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// inverse Hessian, root: blk = sym(D M D + children), D = L L^T
int root_hess_inv(smcp_sym *s, const double *Lt, double *X, int64_t b) {
    smcp_ctx *ctx = s->ctx;
    const int nn = s->root_nn;
    const size_t sq = (size_t)nn * nn;
    double *blk = ROOT_BLK(X, b), *T0 = s->root_ws, *T1 = s->root_ws + sq, *T2 = s->root_ws + 2 * sq;
    const double *L = Lt + s->h_root_boff;
    {
        LaunchScope ls(ctx, "front_elem");
        root_elem_kernel<4><<<elem_grid(s), 256, 0, ctx->stream>>>(root_args(s, b), blk, T0);          // M (full)
    }
    if (launch_gemm(ctx, false, false, L, nn, L, nn, T1, nn, nn, nn, nn, 1.0, 0, 0, 0, "front_gemm_dmma")) return -1;    // D = L L^T
    if (launch_gemm(ctx, false, true, T1, nn, T0, nn, T2, nn, nn, nn, nn, 1.0, 0, 0, 0, "front_gemm_dmma")) return -1;  // D M
    if (launch_gemm(ctx, false, true, T2, nn, T1, nn, T0, nn, nn, nn, nn, 1.0, 0, 0, 0, "front_gemm_dmma")) return -1;  // D M D
    {
        LaunchScope ls(ctx, "front_elem");
        root_elem_kernel<3><<<elem_grid(s), 256, 0, ctx->stream>>>(root_args(s, b), blk, T0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// projected inverse, root: Y = L^{-T} L^{-1}
int root_projinv(smcp_sym *s, double *X, int64_t b) {
    smcp_ctx *ctx = s->ctx;
    const int nn = s->root_nn;
    const size_t sq = (size_t)nn * nn;
    double *blk = ROOT_BLK(X, b), *T0 = s->root_ws, *T1 = s->root_ws + sq;
    {
        LaunchScope ls(ctx, "front_elem");
        set_identity_n<<<elem_grid(s), 256, 0, ctx->stream>>>(T0, nn);
    }
    if (d_trsm_left_lower(ctx, false, blk, nn, nn, T0, nn, nn)) return -1;                                    // L^-1
    if (launch_gemm(ctx, true, true, T0, nn, T0, nn, T1, nn, nn, nn, nn, 1.0, 0, 1, 0, "front_gemm_dmma")) return -1;    // L^-T L^-1 (lower)
    {
        LaunchScope ls(ctx, "front_elem");
        root_elem_kernel<6><<<elem_grid(s), 256, 0, ctx->stream>>>(root_args(s, b), blk, T1);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// completion, root (no separator): X = M^T M (M lower) by a Cholesky of the reversed matrix, L = M^{-1}
int root_completion(smcp_sym *s, double *X, const double *Xin, int64_t b) {
    smcp_ctx *ctx = s->ctx;
    const int nn = s->root_nn;
    const size_t sq = (size_t)nn * nn;
    double *blk = ROOT_BLK(X, b), *T0 = s->root_ws, *T1 = s->root_ws + sq, *T2 = s->root_ws + 2 * sq;
    const double *bin = ROOT_BLK(Xin, b);
    {
        LaunchScope ls(ctx, "front_elem", 3);
        root_reverse_kernel<<<elem_grid(s), 256, 0, ctx->stream>>>(bin, T0, nn, 0);
    }
    if (d_potrf(ctx, T0, nn, nn, nn, s->root_info, 0, 1)) return -1;
    root_flag_kernel<<<1, 1, 0, ctx->stream>>>(s->root_info, s->fail + b);
    {
        LaunchScope ls(ctx, "front_elem", 2);
        root_reverse_kernel<<<elem_grid(s), 256, 0, ctx->stream>>>(T0, T1, nn, 1);     // M
        set_identity_n<<<elem_grid(s), 256, 0, ctx->stream>>>(T2, nn);
    }
    if (d_trsm_left_lower(ctx, false, T1, nn, nn, T2, nn, nn)) return -1;               // M^-1
    {
        LaunchScope ls(ctx, "front_elem");
        root_elem_kernel<6><<<elem_grid(s), 256, 0, ctx->stream>>>(root_args(s, b), blk, T2);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

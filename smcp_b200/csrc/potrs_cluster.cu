// lapack.potrs (src/python/solvers.py:526, 1954: y <- H^{-1} y with the Cholesky factor of the m x m Schur
// complement) as ONE launch of a thread-block CLUSTER.
//
// A triangular solve with one right-hand side is m dependent steps, and a single CTA can only pull the
// factor out of L2 at ~100 GB/s (the round-1 kernel: 0.24 ms per solve at m = 1000, 3.5 ms at m = 10^4).
// Here the 64-row blocks of the vector are owned cyclically by the R CTAs of a cluster (R = 8 or 16):
//   step b:  the owner of block b applies the INVERSE of the 64 x 64 diagonal block (computed once per
//            factorisation by trtri64_kernel, like the triangular solves of MAGMA / cuBLAS) to its fully
//            updated y_b and writes x_b into the shared memory of every CTA of the cluster (DSMEM);
//            cluster barrier;
//            every CTA subtracts L(blk, b) x_b from the blocks blk > b it owns (the block the next step
//            needs first), reading each tile of the factor exactly once, 16 loads in flight per thread.
// The backward sweep with L^T mirrors it.  All sums have a fixed order: bitwise reproducible.
// Inverting only the 64 x 64 DIAGONAL blocks keeps the solve as accurate as substitution up to the
// condition of those blocks (the iterative refinement of the Newton systems, solvers.py:907-913, absorbs
// it); SMCP_B200_POTRS_SUBST=1 selects the substitution kernel of dense.cu instead.
#include "internal.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

#define PC_T 256

// Dinv[b] = L(b, b)^{-1} for every 64 x 64 diagonal block b (lower triangular, column-major 64 x 64,
// zero above the diagonal; rows/columns beyond m: identity).  One CTA per block, one column per thread.
__global__ void __launch_bounds__(64) trtri64_kernel(const double *__restrict__ H, long long ld, long long m, double *__restrict__ Dinv) {
    __shared__ double Ls[64 * 65];
    const int b = blockIdx.x, tid = threadIdx.x;
    const long long k0 = 64LL * b;
    for (int idx = tid; idx < 64 * 64; idx += 64) {
        const int r = idx & 63, c = idx >> 6;
        double v = (r == c) ? 1.0 : 0.0;
        if (r >= c && k0 + r < m && k0 + c < m) v = H[(k0 + r) + (k0 + c) * ld];
        Ls[c * 65 + r] = v;
    }
    __syncthreads();
    // column tid of the inverse: solve L x = e_tid (x_r = 0 for r < tid)
    double x[64];
#pragma unroll
    for (int r = 0; r < 64; ++r) x[r] = 0.0;
    const int c = tid;
#pragma unroll
    for (int r = 0; r < 64; ++r) {
        if (r >= c) {
            double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
            for (int p = 0; p < 64; ++p)
                if (p < r && p >= c) s = fma(-Ls[p * 65 + r], x[p], s);
            x[r] = s / Ls[r * 65 + r];
        }
    }
    double *D = Dinv + (long long)b * 4096;
#pragma unroll
    for (int r = 0; r < 64; ++r) D[r + c * 64] = x[r];
}

struct PotrsArgs {
    const double *L;
    long long ld;
    int m;
    const double *Dinv;
    double *y;
    long long ldy;      // right-hand sides are columns of y, solved one after the other
    int nrhs;
    int do_fwd, do_bwd; // L^-1 and / or L^-T
};

// 64 x 64 tile times a 64-vector: out[r] (-)= sum_c T(r, c) x[c]; the 256 threads split each row's sum in 4
// column groups (16 independent loads per thread), partial sums combined in a fixed order through `part`.
// TRANS: T(r, c) = A[c + r*lda] (the tile of L read transposed).
template <bool TRANS>
__device__ __forceinline__ void tile_gemv(const double *__restrict__ A, long long lda, int rows, int cols, const double *x, double *part, int tid) {
    const int r = tid & 63, g = tid >> 6;
    double s0 = 0.0, s1 = 0.0;
    if (r < rows) {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int c = 16 * g + q;
            v[q] = (c < cols) ? (TRANS ? A[c + (long long)r * lda] : A[r + (long long)c * lda]) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
            s0 = fma(v[q], x[16 * g + q], s0);
            s1 = fma(v[q + 1], x[16 * g + q + 1], s1);
        }
    }
    part[g * 64 + r] = s0 + s1;
}

template <bool BWD>
__device__ __forceinline__ void cluster_sweep(const PotrsArgs &a, cg::cluster_group &cluster, double *ys, double *xb, double *part, double *Ds,
                                              int tid) {
    const int R = (int)cluster.num_blocks(), rk = (int)cluster.block_rank();
    const int nb = (a.m + 63) / 64;
    for (int q = 0; q < nb; ++q) {
        const int b = BWD ? nb - 1 - q : q;
        const int kb = min(64, a.m - 64 * b);
        double *xcur = xb + (q & 1) * 64;
        if (b % R == rk) {
            // x_b = Dinv_b y_b (backward: Dinv_b^T y_b)
            const double *D = a.Dinv + (long long)b * 4096;
            tile_gemv<BWD>(D, 64, 64, 64, ys + (b / R) * 64, part, tid);
            __syncthreads();
            if (tid < 64) {
                const double xv = (tid < kb) ? (part[tid] + part[64 + tid]) + (part[128 + tid] + part[192 + tid]) : 0.0;
                ys[(b / R) * 64 + tid] = xv;
                for (int t = 0; t < R; ++t) cluster.map_shared_rank(xcur, t)[tid] = xv;
            }
        }
        cluster.sync();
        // blocks this CTA owns on the far side of b, nearest first (the next step needs that one), four
        // tiles of the factor per round: 64 loads in flight per thread, one pair of barriers per round
        int o = 1;
        for (;;) {
            int tb[4], nt = 0;
            for (; nt < 4; ++o) {
                const int blk = BWD ? b - o : b + o;
                if (blk < 0 || blk >= nb) break;
                if (blk % R == rk) tb[nt++] = blk;
            }
            if (!nt) break;
            const int r = tid & 63, g = tid >> 6;
            double v[4][16];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (t < nt) {
                    const int blk = tb[t];
                    const int rows = min(64, a.m - 64 * blk);
                    // forward : y_blk -= L(blk, b) x_b       tile entry (r, c) at L[64 blk + r, 64 b + c]
                    // backward: y_blk -= L(b, blk)^T x_b     tile entry (r, c) at L[64 b + c, 64 blk + r]
                    const double *T = BWD ? a.L + 64LL * b + 64LL * blk * a.ld : a.L + 64LL * blk + 64LL * b * a.ld;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int c = 16 * g + e;
                        v[t][e] = (r < rows && c < kb) ? (BWD ? T[c + (long long)r * a.ld] : T[r + (long long)c * a.ld]) : 0.0;
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (t < nt) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int e = 0; e < 16; e += 2) {
                        s0 = fma(v[t][e], xcur[16 * g + e], s0);
                        s1 = fma(v[t][e + 1], xcur[16 * g + e + 1], s1);
                    }
                    part[t * 256 + g * 64 + r] = s0 + s1;
                }
            }
            __syncthreads();
            {
                const int t = tid >> 6, rr = tid & 63;
                if (t < nt) {
                    const double *pp = part + t * 256;
                    ys[(tb[t] / R) * 64 + rr] -= (pp[rr] + pp[64 + rr]) + (pp[128 + rr] + pp[192 + rr]);
                }
            }
            __syncthreads();
        }
    }
    (void)Ds;
}

__global__ void __launch_bounds__(PC_T, 1) potrs_cluster_kernel(PotrsArgs a) {
    extern __shared__ double pcsm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int R = (int)cluster.num_blocks(), rk = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const int nb = (a.m + 63) / 64;
    const int nloc = (nb + R - 1) / R;
    double *ys = pcsm;                  // nloc x 64: the blocks of y this CTA owns
    double *xb = ys + nloc * 64;        // 2 x 64: x of the current step (written by its owner through DSMEM)
    double *part = xb + 128;            // 4 tiles x 4 x 64 partial sums
    cluster.sync();
    for (int rhs = 0; rhs < a.nrhs; ++rhs) {
        double *y = a.y + (long long)rhs * a.ldy;
        for (int idx = tid; idx < nloc * 64; idx += PC_T) {
            const int b = (idx >> 6) * R + rk;
            const long long i = 64LL * b + (idx & 63);
            ys[idx] = (b < nb && i < a.m) ? y[i] : 0.0;
        }
        __syncthreads();
        if (a.do_fwd) cluster_sweep<false>(a, cluster, ys, xb, part, nullptr, tid);
        cluster.sync();
        if (a.do_bwd) cluster_sweep<true>(a, cluster, ys, xb, part, nullptr, tid);
        __syncthreads();
        for (int idx = tid; idx < nloc * 64; idx += PC_T) {
            const int b = (idx >> 6) * R + rk;
            const long long i = 64LL * b + (idx & 63);
            if (b < nb && i < a.m) y[i] = ys[idx];
        }
        cluster.sync();                 // no CTA may run ahead (or exit) while others still write into its shared memory
    }
}

int d_potrs_prepare(smcp_ctx *ctx, const double *H, int64_t ld, int64_t m, double *Dinv) {
    if (m <= 0) return 0;
    LaunchScope ls(ctx, "potrs_trtri");
    trtri64_kernel<<<(unsigned)((m + 63) / 64), 64, 0, ctx->stream>>>(H, ld, m, Dinv);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

bool potrs_cluster_enabled() {
    static const bool off = getenv("SMCP_B200_POTRS_SUBST") && atoi(getenv("SMCP_B200_POTRS_SUBST")) != 0;
    return !off;
}
// One cluster (16 SMs at most) streams the whole factor: fine while it sits in L2 and the solve is a
// latency chain (m = 1000: 0.09 ms against 0.24), slower than the many-SM streamed updates of dense.cu
// once the factor is hundreds of MB (m = 10^4: 5.1 ms against 3.6)
bool potrs_cluster_for(int64_t m) { return potrs_cluster_enabled() && m <= 4096; }

int d_potrs_cluster(smcp_ctx *ctx, const double *H, int64_t m, const double *Dinv, double *y_dev) {
    return d_trs_cluster(ctx, H, m, m, Dinv, y_dev, m, 1, 1, 1, "potrs");
}

// L^-1 and / or L^-T applied to a few right-hand sides (columns of y); L lower triangular n x n (leading
// dimension ld) whose inverted 64 x 64 diagonal blocks are in Dinv (d_potrs_prepare)
int d_trs_cluster(smcp_ctx *ctx, const double *H, int64_t ld, int64_t m, const double *Dinv, double *y_dev, int64_t ldy, int64_t nrhs,
                  int do_fwd, int do_bwd, const char *name) {
    if (m <= 0 || nrhs <= 0) return 0;
    const int nb = (int)((m + 63) / 64);
    int R = nb >= 64 ? 16 : 8;
    if (R > nb) R = std::max(1, nb);
    // cluster sizes must divide the grid; powers of two up to 16
    int Rp = 1;
    while (Rp * 2 <= R) Rp *= 2;
    R = Rp;
    const int nloc = (nb + R - 1) / R;
    const size_t smem = (size_t)(nloc * 64 + 128 + 1024) * sizeof(double);
    static bool attr = false;
    if (!attr) {
        CUDA_TRY(cudaFuncSetAttribute(potrs_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        CUDA_TRY(cudaFuncSetAttribute(potrs_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    if (smem > 200 * 1024) { smcp_set_error("d_potrs_cluster: m too large"); return -2; }
    PotrsArgs a;
    a.L = H; a.ld = ld; a.m = (int)m; a.Dinv = Dinv; a.y = y_dev;
    a.ldy = ldy; a.nrhs = (int)nrhs; a.do_fwd = do_fwd; a.do_bwd = do_bwd;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)R);
    cfg.blockDim = dim3(PC_T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)R;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    LaunchScope ls(ctx, name, 1, 4.0 * (double)m * (double)m * (double)nrhs * (do_fwd + do_bwd));
    CUDA_TRY(cudaLaunchKernelEx(&cfg, potrs_cluster_kernel, a));
    return 0;
}

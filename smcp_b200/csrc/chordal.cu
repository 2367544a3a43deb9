// Supernodal chordal-matrix kernels for sm_100a: cholesky, completion, projected inverse,
// llt, barrier Hessian (forward and inverse), chordal trsm and the level-1 reductions.
//
// These replace the chompack routines SMCP calls once or many times per interior-point
// iteration (reference call sites: src/python/solvers.py:82-97 imports; cholesky e.g. 884,
// 2354; completion 874, 2344; projected_inverse 891, 2361; llt 904, 1721; hessian 483, 524,
// 531, 1913, 1952, 1959 and the inverse map 405, 1735, 2021; trsm 491-492; dot: 52 sites).
//
// Execution model (B200-first, not a translation of chompack's Python recursion):
//  * every routine works on a BATCH of chordal matrices that share one clique tree (the m
//    constraint matrices of the Schur complement, or the candidates of a line search);
//  * one persistent kernel per routine.  Work items are (task, batch element) pairs where a
//    task is a connected piece of the clique tree (a chain or a small subtree) that one CTA
//    walks sequentially; items are handed out through an atomic queue in a topological
//    order and cross-task dependencies are resolved with release/acquire epoch flags in
//    global memory, so there is no host round trip and no launch per tree level.  Because
//    every CTA of the grid is co-resident and items are claimed in topological order, the
//    lowest unfinished item can always run: the scheme cannot deadlock;
//  * frontal matrices are staged in shared memory when they fit (tiny cliques: band and
//    max-cut patterns) and in a per-CTA global scratch otherwise;
//  * extend-add is done by the parent, child after child, so sums are formed in a fixed
//    order (bitwise reproducible; no floating-point atomics anywhere).
#include "internal.cuh"
#include <cstdio>

#define TID ((int)threadIdx.x)
#define NT ((int)blockDim.x)

enum { OP_CHOL = 0, OP_LLT, OP_PROJINV, OP_COMPL, OP_HPREP, OP_HPREP_INV, OP_HFWD_UP, OP_HFWD_DOWN,
       OP_HINV };

struct TreeArgs {
    SymDev S;
    TaskSched T;
    double *X;           // batch x nblk (in/out)
    const double *Xin;   // completion: input copy
    double *upd;         // batch x nupd
    const double *Lt;    // hessian factor
    const double *Yaa;
    double *Raa;
    const double *L0, *Y0;   // hess prep inputs
    double *Lt_out, *Yaa_out;
    int B;
    unsigned *counter, *done;
    unsigned epoch;
    int *fail;
    double *cta_ws;
    long long ws_stride;
    int use_smem;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------
// CTA-cooperative dense helpers on column-major matrices (generic address space).
// Every helper ends with __syncthreads().
// ---------------------------------------------------------------------------------------

// C(m x n) = (acc ? C : 0) + alpha * A(m x k) * B(k x n), element strides given explicitly.
__device__ void mm(double *C, int ldc, int m, int n, int k, double alpha, const double *A, int ars,
                   int acs, const double *B, int brs, int bcs, bool acc, bool lower) {
    int tot = m * n;
    for (int idx = TID; idx < tot; idx += NT) {
        int i = idx % m, j = idx / m;
        if (lower && i < j) continue;
        const double *a = A + (long long)i * ars;
        const double *b = B + (long long)j * bcs;
        double s = 0.0;
        for (int l = 0; l < k; ++l) s = fma(a[(long long)l * acs], b[(long long)l * brs], s);
        double *c = C + i + (long long)j * ldc;
        *c = acc ? (*c + alpha * s) : alpha * s;
    }
    __syncthreads();
}

// Cholesky of the leading n x n block of an mrows x n panel (mrows >= n), lower storage;
// rows n..mrows-1 receive B * L^{-T}.  dpotrf failure rule: pivot <= 0 or NaN.
__device__ void chol_panel(double *A, int lda, int n, int mrows, int *fail) {
    for (int j = 0; j < n; ++j) {
        double d = A[j + (long long)j * lda];
        bool bad = !(d > 0.0);
        if (bad && TID == 0) *fail = 1;
        double s = bad ? 1.0 : sqrt(d);
        for (int i = j + 1 + TID; i < mrows; i += NT) A[i + (long long)j * lda] /= s;
        __syncthreads();
        if (TID == 0) A[j + (long long)j * lda] = s;
        int nc = n - j - 1;
        int nr = mrows - j - 1;
        int tot = nc * nr;
        for (int idx = TID; idx < tot; idx += NT) {
            int r = idx % nr, c = idx / nr;
            int i = j + 1 + r, cc = j + 1 + c;
            if (i >= cc)
                A[i + (long long)cc * lda] = fma(-A[i + (long long)j * lda], A[cc + (long long)j * lda],
                                                 A[i + (long long)cc * lda]);
        }
        __syncthreads();
    }
}

// "Reverse" Cholesky in lower storage: A = M^T M with M lower triangular (in place).
__device__ void rev_chol(double *A, int lda, int n, int *fail) {
    for (int j = n - 1; j >= 0; --j) {
        double d = A[j + (long long)j * lda];
        bool bad = !(d > 0.0);
        if (bad && TID == 0) *fail = 1;
        double s = bad ? 1.0 : sqrt(d);
        for (int i = TID; i < j; i += NT) A[j + (long long)i * lda] /= s;
        __syncthreads();
        if (TID == 0) A[j + (long long)j * lda] = s;
        int tot = j * j;
        for (int idx = TID; idx < tot; idx += NT) {
            int i = idx % j, k = idx / j;
            if (i >= k)
                A[i + (long long)k * lda] = fma(-A[j + (long long)i * lda], A[j + (long long)k * lda],
                                                A[i + (long long)k * lda]);
        }
        __syncthreads();
    }
}

// B(n x nrhs) <- L^{-1} B
__device__ void trsm_ll(const double *L, int ldl, int n, double *B, int ldb, int nrhs) {
    for (int j = 0; j < n; ++j) {
        double d = L[j + (long long)j * ldl];
        for (int c = TID; c < nrhs; c += NT) B[j + (long long)c * ldb] /= d;
        __syncthreads();
        int nr = n - j - 1, tot = nr * nrhs;
        for (int idx = TID; idx < tot; idx += NT) {
            int r = idx % nr, c = idx / nr;
            int i = j + 1 + r;
            B[i + (long long)c * ldb] = fma(-L[i + (long long)j * ldl], B[j + (long long)c * ldb],
                                            B[i + (long long)c * ldb]);
        }
        __syncthreads();
    }
}

// B(n x nrhs) <- L^{-T} B
__device__ void trsm_llt(const double *L, int ldl, int n, double *B, int ldb, int nrhs) {
    for (int j = n - 1; j >= 0; --j) {
        double d = L[j + (long long)j * ldl];
        for (int c = TID; c < nrhs; c += NT) B[j + (long long)c * ldb] /= d;
        __syncthreads();
        int tot = j * nrhs;
        for (int idx = TID; idx < tot; idx += NT) {
            int i = idx % j, c = idx / j;
            B[i + (long long)c * ldb] = fma(-L[j + (long long)i * ldl], B[j + (long long)c * ldb],
                                            B[i + (long long)c * ldb]);
        }
        __syncthreads();
    }
}

// B(m x n) <- B L^{-1}
__device__ void trsm_rl(const double *L, int ldl, int n, double *B, int ldb, int m) {
    for (int c = n - 1; c >= 0; --c) {
        double d = L[c + (long long)c * ldl];
        for (int i = TID; i < m; i += NT) B[i + (long long)c * ldb] /= d;
        __syncthreads();
        int tot = c * m;
        for (int idx = TID; idx < tot; idx += NT) {
            int i = idx % m, r = idx / m;
            B[i + (long long)r * ldb] = fma(-B[i + (long long)c * ldb], L[c + (long long)r * ldl],
                                            B[i + (long long)r * ldb]);
        }
        __syncthreads();
    }
}

// B(m x n) <- B L^{-T}
__device__ void trsm_rlt(const double *L, int ldl, int n, double *B, int ldb, int m) {
    for (int c = 0; c < n; ++c) {
        double d = L[c + (long long)c * ldl];
        for (int i = TID; i < m; i += NT) B[i + (long long)c * ldb] /= d;
        __syncthreads();
        int nc = n - c - 1, tot = nc * m;
        for (int idx = TID; idx < tot; idx += NT) {
            int i = idx % m, r = c + 1 + idx / m;
            B[i + (long long)r * ldb] = fma(-B[i + (long long)c * ldb], L[r + (long long)c * ldl],
                                            B[i + (long long)r * ldb]);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// per-supernode steps
// ---------------------------------------------------------------------------------------
struct Node {
    int k, nn, na, nj;
    long long boff, uoff;
};

__device__ __forceinline__ Node node_of(const SymDev &S, int k) {
    Node q;
    q.k = k;
    q.nn = S.nn[k];
    q.na = S.na[k];
    q.nj = q.nn + q.na;
    q.boff = S.blkptr[k];
    q.uoff = S.updptr[k];
    return q;
}

// X = L L^T (App. A.1): extend-add children, factor the pivot block, Schur update.
__device__ void op_chol(const TreeArgs &a, const Node &q, int b) {
    const SymDev &S = a.S;
    double *blk = a.X + (long long)b * S.nblk + q.boff;
    double *ub = a.upd + (long long)b * S.nupd;
    double *Uk = ub + q.uoff;
    const int nn = q.nn, na = q.na, nj = q.nj;
    for (int idx = TID; idx < na * na; idx += NT) Uk[idx] = 0.0;
    __syncthreads();
    for (int ch = S.chptr[q.k]; ch < S.chptr[q.k + 1]; ++ch) {
        int c = S.chidx[ch];
        int nac = S.na[c];
        const int *rel = S.relidx + S.relptr[c];
        const double *Uc = ub + S.updptr[c];
        for (int idx = TID; idx < nac * nac; idx += NT) {
            int i = idx % nac, j = idx / nac;
            if (i < j) continue;
            int ri = rel[i], rj = rel[j];
            double v = Uc[idx];
            if (rj < nn) blk[ri + (long long)rj * nj] += v;
            else Uk[(ri - nn) + (long long)(rj - nn) * na] += v;
        }
        __syncthreads();
    }
    chol_panel(blk, nj, nn, nj, a.fail + b);
    if (na) mm(Uk, na, na, na, nn, -1.0, blk + nn, 1, nj, blk + nn, nj, 1, true, true);
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        if (i < j) blk[i + (long long)j * nj] = 0.0;
    }
    __syncthreads();
}

// X = P(L L^T) (App. A.6)
__device__ void op_llt(const TreeArgs &a, const Node &q, int b, double *ws) {
    const SymDev &S = a.S;
    double *blk = a.X + (long long)b * S.nblk + q.boff;
    double *ub = a.upd + (long long)b * S.nupd;
    double *Uk = ub + q.uoff;
    const int nn = q.nn, na = q.na, nj = q.nj;
    double *P = ws;   // nj x nj lower
    for (int idx = TID; idx < nj * nj; idx += NT) {
        int i = idx % nj, j = idx / nj;
        if (i < j) continue;
        int kmax = j < nn ? j : nn - 1;
        double s = 0.0;
        for (int c = 0; c <= kmax; ++c) s = fma(blk[i + (long long)c * nj], blk[j + (long long)c * nj], s);
        P[idx] = s;
    }
    __syncthreads();
    for (int ch = S.chptr[q.k]; ch < S.chptr[q.k + 1]; ++ch) {
        int c = S.chidx[ch];
        int nac = S.na[c];
        const int *rel = S.relidx + S.relptr[c];
        const double *Uc = ub + S.updptr[c];
        for (int idx = TID; idx < nac * nac; idx += NT) {
            int i = idx % nac, j = idx / nac;
            if (i < j) continue;
            P[rel[i] + (long long)rel[j] * nj] += Uc[idx];
        }
        __syncthreads();
    }
    for (int idx = TID; idx < nj * nn; idx += NT) {
        int i = idx % nj, j = idx / nj;
        blk[idx] = (i >= j) ? P[i + (long long)j * nj] : 0.0;
    }
    for (int idx = TID; idx < na * na; idx += NT) {
        int i = idx % na, j = idx / na;
        if (i >= j) Uk[idx] = P[(nn + i) + (long long)(nn + j) * nj];
    }
    __syncthreads();
}

__device__ void gather_aa(const SymDev &S, const Node &q, const double *Xb, double *dst) {
    const int *ai = S.aaidx + q.uoff;
    int tot = q.na * q.na;
    for (int idx = TID; idx < tot; idx += NT) dst[idx] = Xb[ai[idx]];
}

// Y = P((L L^T)^{-1}) (App. A.2), root to leaves.
__device__ void op_projinv(const TreeArgs &a, const Node &q, int b, double *ws) {
    const SymDev &S = a.S;
    double *Xb = a.X + (long long)b * S.nblk;
    double *blk = Xb + q.boff;
    const int nn = q.nn, na = q.na, nj = q.nj;
    double *T1 = ws;                 // na x nn : Lt
    double *T2 = T1 + na * nn;       // nn x nn : L^{-1}
    double *T3 = T2 + nn * nn;       // na x na : Y_aa
    for (int idx = TID; idx < na * nn; idx += NT) {
        int i = idx % na, c = idx / na;
        T1[idx] = blk[nn + i + (long long)c * nj];
    }
    for (int idx = TID; idx < nn * nn; idx += NT) T2[idx] = (idx % nn == idx / nn) ? 1.0 : 0.0;
    gather_aa(S, q, Xb, T3);
    __syncthreads();
    if (na) trsm_rl(blk, nj, nn, T1, na, na);
    trsm_ll(blk, nj, nn, T2, nn, nn);
    if (na) mm(blk + nn, nj, na, nn, na, -1.0, T3, 1, na, T1, 1, na, false, false);
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        if (i < j) { blk[i + (long long)j * nj] = 0.0; continue; }
        double s = 0.0, s2 = 0.0;
        for (int r = i; r < nn; ++r) s = fma(T2[r + i * nn], T2[r + j * nn], s);
        // symmetrised -Lt^T Y_an (matches 0.5*(Y+Y^T) of the oracle up to rounding)
        for (int r = 0; r < na; ++r) {
            s2 = fma(T1[r + i * na], blk[nn + r + (long long)j * nj], s2);
            s2 = fma(T1[r + j * na], blk[nn + r + (long long)i * nj], s2);
        }
        ws[na * nn + nn * nn + na * na + idx] = s - 0.5 * s2;
    }
    __syncthreads();
    double *T4 = ws + na * nn + nn * nn + na * na;
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        if (i >= j) blk[i + (long long)j * nj] = T4[idx];
    }
    __syncthreads();
}

// L with P((L L^T)^{-1}) = X (App. A.3); independent per supernode, out of place.
__device__ void op_compl(const TreeArgs &a, const Node &q, int b, double *ws) {
    const SymDev &S = a.S;
    const double *Xi = a.Xin + (long long)b * S.nblk;
    const double *bin = Xi + q.boff;
    double *bout = a.X + (long long)b * S.nblk + q.boff;
    const int nn = q.nn, na = q.na, nj = q.nj;
    double *R = ws;                 // na x na
    double *Z = R + na * na;        // na x nn
    double *Dl = Z + na * nn;       // nn x nn
    double *Li = Dl + nn * nn;      // nn x nn
    gather_aa(S, q, Xi, R);
    for (int idx = TID; idx < na * nn; idx += NT) Z[idx] = bin[nn + idx % na + (long long)(idx / na) * nj];
    for (int idx = TID; idx < nn * nn; idx += NT) {
        Dl[idx] = bin[idx % nn + (long long)(idx / nn) * nj];
        Li[idx] = (idx % nn == idx / nn) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (na) {
        chol_panel(R, na, na, na, a.fail + b);
        trsm_ll(R, na, na, Z, na, nn);
        mm(Dl, nn, nn, nn, na, -1.0, Z, na, 1, Z, 1, na, true, true);
        trsm_llt(R, na, na, Z, na, nn);
    }
    rev_chol(Dl, nn, nn, a.fail + b);
    trsm_ll(Dl, nn, nn, Li, nn, nn);          // Li = M^{-1} = L_nn
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        bout[i + (long long)j * nj] = (i >= j) ? Li[idx] : 0.0;
    }
    for (int idx = TID; idx < na * nn; idx += NT) {
        int i = idx % na, c = idx / na;
        double s = 0.0;
        for (int r = c; r < nn; ++r) s = fma(Z[i + r * na], Li[r + c * nn], s);
        bout[nn + i + (long long)c * nj] = -s;
    }
    __syncthreads();
}

// Hessian factor: Lt block (L_nn copy, L_an L_nn^{-1}) and Y_aa.
__device__ void op_hprep(const TreeArgs &a, const Node &q) {
    const SymDev &S = a.S;
    const double *Lb = a.L0 + q.boff;
    double *Ob = a.Lt_out + q.boff;
    const int nn = q.nn, na = q.na, nj = q.nj;
    for (int idx = TID; idx < nj * nn; idx += NT) {
        int i = idx % nj, j = idx / nj;
        Ob[idx] = (i >= j) ? Lb[idx] : 0.0;
    }
    gather_aa(S, q, a.Y0, a.Yaa_out + q.uoff);
    __syncthreads();
    if (na) trsm_rl(Ob, nj, nn, Ob + nn, nj, na);
}

__device__ void op_hprep_inv(const TreeArgs &a, const Node &q) {
    double *R = a.Raa + q.uoff;
    const double *Y = a.Yaa + q.uoff;
    const int na = q.na;
    for (int idx = TID; idx < na * na; idx += NT) R[idx] = Y[idx];
    __syncthreads();
    if (na) chol_panel(R, na, na, na, a.fail);
}

__device__ void extend_add_full(const SymDev &S, int c, const double *ub, double *Fnn, int ldnn,
                                double *Fan, int ldan, double *Faa, int ldaa, int nn) {
    // adds the full symmetric update matrix of child c into the three blocks of a frontal
    // matrix stored as F_nn (nn x nn, full), F_an (na x nn), F_aa (na x na, full)
    int nac = S.na[c];
    const int *rel = S.relidx + S.relptr[c];
    const double *Uc = ub + S.updptr[c];
    for (int idx = TID; idx < nac * nac; idx += NT) {
        int i = idx % nac, j = idx / nac;
        int ri = rel[i], rj = rel[j];
        double v = Uc[idx];
        if (rj < nn) {
            if (ri < nn) Fnn[ri + (long long)rj * ldnn] += v;
            else Fan[(ri - nn) + (long long)rj * ldan] += v;
        } else if (ri >= nn) {
            Faa[(ri - nn) + (long long)(rj - nn) * ldaa] += v;
        }
    }
    __syncthreads();
}

// forward Hessian, pass 1 + scaling (App. A.4 steps 1-2), leaves to root
__device__ void op_hfwd_up(const TreeArgs &a, const Node &q, int b, double *ws) {
    const SymDev &S = a.S;
    double *blk = a.X + (long long)b * S.nblk + q.boff;
    double *ub = a.upd + (long long)b * S.nupd;
    double *Uk = ub + q.uoff;
    const double *Lb = a.Lt + q.boff;          // L_nn (lower) and Lt
    const double *Ltan = Lb + q.nn;            // Lt(i, r) at Ltan[i + r*nj]
    const double *Yaa = a.Yaa + q.uoff;
    const int nn = q.nn, na = q.na, nj = q.nj;
    double *Fnn = ws;                  // nn x nn full
    double *Fan = Fnn + nn * nn;       // na x nn  (becomes K_an)
    double *Faa = Fan + na * nn;       // na x na full
    double *Fold = Faa + na * na;      // na x nn  copy of F_an before the congruence
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        Fnn[idx] = (i >= j) ? blk[i + (long long)j * nj] : blk[j + (long long)i * nj];
    }
    for (int idx = TID; idx < na * nn; idx += NT) Fan[idx] = blk[nn + idx % na + (long long)(idx / na) * nj];
    for (int idx = TID; idx < na * na; idx += NT) Faa[idx] = 0.0;
    __syncthreads();
    for (int ch = S.chptr[q.k]; ch < S.chptr[q.k + 1]; ++ch)
        extend_add_full(S, S.chidx[ch], ub, Fnn, nn, Fan, na, Faa, na, nn);
    if (na) {
        for (int idx = TID; idx < na * nn; idx += NT) Fold[idx] = Fan[idx];
        __syncthreads();
        // K_an = F_an - Lt F_nn
        mm(Fan, na, na, nn, nn, -1.0, Ltan, 1, nj, Fnn, 1, nn, true, false);
        // U' = F_aa - Lt F_an(old)^T - K_an Lt^T   (full symmetric)
        for (int idx = TID; idx < na * na; idx += NT) {
            int i = idx % na, j = idx / na;
            double s = 0.0;
            for (int r = 0; r < nn; ++r) {
                s = fma(Ltan[i + (long long)r * nj], Fold[j + r * na], s);
                s = fma(Fan[i + r * na], Ltan[j + (long long)r * nj], s);
            }
            Uk[idx] = Faa[idx] - s;
        }
        __syncthreads();
    }
    // M_nn = D^{-1} K_nn D^{-1}, D = L L^T
    trsm_ll(Lb, nj, nn, Fnn, nn, nn);
    trsm_rlt(Lb, nj, nn, Fnn, nn, nn);
    trsm_llt(Lb, nj, nn, Fnn, nn, nn);
    trsm_rl(Lb, nj, nn, Fnn, nn, nn);
    if (na) {
        // M_an = Y_aa K_an D^{-1}
        trsm_rlt(Lb, nj, nn, Fan, na, na);
        trsm_rl(Lb, nj, nn, Fan, na, na);
        mm(blk + nn, nj, na, nn, na, 1.0, Yaa, 1, na, Fan, 1, na, false, false);
    }
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        blk[i + (long long)j * nj] = (i >= j) ? 0.5 * (Fnn[idx] + Fnn[j + i * nn]) : 0.0;
    }
    __syncthreads();
}

// forward Hessian, pass 3 (App. A.4 step 3), root to leaves
__device__ void op_hfwd_down(const TreeArgs &a, const Node &q, int b, double *ws) {
    const int nn = q.nn, na = q.na, nj = q.nj;
    if (!na) return;
    const SymDev &S = a.S;
    double *Xb = a.X + (long long)b * S.nblk;
    double *blk = Xb + q.boff;
    const double *Ltan = a.Lt + q.boff + nn;
    double *Zaa = ws;                 // na x na
    double *Mold = Zaa + na * na;     // na x nn
    double *Tn = Mold + na * nn;      // nn x nn
    gather_aa(S, q, Xb, Zaa);
    for (int idx = TID; idx < na * nn; idx += NT) Mold[idx] = blk[nn + idx % na + (long long)(idx / na) * nj];
    __syncthreads();
    // Z_an = M_an - Z_aa Lt
    for (int idx = TID; idx < na * nn; idx += NT) {
        int i = idx % na, c = idx / na;
        double s = 0.0;
        for (int r = 0; r < na; ++r) s = fma(Zaa[i + r * na], Ltan[r + (long long)c * nj], s);
        blk[nn + i + (long long)c * nj] = Mold[idx] - s;
    }
    __syncthreads();
    // Z_nn = M_nn - Lt^T M_an - Z_an^T Lt   (symmetrised)
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        if (i < j) continue;
        double s = 0.0;
        for (int r = 0; r < na; ++r) {
            double li = Ltan[r + (long long)i * nj], lj = Ltan[r + (long long)j * nj];
            s = fma(li, Mold[r + j * na], s);
            s = fma(blk[nn + r + (long long)i * nj], lj, s);
            s = fma(lj, Mold[r + i * na], s);
            s = fma(blk[nn + r + (long long)j * nj], li, s);
        }
        Tn[idx] = blk[i + (long long)j * nj] - 0.5 * s;
    }
    __syncthreads();
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        if (i >= j) blk[i + (long long)j * nj] = Tn[idx];
    }
    __syncthreads();
}

// inverse Hessian (App. A.5), one sweep leaves to root
__device__ void op_hinv(const TreeArgs &a, const Node &q, int b, double *ws) {
    const SymDev &S = a.S;
    double *Xb = a.X + (long long)b * S.nblk;
    double *blk = Xb + q.boff;
    double *ub = a.upd + (long long)b * S.nupd;
    double *Uk = ub + q.uoff;
    const double *Lb = a.Lt + q.boff;
    const double *Ltan = Lb + q.nn;
    const int nn = q.nn, na = q.na, nj = q.nj;
    double *T1 = ws;                 // na x nn : M_an, later F_an
    double *T2 = T1 + na * nn;       // nn x nn : M_nn, later K_nn / F_nn
    double *T3 = T2 + nn * nn;       // na x na : Z_aa, later F_aa
    double *T4 = T3 + na * na;       // nn x nn : D
    double *T5 = T4 + nn * nn;       // nn x nn : temp
    double *T6 = T5 + nn * nn;       // na x nn : K_an
    gather_aa(S, q, Xb, T3);
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        int kmax = i < j ? i : j;
        double s = 0.0;
        for (int r = 0; r <= kmax; ++r) s = fma(Lb[i + (long long)r * nj], Lb[j + (long long)r * nj], s);
        T4[idx] = s;
    }
    __syncthreads();
    // M_an = Z_an + Z_aa Lt
    for (int idx = TID; idx < na * nn; idx += NT) {
        int i = idx % na, c = idx / na;
        double s = 0.0;
        for (int r = 0; r < na; ++r) s = fma(T3[i + r * na], Ltan[r + (long long)c * nj], s);
        T1[idx] = blk[nn + i + (long long)c * nj] + s;
    }
    __syncthreads();
    // M_nn = Z_nn + Lt^T Z_an + M_an^T Lt
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        double s = (i >= j) ? blk[i + (long long)j * nj] : blk[j + (long long)i * nj];
        for (int r = 0; r < na; ++r) {
            s = fma(Ltan[r + (long long)i * nj], blk[nn + r + (long long)j * nj], s);
            s = fma(T1[r + i * na], Ltan[r + (long long)j * nj], s);
        }
        T2[idx] = s;
    }
    __syncthreads();
    // K_nn = D M_nn D
    mm(T5, nn, nn, nn, nn, 1.0, T4, 1, nn, T2, 1, nn, false, false);
    mm(T2, nn, nn, nn, nn, 1.0, T5, 1, nn, T4, 1, nn, false, false);
    if (na) {
        // K_an = Y_aa^{-1} M_an D
        mm(T6, na, na, nn, nn, 1.0, T1, 1, na, T4, 1, nn, false, false);
        const double *R = a.Raa + q.uoff;
        trsm_ll(R, na, na, T6, na, nn);
        trsm_llt(R, na, na, T6, na, nn);
        // F_an = K_an + Lt K_nn
        for (int idx = TID; idx < na * nn; idx += NT) {
            int i = idx % na, c = idx / na;
            double s = 0.0;
            for (int r = 0; r < nn; ++r) s = fma(Ltan[i + (long long)r * nj], T2[r + c * nn], s);
            T1[idx] = T6[idx] + s;
        }
        __syncthreads();
        // F_aa = Lt K_an^T + F_an Lt^T
        for (int idx = TID; idx < na * na; idx += NT) {
            int i = idx % na, j = idx / na;
            double s = 0.0;
            for (int r = 0; r < nn; ++r) {
                s = fma(Ltan[i + (long long)r * nj], T6[j + r * na], s);
                s = fma(T1[i + r * na], Ltan[j + (long long)r * nj], s);
            }
            T3[idx] = s;
        }
        __syncthreads();
    }
    for (int ch = S.chptr[q.k]; ch < S.chptr[q.k + 1]; ++ch)
        extend_add_full(S, S.chidx[ch], ub, T2, nn, T1, na, T3, na, nn);
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        blk[i + (long long)j * nj] = (i >= j) ? 0.5 * (T2[idx] + T2[j + i * nn]) : 0.0;
    }
    for (int idx = TID; idx < na * nn; idx += NT) blk[nn + idx % na + (long long)(idx / na) * nj] = T1[idx];
    for (int idx = TID; idx < na * na; idx += NT) Uk[idx] = T3[idx];
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// persistent dependency-driven kernel
// ---------------------------------------------------------------------------------------
template <int OP>
__global__ void tree_kernel(TreeArgs a) {
    extern __shared__ double smem_ws[];
    __shared__ int s_item;
    double *ws = a.use_smem ? smem_ws : a.cta_ws + (long long)blockIdx.x * a.ws_stride;
    const int total = a.T.ntask * a.B;
    for (;;) {
        if (TID == 0) s_item = (int)atomicAdd(a.counter, 1u);
        __syncthreads();
        int item = s_item;
        __syncthreads();
        if (item >= total) break;
        int t = item / a.B, b = item % a.B;
        if (TID == 0) {
            for (int d = a.T.dep_ptr[t]; d < a.T.dep_ptr[t + 1]; ++d) {
                const unsigned *flag = a.done + (long long)a.T.dep_idx[d] * a.B + b;
                while (ld_acquire(flag) != a.epoch) __nanosleep(32);
            }
        }
        __syncthreads();
        __threadfence();
        for (int p = a.T.task_ptr[t]; p < a.T.task_ptr[t + 1]; ++p) {
            Node q = node_of(a.S, a.T.task_sn[p]);
            if (OP == OP_CHOL) op_chol(a, q, b);
            else if (OP == OP_LLT) op_llt(a, q, b, ws);
            else if (OP == OP_PROJINV) op_projinv(a, q, b, ws);
            else if (OP == OP_COMPL) op_compl(a, q, b, ws);
            else if (OP == OP_HPREP) op_hprep(a, q);
            else if (OP == OP_HPREP_INV) op_hprep_inv(a, q);
            else if (OP == OP_HFWD_UP) op_hfwd_up(a, q, b, ws);
            else if (OP == OP_HFWD_DOWN) op_hfwd_down(a, q, b, ws);
            else if (OP == OP_HINV) op_hinv(a, q, b, ws);
            __syncthreads();
        }
        __threadfence();
        __syncthreads();
        if (TID == 0) st_release(a.done + (long long)t * a.B + b, a.epoch);
    }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
static const size_t SMEM_WS_LIMIT = 96 * 1024;

static size_t ws_doubles(const smcp_sym *s) {
    size_t nj = (size_t)s->max_nj;
    return 4 * nj * nj + 8;
}

int sym_ensure(smcp_sym *s, int64_t batch, bool need_tmp) {
    smcp_ctx *ctx = s->ctx;
    size_t need_done = (size_t)(s->d.nsn) * (size_t)batch * sizeof(unsigned);
    if (need_done > s->done_cap) {
        if (s->done) cudaFree(s->done);
        CUDA_TRY(cudaMalloc(&s->done, need_done));
        CUDA_TRY(cudaMemsetAsync(s->done, 0, need_done, ctx->stream));
        s->done_cap = need_done;
    }
    if (grow((void **)&s->fail, &s->fail_cap, (size_t)batch * sizeof(int))) return -1;
    if (grow((void **)&s->upd, &s->upd_cap, (size_t)batch * (size_t)(s->d.nupd + 1) * sizeof(double))) return -1;
    if (need_tmp && grow((void **)&s->tmp, &s->tmp_cap, (size_t)batch * (size_t)s->d.nblk * sizeof(double))) return -1;
    return 0;
}

template <int OP>
static int launch_tree(smcp_sym *s, TreeArgs &a, const TaskSched &T, int64_t batch, int threads, const char *name) {
    smcp_ctx *ctx = s->ctx;
    a.S = s->d;
    a.T = T;
    a.B = (int)batch;
    a.counter = s->counter;
    a.done = s->done;
    a.epoch = ++s->epoch;
    a.fail = s->fail;
    size_t wsd = ws_doubles(s);
    size_t smem = wsd * sizeof(double);
    a.use_smem = smem <= SMEM_WS_LIMIT;
    auto kern = tree_kernel<OP>;
    if (a.use_smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
        smem = 0;
    }
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) { smcp_set_error("tree kernel does not fit on an SM"); return -1; }
    long long items = (long long)T.ntask * batch;
    long long grid = (long long)per_sm * ctx->num_sms;
    if (grid > items) grid = items;
    if (grid < 1) grid = 1;
    if (!a.use_smem) {
        if (grow((void **)&s->cta_ws, &s->cta_ws_cap, (size_t)grid * wsd * sizeof(double))) return -1;
        a.cta_ws = s->cta_ws;
        a.ws_stride = (long long)wsd;
    }
    CUDA_TRY(cudaMemsetAsync(s->counter, 0, sizeof(unsigned), ctx->stream));
    {
        LaunchScope ls(ctx, name, 1, (double)batch);
        kern<<<(unsigned)grid, threads, smem, ctx->stream>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int pick_threads(const smcp_sym *s) {
    int nj = s->max_nj;
    if (nj <= 8) return 32;
    if (nj <= 16) return 64;
    if (nj <= 48) return 128;
    return 256;
}

static int fetch_fail(smcp_sym *s, int64_t batch, int32_t *info_host) {
    smcp_ctx *ctx = s->ctx;
    CUDA_TRY(cudaMemcpyAsync(info_host, s->fail, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int k_cholesky(smcp_sym *s, double *x, int64_t batch, int32_t *info_host) {
    if (sym_ensure(s, batch, false)) return -1;
    CUDA_TRY(cudaMemsetAsync(s->fail, 0, (size_t)batch * sizeof(int), s->ctx->stream));
    TreeArgs a = {};
    a.X = x;
    a.upd = s->upd;
    if (launch_tree<OP_CHOL>(s, a, s->up, batch, pick_threads(s), batch > 1 ? "cholesky_batch" : "cholesky")) return -1;
    if (info_host) return fetch_fail(s, batch, info_host);
    return 0;
}

int k_llt(smcp_sym *s, double *x, int64_t batch) {
    if (sym_ensure(s, batch, false)) return -1;
    TreeArgs a = {};
    a.X = x;
    a.upd = s->upd;
    return launch_tree<OP_LLT>(s, a, s->up, batch, pick_threads(s), "llt");
}

int k_projinv(smcp_sym *s, double *x, int64_t batch) {
    if (sym_ensure(s, batch, false)) return -1;
    TreeArgs a = {};
    a.X = x;
    return launch_tree<OP_PROJINV>(s, a, s->down, batch, pick_threads(s), "projected_inverse");
}

static TaskSched flat_sched(const smcp_sym *s) { return s->flat; }

int k_completion(smcp_sym *s, double *x, int64_t batch, int32_t *info_host) {
    if (sym_ensure(s, batch, true)) return -1;
    smcp_ctx *ctx = s->ctx;
    CUDA_TRY(cudaMemsetAsync(s->fail, 0, (size_t)batch * sizeof(int), ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(s->tmp, x, (size_t)batch * s->d.nblk * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    TreeArgs a = {};
    a.X = x;
    a.Xin = s->tmp;
    if (launch_tree<OP_COMPL>(s, a, flat_sched(s), batch, pick_threads(s), batch > 1 ? "completion_batch" : "completion")) return -1;
    if (info_host) return fetch_fail(s, batch, info_host);
    return 0;
}

int k_hess_prep(smcp_hess *h, const double *L, const double *Y) {
    smcp_sym *s = h->sym;
    if (sym_ensure(s, 1, false)) return -1;
    TreeArgs a = {};
    a.L0 = L;
    a.Y0 = Y;
    a.Lt_out = h->Lt;
    a.Yaa_out = h->Yaa;
    return launch_tree<OP_HPREP>(s, a, flat_sched(s), 1, pick_threads(s), "hessian_prep");
}

int k_hess_prep_inv(smcp_hess *h) {
    smcp_sym *s = h->sym;
    if (sym_ensure(s, 1, false)) return -1;
    CUDA_TRY(cudaMemsetAsync(s->fail, 0, sizeof(int), s->ctx->stream));
    TreeArgs a = {};
    a.Yaa = h->Yaa;
    a.Raa = h->Raa;
    return launch_tree<OP_HPREP_INV>(s, a, flat_sched(s), 1, pick_threads(s), "hessian_prep_inv");
}

int k_hess_apply(smcp_hess *h, double *U, int64_t batch, int inv) {
    smcp_sym *s = h->sym;
    if (sym_ensure(s, batch, false)) return -1;
    TreeArgs a = {};
    a.X = U;
    a.upd = s->upd;
    a.Lt = h->Lt;
    a.Yaa = h->Yaa;
    a.Raa = h->Raa;
    int threads = pick_threads(s);
    if (!inv) {
        const bool big = batch >= 32;
        if (launch_tree<OP_HFWD_UP>(s, a, s->up, batch, threads, big ? "hessian_up_batch" : "hessian_up")) return -1;
        return launch_tree<OP_HFWD_DOWN>(s, a, s->down, batch, threads, big ? "hessian_down_batch" : "hessian_down");
    }
    if (!h->have_Raa) {
        if (k_hess_prep_inv(h)) return -1;
        h->have_Raa = true;
    }
    return launch_tree<OP_HINV>(s, a, s->up, batch, threads, batch >= 32 ? "hessian_inv_batch" : "hessian_inv");
}

// ---------------------------------------------------------------------------------------
// chordal trsm: B <- L^{-1} B / L^{-T} B, B dense n x nrhs (rows in internal order)
// One CTA per block of right-hand sides walks the whole tree (columns are independent).
// ---------------------------------------------------------------------------------------
__global__ void trsm_kernel(SymDev S, const double *L, double *B, long long ldb, int nrhs, int trans, int cols_per_cta) {
    int c0 = blockIdx.x * cols_per_cta;
    int nc = min(cols_per_cta, nrhs - c0);
    if (nc <= 0) return;
    double *Bc = B + (long long)c0 * ldb;
    if (!trans) {
        for (int k = 0; k < S.nsn; ++k) {
            int nn = S.nn[k], na = S.na[k], nj = nn + na;
            const double *blk = L + S.blkptr[k];
            const int *rows = S.rowidx + S.rowptr[k];
            int r0 = rows[0];
            // solve with L_nn on rows r0..r0+nn-1 (contiguous)
            for (int j = 0; j < nn; ++j) {
                double d = blk[j + (long long)j * nj];
                for (int c = TID; c < nc; c += NT) Bc[r0 + j + (long long)c * ldb] /= d;
                __syncthreads();
                int nr = nn - j - 1, tot = nr * nc;
                for (int idx = TID; idx < tot; idx += NT) {
                    int r = idx % nr, c = idx / nr;
                    Bc[r0 + j + 1 + r + (long long)c * ldb] -= blk[j + 1 + r + (long long)j * nj] * Bc[r0 + j + (long long)c * ldb];
                }
                __syncthreads();
            }
            // B_alpha -= L_an B_nu
            for (int idx = TID; idx < na * nc; idx += NT) {
                int i = idx % na, c = idx / na;
                double s = 0.0;
                for (int r = 0; r < nn; ++r) s = fma(blk[nn + i + (long long)r * nj], Bc[r0 + r + (long long)c * ldb], s);
                Bc[rows[nn + i] + (long long)c * ldb] -= s;
            }
            __syncthreads();
        }
    } else {
        for (int k = S.nsn - 1; k >= 0; --k) {
            int nn = S.nn[k], na = S.na[k], nj = nn + na;
            const double *blk = L + S.blkptr[k];
            const int *rows = S.rowidx + S.rowptr[k];
            int r0 = rows[0];
            for (int idx = TID; idx < nn * nc; idx += NT) {
                int r = idx % nn, c = idx / nn;
                double s = 0.0;
                for (int i = 0; i < na; ++i) s = fma(blk[nn + i + (long long)r * nj], Bc[rows[nn + i] + (long long)c * ldb], s);
                Bc[r0 + r + (long long)c * ldb] -= s;
            }
            __syncthreads();
            for (int j = nn - 1; j >= 0; --j) {
                double d = blk[j + (long long)j * nj];
                for (int c = TID; c < nc; c += NT) Bc[r0 + j + (long long)c * ldb] /= d;
                __syncthreads();
                int tot = j * nc;
                for (int idx = TID; idx < tot; idx += NT) {
                    int i = idx % j, c = idx / j;
                    Bc[r0 + i + (long long)c * ldb] -= blk[j + (long long)i * nj] * Bc[r0 + j + (long long)c * ldb];
                }
                __syncthreads();
            }
        }
    }
}

int k_trsm(smcp_sym *s, const double *L, double *B, int64_t ldb, int64_t nrhs, int trans) {
    smcp_ctx *ctx = s->ctx;
    int cols = 8;
    int grid = (int)((nrhs + cols - 1) / cols);
    if (grid < 1) return 0;
    {
        LaunchScope ls(ctx, "chordal_trsm");
        trsm_kernel<<<grid, 128, 0, ctx->stream>>>(s->d, L, B, ldb, (int)nrhs, trans, cols);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------
// level-1 kernels
// ---------------------------------------------------------------------------------------
__global__ void axpy_kernel(double a, const double *__restrict__ x, double *__restrict__ y, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = __dadd_rn(y[i], __dmul_rn(a, x[i]));   // no FMA: matches y += a*x
}
__global__ void scal_kernel(double a, double *x, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) x[i] *= a;
}
// out[k*n + i] = x[i] + gam[k]*dx[i]
__global__ void axpy_batch_kernel(const double *__restrict__ x, const double *__restrict__ dx,
                                  const double *__restrict__ gam, double *__restrict__ out, long long n, int count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double xv = x[i], dv = dx[i];
        for (int k = 0; k < count; ++k) out[(long long)k * n + i] = __dadd_rn(xv, __dmul_rn(gam[k], dv));
    }
}

static int grid_for(smcp_ctx *ctx, long long n, int threads) {
    long long g = (n + threads - 1) / threads;
    long long cap = (long long)ctx->num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int k_axpy(smcp_sym *s, double a, const double *x, double *y, int64_t len) {
    smcp_ctx *ctx = s->ctx;
    {
        LaunchScope ls(ctx, "level1");
        axpy_kernel<<<grid_for(ctx, len, 256), 256, 0, ctx->stream>>>(a, x, y, len);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int k_scal(smcp_sym *s, double a, double *x, int64_t len) {
    smcp_ctx *ctx = s->ctx;
    {
        LaunchScope ls(ctx, "level1");
        scal_kernel<<<grid_for(ctx, len, 256), 256, 0, ctx->stream>>>(a, x, len);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int k_axpy_batch(smcp_sym *s, const double *x, const double *dx, const double *gam_dev, double *out, int64_t count) {
    smcp_ctx *ctx = s->ctx;
    {
        LaunchScope ls(ctx, "level1");
        axpy_batch_kernel<<<grid_for(ctx, s->d.nblk, 256), 256, 0, ctx->stream>>>(x, dx, gam_dev, out, s->d.nblk, (int)count);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// deterministic two-stage weighted dot: fixed grid, fixed per-thread ranges, tree reduce
#define RED_BLOCKS 256
#define RED_THREADS 256
__device__ double block_reduce_sum(double v) {
    __shared__ double sh[RED_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = (threadIdx.x < blockDim.x / 32) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
}
__global__ void dot_stage1(const double *__restrict__ x, const double *__restrict__ y,
                           const double *__restrict__ w, long long n, double *part) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    double s = 0.0;
    for (; i < n; i += stride) s = fma(x[i] * w[i], y[i], s);
    s = block_reduce_sum(s);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void sum_stage2(const double *part, int n, double *out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
    s = block_reduce_sum(s);
    if (threadIdx.x == 0) *out = s;
}
__global__ void logdiag_stage1(const double *__restrict__ x, const int *__restrict__ diag, int n, long long stride_b, double *part) {
    // blockIdx.y = batch element
    const double *xb = x + (long long)blockIdx.y * stride_b;
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += log(xb[diag[i]]);
    s = block_reduce_sum(s);
    if (threadIdx.x == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = s;
}
__global__ void sum_stage2_batch(const double *part, int n, double *out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[blockIdx.x * n + i];
    s = block_reduce_sum(s);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

int k_dot(smcp_sym *s, const double *x, const double *y, double *out_host) {
    smcp_ctx *ctx = s->ctx;
    if (grow((void **)&s->red, &s->red_cap, (RED_BLOCKS + 8) * sizeof(double))) return -1;
    {
        LaunchScope ls(ctx, "reduce", 2);
        dot_stage1<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(x, y, s->d.wdot, s->d.nblk, s->red);
        sum_stage2<<<1, RED_THREADS, 0, ctx->stream>>>(s->red, RED_BLOCKS, s->red + RED_BLOCKS);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_host, s->red + RED_BLOCKS, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int k_sumlogdiag(smcp_sym *s, const double *x, int64_t batch, double *out_host) {
    smcp_ctx *ctx = s->ctx;
    const int nb = 32;
    if (grow((void **)&s->red, &s->red_cap, ((size_t)batch * (nb + 1) + RED_BLOCKS + 8) * sizeof(double))) return -1;
    double *part = s->red;
    double *outd = s->red + (size_t)batch * nb;
    {
        LaunchScope ls(ctx, "reduce", 2);
        logdiag_stage1<<<dim3(nb, (unsigned)batch), RED_THREADS, 0, ctx->stream>>>(x, s->d.diagblk, s->d.n, s->d.nblk, part);
        sum_stage2_batch<<<(unsigned)batch, RED_THREADS, 0, ctx->stream>>>(part, nb, outd);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_host, outd, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

__global__ void scatter_vec_kernel(double *dst, const double *vec, const int *map, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[map[i]] = vec[i];
}
__global__ void gather_vec_kernel(const double *src, double *vec, const int *map, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) vec[i] = src[map[i]];
}
int k_scatter_vec(smcp_sym *s, double *dst, const double *dev_vec) {
    smcp_ctx *ctx = s->ctx;
    CUDA_TRY(cudaMemsetAsync(dst, 0, (size_t)s->d.nblk * sizeof(double), ctx->stream));
    {
        LaunchScope ls(ctx, "level1");
        scatter_vec_kernel<<<grid_for(ctx, s->d.nvp, 256), 256, 0, ctx->stream>>>(dst, dev_vec, s->d.vec2blk, s->d.nvp);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int k_gather_vec(smcp_sym *s, const double *src, double *dev_vec) {
    smcp_ctx *ctx = s->ctx;
    {
        LaunchScope ls(ctx, "level1");
        gather_vec_kernel<<<grid_for(ctx, s->d.nvp, 256), 256, 0, ctx->stream>>>(src, dev_vec, s->d.vec2blk, s->d.nvp);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Per-supernode dense steps of the chordal recursions, shared by the
// CTA-per-supernode kernels (chordal.cu, W = false: a whole CTA cooperates on one frontal
// matrix) and the warp-per-supernode kernels for tiny cliques (chordal_small.cu, W = true:
// one warp per frontal matrix, __syncwarp instead of __syncthreads).
#pragma once
#include "internal.cuh"

#define TID (W ? (int)(threadIdx.x & 31) : (int)threadIdx.x)
#define NT (W ? 32 : (int)blockDim.x)
#define SYNC() do { if (W) __syncwarp(); else __syncthreads(); } while (0)

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
    const int *skipflag; // per supernode: 1 = left to the dense top-set path (bigfront.cu); may be null
    int half;            // half factors of the Hessian (chompack's adj=False/True): 0 = full map;
                         // forward sweeps: 1 = G (up) / G^adj (down); inverse sweep: 1 = G^-adj, 2 = G^-1
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
template <bool W>
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
    SYNC();
}

// Cholesky of the leading n x n block of an mrows x n panel (mrows >= n), lower storage;
// rows n..mrows-1 receive B * L^{-T}.  dpotrf failure rule: pivot <= 0 or NaN.
template <bool W>
__device__ void chol_panel(double *A, int lda, int n, int mrows, int *fail) {
    for (int j = 0; j < n; ++j) {
        double d = A[j + (long long)j * lda];
        bool bad = !(d > 0.0);
        if (bad && TID == 0) *fail = 1;
        double s = bad ? 1.0 : sqrt(d);
        for (int i = j + 1 + TID; i < mrows; i += NT) A[i + (long long)j * lda] /= s;
        SYNC();
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
        SYNC();
    }
}

// "Reverse" Cholesky in lower storage: A = M^T M with M lower triangular (in place).
template <bool W>
__device__ void rev_chol(double *A, int lda, int n, int *fail) {
    for (int j = n - 1; j >= 0; --j) {
        double d = A[j + (long long)j * lda];
        bool bad = !(d > 0.0);
        if (bad && TID == 0) *fail = 1;
        double s = bad ? 1.0 : sqrt(d);
        for (int i = TID; i < j; i += NT) A[j + (long long)i * lda] /= s;
        SYNC();
        if (TID == 0) A[j + (long long)j * lda] = s;
        int tot = j * j;
        for (int idx = TID; idx < tot; idx += NT) {
            int i = idx % j, k = idx / j;
            if (i >= k)
                A[i + (long long)k * lda] = fma(-A[j + (long long)i * lda], A[j + (long long)k * lda],
                                                A[i + (long long)k * lda]);
        }
        SYNC();
    }
}

// B(n x nrhs) <- L^{-1} B
template <bool W>
__device__ void trsm_ll(const double *L, int ldl, int n, double *B, int ldb, int nrhs) {
    for (int j = 0; j < n; ++j) {
        double d = L[j + (long long)j * ldl];
        for (int c = TID; c < nrhs; c += NT) B[j + (long long)c * ldb] /= d;
        SYNC();
        int nr = n - j - 1, tot = nr * nrhs;
        for (int idx = TID; idx < tot; idx += NT) {
            int r = idx % nr, c = idx / nr;
            int i = j + 1 + r;
            B[i + (long long)c * ldb] = fma(-L[i + (long long)j * ldl], B[j + (long long)c * ldb],
                                            B[i + (long long)c * ldb]);
        }
        SYNC();
    }
}

// B(n x nrhs) <- L^{-T} B
template <bool W>
__device__ void trsm_llt(const double *L, int ldl, int n, double *B, int ldb, int nrhs) {
    for (int j = n - 1; j >= 0; --j) {
        double d = L[j + (long long)j * ldl];
        for (int c = TID; c < nrhs; c += NT) B[j + (long long)c * ldb] /= d;
        SYNC();
        int tot = j * nrhs;
        for (int idx = TID; idx < tot; idx += NT) {
            int i = idx % j, c = idx / j;
            B[i + (long long)c * ldb] = fma(-L[j + (long long)i * ldl], B[j + (long long)c * ldb],
                                            B[i + (long long)c * ldb]);
        }
        SYNC();
    }
}

// B(m x n) <- B L^{-1}
template <bool W>
__device__ void trsm_rl(const double *L, int ldl, int n, double *B, int ldb, int m) {
    for (int c = n - 1; c >= 0; --c) {
        double d = L[c + (long long)c * ldl];
        for (int i = TID; i < m; i += NT) B[i + (long long)c * ldb] /= d;
        SYNC();
        int tot = c * m;
        for (int idx = TID; idx < tot; idx += NT) {
            int i = idx % m, r = idx / m;
            B[i + (long long)r * ldb] = fma(-B[i + (long long)c * ldb], L[c + (long long)r * ldl],
                                            B[i + (long long)r * ldb]);
        }
        SYNC();
    }
}

// B(m x n) <- B L^{-T}
template <bool W>
__device__ void trsm_rlt(const double *L, int ldl, int n, double *B, int ldb, int m) {
    for (int c = 0; c < n; ++c) {
        double d = L[c + (long long)c * ldl];
        for (int i = TID; i < m; i += NT) B[i + (long long)c * ldb] /= d;
        SYNC();
        int nc = n - c - 1, tot = nc * m;
        for (int idx = TID; idx < tot; idx += NT) {
            int i = idx % m, r = c + 1 + idx / m;
            B[i + (long long)r * ldb] = fma(-B[i + (long long)c * ldb], L[r + (long long)c * ldl],
                                            B[i + (long long)r * ldb]);
        }
        SYNC();
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
template <bool W>
__device__ void op_chol(const TreeArgs &a, const Node &q, int b) {
    const SymDev &S = a.S;
    double *blk = a.X + (long long)b * S.nblk + q.boff;
    double *ub = a.upd + (long long)b * S.nupd;
    double *Uk = ub + q.uoff;
    const int nn = q.nn, na = q.na, nj = q.nj;
    for (int idx = TID; idx < na * na; idx += NT) Uk[idx] = 0.0;
    SYNC();
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
        SYNC();
    }
    chol_panel<W>(blk, nj, nn, nj, a.fail + b);
    if (na) mm<W>(Uk, na, na, na, nn, -1.0, blk + nn, 1, nj, blk + nn, nj, 1, true, true);
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        if (i < j) blk[i + (long long)j * nj] = 0.0;
    }
    SYNC();
}

// X = P(L L^T) (App. A.6)
template <bool W>
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
    SYNC();
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
        SYNC();
    }
    for (int idx = TID; idx < nj * nn; idx += NT) {
        int i = idx % nj, j = idx / nj;
        blk[idx] = (i >= j) ? P[i + (long long)j * nj] : 0.0;
    }
    for (int idx = TID; idx < na * na; idx += NT) {
        int i = idx % na, j = idx / na;
        if (i >= j) Uk[idx] = P[(nn + i) + (long long)(nn + j) * nj];
    }
    SYNC();
}

template <bool W>
__device__ void gather_aa(const SymDev &S, const Node &q, const double *Xb, double *dst) {
    const int *ai = S.aaidx + q.uoff;
    int tot = q.na * q.na;
    for (int idx = TID; idx < tot; idx += NT) dst[idx] = Xb[ai[idx]];
}

// Y = P((L L^T)^{-1}) (App. A.2), root to leaves.
template <bool W>
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
    gather_aa<W>(S, q, Xb, T3);
    SYNC();
    if (na) trsm_rl<W>(blk, nj, nn, T1, na, na);
    trsm_ll<W>(blk, nj, nn, T2, nn, nn);
    if (na) mm<W>(blk + nn, nj, na, nn, na, -1.0, T3, 1, na, T1, 1, na, false, false);
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
    SYNC();
    double *T4 = ws + na * nn + nn * nn + na * na;
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        if (i >= j) blk[i + (long long)j * nj] = T4[idx];
    }
    SYNC();
}

// L with P((L L^T)^{-1}) = X (App. A.3); independent per supernode, out of place.
template <bool W>
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
    gather_aa<W>(S, q, Xi, R);
    for (int idx = TID; idx < na * nn; idx += NT) Z[idx] = bin[nn + idx % na + (long long)(idx / na) * nj];
    for (int idx = TID; idx < nn * nn; idx += NT) {
        Dl[idx] = bin[idx % nn + (long long)(idx / nn) * nj];
        Li[idx] = (idx % nn == idx / nn) ? 1.0 : 0.0;
    }
    SYNC();
    if (na) {
        chol_panel<W>(R, na, na, na, a.fail + b);
        trsm_ll<W>(R, na, na, Z, na, nn);
        mm<W>(Dl, nn, nn, nn, na, -1.0, Z, na, 1, Z, 1, na, true, true);
        trsm_llt<W>(R, na, na, Z, na, nn);
    }
    rev_chol<W>(Dl, nn, nn, a.fail + b);
    trsm_ll<W>(Dl, nn, nn, Li, nn, nn);          // Li = M^{-1} = L_nn
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
    SYNC();
}

// Hessian factor: Lt block (L_nn copy, L_an L_nn^{-1}) and Y_aa.
template <bool W>
__device__ void op_hprep(const TreeArgs &a, const Node &q) {
    const SymDev &S = a.S;
    const double *Lb = a.L0 + q.boff;
    double *Ob = a.Lt_out + q.boff;
    const int nn = q.nn, na = q.na, nj = q.nj;
    for (int idx = TID; idx < nj * nn; idx += NT) {
        int i = idx % nj, j = idx / nj;
        Ob[idx] = (i >= j) ? Lb[idx] : 0.0;
    }
    gather_aa<W>(S, q, a.Y0, a.Yaa_out + q.uoff);
    SYNC();
    if (na) trsm_rl<W>(Ob, nj, nn, Ob + nn, nj, na);
}

template <bool W>
__device__ void op_hprep_inv(const TreeArgs &a, const Node &q) {
    double *R = a.Raa + q.uoff;
    const double *Y = a.Yaa + q.uoff;
    const int na = q.na;
    for (int idx = TID; idx < na * na; idx += NT) R[idx] = Y[idx];
    SYNC();
    if (na) chol_panel<W>(R, na, na, na, a.fail);
    for (int idx = TID; idx < na * na; idx += NT)
        if (idx % na < idx / na) R[idx] = 0.0;          // strictly upper part: the half factors use R as a dense block
    SYNC();
}

template <bool W>
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
    SYNC();
}

// forward Hessian, pass 1 + scaling (App. A.4 steps 1-2), leaves to root
template <bool W>
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
    SYNC();
    for (int ch = S.chptr[q.k]; ch < S.chptr[q.k + 1]; ++ch)
        extend_add_full<W>(S, S.chidx[ch], ub, Fnn, nn, Fan, na, Faa, na, nn);
    if (na) {
        for (int idx = TID; idx < na * nn; idx += NT) Fold[idx] = Fan[idx];
        SYNC();
        // K_an = F_an - Lt F_nn
        mm<W>(Fan, na, na, nn, nn, -1.0, Ltan, 1, nj, Fnn, 1, nn, true, false);
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
        SYNC();
    }
    // M_nn = D^{-1} K_nn D^{-1}, D = L L^T   (half factor G: L^{-1} K_nn L^{-T})
    trsm_ll<W>(Lb, nj, nn, Fnn, nn, nn);
    trsm_rlt<W>(Lb, nj, nn, Fnn, nn, nn);
    if (!a.half) {
        trsm_llt<W>(Lb, nj, nn, Fnn, nn, nn);
        trsm_rl<W>(Lb, nj, nn, Fnn, nn, nn);
    }
    if (na) {
        // M_an = Y_aa K_an D^{-1}   (half factor G: R^T K_an L^{-T}, Y_aa = R R^T)
        trsm_rlt<W>(Lb, nj, nn, Fan, na, na);
        if (!a.half) {
            trsm_rl<W>(Lb, nj, nn, Fan, na, na);
            mm<W>(blk + nn, nj, na, nn, na, 1.0, Yaa, 1, na, Fan, 1, na, false, false);
        } else {
            mm<W>(blk + nn, nj, na, nn, na, 1.0, a.Raa + q.uoff, na, 1, Fan, 1, na, false, false);
        }
    }
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        blk[i + (long long)j * nj] = (i >= j) ? 0.5 * (Fnn[idx] + Fnn[j + i * nn]) : 0.0;
    }
    SYNC();
}

// forward Hessian, pass 3 (App. A.4 step 3), root to leaves
template <bool W>
__device__ void op_hfwd_down(const TreeArgs &a, const Node &q, int b, double *ws) {
    const int nn = q.nn, na = q.na, nj = q.nj;
    const SymDev &S = a.S;
    double *Xb = a.X + (long long)b * S.nblk;
    double *blk = Xb + q.boff;
    if (a.half) {
        // G^adj: the adjoint of the half scaling on this supernode's own block, then pass 3:
        // M_nn = L^{-T} V_nn L^{-1}, M_an = R V_an L^{-1}
        const double *Lb = a.Lt + q.boff;
        double *Vnn = ws;                 // nn x nn full
        double *Van = Vnn + nn * nn;      // na x nn
        for (int idx = TID; idx < nn * nn; idx += NT) {
            int i = idx % nn, j = idx / nn;
            Vnn[idx] = (i >= j) ? blk[i + (long long)j * nj] : blk[j + (long long)i * nj];
        }
        for (int idx = TID; idx < na * nn; idx += NT) Van[idx] = blk[nn + idx % na + (long long)(idx / na) * nj];
        SYNC();
        trsm_llt<W>(Lb, nj, nn, Vnn, nn, nn);
        trsm_rl<W>(Lb, nj, nn, Vnn, nn, nn);
        for (int idx = TID; idx < nn * nn; idx += NT) {
            int i = idx % nn, j = idx / nn;
            blk[i + (long long)j * nj] = (i >= j) ? 0.5 * (Vnn[idx] + Vnn[j + i * nn]) : 0.0;
        }
        if (na) {
            trsm_rl<W>(Lb, nj, nn, Van, na, na);
            mm<W>(blk + nn, nj, na, nn, na, 1.0, a.Raa + q.uoff, 1, na, Van, 1, na, false, false);
        }
        SYNC();
    }
    if (!na) return;
    const double *Ltan = a.Lt + q.boff + nn;
    double *Zaa = ws;                 // na x na
    double *Mold = Zaa + na * na;     // na x nn
    double *Tn = Mold + na * nn;      // nn x nn
    gather_aa<W>(S, q, Xb, Zaa);
    for (int idx = TID; idx < na * nn; idx += NT) Mold[idx] = blk[nn + idx % na + (long long)(idx / na) * nj];
    SYNC();
    // Z_an = M_an - Z_aa Lt
    for (int idx = TID; idx < na * nn; idx += NT) {
        int i = idx % na, c = idx / na;
        double s = 0.0;
        for (int r = 0; r < na; ++r) s = fma(Zaa[i + r * na], Ltan[r + (long long)c * nj], s);
        blk[nn + i + (long long)c * nj] = Mold[idx] - s;
    }
    SYNC();
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
    SYNC();
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        if (i >= j) blk[i + (long long)j * nj] = Tn[idx];
    }
    SYNC();
}

// inverse Hessian (App. A.5), one sweep leaves to root
template <bool W>
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
    gather_aa<W>(S, q, Xb, T3);
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        int kmax = i < j ? i : j;
        double s = 0.0;
        for (int r = 0; r <= kmax; ++r) s = fma(Lb[i + (long long)r * nj], Lb[j + (long long)r * nj], s);
        // half factors use L itself (lower triangular, full storage) instead of D = L L^T
        T4[idx] = a.half ? ((i >= j) ? Lb[i + (long long)j * nj] : 0.0) : s;
    }
    SYNC();
    if (a.half != 2) {
        // M_an = Z_an + Z_aa Lt
        for (int idx = TID; idx < na * nn; idx += NT) {
            int i = idx % na, c = idx / na;
            double s = 0.0;
            for (int r = 0; r < na; ++r) s = fma(T3[i + r * na], Ltan[r + (long long)c * nj], s);
            T1[idx] = blk[nn + i + (long long)c * nj] + s;
        }
        SYNC();
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
        SYNC();
    } else {
        // G^-1: the block itself is (V_nn, V_an)
        for (int idx = TID; idx < na * nn; idx += NT) T1[idx] = blk[nn + idx % na + (long long)(idx / na) * nj];
        for (int idx = TID; idx < nn * nn; idx += NT) {
            int i = idx % nn, j = idx / nn;
            T2[idx] = (i >= j) ? blk[i + (long long)j * nj] : blk[j + (long long)i * nj];
        }
        SYNC();
    }
    if (a.half == 1) {
        // G^-adj: V_nn = L^T M_nn L, V_an = R^-1 M_an L; written to the block, nothing is passed up
        mm<W>(T5, nn, nn, nn, nn, 1.0, T4, nn, 1, T2, 1, nn, false, false);       // L^T M
        mm<W>(T2, nn, nn, nn, nn, 1.0, T5, 1, nn, T4, 1, nn, false, false);       // (L^T M) L
        if (na) {
            mm<W>(T6, na, na, nn, nn, 1.0, T1, 1, na, T4, 1, nn, false, false);   // M_an L
            trsm_ll<W>(a.Raa + q.uoff, na, na, T6, na, nn);
        }
        for (int idx = TID; idx < nn * nn; idx += NT) {
            int i = idx % nn, j = idx / nn;
            blk[i + (long long)j * nj] = (i >= j) ? 0.5 * (T2[idx] + T2[j + i * nn]) : 0.0;
        }
        for (int idx = TID; idx < na * nn; idx += NT) blk[nn + idx % na + (long long)(idx / na) * nj] = T6[idx];
        SYNC();
        return;
    }
    // K_nn = D M_nn D   (G^-1: L V_nn L^T)
    if (a.half == 2) {
        mm<W>(T5, nn, nn, nn, nn, 1.0, T4, 1, nn, T2, 1, nn, false, false);       // L V
        mm<W>(T2, nn, nn, nn, nn, 1.0, T5, 1, nn, T4, nn, 1, false, false);       // (L V) L^T
    } else {
        mm<W>(T5, nn, nn, nn, nn, 1.0, T4, 1, nn, T2, 1, nn, false, false);
        mm<W>(T2, nn, nn, nn, nn, 1.0, T5, 1, nn, T4, 1, nn, false, false);
    }
    if (na) {
        // K_an = Y_aa^{-1} M_an D   (G^-1: R^-T V_an L^T)
        const double *R = a.Raa + q.uoff;
        if (a.half == 2) {
            mm<W>(T6, na, na, nn, nn, 1.0, T1, 1, na, T4, nn, 1, false, false);   // V_an L^T
        } else {
            mm<W>(T6, na, na, nn, nn, 1.0, T1, 1, na, T4, 1, nn, false, false);
            trsm_ll<W>(R, na, na, T6, na, nn);
        }
        trsm_llt<W>(R, na, na, T6, na, nn);
        // F_an = K_an + Lt K_nn
        for (int idx = TID; idx < na * nn; idx += NT) {
            int i = idx % na, c = idx / na;
            double s = 0.0;
            for (int r = 0; r < nn; ++r) s = fma(Ltan[i + (long long)r * nj], T2[r + c * nn], s);
            T1[idx] = T6[idx] + s;
        }
        SYNC();
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
        SYNC();
    }
    for (int ch = S.chptr[q.k]; ch < S.chptr[q.k + 1]; ++ch)
        extend_add_full<W>(S, S.chidx[ch], ub, T2, nn, T1, na, T3, na, nn);
    for (int idx = TID; idx < nn * nn; idx += NT) {
        int i = idx % nn, j = idx / nn;
        blk[i + (long long)j * nj] = (i >= j) ? 0.5 * (T2[idx] + T2[j + i * nn]) : 0.0;
    }
    for (int idx = TID; idx < na * nn; idx += NT) blk[nn + idx % na + (long long)(idx / na) * nj] = T1[idx];
    for (int idx = TID; idx < na * na; idx += NT) Uk[idx] = T3[idx];
    SYNC();
}
#undef TID
#undef NT
#undef SYNC

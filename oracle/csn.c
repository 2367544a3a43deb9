/* TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): plain-C restatement of oracle/supernodal.py:trsm
 * (chompack.trsm at /root/reference/src/python/solvers.py:491-492, 1921-1922), used by
 * bench.py's cpu_baseline / reference arm so that the per-column loop of the sparse-constraint Schur assembly is
 * timed on compiled code with all host cores (like chompack's C) instead of on a NumPy loop over 800 supernodes.
 * tests/test_oracle_drivers.py::test_c_trsm_matches_numpy pins it on the NumPy version.
 *
 *   B (n x k, ROW-major, rows in the internal order of the symbolic object) <- L^-1 B (trans = 0) or L^-T B (trans = 1)
 *   supernode s: nn[s] columns, row list rowidx[rowptr[s] .. rowptr[s+1]) (own rows first), dense column-major block
 *   of shape (nn+na) x nn at blkval[blkptr[s]].
 * A right-hand side never interacts with another one: the threads split the COLUMNS of B and run the whole sweep
 * without synchronisation.
 */
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void sweep(int64_t nsn, const int64_t *nn_, const int64_t *rowptr, const int64_t *rowidx, const int64_t *blkptr,
                  const double *L, double *B, int64_t k, int64_t c0, int64_t c1, int trans) {
    if (!trans) {
        for (int64_t s = 0; s < nsn; ++s) {
            const int64_t nn = nn_[s], nj = rowptr[s + 1] - rowptr[s], na = nj - nn;
            const int64_t *rows = rowidx + rowptr[s];
            const double *blk = L + blkptr[s];
            for (int64_t j = 0; j < nn; ++j) {
                double *bj = B + rows[j] * k;
                const double d = blk[j + j * nj];
                for (int64_t c = c0; c < c1; ++c) bj[c] /= d;
                for (int64_t r = j + 1; r < nn; ++r) {
                    const double l = blk[r + j * nj];
                    double *br = B + rows[r] * k;
                    for (int64_t c = c0; c < c1; ++c) br[c] -= l * bj[c];
                }
            }
            for (int64_t i = 0; i < na; ++i) {
                double *bi = B + rows[nn + i] * k;
                for (int64_t j = 0; j < nn; ++j) {
                    const double l = blk[nn + i + j * nj];
                    const double *bj = B + rows[j] * k;
                    for (int64_t c = c0; c < c1; ++c) bi[c] -= l * bj[c];
                }
            }
        }
    } else {
        for (int64_t s = nsn - 1; s >= 0; --s) {
            const int64_t nn = nn_[s], nj = rowptr[s + 1] - rowptr[s], na = nj - nn;
            const int64_t *rows = rowidx + rowptr[s];
            const double *blk = L + blkptr[s];
            for (int64_t i = 0; i < na; ++i) {
                const double *bi = B + rows[nn + i] * k;
                for (int64_t j = 0; j < nn; ++j) {
                    const double l = blk[nn + i + j * nj];
                    double *bj = B + rows[j] * k;
                    for (int64_t c = c0; c < c1; ++c) bj[c] -= l * bi[c];
                }
            }
            for (int64_t j = nn - 1; j >= 0; --j) {
                double *bj = B + rows[j] * k;
                const double d = blk[j + j * nj];
                for (int64_t c = c0; c < c1; ++c) bj[c] /= d;
                for (int64_t r = 0; r < j; ++r) {
                    const double l = blk[j + r * nj];
                    double *br = B + rows[r] * k;
                    for (int64_t c = c0; c < c1; ++c) br[c] -= l * bj[c];
                }
            }
        }
    }
}

/* returns the number of threads used */
int csn_trsm(int64_t nsn, const int64_t *nn, const int64_t *rowptr, const int64_t *rowidx, const int64_t *blkptr,
             const double *L, double *B, int64_t k, int trans) {
    int used = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
        const int nt = omp_get_num_threads(), t = omp_get_thread_num();
        /* at least 8 columns per thread: shorter slices waste the vector units */
        int64_t parts = (k + 7) / 8;
        if (parts > nt) parts = nt;
        if (parts < 1) parts = 1;
        if (t == 0) used = (int)parts;
        if (t < parts) {
            const int64_t c0 = k * t / parts, c1 = k * (t + 1) / parts;
            sweep(nsn, nn, rowptr, rowidx, blkptr, L, B, k, c0, c1, trans);
        }
    }
#else
    sweep(nsn, nn, rowptr, rowidx, blkptr, L, B, k, 0, k, trans);
#endif
    return used;
}

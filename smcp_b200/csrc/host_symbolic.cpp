// Host-side symbolic analysis in native code: the once-per-solve step
// immediately before the GPU hot path (reference call sites src/python/solvers.py:278-319,
// 1542-1560: cvxopt.amd.order / chompack.maxcardsearch / chompack.symbolic — not vendored).
//
// These are the native twins of smcp_b200/symbolic.py (min_degree, maxcardsearch, embed): the
// SAME tie-breaking rules, so both produce bit-identical orderings and filled patterns (tested
// in tests/test_host_symbolic.py).  The Python versions stay as the readable specification; the
// drivers call the native ones (n = 20 000 max-cut: minimum degree 553 s in Python).
//
// Input everywhere: lower-triangular CCS pattern (colptr[n+1], rowind[nnz]) with int64 indices
// like CVXOPT's int_t (src/C/cvxopt.h:46).  No CUDA in this file.
#include "../../include/smcp_b200.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <queue>
#include <utility>
#include <vector>

namespace {

// symmetric adjacency without the diagonal, neighbours sorted increasingly
void build_adjacency(int64_t n, const int64_t *colptr, const int64_t *rowind, std::vector<int64_t> &ptr,
                     std::vector<int32_t> &idx) {
    ptr.assign(n + 1, 0);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t q = colptr[j]; q < colptr[j + 1]; ++q) {
            const int64_t i = rowind[q];
            if (i == j) continue;
            ++ptr[i + 1];
            ++ptr[j + 1];
        }
    for (int64_t v = 0; v < n; ++v) ptr[v + 1] += ptr[v];
    idx.resize((size_t)ptr[n]);
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t q = colptr[j]; q < colptr[j + 1]; ++q) {
            const int64_t i = rowind[q];
            if (i == j) continue;
            idx[(size_t)fill[i]++] = (int32_t)j;
            idx[(size_t)fill[j]++] = (int32_t)i;
        }
    for (int64_t v = 0; v < n; ++v) {
        std::sort(idx.begin() + ptr[v], idx.begin() + ptr[v + 1]);
        // duplicates cannot occur for a valid CCS pattern, but be safe
    }
}

}  // namespace

extern "C" {

// Exact minimum (external) degree on the elimination graph, ties by smallest vertex index:
// at every step the not-yet-eliminated vertex with the smallest (degree, index) goes next and its
// neighbourhood becomes a clique.  (smcp_b200/symbolic.py:min_degree; stand-in for
// cvxopt.amd.order, solvers.py:192-198, 278-279.)
int smcp_host_min_degree(int64_t n, const int64_t *colptr, const int64_t *rowind, int64_t *perm) {
    if (n < 0 || !colptr || !perm) return -1;
    if (n > INT32_MAX) return -1;
    std::vector<int64_t> ptr;
    std::vector<int32_t> idx;
    build_adjacency(n, colptr, rowind, ptr, idx);
    std::vector<std::vector<int32_t>> adj((size_t)n);
    for (int64_t v = 0; v < n; ++v) adj[(size_t)v].assign(idx.begin() + ptr[v], idx.begin() + ptr[v + 1]);
    idx.clear();
    idx.shrink_to_fit();
    std::vector<int32_t> deg((size_t)n);
    std::vector<char> done((size_t)n, 0);
    typedef std::pair<int32_t, int32_t> Key;      // (degree, vertex)
    std::priority_queue<Key, std::vector<Key>, std::greater<Key>> heap;
    for (int64_t v = 0; v < n; ++v) {
        deg[(size_t)v] = (int32_t)adj[(size_t)v].size();
        heap.push(Key(deg[(size_t)v], (int32_t)v));
    }
    std::vector<int32_t> merged;
    int64_t k = 0, remaining = n;
    while (!heap.empty()) {
        const Key top = heap.top();
        heap.pop();
        const int32_t v = top.second;
        if (done[(size_t)v] || top.first != deg[(size_t)v]) continue;
        if ((int64_t)top.first == remaining - 1) {
            // the rest of the graph is one clique: every remaining vertex keeps the same degree
            // after each elimination, so the order is by index from here on
            for (int64_t u = 0; u < n; ++u)
                if (!done[(size_t)u]) perm[k++] = u;
            break;
        }
        done[(size_t)v] = 1;
        perm[k++] = v;
        --remaining;
        std::vector<int32_t> nb;
        nb.swap(adj[(size_t)v]);
        for (size_t a = 0; a < nb.size(); ++a) {
            const int32_t u = nb[a];
            std::vector<int32_t> &au = adj[(size_t)u];
            // au <- (au ∪ nb) \ {u, v}
            merged.clear();
            merged.reserve(au.size() + nb.size());
            size_t i = 0, j = 0;
            while (i < au.size() || j < nb.size()) {
                int32_t w;
                if (j >= nb.size() || (i < au.size() && au[i] < nb[j])) w = au[i++];
                else if (i >= au.size() || nb[j] < au[i]) w = nb[j++];
                else { w = au[i]; ++i; ++j; }
                if (w != u && w != v) merged.push_back(w);
            }
            au.assign(merged.begin(), merged.end());
        }
        for (size_t a = 0; a < nb.size(); ++a) {
            const int32_t u = nb[a];
            deg[(size_t)u] = (int32_t)adj[(size_t)u].size();
            heap.push(Key(deg[(size_t)u], u));
        }
    }
    return k == n ? 0 : -2;
}

// Maximum cardinality search; returns the REVERSE visiting order (a perfect elimination ordering
// iff the pattern is chordal).  Ties towards the largest vertex index.
// (smcp_b200/symbolic.py:maxcardsearch; chompack.maxcardsearch, solvers.py:301, 1542.)
int smcp_host_maxcardsearch(int64_t n, const int64_t *colptr, const int64_t *rowind, int64_t *order) {
    if (n < 0 || !colptr || !order) return -1;
    if (n > INT32_MAX) return -1;
    std::vector<int64_t> ptr;
    std::vector<int32_t> idx;
    build_adjacency(n, colptr, rowind, ptr, idx);
    std::vector<int32_t> weight((size_t)n, 0);
    std::vector<char> visited((size_t)n, 0);
    typedef std::pair<int32_t, int32_t> Key;      // (weight, vertex): max-heap on both
    std::priority_queue<Key> heap;
    for (int64_t v = 0; v < n; ++v) heap.push(Key(0, (int32_t)v));
    int64_t k = n - 1;
    while (!heap.empty()) {
        const Key top = heap.top();
        heap.pop();
        const int32_t v = top.second;
        if (visited[(size_t)v] || top.first != weight[(size_t)v]) continue;
        visited[(size_t)v] = 1;
        order[k--] = v;
        for (int64_t q = ptr[v]; q < ptr[v + 1]; ++q) {
            const int32_t u = idx[(size_t)q];
            if (!visited[(size_t)u]) {
                ++weight[(size_t)u];
                heap.push(Key(weight[(size_t)u], u));
            }
        }
    }
    return 0;
}

// Symbolic Cholesky of a lower-triangular pattern that is ALREADY in elimination order:
// elimination tree and column structures of the filled (chordal) pattern.  Two calls:
// frowind == NULL returns the column counts in fcolptr (so the caller can allocate), the second
// call fills frowind (sorted rows, diagonal first).  parent may be NULL.
// (smcp_b200/symbolic.py:embed; the embedding step of chompack.symbolic, solvers.py:305-308.)
int smcp_host_embed(int64_t n, const int64_t *colptr, const int64_t *rowind, int64_t *fcolptr, int64_t *frowind,
                    int64_t *parent) {
    if (n < 0 || !colptr || !fcolptr) return -1;
    // struct(j) = rows > j of column j  ∪  ⋃_{children c} struct(c) \ {j}; children's structures
    // are released as soon as they are merged into the parent when only counting.
    std::vector<std::vector<int64_t>> st((size_t)n);
    std::vector<std::vector<int64_t>> children((size_t)n);
    std::vector<int64_t> mark((size_t)n, -1);
    fcolptr[0] = 0;
    for (int64_t j = 0; j < n; ++j) {
        std::vector<int64_t> &s = st[(size_t)j];
        mark[(size_t)j] = j;
        for (int64_t q = colptr[j]; q < colptr[j + 1]; ++q) {
            const int64_t i = rowind[q];
            if (i > j && mark[(size_t)i] != j) { mark[(size_t)i] = j; s.push_back(i); }
        }
        for (size_t c = 0; c < children[(size_t)j].size(); ++c) {
            std::vector<int64_t> &sc = st[(size_t)children[(size_t)j][c]];
            for (size_t a = 0; a < sc.size(); ++a) {
                const int64_t i = sc[a];
                if (i > j && mark[(size_t)i] != j) { mark[(size_t)i] = j; s.push_back(i); }
            }
            if (!frowind) std::vector<int64_t>().swap(sc);
        }
        std::sort(s.begin(), s.end());
        if (!s.empty()) {
            if (parent) parent[j] = s[0];
            children[(size_t)s[0]].push_back(j);
        } else if (parent) parent[j] = -1;
        fcolptr[j + 1] = fcolptr[j] + 1 + (int64_t)s.size();
        if (frowind) {
            int64_t *out = frowind + fcolptr[j];
            out[0] = j;
            std::copy(s.begin(), s.end(), out + 1);
            // children of j are final; free them
            for (size_t c = 0; c < children[(size_t)j].size(); ++c) std::vector<int64_t>().swap(st[(size_t)children[(size_t)j][c]]);
        }
    }
    return 0;
}

// alpha x alpha gather map of the clique tree (smcp_b200/symbolic.py:Symbolic, "aaidx"): for every
// supernode k with separator alpha_k, the blkval offset of entry (alpha_i, alpha_j) of the chordal
// matrix, column-major na_k x na_k at updptr[k].  An entry lives in the parent's block when its
// column is one of the parent's own columns, otherwise it is found through the parent's map
// (top-down: parents have larger indices).  nupd reaches 1.7e8 on the n = 20 000 max-cut pattern,
// where the NumPy version spent 8 of the 10 seconds of the whole symbolic phase.
int smcp_host_aaidx(int64_t nsn, const int64_t *snpar, const int64_t *nn, const int64_t *nj, const int64_t *relptr,
                    const int64_t *relidx, const int64_t *blkptr, const int64_t *updptr, int64_t *aaidx) {
    if (nsn < 0 || !aaidx) return -1;
    for (int64_t k = nsn - 1; k >= 0; --k) {
        const int64_t a = nj[k] - nn[k];
        if (a == 0) continue;
        const int64_t pk = snpar[k];
        if (pk < 0) return -2;                         // a separator needs a parent
        const int64_t *rel = relidx + relptr[k];
        const int64_t nnp = nn[pk], njp = nj[pk], nap = njp - nnp;
        const int64_t *par = aaidx + updptr[pk];
        int64_t *out = aaidx + updptr[k];
        for (int64_t j = 0; j < a; ++j) {
            const int64_t rj = rel[j];
            for (int64_t i = 0; i < a; ++i) {
                const int64_t ri = rel[i];
                const int64_t lo_r = ri > rj ? ri : rj, lo_c = ri > rj ? rj : ri;
                out[j * a + i] = (lo_c < nnp) ? blkptr[pk] + lo_c * njp + lo_r : par[(lo_c - nnp) * nap + (lo_r - nnp)];
            }
        }
    }
    return 0;
}

// Post-order of a forest (parent = -1: root), children visited in increasing index order -- the rule of
// smcp_b200/symbolic.py:_postorder.
static void forest_postorder(int64_t n, const int64_t *parent, std::vector<int64_t> &post) {
    std::vector<int64_t> head(n, -1), next(n, -1), tail(n, -1), roots;
    for (int64_t v = 0; v < n; ++v) {
        const int64_t p = parent[v];
        if (p < 0) { roots.push_back(v); continue; }
        if (head[p] < 0) head[p] = v; else next[tail[p]] = v;
        tail[p] = v;
    }
    post.clear();
    post.reserve(n);
    std::vector<int64_t> stack, cur(head);
    for (int64_t r : roots) {
        stack.push_back(r);
        while (!stack.empty()) {
            const int64_t v = stack.back();
            const int64_t c = cur[v];
            if (c >= 0) {
                cur[v] = next[c];
                stack.push_back(c);
            } else {
                post.push_back(v);
                stack.pop_back();
            }
        }
    }
}

// Supernode partition, post-ordered relabelling, row lists and relative indices of a filled (chordal)
// pattern given in a perfect elimination ordering: the native twin of the first half of
// smcp_b200/symbolic.py:Symbolic.__init__ (maximal supernodes, Pothen-Sun first-qualifying-child rule;
// stand-in for chompack.symbolic, solvers.py:314, 1555).  Outputs (caller-allocated): nsn; perm[n];
// snptr[n+1]; snpar[n]; rowptr[n+1]; rowidx[n + nnz]; relptr[n+1]; relidx[nnz].  Bit-identical to the
// NumPy specification (tests/test_host_symbolic.py).  Returns -2 when the pattern is not chordal in this order.
int smcp_host_supernodes(int64_t n, const int64_t *colptr, const int64_t *rowind, int64_t *nsn_out, int64_t *perm,
                         int64_t *snptr, int64_t *snpar, int64_t *rowptr, int64_t *rowidx, int64_t *relptr, int64_t *relidx) {
    if (n < 0 || !colptr || !nsn_out) return -1;
    std::vector<int64_t> parent(n, -1), colcount(n);
    for (int64_t j = 0; j < n; ++j) {
        colcount[j] = colptr[j + 1] - colptr[j];
        if (colcount[j] < 1 || rowind[colptr[j]] != j) return -2;          // diagonal first
        if (colcount[j] > 1) parent[j] = rowind[colptr[j] + 1];
    }
    std::vector<int64_t> post;
    forest_postorder(n, parent.data(), post);
    std::vector<int64_t> rep(n, -1);
    for (int64_t j : post) {
        if (rep[j] < 0) rep[j] = j;
        const int64_t pj = parent[j];
        if (pj >= 0 && rep[pj] < 0 && colcount[j] - 1 == colcount[pj]) rep[pj] = rep[j];
    }
    // supernodes numbered by increasing representative; members in path order (the post-order visits them so)
    std::vector<int64_t> snid(n, -1);
    int64_t nsn = 0;
    for (int64_t v = 0; v < n; ++v)
        if (rep[v] == v) snid[v] = nsn++;
    std::vector<int64_t> cnt(nsn + 1, 0);
    for (int64_t v = 0; v < n; ++v) ++cnt[snid[rep[v]] + 1];
    for (int64_t k = 0; k < nsn; ++k) cnt[k + 1] += cnt[k];
    std::vector<int64_t> mem(n), fill(cnt.begin(), cnt.end() - 1);
    for (int64_t j : post) mem[fill[snid[rep[j]]]++] = j;
    std::vector<int64_t> snpar0(nsn, -1);
    for (int64_t k = 0; k < nsn; ++k) {
        const int64_t top = mem[cnt[k + 1] - 1];
        if (parent[top] >= 0) snpar0[k] = snid[rep[parent[top]]];
    }
    std::vector<int64_t> snpost;
    forest_postorder(nsn, snpar0.data(), snpost);
    std::vector<int64_t> renum(nsn);
    for (int64_t knew = 0; knew < nsn; ++knew) renum[snpost[knew]] = knew;
    std::vector<int64_t> iperm(n);
    int64_t k0 = 0;
    snptr[0] = 0;
    for (int64_t knew = 0; knew < nsn; ++knew) {
        const int64_t kold = snpost[knew];
        for (int64_t q = cnt[kold]; q < cnt[kold + 1]; ++q) {
            perm[k0] = mem[q];
            iperm[mem[q]] = k0;
            ++k0;
        }
        snptr[knew + 1] = k0;
        snpar[knew] = snpar0[kold] >= 0 ? renum[snpar0[kold]] : -1;
    }
    rowptr[0] = 0;
    relptr[0] = 0;
    for (int64_t k = 0; k < nsn; ++k) {
        const int64_t nn = snptr[k + 1] - snptr[k];
        const int64_t top = perm[snptr[k + 1] - 1];
        const int64_t na = colcount[top] - 1;
        int64_t *r = rowidx + rowptr[k];
        for (int64_t i = 0; i < nn; ++i) r[i] = snptr[k] + i;
        for (int64_t i = 0; i < na; ++i) r[nn + i] = iperm[rowind[colptr[top] + 1 + i]];
        std::sort(r + nn, r + nn + na);
        rowptr[k + 1] = rowptr[k] + nn + na;
        relptr[k + 1] = relptr[k] + na;
    }
    for (int64_t k = 0; k < nsn; ++k) {
        const int64_t pk = snpar[k];
        if (pk < 0) continue;
        const int64_t nn = snptr[k + 1] - snptr[k], na = rowptr[k + 1] - rowptr[k] - nn;
        const int64_t *a = rowidx + rowptr[k] + nn;
        const int64_t *g = rowidx + rowptr[pk], *gend = rowidx + rowptr[pk + 1];
        for (int64_t i = 0; i < na; ++i) {
            const int64_t *it = std::lower_bound(g, gend, a[i]);
            if (it == gend || *it != a[i]) return -2;
            relidx[relptr[k] + i] = it - g;
        }
    }
    *nsn_out = nsn;
    return 0;
}

}  // extern "C"

// Host symbolic analysis: pattern of the KKT matrix, elimination tree, postorder, column counts,
// relaxed supernodes, supernodal row structure.  See symbolic.hpp for what this replaces in the
// reference (CHOLMOD / LDLFactorizations analyse phase) -- it has no counterpart under
// /root/reference; algorithms are the published ones (Liu 1990; Gilbert-Ng-Peyton 1994;
// Ashcraft-Grimes 1989).
#include "symbolic.hpp"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <stdexcept>

namespace tlp {

// ------------------------------------------------------------------------------------------
// Patterns
// ------------------------------------------------------------------------------------------

// lower(A*A' + I): for each row r of A (via CSR), union of the columns' row sets restricted to >= r.
SymPattern pattern_k1(int64_t m, int64_t n, const int64_t* colptr, const int32_t* rowidx) {
    SymPattern P;
    P.N = (int32_t)m;
    // CSR of A (pattern only)
    std::vector<int64_t> rp(m + 1, 0);
    const int64_t nnz = colptr[n];
    for (int64_t p = 0; p < nnz; ++p) rp[rowidx[p] + 1]++;
    for (int64_t i = 0; i < m; ++i) rp[i + 1] += rp[i];
    std::vector<int32_t> rc(nnz);
    {
        std::vector<int64_t> nxt(rp.begin(), rp.end() - 1);
        for (int64_t j = 0; j < n; ++j)
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) rc[nxt[rowidx[p]]++] = (int32_t)j;
    }
    P.colptr.assign(m + 1, 0);
    std::vector<int32_t> mark(m, -1), buf;
    for (int32_t c = 0; c < (int32_t)m; ++c) {
        buf.clear();
        mark[c] = c;
        buf.push_back(c);
        for (int64_t q = rp[c]; q < rp[c + 1]; ++q) {
            const int32_t j = rc[q];
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                const int32_t i = rowidx[p];
                if (i > c && mark[i] != c) { mark[i] = c; buf.push_back(i); }
            }
        }
        std::sort(buf.begin(), buf.end());
        P.rowidx.insert(P.rowidx.end(), buf.begin(), buf.end());
        P.colptr[c + 1] = (int64_t)P.rowidx.size();
    }
    return P;
}

// lower([-D A'; A R]) with the x-block (n) first, then the y-block (m): column j<n holds the
// diagonal and rows n+i for i in A(:,j); column n+i holds only its diagonal.
SymPattern pattern_k2(int64_t m, int64_t n, const int64_t* colptr, const int32_t* rowidx) {
    SymPattern P;
    P.N = (int32_t)(n + m);
    P.colptr.assign(n + m + 1, 0);
    P.rowidx.reserve(colptr[n] + n + m);
    std::vector<int32_t> buf;
    for (int64_t j = 0; j < n; ++j) {
        P.rowidx.push_back((int32_t)j);
        buf.clear();
        for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) buf.push_back((int32_t)(n + rowidx[p]));
        std::sort(buf.begin(), buf.end());
        buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
        P.rowidx.insert(P.rowidx.end(), buf.begin(), buf.end());
        P.colptr[j + 1] = (int64_t)P.rowidx.size();
    }
    for (int64_t i = 0; i < m; ++i) {
        P.rowidx.push_back((int32_t)(n + i));
        P.colptr[n + i + 1] = (int64_t)P.rowidx.size();
    }
    return P;
}

SymPattern permute_pattern(const SymPattern& P, const std::vector<int32_t>& iperm) {
    const int32_t N = P.N;
    SymPattern Q;
    Q.N = N;
    Q.colptr.assign(N + 1, 0);
    // count: entry (i,j) -> (max(pi,pj), min(pi,pj))
    for (int32_t j = 0; j < N; ++j)
        for (int64_t p = P.colptr[j]; p < P.colptr[j + 1]; ++p) {
            int32_t a = iperm[P.rowidx[p]], b = iperm[j];
            Q.colptr[std::min(a, b) + 1]++;
        }
    for (int32_t j = 0; j < N; ++j) Q.colptr[j + 1] += Q.colptr[j];
    Q.rowidx.resize(Q.colptr[N]);
    std::vector<int64_t> nxt(Q.colptr.begin(), Q.colptr.end() - 1);
    for (int32_t j = 0; j < N; ++j)
        for (int64_t p = P.colptr[j]; p < P.colptr[j + 1]; ++p) {
            int32_t a = iperm[P.rowidx[p]], b = iperm[j];
            Q.rowidx[nxt[std::min(a, b)]++] = std::max(a, b);
        }
    for (int32_t j = 0; j < N; ++j) std::sort(Q.rowidx.begin() + Q.colptr[j], Q.rowidx.begin() + Q.colptr[j + 1]);
    return Q;
}

// ------------------------------------------------------------------------------------------
// Elimination tree (Liu): process rows in order; for row i, walk from every column j<i with
// K(i,j) != 0 up the partially built tree (with path compression through `anc`) and hook the
// root under i.  Works from the lower-CSC pattern by bucketing entries by row first.
// ------------------------------------------------------------------------------------------
std::vector<int32_t> etree_lower(const SymPattern& P) {
    const int32_t N = P.N;
    std::vector<int32_t> parent(N, -1), anc(N, -1);
    // CSR view of the strictly-lower part: row i -> columns j < i
    std::vector<int64_t> rp(N + 1, 0);
    for (int32_t j = 0; j < N; ++j)
        for (int64_t p = P.colptr[j]; p < P.colptr[j + 1]; ++p)
            if (P.rowidx[p] > j) rp[P.rowidx[p] + 1]++;
    for (int32_t i = 0; i < N; ++i) rp[i + 1] += rp[i];
    std::vector<int32_t> rc(rp[N]);
    {
        std::vector<int64_t> nxt(rp.begin(), rp.end() - 1);
        for (int32_t j = 0; j < N; ++j)
            for (int64_t p = P.colptr[j]; p < P.colptr[j + 1]; ++p)
                if (P.rowidx[p] > j) rc[nxt[P.rowidx[p]]++] = j;
    }
    for (int32_t i = 0; i < N; ++i) {
        for (int64_t q = rp[i]; q < rp[i + 1]; ++q) {
            int32_t j = rc[q];
            while (j != -1 && j < i) {
                int32_t jn = anc[j];
                anc[j] = i;
                if (jn == -1) parent[j] = i;
                j = jn;
            }
        }
    }
    return parent;
}

std::vector<int32_t> postorder(const std::vector<int32_t>& parent) {
    const int32_t N = (int32_t)parent.size();
    std::vector<int32_t> head(N, -1), next(N, -1), post;
    post.reserve(N);
    for (int32_t j = N - 1; j >= 0; --j)
        if (parent[j] >= 0) { next[j] = head[parent[j]]; head[parent[j]] = j; }
    std::vector<int32_t> stack;
    for (int32_t r = 0; r < N; ++r) {
        if (parent[r] >= 0) continue;
        stack.push_back(r);
        while (!stack.empty()) {
            int32_t v = stack.back();
            int32_t c = head[v];
            if (c >= 0) { head[v] = next[c]; stack.push_back(c); }
            else { post.push_back(v); stack.pop_back(); }
        }
    }
    return post;
}

// ------------------------------------------------------------------------------------------
// Column counts (Gilbert-Ng-Peyton skeleton/least-common-ancestor method).  P must already be
// postordered (node k is the k-th node of a postorder of its own etree), so that the first
// descendant of j is j - (subtree size) + 1.
// ------------------------------------------------------------------------------------------
std::vector<int32_t> column_counts(const SymPattern& P, const std::vector<int32_t>& parent) {
    const int32_t N = P.N;
    std::vector<int32_t> first(N), size(N, 1), delta(N), maxfirst(N, -1), prevleaf(N, -1), uf(N);
    for (int32_t j = 0; j < N; ++j)
        if (parent[j] >= 0) {
            if (parent[j] <= j) throw std::runtime_error("column_counts: pattern is not postordered");
            size[parent[j]] += size[j];
        }
    for (int32_t j = 0; j < N; ++j) {
        first[j] = j - size[j] + 1;
        delta[j] = (size[j] == 1) ? 1 : 0;      // leaves start at 1
        uf[j] = j;
    }
    auto find = [&](int32_t x) {
        int32_t r = x;
        while (uf[r] != r) r = uf[r];
        while (uf[x] != r) { int32_t nx = uf[x]; uf[x] = r; x = nx; }
        return r;
    };
    for (int32_t j = 0; j < N; ++j) {
        if (parent[j] >= 0) delta[parent[j]]--;  // j is not a root: parent loses the column j itself
        for (int64_t p = P.colptr[j]; p < P.colptr[j + 1]; ++p) {
            const int32_t i = P.rowidx[p];
            if (i <= j) continue;
            // is j a leaf of the row subtree of i ?
            if (first[j] <= maxfirst[i]) continue;
            maxfirst[i] = first[j];
            const int32_t jprev = prevleaf[i];
            prevleaf[i] = j;
            delta[j]++;
            if (jprev >= 0) delta[find(jprev)]--;  // overlap counted at the least common ancestor
        }
        if (parent[j] >= 0) uf[j] = parent[j];
    }
    std::vector<int32_t> cc(delta);
    for (int32_t j = 0; j < N; ++j)
        if (parent[j] >= 0) cc[parent[j]] += cc[j];
    return cc;
}

// ------------------------------------------------------------------------------------------
// Full analysis
// ------------------------------------------------------------------------------------------
void analyze_pattern(const SymPattern& P0, const SymOptions& opt, const int8_t* orig_sign, Symbolic& S) {
    const int32_t N = P0.N;
    S.N = N;
    static const bool trace = getenv("TLPB200_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_mark = now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        const double t = now();
        fprintf(stderr, "tlpb200 setup:   %-20s %8.3f s\n", what, t - t_mark);
        t_mark = t;
    };

    // 1. fill-reducing ordering
    std::vector<int32_t> perm1;
    if (opt.ordering == 1) perm1 = amd_order(P0, opt.dense_row_factor);
    else { perm1.resize(N); std::iota(perm1.begin(), perm1.end(), 0); }
    if ((int32_t)perm1.size() != N) throw std::runtime_error("ordering did not return a permutation");
    std::vector<int32_t> ip1(N, -1);
    for (int32_t k = 0; k < N; ++k) {
        if (perm1[k] < 0 || perm1[k] >= N || ip1[perm1[k]] != -1) throw std::runtime_error("ordering is not a permutation");
        ip1[perm1[k]] = k;
    }
    lap("ordering (AMD)");
    SymPattern P1 = permute_pattern(P0, ip1);

    // 2. etree + postorder, composed into the permutation
    std::vector<int32_t> par1 = etree_lower(P1);
    std::vector<int32_t> post = postorder(par1);      // post[k] = node (in P1 numbering) visited k-th
    std::vector<int32_t> ipost(N);
    for (int32_t k = 0; k < N; ++k) ipost[post[k]] = k;
    S.perm.resize(N);
    S.iperm.resize(N);
    for (int32_t k = 0; k < N; ++k) S.perm[k] = perm1[post[k]];
    for (int32_t k = 0; k < N; ++k) S.iperm[S.perm[k]] = k;
    SymPattern Pp = permute_pattern(P1, ipost);
    S.parent.assign(N, -1);
    for (int32_t v = 0; v < N; ++v)
        if (par1[v] >= 0) S.parent[ipost[v]] = ipost[par1[v]];

    lap("etree + postorder");
    // 3. column counts
    S.colcount = column_counts(Pp, S.parent);
    lap("column counts");
    S.nnzL = 0;
    S.flops = 0.0;
    for (int32_t j = 0; j < N; ++j) { S.nnzL += S.colcount[j]; S.flops += (double)S.colcount[j] * (double)S.colcount[j]; }

    // 4. fundamental supernodes: j+1 joins j when parent[j]==j+1 and count[j+1]==count[j]-1
    std::vector<int32_t> first;     // first column per supernode
    first.push_back(0);
    for (int32_t j = 0; j + 1 < N; ++j)
        if (!(S.parent[j] == j + 1 && S.colcount[j + 1] == S.colcount[j] - 1)) first.push_back(j + 1);
    int32_t ns = N > 0 ? (int32_t)first.size() : 0;
    if (N == 0) first.clear();
    first.push_back(N);

    // 5. relaxed amalgamation (bottom-up; a child can only merge into its parent when its columns
    //    immediately precede the parent's, i.e. it is the last child in postorder)
    {
        std::vector<int32_t> c2s(N);
        for (int32_t s = 0; s < ns; ++s) for (int32_t j = first[s]; j < first[s + 1]; ++j) c2s[j] = s;
        std::vector<int32_t> f(ns), l(ns), nbelow(ns), spar(ns);
        std::vector<double> zeros(ns, 0.0);
        std::vector<char> dead(ns, 0);
        for (int32_t s = 0; s < ns; ++s) {
            f[s] = first[s];
            l[s] = first[s + 1];
            nbelow[s] = S.colcount[f[s]] - (l[s] - f[s]);
            int32_t pj = S.parent[l[s] - 1];
            spar[s] = pj >= 0 ? c2s[pj] : -1;
        }
        // a dead supernode has been merged into its parent, so "the live owner of x" is found by
        // following parent links through dead nodes
        auto live = [&](int32_t x) { while (x >= 0 && dead[x]) x = spar[x]; return x; };
        for (int32_t p = 0; p < ns; ++p) {
            while (true) {
                if (f[p] == 0) break;
                int32_t c = live(c2s[f[p] - 1]);     // live supernode whose columns end at f[p]
                if (c == p) break;
                if (live(spar[c]) != p) break;
                const double nc = (double)(l[c] - f[c]), np = (double)(l[p] - f[p]);
                const double newz = zeros[c] + zeros[p] + nc * (np + (double)nbelow[p] - (double)nbelow[c]);
                const double ncol = nc + np;
                const double tot = ncol * (ncol + 1) / 2 + ncol * (double)nbelow[p];
                const double frac = newz / tot;
                bool merge = (ncol <= opt.relax_always) ||
                             (ncol <= opt.relax_ncol1 && frac < opt.relax_frac1) ||
                             (ncol <= opt.relax_ncol2 && frac < opt.relax_frac2) ||
                             (frac < opt.relax_frac3);
                if (!merge) break;
                f[p] = f[c];
                zeros[p] = newz;
                dead[c] = 1;
                spar[c] = p;
            }
        }
        std::vector<int32_t> nf;
        for (int32_t s = 0; s < ns; ++s) if (!dead[s]) nf.push_back(f[s]);
        std::sort(nf.begin(), nf.end());
        nf.push_back(N);
        first.swap(nf);
        ns = (int32_t)first.size() - 1;
    }
    S.nsuper = ns;
    S.sn_first = first;
    S.col2sn.resize(N);
    for (int32_t s = 0; s < ns; ++s) for (int32_t j = first[s]; j < first[s + 1]; ++j) S.col2sn[j] = s;

    lap("supernodes + relax");
    // 6. supernodal row structure: rows(s) = own columns, then sorted union of
    //    {pattern rows >= last col} and {children's below rows} minus own columns
    S.sn_parent.assign(ns, -1);
    S.sn_rowptr.assign(ns + 1, 0);
    S.sn_rows.clear();
    {
        std::vector<int32_t> mark(N, -1), buf;
        std::vector<int32_t> chead(ns, -1), cnext(ns, -1);
        for (int32_t s = 0; s < ns; ++s) {
            const int32_t f = first[s], l = first[s + 1];
            buf.clear();
            for (int32_t j = f; j < l; ++j)
                for (int64_t p = Pp.colptr[j]; p < Pp.colptr[j + 1]; ++p) {
                    int32_t i = Pp.rowidx[p];
                    if (i >= l && mark[i] != s) { mark[i] = s; buf.push_back(i); }
                }
            for (int32_t c = chead[s]; c >= 0; c = cnext[c]) {
                const int64_t b = S.sn_rowptr[c] + (first[c + 1] - first[c]), e = S.sn_rowptr[c + 1];
                for (int64_t q = b; q < e; ++q) {
                    int32_t i = S.sn_rows[q];
                    if (i >= l && mark[i] != s) { mark[i] = s; buf.push_back(i); }
                }
            }
            std::sort(buf.begin(), buf.end());
            for (int32_t j = f; j < l; ++j) S.sn_rows.push_back(j);
            S.sn_rows.insert(S.sn_rows.end(), buf.begin(), buf.end());
            S.sn_rowptr[s + 1] = (int64_t)S.sn_rows.size();
            if (!buf.empty()) {
                int32_t ps = S.col2sn[buf[0]];
                S.sn_parent[s] = ps;
                cnext[s] = chead[ps];
                chead[ps] = s;
            }
        }
    }

    lap("row structure");
    // 7. panel storage
    S.sn_xptr.assign(ns + 1, 0);
    S.max_ncol = S.max_nrow = 0;
    S.nnzL_relaxed = 0;
    for (int32_t s = 0; s < ns; ++s) {
        const int64_t ncol = first[s + 1] - first[s];
        const int64_t nrow = S.sn_rowptr[s + 1] - S.sn_rowptr[s];
        S.sn_xptr[s + 1] = S.sn_xptr[s] + nrow * ncol;
        S.max_ncol = std::max<int32_t>(S.max_ncol, (int32_t)ncol);
        S.max_nrow = std::max<int32_t>(S.max_nrow, (int32_t)nrow);
        S.nnzL_relaxed += ncol * (ncol + 1) / 2 + ncol * (nrow - ncol);
    }
    S.lx_size = S.sn_xptr[ns];
    S.diagpos.resize(N);
    S.sign.resize(N);
    for (int32_t s = 0; s < ns; ++s) {
        const int64_t nrow = S.sn_rowptr[s + 1] - S.sn_rowptr[s];
        for (int32_t j = first[s]; j < first[s + 1]; ++j) {
            const int64_t lc = j - first[s];
            S.diagpos[j] = S.sn_xptr[s] + lc * nrow + lc;
        }
    }
    for (int32_t q = 0; q < N; ++q) S.sign[q] = orig_sign ? orig_sign[S.perm[q]] : (int8_t)1;
}

}  // namespace tlp

// Approximate-minimum-degree ordering on a quotient graph (host, once per setup()).
//
// Stands in for the fill-reducing ordering CHOLMOD computes inside
// cholesky(Symmetric(K)) / ldlt(Symmetric(K))  (/root/reference/src/KKT/Cholmod/spd.jl:17,
// sqd.jl:19) and LDLFactorizations' ldl_analyze (ldlfact.jl:77).  Written from the published
// algorithm (Amestoy, Davis, Duff, "An approximate minimum degree ordering algorithm",
// SIMAX 1996): quotient graph with elements, approximate external degrees via the |Le \ Lp|
// trick, aggressive element absorption, mass elimination, hash-based supervariable detection,
// dense-row deferral.  Storage is plain std::vector per node rather than AMD's in-place
// compressed workspace -- setup() is not on the per-iteration path.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "symbolic.hpp"

namespace tlp {

namespace {

struct DegLists {
    std::vector<int32_t> head, next, prev;
    explicit DegLists(int32_t n) : head(n + 1, -1), next(n, -1), prev(n, -1) {}
    void insert(int32_t v, int32_t d) {
        next[v] = head[d];
        prev[v] = -1;
        if (head[d] >= 0) prev[head[d]] = v;
        head[d] = v;
    }
    void remove(int32_t v, int32_t d) {
        if (prev[v] >= 0) next[prev[v]] = next[v]; else head[d] = next[v];
        if (next[v] >= 0) prev[next[v]] = prev[v];
        next[v] = prev[v] = -1;
    }
};

}  // namespace

std::vector<int32_t> amd_order(const SymPattern& P, int dense_row_factor) {
    const int32_t n = P.N;
    std::vector<int32_t> result;
    result.reserve(n);
    if (n == 0) return result;

    // ---- full adjacency (both triangles, diagonal dropped)
    std::vector<int32_t> deg(n, 0);
    for (int32_t j = 0; j < n; ++j)
        for (int64_t p = P.colptr[j]; p < P.colptr[j + 1]; ++p) {
            int32_t i = P.rowidx[p];
            if (i != j) { deg[i]++; deg[j]++; }
        }
    // dense rows are removed from the graph and ordered last
    double thr = dense_row_factor > 0 ? std::max(16.0, dense_row_factor * std::sqrt((double)n)) : 2.0 * n;
    std::vector<char> dense(n, 0);
    int32_t ndense = 0;
    for (int32_t i = 0; i < n; ++i)
        if ((double)deg[i] > thr) { dense[i] = 1; ndense++; }

    std::vector<std::vector<int32_t>> av(n), ae(n), le(n);
    for (int32_t i = 0; i < n; ++i)
        if (!dense[i]) av[i].reserve(deg[i]);
    for (int32_t j = 0; j < n; ++j)
        for (int64_t p = P.colptr[j]; p < P.colptr[j + 1]; ++p) {
            int32_t i = P.rowidx[p];
            if (i == j || dense[i] || dense[j]) continue;
            av[i].push_back(j);
            av[j].push_back(i);
        }

    std::vector<int32_t> nv(n, 1);          // supervariable weight (0 = not a live variable)
    std::vector<int32_t> weight(n, 1);      // weight at the time of elimination / merge
    std::vector<int32_t> degree(n, 0);
    std::vector<char> elem_alive(n, 0);
    std::vector<int64_t> esize(n, 0);
    std::vector<int64_t> w(n, 0);
    int64_t wbase = 1;
    std::vector<int32_t> mark(n, -1);
    int32_t tag = 0;
    std::vector<int32_t> merged_into(n, -1);
    std::vector<int32_t> order;             // principal pivots in elimination order
    order.reserve(n);

    DegLists dl(n);
    int64_t remaining = 0;                  // total weight of live variables
    for (int32_t i = 0; i < n; ++i) {
        if (dense[i]) { nv[i] = 0; continue; }
        degree[i] = (int32_t)av[i].size();
        dl.insert(i, degree[i]);
        remaining += 1;
    }

    std::vector<int32_t> Lp, keepE, keepV;
    std::vector<int64_t> degout;            // outside degree of members of Lp (parallel to Lp)
    std::vector<std::pair<uint64_t, int32_t>> hashes;
    std::vector<int32_t> mark2(n, -1);
    int32_t tag2 = 0;
    int32_t mindeg = 0;

    while (remaining > 0) {
        while (mindeg <= n && dl.head[mindeg] < 0) ++mindeg;
        const int32_t p = dl.head[mindeg];
        dl.remove(p, degree[p]);

        // ---- form Lp = (A_p  U  union of L_e, e in E_p) \ {p}
        ++tag;
        mark[p] = tag;
        Lp.clear();
        for (int32_t v : av[p])
            if (nv[v] > 0 && mark[v] != tag) { mark[v] = tag; Lp.push_back(v); }
        for (int32_t e : ae[p]) {
            if (!elem_alive[e]) continue;
            for (int32_t v : le[e])
                if (nv[v] > 0 && mark[v] != tag) { mark[v] = tag; Lp.push_back(v); }
            elem_alive[e] = 0;                          // e is absorbed into p
            std::vector<int32_t>().swap(le[e]);
        }
        std::vector<int32_t>().swap(av[p]);
        std::vector<int32_t>().swap(ae[p]);
        const int32_t nvp = nv[p];
        weight[p] = nvp;
        nv[p] = 0;
        remaining -= nvp;
        order.push_back(p);

        int64_t wLp = 0;
        for (int32_t v : Lp) {
            wLp += nv[v];
            dl.remove(v, degree[v]);
        }

        // ---- w[e] - wbase = |L_e \ L_p| for every element adjacent to a member of Lp
        for (int32_t v : Lp)
            for (int32_t e : ae[v]) {
                if (!elem_alive[e]) continue;
                if (w[e] < wbase) w[e] = wbase + esize[e];
                w[e] -= nv[v];
            }

        // ---- first pass: prune lists, outside degree, mass elimination, hash
        degout.assign(Lp.size(), 0);
        hashes.clear();
        for (size_t idx = 0; idx < Lp.size(); ++idx) {
            const int32_t v = Lp[idx];
            int64_t d = 0;
            uint64_t h = 0;
            keepE.clear();
            for (int32_t e : ae[v]) {
                if (!elem_alive[e]) continue;
                int64_t we = w[e] - wbase;
                if (we <= 0) {                          // L_e is a subset of L_p: aggressive absorption
                    elem_alive[e] = 0;
                    std::vector<int32_t>().swap(le[e]);
                    continue;
                }
                keepE.push_back(e);
                d += we;
                h += (uint64_t)e;
            }
            keepV.clear();
            for (int32_t u : av[v]) {
                if (nv[u] <= 0 || mark[u] == tag) continue;   // dead, merged, or now covered by element p
                keepV.push_back(u);
                d += nv[u];
                h += (uint64_t)u;
            }
            if (d == 0) {
                // v is adjacent to nothing but element p: indistinguishable from p, eliminate now
                merged_into[v] = p;
                weight[v] = nv[v];
                wLp -= nv[v];
                remaining -= nv[v];
                nv[v] = 0;
                std::vector<int32_t>().swap(av[v]);
                std::vector<int32_t>().swap(ae[v]);
                continue;
            }
            keepE.push_back(p);
            h += (uint64_t)p;
            ae[v].assign(keepE.begin(), keepE.end());
            av[v].assign(keepV.begin(), keepV.end());
            degout[idx] = d;
            hashes.emplace_back(h, (int32_t)idx);
        }

        // ---- supervariable detection among the members of Lp
        std::sort(hashes.begin(), hashes.end());
        for (size_t a = 0; a < hashes.size();) {
            size_t b = a;
            while (b < hashes.size() && hashes[b].first == hashes[a].first) ++b;
            for (size_t x = a; x < b; ++x) {
                const int32_t i = Lp[hashes[x].second];
                if (nv[i] <= 0) continue;
                ++tag2;
                for (int32_t u : av[i]) mark2[u] = tag2;
                for (int32_t e : ae[i]) mark2[e] = tag2;   // variables and elements share the id space
                for (size_t y = x + 1; y < b; ++y) {
                    const int32_t j = Lp[hashes[y].second];
                    if (nv[j] <= 0) continue;
                    if (av[j].size() != av[i].size() || ae[j].size() != ae[i].size()) continue;
                    bool same = true;
                    for (int32_t u : av[j]) if (mark2[u] != tag2) { same = false; break; }
                    if (same) for (int32_t e : ae[j]) if (mark2[e] != tag2) { same = false; break; }
                    if (!same) continue;
                    // j is indistinguishable from i
                    merged_into[j] = i;
                    weight[j] = nv[j];
                    nv[i] += nv[j];
                    nv[j] = 0;
                    std::vector<int32_t>().swap(av[j]);
                    std::vector<int32_t>().swap(ae[j]);
                }
            }
            a = b;
        }

        // ---- second pass: final approximate degrees, new element, degree lists
        std::vector<int32_t>& Lnew = le[p];
        Lnew.clear();
        int64_t sz = 0;
        for (size_t idx = 0; idx < Lp.size(); ++idx) {
            const int32_t v = Lp[idx];
            if (nv[v] <= 0) continue;
            const int64_t ext = wLp - nv[v];
            int64_t d = std::min<int64_t>((int64_t)degree[v] + ext, degout[idx] + ext);
            d = std::min<int64_t>(d, remaining - nv[v]);
            if (d < 0) d = 0;
            degree[v] = (int32_t)d;
            dl.insert(v, degree[v]);
            if (degree[v] < mindeg) mindeg = degree[v];
            Lnew.push_back(v);
            sz += nv[v];
        }
        if (!Lnew.empty()) {
            elem_alive[p] = 1;
            esize[p] = sz;
        }
        wbase += (int64_t)n + 2;
    }

    // ---- emit: each principal pivot followed by everything merged into it (recursively)
    std::vector<int32_t> child_head(n, -1), child_next(n, -1);
    for (int32_t v = n - 1; v >= 0; --v)
        if (merged_into[v] >= 0) { child_next[v] = child_head[merged_into[v]]; child_head[merged_into[v]] = v; }
    std::vector<int32_t> stack;
    for (int32_t p : order) {
        stack.push_back(p);
        while (!stack.empty()) {
            int32_t v = stack.back();
            stack.pop_back();
            result.push_back(v);
            for (int32_t c = child_head[v]; c >= 0; c = child_next[c]) stack.push_back(c);
        }
    }
    // dense rows last, by increasing degree
    std::vector<int32_t> dr;
    for (int32_t i = 0; i < n; ++i) if (dense[i]) dr.push_back(i);
    std::stable_sort(dr.begin(), dr.end(), [&](int32_t a, int32_t b) { return deg[a] < deg[b]; });
    for (int32_t i : dr) result.push_back(i);
    return result;   // result[k] = original index eliminated k-th  (perm[new] = old)
}

}  // namespace tlp

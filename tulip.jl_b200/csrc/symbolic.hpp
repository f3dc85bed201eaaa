// Host-side symbolic analysis for the B200 KKT backend (integer work, once per setup()).
//
// Replaces what the reference delegates to CHOLMOD's analyse phase inside
//   cholesky(Symmetric(K))  -- /root/reference/src/KKT/Cholmod/spd.jl:17
//   ldlt(Symmetric(K))      -- /root/reference/src/KKT/Cholmod/sqd.jl:19
//   ldl_analyze             -- /root/reference/src/KKT/LDLFactorizations/ldlfact.jl:77
// i.e. fill-reducing ordering, elimination tree, column counts, supernode partition.
// Everything here is written from the published algorithms (Amestoy-Davis-Duff approximate
// minimum degree; Liu's elimination tree; Gilbert-Ng-Peyton column counts; relaxed supernode
// amalgamation as described by Ashcraft-Grimes) -- none of it exists in the reference tree.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tlp {

struct SymOptions {
    int ordering = 1;          // 0 = natural, 1 = approximate minimum degree
    int relax_always = 8;      // merge child into parent whenever merged width <= this
    int relax_ncol1 = 32;      // ... or width <= relax_ncol1 and zero fraction < relax_frac1
    double relax_frac1 = 0.50;
    int relax_ncol2 = 128;     // ... or width <= relax_ncol2 and zero fraction < relax_frac2
    double relax_frac2 = 0.25;
    double relax_frac3 = 0.10; // ... or zero fraction < relax_frac3 (any width)
    int dense_row_factor = 10; // AMD: rows with degree > factor*sqrt(N) are ordered last
};

// Lower-triangular pattern (CSC, diagonal included, row indices sorted) of a symmetric matrix.
struct SymPattern {
    int32_t N = 0;
    std::vector<int64_t> colptr;
    std::vector<int32_t> rowidx;
};

struct Symbolic {
    int system = 1;            // 1 = K1 normal equations, 2 = K2 augmented
    int64_t m = 0, n = 0;      // shape of A
    int32_t N = 0;             // order of the factored matrix (m for K1, n+m for K2)

    std::vector<int32_t> perm;     // perm[new] = old   (fill-reducing ordering composed with postorder)
    std::vector<int32_t> iperm;    // iperm[old] = new
    std::vector<int32_t> parent;   // elimination tree of the permuted matrix (postordered)
    std::vector<int32_t> colcount; // nnz(L(:,j)) including the diagonal
    int64_t nnzL = 0;              // sum colcount
    double flops = 0.0;            // sum colcount^2

    int32_t nsuper = 0;
    std::vector<int32_t> sn_first;   // [nsuper+1] first column of each supernode (contiguous columns)
    std::vector<int32_t> sn_parent;  // [nsuper]   supernodal elimination tree (-1 = root)
    std::vector<int32_t> col2sn;     // [N]
    std::vector<int64_t> sn_rowptr;  // [nsuper+1] offsets into sn_rows
    std::vector<int32_t> sn_rows;    // row list of each supernode: own columns first, then sorted below rows
    std::vector<int64_t> sn_xptr;    // [nsuper+1] offsets of each (nrow x ncol, column-major) panel in Lx
    int64_t lx_size = 0;             // total doubles of panel storage
    int64_t nnzL_relaxed = 0;        // structural non-zeros after amalgamation (trapezoids)
    int32_t max_ncol = 0, max_nrow = 0;

    std::vector<int8_t> sign;        // [N] expected pivot sign in permuted order (+1; -1 for the K2 x-block)
    std::vector<int64_t> diagpos;    // [N] position in Lx of the diagonal entry of permuted column q
};

// Ordering + etree + counts + supernodes + supernodal row structure of pattern P.
// `orig_sign[v]` = expected pivot sign of original index v (nullptr = all +1).
void analyze_pattern(const SymPattern& P, const SymOptions& opt, const int8_t* orig_sign, Symbolic& S);

// pattern of lower(A*A' + I) for K1 and of lower([-I A'; A I]) for K2.  A is CSC, 0-based.
SymPattern pattern_k1(int64_t m, int64_t n, const int64_t* colptr, const int32_t* rowidx);
SymPattern pattern_k2(int64_t m, int64_t n, const int64_t* colptr, const int32_t* rowidx);

// --- pieces of the analysis exposed for tests -------------------------------------------------
std::vector<int32_t> amd_order(const SymPattern& P, int dense_row_factor);
std::vector<int32_t> etree_lower(const SymPattern& P);               // parent[] of a lower-CSC pattern
std::vector<int32_t> postorder(const std::vector<int32_t>& parent);  // post[k] = node visited k-th
std::vector<int32_t> column_counts(const SymPattern& P, const std::vector<int32_t>& parent);  // needs postordered P
SymPattern permute_pattern(const SymPattern& P, const std::vector<int32_t>& iperm);

}  // namespace tlp

/* oracle/cpu_supernodal.c -- CPU supernodal Cholesky / signed LDL' (TEST INFRASTRUCTURE + CPU baseline).
 *
 * The reference's sparse factorisation is third-party code that is NOT under /root/reference:
 * SuiteSparse CHOLMOD through Julia's SparseArrays stdlib (src/KKT/Cholmod/spd.jl:17,46 cholesky /
 * cholesky!, sqd.jl:19,53 ldlt / ldlt!; version = the running Julia's, not pinned) and
 * LDLFactorizations.jl 0.10 (ldlfact.jl:77,113).  This file restates the *published* algorithm of
 * CHOLMOD's supernodal numeric phase (Chen, Davis, Hager, Rajamanickam, ACM TOMS 35(3), 2008,
 * section 5: left-looking supernodal factorisation, each supernode updated by its descendants
 * through dense SYRK/GEMM + scatter with a relative map, then dense POTRF + TRSM), with BLAS-level
 * threading only, as CHOLMOD does.  Dense kernels come from the OpenBLAS that SciPy bundles
 * (dlopen'ed; symbols scipy_dpotrf_ ...).  Parity status: checked in tests/ against dense LAPACK
 * and SuperLU on the same matrices -- NOT against CHOLMOD itself (unavailable here), so every
 * timing from this file is labelled "CPU port (own supernodal Cholesky + OpenBLAS) -- NOT CHOLMOD".
 *
 * The quasi-definite K2 case uses the signed variant K = L S L', S = diag(+-1) (what an LDL'
 * without pivoting computes, up to scaling of the columns of L by sqrt|d_j|).
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void (*dpotrf_t)(const char*, const int*, double*, const int*, int*);
typedef void (*dtrsm_t)(const char*, const char*, const char*, const char*, const int*, const int*, const double*,
                        const double*, const int*, double*, const int*);
typedef void (*dsyrk_t)(const char*, const char*, const int*, const int*, const double*, const double*, const int*,
                        const double*, double*, const int*);
typedef void (*dgemm_t)(const char*, const char*, const int*, const int*, const int*, const double*, const double*,
                        const int*, const double*, const int*, const double*, double*, const int*);
typedef void (*dgemv_t)(const char*, const int*, const int*, const double*, const double*, const int*, const double*,
                        const int*, const double*, double*, const int*);
typedef void (*dtrsv_t)(const char*, const char*, const char*, const int*, const double*, const int*, double*,
                        const int*);
typedef void (*setthr_t)(int);

static dpotrf_t p_dpotrf;
static dtrsm_t p_dtrsm;
static dsyrk_t p_dsyrk;
static dgemm_t p_dgemm;
static dgemv_t p_dgemv;
static dtrsv_t p_dtrsv;
static setthr_t p_setthr;

int cpu_sn_init(const char* blas_path, int nthreads) {
    void* h = dlopen(blas_path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "cpu_sn_init: %s\n", dlerror()); return 1; }
    p_dpotrf = (dpotrf_t)dlsym(h, "scipy_dpotrf_");
    p_dtrsm = (dtrsm_t)dlsym(h, "scipy_dtrsm_");
    p_dsyrk = (dsyrk_t)dlsym(h, "scipy_dsyrk_");
    p_dgemm = (dgemm_t)dlsym(h, "scipy_dgemm_");
    p_dgemv = (dgemv_t)dlsym(h, "scipy_dgemv_");
    p_dtrsv = (dtrsv_t)dlsym(h, "scipy_dtrsv_");
    p_setthr = (setthr_t)dlsym(h, "scipy_openblas_set_num_threads");
    if (!p_dpotrf || !p_dtrsm || !p_dsyrk || !p_dgemm || !p_dgemv || !p_dtrsv) return 2;
    if (p_setthr && nthreads > 0) p_setthr(nthreads);
    return 0;
}

/* scatter the lower triangle (CSC, permuted numbering, rows sorted or not) into zeroed panels */
int cpu_sn_scatter(int32_t N, int32_t nsuper, const int32_t* sn_first, const int64_t* sn_rowptr,
                   const int32_t* sn_rows, const int64_t* sn_xptr, const int64_t* colptr, const int32_t* rowidx,
                   const double* val, double* Lx) {
    int32_t* map = (int32_t*)malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
    if (!map) return 1;
    memset(Lx, 0, sizeof(double) * (size_t)sn_xptr[nsuper]);
    for (int32_t s = 0; s < nsuper; ++s) {
        const int64_t rp = sn_rowptr[s];
        const int32_t nrow = (int32_t)(sn_rowptr[s + 1] - rp);
        for (int32_t i = 0; i < nrow; ++i) map[sn_rows[rp + i]] = i;
        for (int32_t j = sn_first[s]; j < sn_first[s + 1]; ++j) {
            double* col = Lx + sn_xptr[s] + (int64_t)(j - sn_first[s]) * nrow;
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p)
                if (rowidx[p] >= j) col[map[rowidx[p]]] += val[p];
        }
    }
    free(map);
    return 0;
}

/* C[i, j] = beta C[i, j] + alpha sum_k A[i, k] A[j, k]  for j < ncols, j <= i < nrows (lower trapezoid; A is nrows x k).
 * One dsyrk + one dgemm when the output is small; block columns of CB otherwise: SciPy's OpenBLAS (0.3.x, LP64) crashes in
 * dsyrk_thread_LN / dpotrf_L_parallel once n * ld of a single call passes 2^31 (reproduced at n = 50 870), so no single
 * call is allowed to address that much. */
static void syrk_gemm_lower(int nrows, int ncols, int k, double alpha, const double* A, int lda, double beta, double* C, int ldc) {
    const int CB = ((double)ncols * (double)ldc < 1.5e9 && (double)nrows * (double)lda < 1.5e9) ? ncols : 2048;
    for (int c0 = 0; c0 < ncols; c0 += CB) {
        const int cw = (ncols - c0 < CB) ? ncols - c0 : CB;
        double* Cc = C + (size_t)c0 * ldc + c0;
        p_dsyrk("L", "N", &cw, &k, &alpha, A + c0, &lda, &beta, Cc, &ldc);
        const int nr = nrows - c0 - cw;
        if (nr > 0) p_dgemm("N", "T", &nr, &cw, &k, &alpha, A + c0 + cw, &lda, A + c0, &lda, &beta, Cc + cw, &ldc);
    }
}

/* signed factorisation of an nrow x ns trapezoid (column-major, ld), K = L S L' */
static int factor_panel(double* P, int ld, int nrow, int ns, const int8_t* sg, int32_t gcol0, int32_t* bad) {
    int allpos = 1, info = 0;
    for (int j = 0; j < ns; ++j) if (sg[j] < 0) { allpos = 0; break; }
    const double one = 1.0;
    if (allpos && (double)ns * (double)ld >= 1.5e9) {
        /* see syrk_gemm_lower: a panel this large (the north-star config's root supernode has 84 770 columns) is factored by
         * the textbook blocked right-looking algorithm on top of calls that each address < 2^31 elements: dpotrf on a NB x NB
         * diagonal block, dtrsm on the rows below, block-column dsyrk / dgemm on the trailing part. */
        const int NB = 2048;
        const double mone = -1.0;
        for (int jb = 0; jb < ns; jb += NB) {
            const int nb = (ns - jb < NB) ? ns - jb : NB;
            double* D = P + (size_t)jb * ld + jb;
            p_dpotrf("L", &nb, D, &ld, &info);
            if (info != 0) { if (*bad < 0) *bad = gcol0 + jb + (info > 0 ? info - 1 : 0); return 1; }
            for (int j = 0; j < nb; ++j) if (!(D[(size_t)j * ld + j] > 0.0)) { if (*bad < 0) *bad = gcol0 + jb + j; return 1; }
            const int below = nrow - jb - nb;
            if (below > 0) {
                p_dtrsm("R", "L", "T", "N", &below, &nb, &one, D, &ld, D + nb, &ld);
                const int ntr = ns - jb - nb;        /* trailing columns of the panel */
                if (ntr > 0) syrk_gemm_lower(below, ntr, nb, mone, D + nb, ld, one, D + (size_t)nb * ld + nb, ld);
            }
        }
        return 0;
    }
    if (allpos) {
        p_dpotrf("L", &ns, P, &ld, &info);
        if (info != 0) { if (*bad < 0) *bad = gcol0 + (info > 0 ? info - 1 : 0); return 1; }
        /* dpotrf does not flag NaN pivots */
        for (int j = 0; j < ns; ++j) if (!(P[(size_t)j * ld + j] > 0.0)) { if (*bad < 0) *bad = gcol0 + j; return 1; }
        const int nr = nrow - ns;
        if (nr > 0) p_dtrsm("R", "L", "T", "N", &nr, &ns, &one, P, &ld, P + ns, &ld);
        return 0;
    }
    /* blocked right-looking signed variant */
    const int NB = 64;
    double* W = (double*)malloc(sizeof(double) * (size_t)NB * (size_t)(nrow > 0 ? nrow : 1));
    if (!W) return 2;
    for (int jb = 0; jb < ns; jb += NB) {
        const int nb = (ns - jb < NB) ? ns - jb : NB;
        double* D = P + (size_t)jb * ld + jb;
        for (int j = 0; j < nb; ++j) {              /* unblocked on the diagonal block */
            double d = D[(size_t)j * ld + j];
            const double sj = (double)sg[jb + j];
            if (!(d * sj > 0.0)) { if (*bad < 0) *bad = gcol0 + jb + j; free(W); return 1; }
            const double ljj = sqrt(d * sj), inv = 1.0 / (sj * ljj);
            D[(size_t)j * ld + j] = ljj;
            for (int i = j + 1; i < nb; ++i) D[(size_t)j * ld + i] *= inv;
            for (int k = j + 1; k < nb; ++k) {
                const double f = sj * D[(size_t)j * ld + k];
                for (int i = k; i < nb; ++i) D[(size_t)k * ld + i] -= D[(size_t)j * ld + i] * f;
            }
        }
        const int below = nrow - jb - nb;
        if (below > 0) {
            double* B = D + nb;                      /* rows below the diagonal block */
            p_dtrsm("R", "L", "T", "N", &below, &nb, &one, D, &ld, B, &ld);
            for (int j = 0; j < nb; ++j)
                if (sg[jb + j] < 0) for (int i = 0; i < below; ++i) B[(size_t)j * ld + i] = -B[(size_t)j * ld + i];
            const int ncr = ns - jb - nb;            /* remaining columns of this panel */
            if (ncr > 0) {
                /* W = B[0:ncr, :] * S  (the rows that are columns of the trailing part) */
                for (int j = 0; j < nb; ++j) {
                    const double sj = (double)sg[jb + j];
                    for (int i = 0; i < ncr; ++i) W[(size_t)j * ncr + i] = B[(size_t)j * ld + i] * sj;
                }
                const double mone = -1.0;
                /* trailing(below x ncr) -= B(below x nb) * W(ncr x nb)'  (upper part of the square is unused) */
                p_dgemm("N", "T", &below, &ncr, &nb, &mone, B, &ld, W, &ncr, &one, D + (size_t)nb * ld + nb, &ld);
            }
        }
    }
    free(W);
    return 0;
}

/* left-looking supernodal factorisation.  Lx holds the assembled panels on entry, L on exit.
 * Returns 0 ok, 1 bad pivot (*bad = permuted column), 2 out of memory. */
int cpu_sn_factor(int32_t N, int32_t nsuper, const int32_t* sn_first, const int64_t* sn_rowptr,
                  const int32_t* sn_rows, const int64_t* sn_xptr, const int32_t* col2sn, const int8_t* sign,
                  double* Lx, int32_t* bad) {
    *bad = -1;
    int32_t* map = (int32_t*)malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
    int32_t* head = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nsuper > 0 ? nsuper : 1));
    int32_t* next = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nsuper > 0 ? nsuper : 1));
    int32_t* kpos = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nsuper > 0 ? nsuper : 1));
    size_t csize = 0, wsize = 0;
    double *C = NULL, *W = NULL;
    int rc = 0;
    if (!map || !head || !next || !kpos) { rc = 2; goto done; }
    for (int32_t s = 0; s < nsuper; ++s) head[s] = -1;
    for (int32_t t = 0; t < nsuper; ++t) {
        const int32_t ft = sn_first[t], lt = sn_first[t + 1], nst = lt - ft;
        const int64_t rpt = sn_rowptr[t];
        const int32_t nrowt = (int32_t)(sn_rowptr[t + 1] - rpt);
        double* Pt = Lx + sn_xptr[t];
        for (int32_t i = 0; i < nrowt; ++i) map[sn_rows[rpt + i]] = i;
        int32_t d = head[t];
        while (d >= 0) {
            const int32_t dnext = next[d];
            const int32_t fd = sn_first[d], nsd = sn_first[d + 1] - fd;
            const int64_t rpd = sn_rowptr[d];
            const int32_t nrowd = (int32_t)(sn_rowptr[d + 1] - rpd);
            const int32_t* rd = sn_rows + rpd;
            const int32_t k1 = kpos[d];
            int32_t k2 = k1;
            while (k2 < nrowd && rd[k2] < lt) ++k2;
            const int nd1 = k2 - k1, nd = nrowd - k1;
            const double* Ld = Lx + sn_xptr[d] + k1;     /* rows k1.. of d's panel, ld = nrowd */
            if ((size_t)nd * nd1 > csize) { free(C); csize = (size_t)nd * nd1; C = (double*)malloc(sizeof(double) * csize); if (!C) { rc = 2; goto done; } }
            int mixed = 0;
            for (int j = 0; j < nsd; ++j) if (sign[fd + j] < 0) { mixed = 1; break; }
            const double one = 1.0, zero = 0.0;
            if (!mixed) {
                syrk_gemm_lower(nd, nd1, nsd, one, Ld, nrowd, zero, C, nd);
            } else {
                if ((size_t)nd1 * nsd > wsize) { free(W); wsize = (size_t)nd1 * nsd; W = (double*)malloc(sizeof(double) * wsize); if (!W) { rc = 2; goto done; } }
                for (int j = 0; j < nsd; ++j) {
                    const double sj = (double)sign[fd + j];
                    for (int i = 0; i < nd1; ++i) W[(size_t)j * nd1 + i] = Ld[(size_t)j * nrowd + i] * sj;
                }
                p_dgemm("N", "T", &nd, &nd1, &nsd, &one, Ld, &nrowd, W, &nd1, &zero, C, &nd);
            }
            for (int j = 0; j < nd1; ++j) {
                double* tc = Pt + (size_t)(rd[k1 + j] - ft) * nrowt;
                const double* cc = C + (size_t)j * nd;
                for (int i = j; i < nd; ++i) tc[map[rd[k1 + i]]] -= cc[i];
            }
            kpos[d] = k2;
            if (k2 < nrowd) { const int32_t nt = col2sn[rd[k2]]; next[d] = head[nt]; head[nt] = d; }
            d = dnext;
        }
        rc = factor_panel(Pt, nrowt, nrowt, nst, sign + ft, ft, bad);
        if (rc) goto done;
        if (nrowt > nst) {
            kpos[t] = nst;
            const int32_t nt = col2sn[sn_rows[rpt + nst]];
            next[t] = head[nt];
            head[nt] = t;
        }
    }
done:
    free(map); free(head); free(next); free(kpos); free(C); free(W);
    return rc;
}

/* x := (L S L')^{-1} x in permuted numbering; cpu_sn_solve_part: mode 0 = both sweeps, 1 = x := L^{-1} x only,
 * 2 = x := L^{-T} S x only (the factorised Schur-complement form of the dense-column tests needs the halves) */
int cpu_sn_solve_part(int32_t N, int32_t nsuper, const int32_t* sn_first, const int64_t* sn_rowptr,
                      const int32_t* sn_rows, const int64_t* sn_xptr, const int8_t* sign, const double* Lx, double* x, int mode);
int cpu_sn_solve(int32_t N, int32_t nsuper, const int32_t* sn_first, const int64_t* sn_rowptr,
                 const int32_t* sn_rows, const int64_t* sn_xptr, const int8_t* sign, const double* Lx, double* x) {
    return cpu_sn_solve_part(N, nsuper, sn_first, sn_rowptr, sn_rows, sn_xptr, sign, Lx, x, 0);
}
int cpu_sn_solve_part(int32_t N, int32_t nsuper, const int32_t* sn_first, const int64_t* sn_rowptr,
                      const int32_t* sn_rows, const int64_t* sn_xptr, const int8_t* sign, const double* Lx, double* x, int mode) {
    double* w = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
    if (!w) return 2;
    const int ione = 1;
    const double one = 1.0, mone = -1.0, zero = 0.0;
    for (int32_t s = 0; s < nsuper && mode != 2; ++s) {
        const int32_t f = sn_first[s];
        const int ns = sn_first[s + 1] - f;
        const int64_t rp = sn_rowptr[s];
        const int nrow = (int)(sn_rowptr[s + 1] - rp), nr = nrow - ns;
        const double* P = Lx + sn_xptr[s];
        p_dtrsv("L", "N", "N", &ns, P, &nrow, x + f, &ione);
        if (nr > 0) {
            p_dgemv("N", &nr, &ns, &one, P + ns, &nrow, x + f, &ione, &zero, w, &ione);
            for (int i = 0; i < nr; ++i) x[sn_rows[rp + ns + i]] -= w[i];
        }
    }
    if (mode == 1) { free(w); return 0; }
    for (int32_t q = 0; q < N; ++q) if (sign[q] < 0) x[q] = -x[q];
    for (int32_t s = nsuper - 1; s >= 0; --s) {
        const int32_t f = sn_first[s];
        const int ns = sn_first[s + 1] - f;
        const int64_t rp = sn_rowptr[s];
        const int nrow = (int)(sn_rowptr[s + 1] - rp), nr = nrow - ns;
        const double* P = Lx + sn_xptr[s];
        if (nr > 0) {
            for (int i = 0; i < nr; ++i) w[i] = x[sn_rows[rp + ns + i]];
            p_dgemv("T", &nr, &ns, &mone, P + ns, &nrow, w, &ione, &one, x + f, &ione);
        }
        p_dtrsv("L", "T", "N", &ns, P, &nrow, x + f, &ione);
    }
    free(w);
    return 0;
}

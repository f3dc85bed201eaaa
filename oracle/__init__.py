"""oracle/ -- CPU restatement of Tulip.jl's KKT hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from this package.  The product
(``tulip.jl_b200/``) never imports it and has no CPU fallback.

Parity status: **partially pinned**.  The reference (pure Julia, v0.9.8) cannot run
in this container (no Julia runtime) and the arithmetic of its sparse path lives
in un-vendored third-party code (SuiteSparse CHOLMOD through Julia's stdlib,
LDLFactorizations.jl 0.10; no Manifest => versions not pinned).  What *is* pinned,
and checked in ``tests/test_oracle.py``:

* the KKT conformance vector of ``src/KKT/Test/test.jl:9-46``
  (A=[1 0 1 0; 0 1 0 1], theta=regP=regD=1, xi_p=xi_d=1 => dx=0, dy=(1,1),
  residuals <= sqrt(eps));
* the end-to-end answers the reference's own tests assert for its four example
  LPs (``examples/optimal.jl:37-62`` obj 3/2, x=(1/2,1/2), y=(3/2,-1/2);
  ``examples/freevars.jl``; ``examples/infeasible.jl`` -> Trm_PrimalInfeasible;
  ``examples/unbounded.jl`` -> Trm_DualInfeasible).

Factor values, orderings and fill are NOT pinned by any reference test
(SURVEY.md section 8c), so "parity" for those means: same linear-system
solution (1e-8 rel) and same IPM trajectory, not the same factor.
"""

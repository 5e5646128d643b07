/*
 * oracle/ldl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Symmetric indefinite LDL^T with Bunch-Kaufman partial pivoting (the
 * algorithm of LAPACK's xSYTF2, restated) on dense storage, with the
 * rank-1 / rank-2 trailing updates restricted to the nonzero rows of the
 * pivot column(s) so that a banded / stage-ordered KKT matrix costs
 * O(n * band^2) instead of O(n^3).  Returns the inertia, which is what the
 * interior-point method needs (Ipopt gets it from MUMPS; see
 * assets/document/ipopt_install/ipopt_x86_install_tutorial.md:15-22).
 */
#ifndef ORACLE_LDL_H
#define ORACLE_LDL_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldl_fact {
    int n;
    double *A;     /* n*n row-major, full symmetric, overwritten */
    int *kp;       /* interchange partner per pivot step (position k or k+1) */
    int *kind;     /* 1: 1x1 pivot at k; 2: 2x2 pivot starting at k; 0: second column of a 2x2 */
    double *D;     /* 3 per position: d11, d21, d22 (2x2) or d11 (1x1) */
    int *Lp;       /* n+1 column pointers */
    int *Li;       /* row indices */
    double *Lx;    /* values */
    int Lcap;
    int npos, nneg, nzero;
    int *idx;      /* scratch n */
    double *c1, *c2; /* scratch n */
} ldl_fact;

ldl_fact *ldl_alloc(int n);
void ldl_free(ldl_fact *F);
/* Factor F->A in place (caller filled all n*n entries, both triangles).
 * dense_updates != 0 ignores sparsity (plain O(n^3) Bunch-Kaufman). */
int ldl_factor(ldl_fact *F, int dense_updates);
/* Solve A x = b in place. */
void ldl_solve(const ldl_fact *F, double *b);

/* Reverse Cuthill-McKee ordering of a symmetric pattern given as COO pairs
 * (any triangle, duplicates allowed).  order[pos] = original index. */
void rcm_order(int n, int nnz, const int *ri, const int *ci, int *order);

#ifdef __cplusplus
}
#endif
#endif

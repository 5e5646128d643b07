/*
 * oracle/ldl.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See ldl.h.
 */
#include "ldl.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

ldl_fact *ldl_alloc(int n)
{
    ldl_fact *F = (ldl_fact *)calloc(1, sizeof(ldl_fact));
    F->n = n;
    F->A = (double *)malloc(sizeof(double) * (size_t)n * n);
    F->kp = (int *)malloc(sizeof(int) * n);
    F->kind = (int *)malloc(sizeof(int) * n);
    F->D = (double *)malloc(sizeof(double) * 3 * n);
    F->Lp = (int *)malloc(sizeof(int) * (n + 1));
    F->Lcap = 64 * n + 64;
    F->Li = (int *)malloc(sizeof(int) * F->Lcap);
    F->Lx = (double *)malloc(sizeof(double) * F->Lcap);
    F->idx = (int *)malloc(sizeof(int) * n);
    F->c1 = (double *)malloc(sizeof(double) * n);
    F->c2 = (double *)malloc(sizeof(double) * n);
    return F;
}

void ldl_free(ldl_fact *F)
{
    if (!F) return;
    free(F->A); free(F->kp); free(F->kind); free(F->D); free(F->Lp);
    free(F->Li); free(F->Lx); free(F->idx); free(F->c1); free(F->c2);
    free(F);
}

static void l_reserve(ldl_fact *F, int need)
{
    if (need <= F->Lcap) return;
    while (F->Lcap < need) F->Lcap *= 2;
    F->Li = (int *)realloc(F->Li, sizeof(int) * F->Lcap);
    F->Lx = (double *)realloc(F->Lx, sizeof(double) * F->Lcap);
}

/* symmetric interchange of positions p < q inside the trailing block [k0, n) */
static void sym_swap(double *A, int n, int k0, int p, int q)
{
    if (p == q) return;
    double *rp = A + (size_t)p * n, *rq = A + (size_t)q * n;
    for (int j = k0; j < n; j++) { double t = rp[j]; rp[j] = rq[j]; rq[j] = t; }
    for (int i = k0; i < n; i++) {
        double *r = A + (size_t)i * n;
        double t = r[p]; r[p] = r[q]; r[q] = t;
    }
}

int ldl_factor(ldl_fact *F, int dense_updates)
{
    const int n = F->n;
    double *A = F->A;
    const double alpha = (1.0 + sqrt(17.0)) / 8.0;
    int lnz = 0;
    F->npos = F->nneg = F->nzero = 0;
    int k = 0;
    while (k < n) {
        int kstep = 1;
        double *rk = A + (size_t)k * n;
        double absakk = fabs(rk[k]);
        int imax = -1;
        double colmax = 0.0;
        for (int i = k + 1; i < n; i++) {
            double v = fabs(rk[i]); /* symmetric: row k == column k */
            if (v > colmax) { colmax = v; imax = i; }
        }
        int kp = k;
        if (!(fmax(absakk, colmax) > 0.0)) {
            /* zero column: singular pivot */
            kp = k;
        } else if (absakk >= alpha * colmax) {
            kp = k;
        } else {
            const double *ri = A + (size_t)imax * n;
            double rowmax = 0.0;
            for (int j = k; j < n; j++) {
                if (j == imax) continue;
                double v = fabs(ri[j]);
                if (v > rowmax) rowmax = v;
            }
            if (absakk >= alpha * colmax * (colmax / rowmax)) {
                kp = k;
            } else if (fabs(ri[imax]) >= alpha * rowmax) {
                kp = imax;
            } else {
                kp = imax;
                kstep = 2;
            }
        }
        const int kk = k + kstep - 1;
        if (kp != kk) sym_swap(A, n, k, kk, kp);
        F->kp[k] = kp;

        if (kstep == 1) {
            const double d = rk[k];
            F->kind[k] = 1;
            F->D[3 * k] = d;
            F->Lp[k] = lnz;
            if (d > 0.0) F->npos++; else if (d < 0.0) F->nneg++; else F->nzero++;
            if (d != 0.0) {
                int cnt = 0;
                for (int i = k + 1; i < n; i++) {
                    double v = rk[i];
                    if (dense_updates || v != 0.0) { F->idx[cnt] = i; F->c1[cnt] = v; cnt++; }
                }
                l_reserve(F, lnz + cnt);
                const double r = 1.0 / d;
                for (int a = 0; a < cnt; a++) {
                    const int i = F->idx[a];
                    const double li = F->c1[a] * r;
                    double *rowi = A + (size_t)i * n;
                    for (int b = 0; b < cnt; b++) rowi[F->idx[b]] -= li * F->c1[b];
                    F->Li[lnz] = i; F->Lx[lnz] = li; lnz++;
                }
            }
        } else {
            double *rk1 = A + (size_t)(k + 1) * n;
            const double d11 = rk[k], d21 = rk1[k], d22 = rk1[k + 1];
            F->kind[k] = 2; F->kind[k + 1] = 0;
            F->kp[k + 1] = kp;
            F->D[3 * k] = d11; F->D[3 * k + 1] = d21; F->D[3 * k + 2] = d22;
            const double det = d11 * d22 - d21 * d21;
            if (det < 0.0) { F->npos++; F->nneg++; }
            else if (det > 0.0) { if (d11 + d22 > 0.0) F->npos += 2; else F->nneg += 2; }
            else { F->nzero++; if (d11 + d22 > 0.0) F->npos++; else if (d11 + d22 < 0.0) F->nneg++; else F->nzero++; }
            int cnt = 0;
            for (int i = k + 2; i < n; i++) {
                double v1 = rk[i], v2 = rk1[i];
                if (dense_updates || v1 != 0.0 || v2 != 0.0) {
                    F->idx[cnt] = i; F->c1[cnt] = v1; F->c2[cnt] = v2; cnt++;
                }
            }
            l_reserve(F, lnz + 2 * cnt);
            F->Lp[k] = lnz;
            const double idet = (det != 0.0) ? 1.0 / det : 0.0;
            /* column k of L */
            int l1 = lnz, l2 = lnz + cnt;
            for (int a = 0; a < cnt; a++) {
                const int i = F->idx[a];
                const double w1 = (d22 * F->c1[a] - d21 * F->c2[a]) * idet;
                const double w2 = (d11 * F->c2[a] - d21 * F->c1[a]) * idet;
                double *rowi = A + (size_t)i * n;
                for (int b = 0; b < cnt; b++)
                    rowi[F->idx[b]] -= w1 * F->c1[b] + w2 * F->c2[b];
                F->Li[l1 + a] = i; F->Lx[l1 + a] = w1;
                F->Li[l2 + a] = i; F->Lx[l2 + a] = w2;
            }
            F->Lp[k + 1] = l2;
            lnz += 2 * cnt;
        }
        k += kstep;
    }
    F->Lp[n] = lnz;
    return 0;
}

void ldl_solve(const ldl_fact *F, double *b)
{
    const int n = F->n;
    int k = 0;
    while (k < n) {
        if (F->kind[k] == 1) {
            int kp = F->kp[k];
            if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
            const double bk = b[k];
            for (int p = F->Lp[k]; p < F->Lp[k + 1]; p++) b[F->Li[p]] -= F->Lx[p] * bk;
            const double d = F->D[3 * k];
            b[k] = (d != 0.0) ? bk / d : 0.0;
            k += 1;
        } else {
            int kp = F->kp[k];
            if (kp != k + 1) { double t = b[k + 1]; b[k + 1] = b[kp]; b[kp] = t; }
            const double b1 = b[k], b2 = b[k + 1];
            for (int p = F->Lp[k]; p < F->Lp[k + 1]; p++) b[F->Li[p]] -= F->Lx[p] * b1;
            for (int p = F->Lp[k + 1]; p < F->Lp[k + 2]; p++) b[F->Li[p]] -= F->Lx[p] * b2;
            const double d11 = F->D[3 * k], d21 = F->D[3 * k + 1], d22 = F->D[3 * k + 2];
            const double det = d11 * d22 - d21 * d21;
            const double idet = (det != 0.0) ? 1.0 / det : 0.0;
            b[k] = (d22 * b1 - d21 * b2) * idet;
            b[k + 1] = (d11 * b2 - d21 * b1) * idet;
            k += 2;
        }
    }
    /* backward: find pivot starts from the end */
    k = n - 1;
    while (k >= 0) {
        if (F->kind[k] == 1) {
            double s = b[k];
            for (int p = F->Lp[k]; p < F->Lp[k + 1]; p++) s -= F->Lx[p] * b[F->Li[p]];
            b[k] = s;
            int kp = F->kp[k];
            if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
            k -= 1;
        } else {
            /* kind[k]==0: second column of the 2x2 that starts at k-1 */
            const int k0 = k - 1;
            double s1 = b[k0], s2 = b[k];
            for (int p = F->Lp[k0]; p < F->Lp[k0 + 1]; p++) s1 -= F->Lx[p] * b[F->Li[p]];
            for (int p = F->Lp[k0 + 1]; p < F->Lp[k0 + 2]; p++) s2 -= F->Lx[p] * b[F->Li[p]];
            b[k0] = s1; b[k] = s2;
            int kp = F->kp[k0];
            if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
            k -= 2;
        }
    }
}

/* ---- reverse Cuthill-McKee ---- */
static int cmp_deg_ctx(const void *a, const void *b, void *ctx)
{
    const int *deg = (const int *)ctx;
    int ia = *(const int *)a, ib = *(const int *)b;
    if (deg[ia] != deg[ib]) return deg[ia] < deg[ib] ? -1 : 1;
    return ia < ib ? -1 : (ia > ib);
}

static void sort_by_deg(int *v, int cnt, const int *deg)
{
    /* insertion sort: lists are tiny */
    for (int i = 1; i < cnt; i++) {
        int x = v[i], j = i - 1;
        while (j >= 0 && cmp_deg_ctx(&v[j], &x, (void *)deg) > 0) { v[j + 1] = v[j]; j--; }
        v[j + 1] = x;
    }
}

static int bfs_levels(int n, const int *ap, const int *ai, int start, int *mark, int stamp,
                      int *queue, int *last_level_min_deg_node, const int *deg)
{
    /* returns eccentricity; writes a min-degree node of the last level */
    int head = 0, tail = 0, depth = 0;
    queue[tail++] = start; mark[start] = stamp;
    int level_end = 1, last_begin = 0;
    while (head < tail) {
        if (head == level_end) { depth++; last_begin = head; level_end = tail; }
        int u = queue[head++];
        for (int p = ap[u]; p < ap[u + 1]; p++) {
            int v = ai[p];
            if (mark[v] != stamp) { mark[v] = stamp; queue[tail++] = v; }
        }
    }
    int best = queue[last_begin];
    for (int i = last_begin; i < tail; i++)
        if (deg[queue[i]] < deg[best]) best = queue[i];
    *last_level_min_deg_node = best;
    (void)n;
    return depth;
}

void rcm_order(int n, int nnz, const int *ri, const int *ci, int *order)
{
    int *cnt = (int *)calloc(n + 1, sizeof(int));
    for (int k = 0; k < nnz; k++)
        if (ri[k] != ci[k]) { cnt[ri[k] + 1]++; cnt[ci[k] + 1]++; }
    int *ap = (int *)malloc(sizeof(int) * (n + 1));
    ap[0] = 0;
    for (int i = 0; i < n; i++) ap[i + 1] = ap[i] + cnt[i + 1];
    int *ai = (int *)malloc(sizeof(int) * (ap[n] > 0 ? ap[n] : 1));
    int *fill = (int *)calloc(n, sizeof(int));
    for (int k = 0; k < nnz; k++)
        if (ri[k] != ci[k]) {
            ai[ap[ri[k]] + fill[ri[k]]++] = ci[k];
            ai[ap[ci[k]] + fill[ci[k]]++] = ri[k];
        }
    int *deg = (int *)malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) deg[i] = ap[i + 1] - ap[i];
    int *mark = (int *)calloc(n, sizeof(int));
    int *queue = (int *)malloc(sizeof(int) * n);
    int *visited = (int *)calloc(n, sizeof(int));
    int stamp = 0, nord = 0;
    for (int s = 0; s < n; s++) {
        if (visited[s]) continue;
        /* pseudo-peripheral start inside this component */
        int start = s, ecc = -1;
        for (int it = 0; it < 8; it++) {
            int cand;
            int e = bfs_levels(n, ap, ai, start, mark, ++stamp, queue, &cand, deg);
            if (e <= ecc) break;
            ecc = e; start = cand;
        }
        /* Cuthill-McKee BFS, neighbours by increasing degree */
        int head = nord;
        order[nord++] = start; visited[start] = 1;
        while (head < nord) {
            int u = order[head++];
            int b0 = nord;
            for (int p = ap[u]; p < ap[u + 1]; p++) {
                int v = ai[p];
                if (!visited[v]) { visited[v] = 1; order[nord++] = v; }
            }
            sort_by_deg(order + b0, nord - b0, deg);
        }
    }
    /* reverse */
    for (int i = 0, j = n - 1; i < j; i++, j--) { int t = order[i]; order[i] = order[j]; order[j] = t; }
    free(cnt); free(ap); free(ai); free(fill); free(deg); free(mark); free(queue); free(visited);
}

/*
 * icp_oracle.c - CPU restatement (plain C, FP64) of the icp-proposal hot path.
 *
 * TEST INFRASTRUCTURE ONLY - see icp_oracle.h. PARITY UNPINNED (no reference golden vectors exist,
 * the Scala reference cannot run here). Every function cites the reference file:line it follows;
 * paths are relative to /root/reference/src/main/scala. Scalismo 0.90.0 / Breeze internals are
 * restated from SURVEY.md Appendix A ([S-recall]).
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -ffp-contract=off).
 */
#include "icp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define LOG_2PI 1.8378770664093454835606594728112

/* ============================================================================================ */
/* small vector helpers                                                                         */
/* ============================================================================================ */
static inline void v_sub(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline double v_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void v_cross(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double v_norm(const double *a) { return sqrt(v_dot(a, a)); }
static inline void v_normalize(double *a)
{
    double n = v_norm(a); /* EuclideanVector.normalize: v / norm (NaN for the zero vector) */
    a[0] /= n; a[1] /= n; a[2] /= n;
}

/* ============================================================================================ */
/* exact point-triangle closest point (Appendix A10: classified vertex / edge / interior)       */
/* ============================================================================================ */
double orc_point_triangle_d2(const double *p, const double *a, const double *b, const double *c,
                             double *cp, int32_t *feat)
{
    double ab[3], ac[3], ap[3], bp[3], cpv[3], r[3], d[3];
    int f;
    v_sub(b, a, ab); v_sub(c, a, ac); v_sub(p, a, ap);
    double d1 = v_dot(ab, ap), d2 = v_dot(ac, ap);
    if (d1 <= 0.0 && d2 <= 0.0) { memcpy(r, a, 24); f = 0; goto done; }
    v_sub(p, b, bp);
    double d3 = v_dot(ab, bp), d4 = v_dot(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) { memcpy(r, b, 24); f = 0; goto done; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
        double v = d1 / (d1 - d3);
        for (int k = 0; k < 3; k++) r[k] = a[k] + v * ab[k];
        f = 1; goto done;
    }
    v_sub(p, c, cpv);
    double d5 = v_dot(ab, cpv), d6 = v_dot(ac, cpv);
    if (d6 >= 0.0 && d5 <= d6) { memcpy(r, c, 24); f = 0; goto done; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
        double w = d2 / (d2 - d6);
        for (int k = 0; k < 3; k++) r[k] = a[k] + w * ac[k];
        f = 1; goto done;
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int k = 0; k < 3; k++) r[k] = b[k] + w * (c[k] - b[k]);
        f = 1; goto done;
    }
    {
        double denom = 1.0 / (va + vb + vc);
        double v = vb * denom, w = vc * denom;
        for (int k = 0; k < 3; k++) r[k] = a[k] + ab[k] * v + ac[k] * w;
        f = 2;
    }
done:
    v_sub(p, r, d);
    if (cp) memcpy(cp, r, 24);
    if (feat) *feat = f;
    return v_dot(d, d);
}

void orc_closest_point_brute(int nv, const double *verts, int nt, const int32_t *tris, int nq,
                             const double *q, int32_t *tri, int32_t *feat, double *cp, double *d2)
{
    (void)nv;
    for (int i = 0; i < nq; i++) {
        double best = INFINITY, bcp[3] = {0, 0, 0};
        int32_t bt = -1, bf = -1;
        for (int t = 0; t < nt; t++) {
            double c[3]; int32_t f;
            double d = orc_point_triangle_d2(q + 3 * i, verts + 3 * tris[3 * t], verts + 3 * tris[3 * t + 1],
                                             verts + 3 * tris[3 * t + 2], c, &f);
            if (d < best) { best = d; bt = t; bf = f; memcpy(bcp, c, 24); }
        }
        if (tri) tri[i] = bt;
        if (feat) feat[i] = bf;
        if (cp) memcpy(cp + 3 * i, bcp, 24);
        if (d2) d2[i] = best;
    }
}

void orc_closest_vertex_brute(int nv, const double *verts, int nq, const double *q, int32_t *id, double *d2)
{
    for (int i = 0; i < nq; i++) {
        double best = INFINITY; int32_t b = -1;
        for (int v = 0; v < nv; v++) {
            double d[3]; v_sub(q + 3 * i, verts + 3 * v, d);
            double dd = v_dot(d, d);
            if (dd < best) { best = dd; b = v; }
        }
        if (id) id[i] = b;
        if (d2) d2[i] = best;
    }
}

/* ============================================================================================ */
/* mesh with helper structures                                                                   */
/* ============================================================================================ */
typedef struct { double c[3], r; int left, right, tri; } sph_node;
typedef struct { int dim, pt, left, right; } kd_node;

struct orc_mesh {
    int nv, nt;
    double *v; int32_t *t;
    sph_node *sn; int n_sn;
    kd_node *kd; int n_kd, kd_root;
    uint8_t *boundary;
    int *adj_off, *adj; /* vertex -> triangles CSR, ascending triangle id */
};

static const double *g_sort_v; static int g_sort_dim;
static const double *g_sort_cent;
static int cmp_cent(const void *a, const void *b)
{
    double x = g_sort_cent[3 * (*(const int *)a) + g_sort_dim], y = g_sort_cent[3 * (*(const int *)b) + g_sort_dim];
    if (x < y) return -1; if (x > y) return 1;
    return (*(const int *)a) - (*(const int *)b);
}
static int cmp_vert(const void *a, const void *b)
{
    double x = g_sort_v[3 * (*(const int *)a) + g_sort_dim], y = g_sort_v[3 * (*(const int *)b) + g_sort_dim];
    if (x < y) return -1; if (x > y) return 1;
    return (*(const int *)a) - (*(const int *)b);
}

static int sph_build(orc_mesh *m, int *idx, int n, const double *cent)
{
    int me = m->n_sn++;
    sph_node *nd = &m->sn[me];
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            const double *p = m->v + 3 * m->t[3 * idx[i] + k];
            for (int d = 0; d < 3; d++) { if (p[d] < lo[d]) lo[d] = p[d]; if (p[d] > hi[d]) hi[d] = p[d]; }
        }
    for (int d = 0; d < 3; d++) nd->c[d] = 0.5 * (lo[d] + hi[d]);
    double r2 = 0;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            double dv[3]; v_sub(m->v + 3 * m->t[3 * idx[i] + k], nd->c, dv);
            double dd = v_dot(dv, dv); if (dd > r2) r2 = dd;
        }
    nd->r = sqrt(r2) * (1.0 + 1e-12);
    if (n == 1) { nd->tri = idx[0]; nd->left = nd->right = -1; return me; }
    nd->tri = -1;
    /* split along the widest centroid extent */
    double clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) { double x = cent[3 * idx[i] + d]; if (x < clo[d]) clo[d] = x; if (x > chi[d]) chi[d] = x; }
    int dim = 0; if (chi[1] - clo[1] > chi[dim] - clo[dim]) dim = 1; if (chi[2] - clo[2] > chi[dim] - clo[dim]) dim = 2;
    g_sort_cent = cent; g_sort_dim = dim;
    qsort(idx, n, sizeof(int), cmp_cent);
    int h = n / 2;
    int l = sph_build(m, idx, h, cent);
    int r = sph_build(m, idx + h, n - h, cent);
    m->sn[me].left = l; m->sn[me].right = r;
    return me;
}

static int kd_build(orc_mesh *m, int *idx, int n, int depth)
{
    if (n <= 0) return -1;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) { double x = m->v[3 * idx[i] + d]; if (x < lo[d]) lo[d] = x; if (x > hi[d]) hi[d] = x; }
    int dim = 0; if (hi[1] - lo[1] > hi[dim] - lo[dim]) dim = 1; if (hi[2] - lo[2] > hi[dim] - lo[dim]) dim = 2;
    (void)depth;
    g_sort_v = m->v; g_sort_dim = dim;
    qsort(idx, n, sizeof(int), cmp_vert);
    int h = n / 2;
    int me = m->n_kd++;
    m->kd[me].dim = dim; m->kd[me].pt = idx[h];
    int l = kd_build(m, idx, h, depth + 1);
    int r = kd_build(m, idx + h + 1, n - h - 1, depth + 1);
    m->kd[me].left = l; m->kd[me].right = r;
    return me;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

orc_mesh *orc_mesh_create(int nv, const double *verts, int nt, const int32_t *tris)
{
    orc_mesh *m = (orc_mesh *)calloc(1, sizeof(orc_mesh));
    m->nv = nv; m->nt = nt;
    m->v = (double *)malloc(sizeof(double) * 3 * nv); memcpy(m->v, verts, sizeof(double) * 3 * nv);
    m->t = (int32_t *)malloc(sizeof(int32_t) * 3 * nt); memcpy(m->t, tris, sizeof(int32_t) * 3 * nt);
    /* bounding-sphere tree over triangles (Appendix A10) */
    double *cent = (double *)malloc(sizeof(double) * 3 * nt);
    int *idx = (int *)malloc(sizeof(int) * (nt > nv ? nt : nv));
    for (int t = 0; t < nt; t++) {
        idx[t] = t;
        for (int d = 0; d < 3; d++)
            cent[3 * t + d] = (verts[3 * tris[3 * t] + d] + verts[3 * tris[3 * t + 1] + d] + verts[3 * tris[3 * t + 2] + d]) / 3.0;
    }
    m->sn = (sph_node *)malloc(sizeof(sph_node) * (2 * nt + 1)); m->n_sn = 0;
    if (nt > 0) sph_build(m, idx, nt, cent);
    free(cent);
    /* KD-tree over vertices (Appendix A11) */
    for (int v = 0; v < nv; v++) idx[v] = v;
    m->kd = (kd_node *)malloc(sizeof(kd_node) * (nv + 1)); m->n_kd = 0;
    m->kd_root = kd_build(m, idx, nv, 0);
    free(idx);
    /* boundary table (Appendix A12): vertex on an edge with exactly one incident triangle */
    uint64_t *ek = (uint64_t *)malloc(sizeof(uint64_t) * 3 * nt);
    for (int t = 0; t < nt; t++)
        for (int k = 0; k < 3; k++) {
            uint64_t a = (uint64_t)tris[3 * t + k], b = (uint64_t)tris[3 * t + (k + 1) % 3];
            ek[3 * t + k] = a < b ? (a << 32) | b : (b << 32) | a;
        }
    qsort(ek, 3 * nt, sizeof(uint64_t), cmp_u64);
    m->boundary = (uint8_t *)calloc(nv > 0 ? nv : 1, 1);
    for (int i = 0; i < 3 * nt;) {
        int j = i; while (j < 3 * nt && ek[j] == ek[i]) j++;
        if (j - i == 1) { m->boundary[ek[i] >> 32] = 1; m->boundary[ek[i] & 0xffffffffu] = 1; }
        i = j;
    }
    free(ek);
    /* vertex -> adjacent triangles */
    m->adj_off = (int *)calloc(nv + 1, sizeof(int));
    for (int t = 0; t < 3 * nt; t++) m->adj_off[tris[t] + 1]++;
    for (int v = 0; v < nv; v++) m->adj_off[v + 1] += m->adj_off[v];
    m->adj = (int *)malloc(sizeof(int) * (3 * nt > 0 ? 3 * nt : 1));
    int *fill = (int *)calloc(nv > 0 ? nv : 1, sizeof(int));
    for (int t = 0; t < nt; t++)
        for (int k = 0; k < 3; k++) { int v = tris[3 * t + k]; m->adj[m->adj_off[v] + fill[v]++] = t; }
    free(fill);
    return m;
}

void orc_mesh_free(orc_mesh *m)
{
    if (!m) return;
    free(m->v); free(m->t); free(m->sn); free(m->kd); free(m->boundary); free(m->adj_off); free(m->adj); free(m);
}

typedef struct { double best; int tri, feat; double cp[3]; } cp_state;

static void sph_query(const orc_mesh *m, int node, const double *q, cp_state *s)
{
    const sph_node *nd = &m->sn[node];
    if (nd->tri >= 0) {
        double c[3]; int32_t f; int t = nd->tri;
        double d = orc_point_triangle_d2(q, m->v + 3 * m->t[3 * t], m->v + 3 * m->t[3 * t + 1], m->v + 3 * m->t[3 * t + 2], c, &f);
        if (d < s->best || (d == s->best && t < s->tri)) { s->best = d; s->tri = t; s->feat = f; memcpy(s->cp, c, 24); }
        return;
    }
    double lb[2]; int ch[2] = {nd->left, nd->right};
    for (int k = 0; k < 2; k++) {
        double dv[3]; v_sub(q, m->sn[ch[k]].c, dv);
        double dist = sqrt(v_dot(dv, dv)) - m->sn[ch[k]].r;
        lb[k] = dist > 0 ? dist * dist * (1.0 - 1e-12) : 0.0;
    }
    int first = lb[0] <= lb[1] ? 0 : 1;
    if (lb[first] <= s->best) sph_query(m, ch[first], q, s);
    if (lb[1 - first] <= s->best) sph_query(m, ch[1 - first], q, s);
}

void orc_mesh_closest_point(const orc_mesh *m, int nq, const double *q, int32_t *tri, int32_t *feat,
                            double *cp, double *d2)
{
    for (int i = 0; i < nq; i++) {
        cp_state s; s.best = INFINITY; s.tri = -1; s.feat = -1; s.cp[0] = s.cp[1] = s.cp[2] = 0;
        if (m->nt > 0) sph_query(m, 0, q + 3 * i, &s);
        if (tri) tri[i] = s.tri;
        if (feat) feat[i] = s.feat;
        if (cp) memcpy(cp + 3 * i, s.cp, 24);
        if (d2) d2[i] = s.best;
    }
}

static void kd_query(const orc_mesh *m, int node, const double *q, double *best, int *bid)
{
    if (node < 0) return;
    const kd_node *nd = &m->kd[node];
    double dv[3]; v_sub(q, m->v + 3 * nd->pt, dv);
    double dd = v_dot(dv, dv);
    if (dd < *best || (dd == *best && nd->pt < *bid)) { *best = dd; *bid = nd->pt; }
    double delta = q[nd->dim] - m->v[3 * nd->pt + nd->dim];
    int near = delta < 0 ? nd->left : nd->right, far = delta < 0 ? nd->right : nd->left;
    kd_query(m, near, q, best, bid);
    if (delta * delta <= *best) kd_query(m, far, q, best, bid);
}

void orc_mesh_closest_vertex(const orc_mesh *m, int nq, const double *q, int32_t *id, double *d2)
{
    for (int i = 0; i < nq; i++) {
        double best = INFINITY; int b = -1;
        kd_query(m, m->kd_root, q + 3 * i, &best, &b);
        if (id) id[i] = b;
        if (d2) d2[i] = best;
    }
}

void orc_mesh_boundary_flags(const orc_mesh *m, uint8_t *flags) { memcpy(flags, m->boundary, m->nv); }

/* Appendix A13: normalised unweighted mean of the unit cell normals of the adjacent triangles */
static void vertex_normal(const orc_mesh *m, int v, double *out)
{
    double s[3] = {0, 0, 0};
    int n = m->adj_off[v + 1] - m->adj_off[v];
    for (int k = m->adj_off[v]; k < m->adj_off[v + 1]; k++) {
        int t = m->adj[k];
        const double *p1 = m->v + 3 * m->t[3 * t], *p2 = m->v + 3 * m->t[3 * t + 1], *p3 = m->v + 3 * m->t[3 * t + 2];
        double u[3], w[3], c[3];
        v_sub(p2, p1, u); v_sub(p3, p1, w); v_cross(u, w, c); v_normalize(c);
        s[0] += c[0]; s[1] += c[1]; s[2] += c[2];
    }
    out[0] = s[0] / n; out[1] = s[1] / n; out[2] = s[2] / n;
    v_normalize(out);
}

void orc_mesh_vertex_normals(const orc_mesh *m, double *normals)
{
    for (int v = 0; v < m->nv; v++) vertex_normal(m, v, normals + 3 * v);
}

/* ============================================================================================ */
/* dense linear algebra (no LAPACK on this box): row-major helpers, symmetric eigensolver        */
/* ============================================================================================ */
/* Optional BLAS / LAPACK backend for the timed CPU baseline (bench.py --impl reference): the reference's Breeze calls
 * netlib-java, which binds a native BLAS when one is installed, so the fair CPU arm runs dgemm / dgemv / dsyevd from an
 * optimised library. orc_use_blas(path) dlopens an OpenBLAS build (the `scipy_`-prefixed one bundled with scipy on this
 * image) and switches mm / mtm / mv / sym_eig to it; the tests keep the self-contained loops below. */
#include <dlfcn.h>
typedef void (*dgemm_fn)(int, int, int, int, int, int, double, const double *, int, const double *, int, double, double *, int);
typedef void (*dgemv_fn)(int, int, int, int, double, const double *, int, const double *, int, double, double *, int);
typedef int (*dsyevd_fn)(int, char, char, int, double *, int, double *);
static dgemm_fn g_dgemm; static dgemv_fn g_dgemv; static dsyevd_fn g_dsyevd;
int orc_use_blas(const char *path)
{
    if (!path) { g_dgemm = NULL; g_dgemv = NULL; g_dsyevd = NULL; return 0; }
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return 1;
    const char *pre[2] = {"scipy_", ""};
    for (int i = 0; i < 2; i++) {
        char nm[64];
        snprintf(nm, sizeof nm, "%scblas_dgemm", pre[i]); dgemm_fn a = (dgemm_fn)dlsym(h, nm);
        snprintf(nm, sizeof nm, "%scblas_dgemv", pre[i]); dgemv_fn b = (dgemv_fn)dlsym(h, nm);
        snprintf(nm, sizeof nm, "%sLAPACKE_dsyevd", pre[i]); dsyevd_fn c = (dsyevd_fn)dlsym(h, nm);
        snprintf(nm, sizeof nm, "%sopenblas_set_num_threads", pre[i]); void (*t)(int) = (void (*)(int))dlsym(h, nm);
        if (a && b && c) {
            if (t) t(1);   /* one core per chain: the reference's MH loop is sequential, chains run in parallel threads */
            g_dgemm = a; g_dgemv = b; g_dsyevd = c;
            return 0;
        }
    }
    return 2;
}
int orc_blas_enabled(void) { return g_dgemm != NULL; }

/* C (m x n) = A (m x k) * B (k x n) */
static void mm(int m, int n, int k, const double *A, const double *B, double *C)
{
    if (g_dgemm) { g_dgemm(101, 111, 111, m, n, k, 1.0, A, k, B, n, 0.0, C, n); return; }
    memset(C, 0, sizeof(double) * (size_t)m * n);
    for (int i = 0; i < m; i++)
        for (int p = 0; p < k; p++) {
            double a = A[(size_t)i * k + p];
            const double *b = B + (size_t)p * n; double *c = C + (size_t)i * n;
            for (int j = 0; j < n; j++) c[j] += a * b[j];
        }
}
/* C (m x n) = A^T * B with A (k x m), B (k x n) */
static void mtm(int m, int n, int k, const double *A, const double *B, double *C)
{
    if (g_dgemm) { g_dgemm(101, 112, 111, m, n, k, 1.0, A, m, B, n, 0.0, C, n); return; }
    memset(C, 0, sizeof(double) * (size_t)m * n);
    for (int p = 0; p < k; p++)
        for (int i = 0; i < m; i++) {
            double a = A[(size_t)p * m + i];
            const double *b = B + (size_t)p * n; double *c = C + (size_t)i * n;
            for (int j = 0; j < n; j++) c[j] += a * b[j];
        }
}
static void mv(int m, int n, const double *A, const double *x, double *y)
{
    if (g_dgemv) { g_dgemv(101, 111, m, n, 1.0, A, n, x, 1, 0.0, y, 1); return; }
    for (int i = 0; i < m; i++) {
        double s = 0; const double *a = A + (size_t)i * n;
        for (int j = 0; j < n; j++) s += a[j] * x[j];
        y[i] = s;
    }
}

/* Symmetric eigen-decomposition A = V diag(w) V^T (Householder tridiagonalisation + implicit QL,
 * the classic tred2/tql2 scheme). For the symmetric positive (semi-)definite matrices of this path
 * it coincides with the SVD Breeze computes (svd / pinv): singular values = eigenvalues,
 * U = V. Output sorted descending like an SVD. V is n x n row-major with eigenvectors as columns. */
static void sym_eig(int n, const double *Ain, double *V, double *w)
{
    if (g_dsyevd) {   /* LAPACK divide and conquer: ascending eigenvalues, eigenvectors in the columns; reversed to descending */
        memcpy(V, Ain, sizeof(double) * (size_t)n * n);
        if (g_dsyevd(101, 'V', 'U', n, V, n, w) == 0) {
            for (int j = 0; j < n / 2; j++) {
                double t = w[j]; w[j] = w[n - 1 - j]; w[n - 1 - j] = t;
                for (int i = 0; i < n; i++) { double u = V[(size_t)i * n + j]; V[(size_t)i * n + j] = V[(size_t)i * n + n - 1 - j]; V[(size_t)i * n + n - 1 - j] = u; }
            }
            return;
        }
    }
    double *d = w, *e = (double *)malloc(sizeof(double) * n);
    memcpy(V, Ain, sizeof(double) * (size_t)n * n);
#define VV(i, j) V[(size_t)(i) * n + (j)]
    for (int j = 0; j < n; j++) d[j] = VV(n - 1, j);
    for (int i = n - 1; i > 0; i--) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; k++) scale += fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; j++) { d[j] = VV(i - 1, j); VV(i, j) = 0.0; VV(j, i) = 0.0; }
        } else {
            for (int k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
            double f = d[i - 1], g = sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g; h -= f * g; d[i - 1] = f - g;
            for (int j = 0; j < i; j++) e[j] = 0.0;
            for (int j = 0; j < i; j++) {
                f = d[j]; VV(j, i) = f; g = e[j] + VV(j, j) * f;
                for (int k = j + 1; k <= i - 1; k++) { g += VV(k, j) * d[k]; e[k] += VV(k, j) * f; }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
            double hh = f / (h + h);
            for (int j = 0; j < i; j++) e[j] -= hh * d[j];
            for (int j = 0; j < i; j++) {
                f = d[j]; g = e[j];
                for (int k = j; k <= i - 1; k++) VV(k, j) -= (f * e[k] + g * d[k]);
                d[j] = VV(i - 1, j); VV(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    for (int i = 0; i < n - 1; i++) {
        VV(n - 1, i) = VV(i, i); VV(i, i) = 1.0;
        double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; k++) d[k] = VV(k, i + 1) / h;
            for (int j = 0; j <= i; j++) {
                double g = 0.0;
                for (int k = 0; k <= i; k++) g += VV(k, i + 1) * VV(k, j);
                for (int k = 0; k <= i; k++) VV(k, j) -= g * d[k];
            }
        }
        for (int k = 0; k <= i; k++) VV(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; j++) { d[j] = VV(n - 1, j); VV(n - 1, j) = 0.0; }
    VV(n - 1, n - 1) = 1.0; e[0] = 0.0;
    /* implicit QL */
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0, eps = 2.220446049250313e-16;
    for (int l = 0; l < n; l++) {
        double t = fabs(d[l]) + fabs(e[l]); if (t > tst1) tst1 = t;
        int m = l;
        while (m < n) { if (fabs(e[m]) <= eps * tst1) break; m++; }
        if (m > l) {
            int iter = 0;
            do {
                iter++;
                double g = d[l], p = (d[l + 1] - g) / (2.0 * e[l]), r = hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r); d[l + 1] = e[l] * (p + r);
                double dl1 = d[l + 1], h = g - d[l];
                for (int i = l + 2; i < n; i++) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i]; h = c * p; r = hypot(p, e[i]);
                    e[i + 1] = s * r; s = e[i] / r; c = p / r; p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; k++) {
                        h = VV(k, i + 1);
                        VV(k, i + 1) = s * VV(k, i) + c * h;
                        VV(k, i) = c * VV(k, i) - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p; d[l] = c * p;
            } while (fabs(e[l]) > eps * tst1 && iter < 200);
        }
        d[l] = d[l] + f; e[l] = 0.0;
    }
    /* sort descending */
    for (int i = 0; i < n - 1; i++) {
        int k = i; double p = d[i];
        for (int j = i + 1; j < n; j++) if (d[j] > p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i]; d[i] = p;
            for (int j = 0; j < n; j++) { double t = VV(j, i); VV(j, i) = VV(j, k); VV(j, k) = t; }
        }
    }
#undef VV
    free(e);
}

/* breeze.linalg.pinv for a symmetric matrix: V diag(1/s) V^T, only exact zeros dropped (A3) */
static void pinv_sym(int n, const double *A, double *Ainv)
{
    double *V = (double *)malloc(sizeof(double) * (size_t)n * n), *w = (double *)malloc(sizeof(double) * n);
    double *T = (double *)malloc(sizeof(double) * (size_t)n * n);
    sym_eig(n, A, V, w);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) T[(size_t)j * n + i] = V[(size_t)i * n + j] * (w[j] == 0.0 ? 0.0 : 1.0 / w[j]); /* T = diag(1/w) V^T */
    mm(n, n, n, V, T, Ainv);
    free(V); free(w); free(T);
}

static void inv3(const double *a, double *o)
{
    double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    o[0] = c00 / det; o[1] = (a[2] * a[7] - a[1] * a[8]) / det; o[2] = (a[1] * a[5] - a[2] * a[4]) / det;
    o[3] = c01 / det; o[4] = (a[0] * a[8] - a[2] * a[6]) / det; o[5] = (a[2] * a[3] - a[0] * a[5]) / det;
    o[6] = c02 / det; o[7] = (a[1] * a[6] - a[0] * a[7]) / det; o[8] = (a[0] * a[4] - a[1] * a[3]) / det;
}

/* lower Cholesky, in place on the lower triangle; returns 0 on success */
static int chol_lower(int n, double *A)
{
    for (int j = 0; j < n; j++) {
        double s = A[(size_t)j * n + j];
        for (int k = 0; k < j; k++) s -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(s > 0.0)) return 1;
        double d = sqrt(s); A[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double t = A[(size_t)i * n + j];
            for (int k = 0; k < j; k++) t -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
            A[(size_t)i * n + j] = t / d;
        }
    }
    return 0;
}

/* ============================================================================================ */
/* model                                                                                         */
/* ============================================================================================ */
struct orc_model {
    int N, T, K;
    double *ref, *mean_def, *U, *var, *Q; /* Q = U diag(sqrt(var)), 3N x K row-major */
    int32_t *tris;
    double *S; /* Appendix A5 constant, lazily computed, used by the closed-form variants only */
};

orc_model *orc_model_create(int N, int T, int K, const double *ref, const double *mean_def,
                            const double *U, const double *variance, const int32_t *tris)
{
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    m->N = N; m->T = T; m->K = K;
    m->ref = (double *)malloc(sizeof(double) * 3 * N); memcpy(m->ref, ref, sizeof(double) * 3 * N);
    m->mean_def = (double *)calloc(3 * N, sizeof(double));
    if (mean_def) memcpy(m->mean_def, mean_def, sizeof(double) * 3 * N);
    m->U = (double *)malloc(sizeof(double) * 3 * N * K); memcpy(m->U, U, sizeof(double) * 3 * N * K);
    m->var = (double *)malloc(sizeof(double) * K); memcpy(m->var, variance, sizeof(double) * K);
    m->Q = (double *)malloc(sizeof(double) * 3 * N * K);
    for (size_t r = 0; r < (size_t)3 * N; r++)
        for (int j = 0; j < K; j++) m->Q[r * K + j] = U[r * K + j] * sqrt(variance[j]);
    m->tris = (int32_t *)malloc(sizeof(int32_t) * 3 * T); memcpy(m->tris, tris, sizeof(int32_t) * 3 * T);
    return m;
}
void orc_model_free(orc_model *m)
{
    if (!m) return;
    free(m->ref); free(m->mean_def); free(m->U); free(m->var); free(m->Q); free(m->tris); free(m->S); free(m);
}
int orc_model_rank(const orc_model *m) { return m->K; }

/* Scalismo Rotation(phi, theta, psi, centre): R = Rz(phi) Ry(theta) Rx(psi) ([S-recall]) */
void orc_pose_matrix(const double *theta, double R[9])
{
    double phi = theta[4], th = theta[5], psi = theta[6];
    double cph = cos(phi), sph = sin(phi), cth = cos(th), sth = sin(th), cps = cos(psi), sps = sin(psi);
    R[0] = cth * cph; R[1] = sps * sth * cph - cps * sph; R[2] = sps * sph + cps * sth * cph;
    R[3] = cth * sph; R[4] = cps * cph + sps * sth * sph; R[5] = cps * sth * sph - sps * cph;
    R[6] = -sth;      R[7] = sps * cth;                   R[8] = cps * cth;
}

/* ModelFittingParameters.transformedMesh (api/sampling/ModelFittingParameters.scala:93-110):
 * x = s * ( R (ref + mean + Q alpha - c) + c + t ) */
void orc_transformed_mesh(const orc_model *m, const double *theta, double *xyz)
{
    int K = m->K; const double *alpha = theta + 10;
    double R[9]; orc_pose_matrix(theta, R);
    double s = theta[0]; const double *t = theta + 1, *c = theta + 7;
    for (int i = 0; i < m->N; i++) {
        double p[3];
        for (int d = 0; d < 3; d++) {
            const double *q = m->Q + (size_t)(3 * i + d) * K; double acc = 0;
            for (int j = 0; j < K; j++) acc += q[j] * alpha[j];
            p[d] = m->ref[3 * i + d] + (m->mean_def[3 * i + d] + acc);
        }
        double pc[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
        for (int d = 0; d < 3; d++)
            xyz[3 * i + d] = s * ((R[3 * d] * pc[0] + R[3 * d + 1] * pc[1] + R[3 * d + 2] * pc[2]) + c[d] + t[d]);
    }
}

/* inverse of poseTransform (NonRigidIcpProposal.scala:142): x -> R^T (x - t - c) + c */
static void inverse_pose(const double *theta, const double *x, double *o)
{
    double R[9]; orc_pose_matrix(theta, R);
    const double *t = theta + 1, *c = theta + 7;
    double y[3] = {x[0] - t[0] - c[0], x[1] - t[1] - c[1], x[2] - t[2] - c[2]};
    for (int d = 0; d < 3; d++) o[d] = (R[d] * y[0] + R[3 + d] * y[1] + R[6 + d] * y[2]) + c[d];
}

/* api/sampling/SurfaceNoiseHelpers.scala:32-60, including the inverted fallback test at :46 */
void orc_surface_noise_cov(const double normal[3], double sd_normal, double sd_tangent, double cov[9])
{
    double n[3] = {normal[0], normal[1], normal[2]}; v_normalize(n);
    const double ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0};
    double cand[3], t1[3], t2[3];
    v_cross(n, ex, cand);
    if (v_dot(cand, cand) < 0.0001) memcpy(t1, cand, 24); else v_cross(n, ey, t1);
    v_normalize(t1);
    v_cross(n, t1, t2); v_normalize(t2);
    double vn = sd_normal * sd_normal, vt = sd_tangent * sd_tangent;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) cov[3 * i + j] = n[i] * vn * n[j] + t1[i] * vt * t1[j] + t2[i] * vt * t2[j];
}

/* ============================================================================================ */
/* LowRankGaussianProcess.regression core (Appendix A3)                                          */
/* ============================================================================================ */
/* Qs: 3n x K rows of the (possibly rotated) scaled basis at the observed points, covs: n x 9,
 * resid: 3n (= y - mean at the observed points). Outputs M, Minv (K x K), coeffs (K). */
static void regression_core(int K, int n, const double *Qs, const double *covs, int iso, double iso_var,
                            const double *resid, double *M, double *Minv, double *coeffs)
{
    size_t n3 = (size_t)3 * n;
    double *QtL = (double *)malloc(sizeof(double) * K * n3);
    for (int i = 0; i < n; i++) {
        double L[9];
        if (iso) { memset(L, 0, sizeof L); L[0] = L[4] = L[8] = 1.0 / iso_var; }
        else inv3(covs + 9 * i, L);
        for (int j = 0; j < K; j++) {
            double q0 = Qs[(size_t)(3 * i) * K + j], q1 = Qs[(size_t)(3 * i + 1) * K + j], q2 = Qs[(size_t)(3 * i + 2) * K + j];
            for (int d = 0; d < 3; d++) QtL[(size_t)j * n3 + 3 * i + d] = q0 * L[d] + q1 * L[3 + d] + q2 * L[6 + d];
        }
    }
    double *Mloc = M ? M : (double *)malloc(sizeof(double) * K * K);
    mm(K, K, (int)n3, QtL, Qs, Mloc);
    for (int j = 0; j < K; j++) Mloc[(size_t)j * K + j] += 1.0;
    double *Mi = Minv ? Minv : (double *)malloc(sizeof(double) * K * K);
    pinv_sym(K, Mloc, Mi);
    if (coeffs) { /* (Minv * QtL) * resid, in the reference's association order */
        double *P = (double *)malloc(sizeof(double) * K * n3);
        mm(K, (int)n3, K, Mi, QtL, P);
        mv(K, (int)n3, P, resid, coeffs);
        free(P);
    }
    if (!M) free(Mloc);
    if (!Minv) free(Mi);
    free(QtL);
}

/* ============================================================================================ */
/* NonRigidIcpProposal                                                                           */
/* ============================================================================================ */
#define POST_CACHE 20
typedef struct {
    int valid; double *theta; /* K+10 */
    int n_obs;
    double *mu, *M, *Minv;    /* K, KxK, KxK */
    int have_basis;           /* lazily: Ubar, lamp, Up (3N x K) */
    double *Ubar, *lamp, *Up;
} post_entry;

struct orc_proposal {
    const orc_model *model; const orc_mesh *target;
    double step, sd_t, sd_n; int direction, boundary_aware;
    int n_ids; int32_t *ids; int n_tp; double *tp;
    post_entry cache[POST_CACHE]; int next; /* Memoize(icpPosterior, 20), :49 */
    double *Mc_inv;                         /* unused cache slot (the reference recomputes) */
};

orc_proposal *orc_proposal_create(const orc_model *model, const orc_mesh *target, double step_length,
                                  double tangential_noise, double noise_along_normal, int direction,
                                  int boundary_aware, int n_ids, const int32_t *ids, int n_tp,
                                  const double *target_points)
{
    orc_proposal *p = (orc_proposal *)calloc(1, sizeof(orc_proposal));
    p->model = model; p->target = target; p->step = step_length; p->sd_t = tangential_noise;
    p->sd_n = noise_along_normal; p->direction = direction; p->boundary_aware = boundary_aware;
    p->n_ids = n_ids; p->ids = (int32_t *)malloc(sizeof(int32_t) * (n_ids > 0 ? n_ids : 1));
    if (n_ids > 0) memcpy(p->ids, ids, sizeof(int32_t) * n_ids);
    p->n_tp = n_tp; p->tp = (double *)malloc(sizeof(double) * 3 * (n_tp > 0 ? n_tp : 1));
    if (n_tp > 0) memcpy(p->tp, target_points, sizeof(double) * 3 * n_tp);
    return p;
}

static void post_entry_free(post_entry *e)
{
    free(e->theta); free(e->mu); free(e->M); free(e->Minv); free(e->Ubar); free(e->lamp); free(e->Up);
    memset(e, 0, sizeof *e);
}
void orc_proposal_free(orc_proposal *p)
{
    if (!p) return;
    for (int i = 0; i < POST_CACHE; i++) post_entry_free(&p->cache[i]);
    free(p->ids); free(p->tp); free(p);
}

/* uncertainDisplacementEstimation (:139-149) -> observation list */
static int icp_observations(const orc_proposal *p, const double *theta, int32_t *obs_ids, double *obs_y, double *obs_cov)
{
    const orc_model *md = p->model; int N = md->N;
    double *xyz = (double *)malloc(sizeof(double) * 3 * N);
    orc_transformed_mesh(md, theta, xyz);                             /* :141 */
    orc_mesh *cur = orc_mesh_create(N, xyz, md->T, md->tris);         /* new TriangleMesh per call */
    int n = 0;
    if (p->direction == ORC_TARGET_SAMPLING) {                        /* :112-131 */
        for (int i = 0; i < p->n_tp; i++) {
            const double *tp = p->tp + 3 * i; int32_t id; double nrm[3];
            orc_mesh_closest_vertex(cur, 1, tp, &id, NULL);           /* :118 */
            int on_b = cur->boundary[id];                             /* :119 */
            if (p->boundary_aware && on_b) continue;                  /* :124 */
            vertex_normal(cur, id, nrm);                              /* :120 */
            orc_surface_noise_cov(nrm, p->sd_n, p->sd_t, obs_cov + 9 * n);
            double ip[3]; inverse_pose(theta, tp, ip);
            obs_ids[n] = id;
            for (int d = 0; d < 3; d++) obs_y[3 * n + d] = ip[d] - md->ref[3 * id + d]; /* :129 */
            n++;
        }
    } else {                                                           /* :88-110 */
        for (int i = 0; i < p->n_ids; i++) {
            int id = p->ids[i]; double cp[3], nrm[3]; int32_t tid;
            orc_mesh_closest_point(p->target, 1, xyz + 3 * id, NULL, NULL, cp, NULL); /* :97 */
            orc_mesh_closest_vertex(p->target, 1, cp, &tid, NULL);                    /* :98 */
            int on_b = p->target->boundary[tid];                                      /* :99 */
            if (p->boundary_aware && on_b) continue;                                  /* :104 */
            vertex_normal(cur, id, nrm);                                              /* :100 */
            orc_surface_noise_cov(nrm, p->sd_n, p->sd_t, obs_cov + 9 * n);
            double ip[3]; inverse_pose(theta, cp, ip);
            obs_ids[n] = id;
            for (int d = 0; d < 3; d++) obs_y[3 * n + d] = ip[d] - md->ref[3 * id + d]; /* :108 */
            n++;
        }
    }
    orc_mesh_free(cur); free(xyz);
    return n;
}

static post_entry *posterior_get(orc_proposal *p, const double *theta)
{
    const orc_model *md = p->model; int K = md->K, L = K + 10;
    for (int i = 0; i < POST_CACHE; i++)
        if (p->cache[i].valid && memcmp(p->cache[i].theta, theta, sizeof(double) * L) == 0) return &p->cache[i];
    post_entry *e = &p->cache[p->next]; p->next = (p->next + 1) % POST_CACHE;
    post_entry_free(e);
    e->theta = (double *)malloc(sizeof(double) * L); memcpy(e->theta, theta, sizeof(double) * L);
    int nmax = p->direction == ORC_TARGET_SAMPLING ? p->n_tp : p->n_ids; if (nmax < 1) nmax = 1;
    int32_t *oid = (int32_t *)malloc(sizeof(int32_t) * nmax);
    double *oy = (double *)malloc(sizeof(double) * 3 * nmax), *oc = (double *)malloc(sizeof(double) * 9 * nmax);
    int n = icp_observations(p, theta, oid, oy, oc);
    double *Qs = (double *)malloc(sizeof(double) * 3 * (n > 0 ? n : 1) * K), *res = (double *)malloc(sizeof(double) * 3 * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            memcpy(Qs + (size_t)(3 * i + d) * K, md->Q + (size_t)(3 * oid[i] + d) * K, sizeof(double) * K);
            res[3 * i + d] = oy[3 * i + d] - md->mean_def[3 * oid[i] + d];
        }
    e->mu = (double *)malloc(sizeof(double) * K); e->M = (double *)malloc(sizeof(double) * K * K);
    e->Minv = (double *)malloc(sizeof(double) * K * K);
    regression_core(K, n, Qs, oc, 0, 0.0, res, e->M, e->Minv, e->mu); /* :152 interpolatedModel.posterior */
    e->n_obs = n; e->valid = 1;
    free(oid); free(oy); free(oc); free(Qs); free(res);
    return e;
}

/* Appendix A4: Sigma' = D Minv D, svd -> rotated eigenfunctions evaluated on all N reference points */
static void posterior_basis(const orc_proposal *p, post_entry *e)
{
    if (e->have_basis) return;
    const orc_model *md = p->model; int K = md->K; size_t n3 = (size_t)3 * md->N;
    double *S = (double *)malloc(sizeof(double) * K * K);
    for (int i = 0; i < K; i++)
        for (int j = 0; j < K; j++) S[(size_t)i * K + j] = sqrt(md->var[i]) * e->Minv[(size_t)i * K + j] * sqrt(md->var[j]);
    /* symmetrise rounding noise so that the eigen-solver sees an exactly symmetric matrix */
    for (int i = 0; i < K; i++)
        for (int j = 0; j < i; j++) { double a = 0.5 * (S[(size_t)i * K + j] + S[(size_t)j * K + i]); S[(size_t)i * K + j] = S[(size_t)j * K + i] = a; }
    e->Ubar = (double *)malloc(sizeof(double) * K * K); e->lamp = (double *)malloc(sizeof(double) * K);
    sym_eig(K, S, e->Ubar, e->lamp);
    /* Sign convention of the singular vectors: Breeze's svd returns whatever LAPACK dgesdd produced, which cannot be
     * known without running the JVM (parity unpinned). Oracle, numpy oracle and device all use the
     * same documented rule instead: the entry of largest magnitude of every vector (lowest index on ties) is positive. */
    for (int j = 0; j < K; j++) {
        int im = 0; double vm = 0.0;
        for (int i = 0; i < K; i++) { double a = fabs(e->Ubar[(size_t)i * K + j]); if (a > vm) { vm = a; im = i; } }
        if (e->Ubar[(size_t)im * K + j] < 0.0)
            for (int i = 0; i < K; i++) e->Ubar[(size_t)i * K + j] = -e->Ubar[(size_t)i * K + j];
    }
    e->Up = (double *)malloc(sizeof(double) * n3 * K);
    mm((int)n3, K, K, md->U, e->Ubar, e->Up);   /* phi'_i(x) = sum_j phi_j(x) Ubar[j,i], N K^2 work */
    e->have_basis = 1;
    free(S);
}

int orc_icp_posterior(const orc_proposal *pc, const double *theta, double *mu, double *M, double *Minv,
                      int32_t *obs_ids, double *obs_y, double *obs_cov)
{
    orc_proposal *p = (orc_proposal *)pc; int K = p->model->K;
    post_entry *e = posterior_get(p, theta);
    if (mu) memcpy(mu, e->mu, sizeof(double) * K);
    if (M) memcpy(M, e->M, sizeof(double) * K * K);
    if (Minv) memcpy(Minv, e->Minv, sizeof(double) * K * K);
    if (obs_ids && obs_y && obs_cov) icp_observations(p, theta, obs_ids, obs_y, obs_cov);
    return e->n_obs;
}

/* StatisticalMeshModel.coefficients(mesh): regression on all N points, noise 1e-5 I (Appendix A5).
 * Qfull: 3N x K scaled basis, resid: 3N. The reference rebuilds M and its pinv on every call. */
static void full_mesh_coefficients(int N, int K, const double *Qfull, const double *resid, double *coeffs)
{
    regression_core(K, N, Qfull, NULL, 1, 1e-5, resid, NULL, NULL, coeffs);
}

/* propose (:53-68) */
void orc_propose(const orc_proposal *pc, const double *theta, const double *z, double *theta_out)
{
    orc_proposal *p = (orc_proposal *)pc; const orc_model *md = p->model; int K = md->K; size_t n3 = (size_t)3 * md->N;
    post_entry *e = posterior_get(p, theta);                    /* :54 */
    posterior_basis(p, e);
    /* posterior.sample(): mean_p + sum_i phi'_i sqrt(lambda'_i) z_i, evaluated at every reference point (:55-57) */
    double *w = (double *)malloc(sizeof(double) * K), *def = (double *)malloc(sizeof(double) * n3), *tmp = (double *)malloc(sizeof(double) * n3);
    for (int i = 0; i < K; i++) w[i] = sqrt(e->lamp[i] > 0 ? e->lamp[i] : 0.0) * z[i];
    mv((int)n3, K, md->Q, e->mu, def);
    mv((int)n3, K, e->Up, w, tmp);
    for (size_t r = 0; r < n3; r++) def[r] = (md->mean_def[r] + def[r]) + tmp[r];
    /* model.coefficients(referenceMesh.transform(f)) (:59): residual = deformation - model mean */
    for (size_t r = 0; r < n3; r++) tmp[r] = def[r] - md->mean_def[r];
    double *anew = (double *)malloc(sizeof(double) * K);
    full_mesh_coefficients(md->N, K, md->Q, tmp, anew);
    memcpy(theta_out, theta, sizeof(double) * (K + 10));
    for (int j = 0; j < K; j++) theta_out[10 + j] = theta[10 + j] + (anew[j] - theta[10 + j]) * p->step; /* :61-62 */
    free(w); free(def); free(tmp); free(anew);
}

static int only_shape_changed(int K, const double *from, const double *to)
{
    (void)K;
    for (int i = 0; i < 10; i++) if (!(to[i] == from[i])) return 0; /* DenseVector != , :72 */
    return 1;
}

/* logTransitionProbability (:71-85) */
double orc_log_transition(const orc_proposal *pc, const double *from, const double *to)
{
    orc_proposal *p = (orc_proposal *)pc; const orc_model *md = p->model; int K = md->K; size_t n3 = (size_t)3 * md->N;
    if (!only_shape_changed(K, from, to)) return -INFINITY;     /* :72-74 */
    post_entry *e = posterior_get(p, from);                     /* :76 */
    posterior_basis(p, e);                                      /* :77 StatisticalMeshModel(referenceMesh, pos) */
    double *comp = (double *)malloc(sizeof(double) * K);
    for (int j = 0; j < K; j++) comp[j] = from[10 + j] + ((to[10 + j] - from[10 + j]) / p->step); /* :79 */
    double *inst = (double *)malloc(sizeof(double) * n3), *pm = (double *)malloc(sizeof(double) * n3);
    mv((int)n3, K, md->Q, comp, inst);                          /* :80 model.instance: displacement = mean + Q c */
    mv((int)n3, K, md->Q, e->mu, pm);                           /* posterior mean displacement = mean + Q mu */
    for (size_t r = 0; r < n3; r++) inst[r] = (md->mean_def[r] + inst[r]) - (md->mean_def[r] + pm[r]);
    double *Qp = (double *)malloc(sizeof(double) * n3 * K);
    for (size_t r = 0; r < n3; r++)
        for (int j = 0; j < K; j++) Qp[r * K + j] = e->Up[r * K + j] * sqrt(e->lamp[j] > 0 ? e->lamp[j] : 0.0);
    double *proj = (double *)malloc(sizeof(double) * K);
    full_mesh_coefficients(md->N, K, Qp, inst, proj);           /* :82 posterior.coefficients(toMesh) */
    double ss = 0; for (int j = 0; j < K; j++) ss += proj[j] * proj[j];
    free(comp); free(inst); free(pm); free(Qp); free(proj);
    return -0.5 * (K * LOG_2PI + ss);                           /* :83 pos.logpdf: N(0, I_K) on the coefficients */
}

/* ---- Appendix A closed forms ------------------------------------------------------------------ */
static void compute_S(const orc_model *md, double **S_out)
{
    /* S = (G/eps + I)^-1 G/eps with G = Q^T Q, eps = 1e-5 (Appendix A5) */
    int K = md->K; size_t n3 = (size_t)3 * md->N;
    double *G = (double *)malloc(sizeof(double) * K * K), *A = (double *)malloc(sizeof(double) * K * K), *Ai = (double *)malloc(sizeof(double) * K * K);
    mtm(K, K, (int)n3, md->Q, md->Q, G);
    for (int i = 0; i < K * K; i++) { G[i] /= 1e-5; A[i] = G[i]; }
    for (int j = 0; j < K; j++) A[(size_t)j * K + j] += 1.0;
    pinv_sym(K, A, Ai);
    *S_out = (double *)malloc(sizeof(double) * K * K);
    mm(K, K, K, Ai, G, *S_out);
    free(G); free(A); free(Ai);
}

/* alpha_new = S (mu + W z), W = L^-T with M = L L^T (any W with W W^T = Minv gives the same law; the
 * reference's W comes from an SVD whose vector signs are implementation-defined, Appendix A6) */
void orc_propose_closed_form(const orc_proposal *pc, const double *theta, const double *z, double *theta_out)
{
    orc_proposal *p = (orc_proposal *)pc; const orc_model *md = p->model; int K = md->K;
    post_entry *e = posterior_get(p, theta);
    double *L = (double *)malloc(sizeof(double) * K * K), *w = (double *)malloc(sizeof(double) * K), *a = (double *)malloc(sizeof(double) * K);
    memcpy(L, e->M, sizeof(double) * K * K);
    if (chol_lower(K, L)) { for (int j = 0; j < K + 10; j++) theta_out[j] = NAN; free(L); free(w); free(a); return; }
    for (int i = K - 1; i >= 0; i--) { /* L^T w = z */
        double s = z[i];
        for (int k = i + 1; k < K; k++) s -= L[(size_t)k * K + i] * w[k];
        w[i] = s / L[(size_t)i * K + i];
    }
    for (int j = 0; j < K; j++) w[j] += e->mu[j];
    if (!md->S) compute_S(md, &((orc_model *)md)->S);
    mv(K, K, md->S, w, a);
    memcpy(theta_out, theta, sizeof(double) * (K + 10));
    for (int j = 0; j < K; j++) theta_out[10 + j] = theta[10 + j] + (a[j] - theta[10 + j]) * p->step;
    free(L); free(w); free(a);
}

/* -1/2 (K ln 2pi + d^T M d), d = alpha_c - mu_from (Appendix A7) */
double orc_log_transition_closed_form(const orc_proposal *pc, const double *from, const double *to)
{
    orc_proposal *p = (orc_proposal *)pc; int K = p->model->K;
    if (!only_shape_changed(K, from, to)) return -INFINITY;
    post_entry *e = posterior_get(p, from);
    double *d = (double *)malloc(sizeof(double) * K), *Md = (double *)malloc(sizeof(double) * K);
    for (int j = 0; j < K; j++) d[j] = (from[10 + j] + ((to[10 + j] - from[10 + j]) / p->step)) - e->mu[j];
    mv(K, K, e->M, d, Md);
    double q = 0; for (int j = 0; j < K; j++) q += d[j] * Md[j];
    free(d); free(Md);
    return -0.5 * (K * LOG_2PI + q);
}

/* ============================================================================================ */
/* registration quality measures                                                                  */
/* ============================================================================================ */
/* api/other/RegistrationComparison.scala:24-49 between transformedMesh(theta) and the target:
 * out = {MeshMetrics.avgDistance (:25), MeshMetrics.hausdorffDistance (:27, Appendix A14), avgDistanceBoundaryAware
 * average and maximum (:31-43)}. An empty filtered list gives NaN for both boundary-aware values (the reference
 * divides 0 by 0 and then throws on .max). */
void orc_registration_metrics(const orc_model *md, const orc_mesh *target, const double *theta, double out[4])
{
    int N = md->N;
    double *xyz = (double *)malloc(sizeof(double) * 3 * N);
    orc_transformed_mesh(md, theta, xyz);
    orc_mesh *cur = orc_mesh_create(N, xyz, md->T, md->tris);
    double sum = 0, mx = 0, sb = 0, mb = -INFINITY; int nb = 0;
    for (int i = 0; i < N; i++) {
        double cp[3], d2; int32_t tid;
        orc_mesh_closest_point(target, 1, xyz + 3 * i, NULL, NULL, cp, &d2);     /* :35 */
        orc_mesh_closest_vertex(target, 1, cp, &tid, NULL);                      /* :36 */
        double d = sqrt(d2);
        sum += d; if (d > mx) mx = d;
        if (!target->boundary[tid]) { sb += d; nb++; if (d > mb) mb = d; }       /* :37-38 */
    }
    double back = 0;
    for (int i = 0; i < target->nv; i++) {
        double d2; orc_mesh_closest_point(cur, 1, target->v + 3 * i, NULL, NULL, NULL, &d2);
        if (sqrt(d2) > back) back = sqrt(d2);
    }
    out[0] = sum / N; out[1] = mx > back ? mx : back;
    out[2] = nb ? sb / nb : NAN; out[3] = nb ? mb : NAN;
    orc_mesh_free(cur); free(xyz);
}

/* Scalismo MeshMetrics.diceCoefficient(a, b) [S-recall] (apps/femur/StdIcpVsChainICPrandomInitComparisonAll.scala:46):
 * uniform samples in the box spanned by both bounding boxes; inside(mesh, p) = vertexNormal(v) . (v - p) > 0 with v the
 * mesh vertex nearest to p (toBinaryImage); 2 |A and B| / (|A| + |B|). unit: n x 3 samples in [0, 1] standing in for the
 * reference's unseeded UniformSampler. */
double orc_dice_coefficient(const orc_model *md, const orc_mesh *target, const double *theta, int n, const double *unit)
{
    int N = md->N;
    double *xyz = (double *)malloc(sizeof(double) * 3 * N);
    orc_transformed_mesh(md, theta, xyz);
    orc_mesh *cur = orc_mesh_create(N, xyz, md->T, md->tris);
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < N; i++) for (int d = 0; d < 3; d++) { if (xyz[3 * i + d] < lo[d]) lo[d] = xyz[3 * i + d]; if (xyz[3 * i + d] > hi[d]) hi[d] = xyz[3 * i + d]; }
    for (int i = 0; i < target->nv; i++) for (int d = 0; d < 3; d++) { double x = target->v[3 * i + d]; if (x < lo[d]) lo[d] = x; if (x > hi[d]) hi[d] = x; }
    double na = 0, nb = 0, nab = 0;
    for (int i = 0; i < n; i++) {
        double p[3]; for (int d = 0; d < 3; d++) p[d] = lo[d] + unit[3 * i + d] * (hi[d] - lo[d]);
        int32_t va, vb; double nrm[3], w[3];
        orc_mesh_closest_vertex(cur, 1, p, &va, NULL);
        vertex_normal(cur, va, nrm); v_sub(cur->v + 3 * va, p, w);
        int ina = v_dot(nrm, w) > 0.0;
        orc_mesh_closest_vertex(target, 1, p, &vb, NULL);
        vertex_normal(target, vb, nrm); v_sub(target->v + 3 * vb, p, w);
        int inb = v_dot(nrm, w) > 0.0;
        na += ina; nb += inb; nab += ina && inb;
    }
    orc_mesh_free(cur); free(xyz);
    return 2.0 * nab / (na + nb);
}

/* ============================================================================================ */
/* deterministic ICP iteration (api/other/IcpBasedSurfaceFitting.scala:55-92), identity pose     */
/* ============================================================================================ */
void orc_std_icp_iteration(const orc_model *md, const orc_mesh *target, int direction, int n_ids,
                           const int32_t *ids, int n_tp, const double *target_points, double sigma2,
                           double step_length, const double *alpha, double *alpha_out)
{
    int K = md->K;
    double *theta = (double *)calloc(K + 10, sizeof(double)); theta[0] = 1.0; memcpy(theta + 10, alpha, sizeof(double) * K);
    orc_std_icp_iteration_theta(md, target, direction, n_ids, ids, n_tp, target_points, sigma2, step_length, theta, alpha_out);
    free(theta);
}

/* The same iteration with a rigid transform (currentTrans, :61): the correspondences are found on
 * model.transform(currentTrans).instance(params) = transformedMesh(theta), but the posterior (:81) is that of the
 * UNTRANSFORMED model with the target points as they are - the reference does not pull them back through the transform. */
void orc_std_icp_iteration_theta(const orc_model *md, const orc_mesh *target, int direction, int n_ids,
                                 const int32_t *ids, int n_tp, const double *target_points, double sigma2,
                                 double step_length, const double *theta_in, double *alpha_out)
{
    int K = md->K, N = md->N; size_t n3 = (size_t)3 * N;
    const double *alpha = theta_in + 10;
    double *theta = (double *)malloc(sizeof(double) * (K + 10)); memcpy(theta, theta_in, sizeof(double) * (K + 10));
    double *xyz = (double *)malloc(sizeof(double) * n3);
    orc_transformed_mesh(md, theta, xyz);                       /* :61 instance */
    int n = direction == ORC_MODEL_SAMPLING ? n_ids : n_tp;
    int32_t *cid = (int32_t *)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
    double *cpt = (double *)malloc(sizeof(double) * 3 * (n > 0 ? n : 1));
    if (direction == ORC_MODEL_SAMPLING) {                      /* :71-74 */
        for (int i = 0; i < n; i++) {
            cid[i] = ids[i];
            orc_mesh_closest_point(target, 1, xyz + 3 * ids[i], NULL, NULL, cpt + 3 * i, NULL);
        }
    } else {                                                    /* :75-79 */
        orc_mesh *cur = orc_mesh_create(N, xyz, md->T, md->tris);
        for (int i = 0; i < n; i++) {
            orc_mesh_closest_vertex(cur, 1, target_points + 3 * i, &cid[i], NULL);
            memcpy(cpt + 3 * i, target_points + 3 * i, 24);
        }
        orc_mesh_free(cur);
    }
    /* model.posterior(corr, sigma2).mean (:81-82): discrete regression with isotropic noise */
    double *Qs = (double *)malloc(sizeof(double) * 3 * (n > 0 ? n : 1) * K), *res = (double *)malloc(sizeof(double) * 3 * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            memcpy(Qs + (size_t)(3 * i + d) * K, md->Q + (size_t)(3 * cid[i] + d) * K, sizeof(double) * K);
            res[3 * i + d] = (cpt[3 * i + d] - md->ref[3 * cid[i] + d]) - md->mean_def[3 * cid[i] + d];
        }
    double *mu = (double *)malloc(sizeof(double) * K);
    regression_core(K, n, Qs, NULL, 1, sigma2, res, NULL, NULL, mu);
    /* model.coefficients(fit) (:84): fit displacement - mean = Q mu */
    double *dq = (double *)malloc(sizeof(double) * n3), *cf = (double *)malloc(sizeof(double) * K);
    mv((int)n3, K, md->Q, mu, dq);
    full_mesh_coefficients(N, K, md->Q, dq, cf);
    for (int j = 0; j < K; j++) alpha_out[j] = alpha[j] + (cf[j] - alpha[j]) * step_length; /* :85 */
    free(theta); free(xyz); free(cid); free(cpt); free(Qs); free(res); free(mu); free(dq); free(cf);
}

/* ============================================================================================ */
/* evaluators                                                                                    */
/* ============================================================================================ */
/* breeze Gaussian(mu, sigma).logPdf, Exponential(rate).logPdf (Appendix A15) */
static double gauss_logpdf(double x, double mu, double sd) { return -((x - mu) * (x - mu)) / (2.0 * sd * sd) - log(sd * sqrt(2.0 * M_PI)); }
static double exp_logpdf(double x, double rate) { return log(rate) - rate * x; }

double orc_eval_prior(int K, const double *theta)
{
    double ss = 0; for (int j = 0; j < K; j++) ss += theta[10 + j] * theta[10 + j];
    return -0.5 * (K * LOG_2PI + ss);
}

double orc_eval_independent(const orc_model *md, const orc_mesh *target, int mode, double g_mean, double g_sd,
                            int n_ids, const int32_t *ids, int n_tp, const double *tp, const double *theta)
{
    int N = md->N; double *xyz = (double *)malloc(sizeof(double) * 3 * N);
    orc_transformed_mesh(md, theta, xyz);                       /* :59 */
    double m2t = 0, t2m = 0;
    if (mode != ORC_TARGET_TO_MODEL) {                          /* :40-46 */
        for (int i = 0; i < n_ids; i++) {
            double d2; orc_mesh_closest_point(target, 1, xyz + 3 * ids[i], NULL, NULL, NULL, &d2);
            m2t += gauss_logpdf(sqrt(d2), g_mean, g_sd);
        }
    }
    if (mode != ORC_MODEL_TO_TARGET) {                          /* :49-54 */
        orc_mesh *cur = orc_mesh_create(N, xyz, md->T, md->tris);
        for (int i = 0; i < n_tp; i++) {
            double d2; orc_mesh_closest_point(cur, 1, tp + 3 * i, NULL, NULL, NULL, &d2);
            t2m += gauss_logpdf(sqrt(d2), g_mean, g_sd);
        }
        orc_mesh_free(cur);
    }
    free(xyz);
    if (mode == ORC_MODEL_TO_TARGET) return m2t;
    if (mode == ORC_TARGET_TO_MODEL) return t2m;
    return 0.5 * m2t + 0.5 * t2m;                               /* :63 */
}

double orc_eval_hausdorff(const orc_model *md, const orc_mesh *target, double rate, const double *theta)
{
    int N = md->N; double *xyz = (double *)malloc(sizeof(double) * 3 * N);
    orc_transformed_mesh(md, theta, xyz);
    orc_mesh *cur = orc_mesh_create(N, xyz, md->T, md->tris);
    double hd = 0;
    for (int i = 0; i < N; i++) { double d2; orc_mesh_closest_point(target, 1, xyz + 3 * i, NULL, NULL, NULL, &d2); double d = sqrt(d2); if (d > hd) hd = d; }
    for (int i = 0; i < target->nv; i++) { double d2; orc_mesh_closest_point(cur, 1, target->v + 3 * i, NULL, NULL, NULL, &d2); double d = sqrt(d2); if (d > hd) hd = d; }
    orc_mesh_free(cur); free(xyz);
    return exp_logpdf(hd, rate);                                /* :33-34 */
}

double orc_eval_collective(const orc_model *md, const orc_mesh *target, int mode, double avg_mean, double avg_sd,
                           double max_rate, int n_ids, const int32_t *ids, int n_tp, const double *tp,
                           const double *theta, int *status, double *avg_max)
{
    int N = md->N; double *xyz = (double *)malloc(sizeof(double) * 3 * N);
    orc_transformed_mesh(md, theta, xyz);
    double a1 = 0, m1 = -INFINITY, a2 = 0, m2 = -INFINITY; int c1 = 0, c2 = 0, bad = 0;
    if (mode != ORC_TARGET_TO_MODEL) {                          /* :41-52 */
        for (int i = 0; i < n_ids; i++) {
            double cp[3], d2; int32_t vid;
            orc_mesh_closest_point(target, 1, xyz + 3 * ids[i], NULL, NULL, cp, &d2);
            orc_mesh_closest_vertex(target, 1, cp, &vid, NULL);
            if (target->boundary[vid]) continue;
            double d = sqrt(d2); a1 += d; if (d > m1) m1 = d; c1++;
        }
        if (c1 == 0) bad = 1;
        a1 /= c1;
    }
    if (mode != ORC_MODEL_TO_TARGET) {                          /* :54-64, cross-mesh boundary lookup :58-59 */
        orc_mesh *cur = orc_mesh_create(N, xyz, md->T, md->tris);
        for (int i = 0; i < n_tp; i++) {
            double cp[3], d2; int32_t vid;
            orc_mesh_closest_point(cur, 1, tp + 3 * i, NULL, NULL, cp, &d2);
            orc_mesh_closest_vertex(cur, 1, cp, &vid, NULL);
            if (vid < target->nv && target->boundary[vid]) continue;
            double d = sqrt(d2); a2 += d; if (d > m2) m2 = d; c2++;
        }
        if (c2 == 0) bad = 1;
        a2 /= c2;
        orc_mesh_free(cur);
    }
    free(xyz);
    double avg, mx;
    if (mode == ORC_MODEL_TO_TARGET) { avg = a1; mx = m1; }
    else if (mode == ORC_TARGET_TO_MODEL) { avg = a2; mx = m2; }
    else { avg = 0.5 * a1 + 0.5 * a2; mx = m1 > m2 ? m1 : m2; } /* :71-75 */
    if (status) *status = bad;
    if (avg_max) { avg_max[0] = avg; avg_max[1] = mx; }
    return gauss_logpdf(avg, avg_mean, avg_sd) + exp_logpdf(mx, max_rate); /* :77 */
}

/* MVN(0, sd^2 I).logpdf(residual) (RandomShapeUpdateProposal.scala:30,38-45); evaluated in the
 * numerically stable log form (Appendix A15) */
double orc_random_walk_log_transition(int K, double sd, const double *from, const double *to)
{
    if (!only_shape_changed(K, from, to)) return -INFINITY;
    double ss = 0; for (int j = 0; j < K; j++) { double r = to[10 + j] - from[10 + j]; ss += r * r; }
    return -0.5 * (K * LOG_2PI + K * log(sd * sd) + ss / (sd * sd));
}

double orc_pose_log_transition(int K, int kind, int axis, double sd, const double *from, const double *to)
{
    int slot = kind == 0 ? 4 + axis : 1 + axis; /* rotation (phi,theta,psi) at 4..6, translation at 1..3 */
    for (int i = 0; i < K + 10; i++) {
        int in_group = kind == 0 ? (i >= 4 && i <= 6) : (i >= 1 && i <= 3);
        if (in_group) continue; /* PoseProposals.scala:48 / :82: everything but the whole rotation / translation group must match */
        if (!(to[i] == from[i])) return -INFINITY;
    }
    return gauss_logpdf(to[slot] - from[slot], 0.0, sd);
}

/* ============================================================================================ */
/* Metropolis-Hastings (Scalismo MetropolisHastings.next, MixtureProposal; Appendix A8/A9)       */
/* ============================================================================================ */
static double eval_distance(const orc_chain_desc *d, const double *theta)
{
    switch (d->eval_kind) {
    case ORC_EVAL_INDEPENDENT: return orc_eval_independent(d->model, d->target, d->eval_mode, d->p0, d->p1, d->n_ids, d->ids, d->n_tp, d->target_points, theta);
    case ORC_EVAL_HAUSDORFF: return orc_eval_hausdorff(d->model, d->target, d->p0, theta);
    case ORC_EVAL_COLLECTIVE: return orc_eval_collective(d->model, d->target, d->eval_mode, d->p0, d->p1, d->p2, d->n_ids, d->ids, d->n_tp, d->target_points, theta, NULL, NULL);
    default: return 0.0;
    }
}

static double comp_log_transition(const orc_chain_desc *d, const orc_component *c, const double *from, const double *to)
{
    int K = d->model->K;
    switch (c->kind) {
    case ORC_PROP_ICP: return d->closed_form ? orc_log_transition_closed_form(c->icp, from, to) : orc_log_transition(c->icp, from, to);
    case ORC_PROP_RANDOM_SHAPE: return orc_random_walk_log_transition(K, c->sd, from, to);
    case ORC_PROP_ROTATION: return orc_pose_log_transition(K, 0, c->axis, c->sd, from, to);
    default: return orc_pose_log_transition(K, 1, c->axis, c->sd, from, to);
    }
}

/* MixtureProposal.logTransitionProbability: ln sum_i w_i exp(l_i), -inf when all are -inf */
static double mixture_log_transition(const orc_chain_desc *d, const double *from, const double *to)
{
    double l[64], mx = -INFINITY, wsum = 0;
    for (int i = 0; i < d->n_components; i++) { l[i] = comp_log_transition(d, &d->components[i], from, to); if (l[i] > mx) mx = l[i]; wsum += d->components[i].weight; }
    for (int i = 0; i < d->n_components; i++) if (isnan(l[i])) return NAN;
    if (mx == -INFINITY) return -INFINITY;
    double s = 0;
    for (int i = 0; i < d->n_components; i++) s += (d->components[i].weight / wsum) * exp(l[i] - mx);
    return log(s) + mx;
}

int orc_chain_run(const orc_chain_desc *d, const double *theta0, int n_steps, const double *u_comp,
                  const double *z, const double *u_acc, int32_t *comp, uint8_t *accepted, double *logv,
                  double *theta_log)
{
    int K = d->model->K, L = K + 10, n_acc = 0;
    double *cur = (double *)malloc(sizeof(double) * L), *prop = (double *)malloc(sizeof(double) * L);
    memcpy(cur, theta0, sizeof(double) * L);
    double cur_prior = d->use_prior ? orc_eval_prior(K, cur) : 0.0, cur_dist = eval_distance(d, cur);
    double wsum = 0; for (int i = 0; i < d->n_components; i++) wsum += d->components[i].weight;
    for (int s = 0; s < n_steps; s++) {
        /* MixtureProposal.propose: first component whose cumulative weight reaches r */
        int ci = d->n_components - 1; double acc = 0;
        for (int i = 0; i < d->n_components; i++) { acc += d->components[i].weight / wsum; if (acc >= u_comp[s]) { ci = i; break; } }
        const orc_component *c = &d->components[ci]; const double *zs = z + (size_t)s * K;
        memcpy(prop, cur, sizeof(double) * L);
        switch (c->kind) {
        case ORC_PROP_ICP: if (d->closed_form) orc_propose_closed_form(c->icp, cur, zs, prop); else orc_propose(c->icp, cur, zs, prop); break;
        case ORC_PROP_RANDOM_SHAPE: for (int j = 0; j < K; j++) prop[10 + j] = cur[10 + j] + c->sd * zs[j]; break;
        case ORC_PROP_ROTATION: prop[4 + c->axis] = cur[4 + c->axis] + c->sd * zs[0]; break;
        default: prop[1 + c->axis] = cur[1 + c->axis] + c->sd * zs[0]; break;
        }
        double prop_prior = d->use_prior ? orc_eval_prior(K, prop) : 0.0, prop_dist = eval_distance(d, prop);
        double t = mixture_log_transition(d, cur, prop) - mixture_log_transition(d, prop, cur);
        double a = (prop_prior + prop_dist) - (cur_prior + cur_dist) - t;
        int ok = (a > 0.0) || (u_acc[s] < exp(a));
        if (ok) { memcpy(cur, prop, sizeof(double) * L); cur_prior = prop_prior; cur_dist = prop_dist; n_acc++; }
        if (comp) comp[s] = ci;
        if (accepted) accepted[s] = (uint8_t)ok;
        if (logv) { logv[3 * s] = cur_prior + cur_dist; logv[3 * s + 1] = cur_prior; logv[3 * s + 2] = cur_dist; }
        if (theta_log) memcpy(theta_log + (size_t)s * L, cur, sizeof(double) * L);
    }
    free(cur); free(prop);
    return n_acc;
}

/* ============================================================================================ */
/* Philox4x32-10 (Salmon et al. 2011), bit-exact integer reference for the device generator      */
/* ============================================================================================ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

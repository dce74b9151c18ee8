/*
 * oracle/svc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the C-SVC dual solver that the
 * reference's hot path reaches through scikit-learn (reference call site
 * python/spark_sklearn/base_search.py:83-87 -> sklearn _fit_and_score -> SVC.fit ->
 * vendored libsvm).  The arithmetic lives in a third-party dependency that is not
 * under /root/reference: scikit-learn (reference pins >=0.18.1,<0.20; this image has
 * 1.9.0), file sklearn/svm/src/libsvm/svm.cpp (below "svm.cpp").  Each function cites
 * the svm.cpp lines it restates.  Pinning: tests/test_oracle.py checks this file
 * against sklearn.svm.SVC itself (dual_coef_, intercept_, n_iter_, predictions) on
 * iris and on seeded synthetic problems, and against the committed goldens in
 * tests/golden/.  The reference's own tests pin no numeric result on this path
 * (SURVEY.md §8c), so the pin is scikit-learn run in this image.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use this file.
 *
 * Numeric contract restated from svm.cpp (and mirrored by the CUDA kernels):
 *   - features are float64 (sklearn upcasts, svm/_base.py:218); dot products, kernel
 *     values, alpha, the gradient G and G_bar are float64;
 *   - Q_ij = (float)(y_i*y_j*K(x_i,x_j)) is rounded to float32 (Qfloat, svm.cpp:79,1446);
 *     QD_i = K(x_i,x_i) stays float64 (svm.cpp:1436-1437);
 *   - no fused multiply-add anywhere (x86-64 baseline build): compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define TAU_ 1e-12
enum { ST_LOWER = 0, ST_UPPER = 1, ST_FREE = 2 };
enum { KERNEL_LINEAR = 0, KERNEL_RBF = 1 };

typedef struct {
    int l;                 /* sub-problem size */
    const double *X;       /* [n][d] row-major float64 dataset */
    int d;
    const int *rows;       /* [l] dataset row of sub-problem element (original order) */
    const float *Kpre;     /* optional [n][ldk] precomputed float32 kernel values (unsigned by y) */
    long ldk;
    int kernel;
    double gamma;
    double *xsq;           /* [l] squared norms (rbf), original order */
    signed char *y0;       /* [l] labels in original order */
    float **qrow;          /* [l] lazily computed Q rows in ORIGINAL element order */
} qmat_t;

/* svm.cpp:416-423 Kernel::dot (dense) -- BLAS ddot there; any summation order is
 * absorbed by the float32 rounding of Q except on ~2^-29 of the entries. */
static double dot_rows(const qmat_t *q, int a, int b)
{
    const double *pa = q->X + (size_t)q->rows[a] * q->d, *pb = q->X + (size_t)q->rows[b] * q->d;
    double s = 0.0;
    for (int k = 0; k < q->d; k++) s += pa[k] * pb[k];
    return s;
}

/* svm.cpp:336-347 kernel_linear / kernel_rbf on original elements a, b */
static double kernel_val(const qmat_t *q, int a, int b)
{
    if (q->kernel == KERNEL_LINEAR) return dot_rows(q, a, b);
    return exp(-q->gamma * (q->xsq[a] + q->xsq[b] - 2 * dot_rows(q, a, b)));
}

/* svm.cpp:1439-1449 SVC_Q::get_Q: one signed, float32-rounded row (original element order) */
static const float *q_row(qmat_t *q, int a)
{
    if (!q->qrow[a]) {
        float *r = (float *)malloc(sizeof(float) * (size_t)q->l);
        if (q->Kpre) {
            const float *kr = q->Kpre + (size_t)q->rows[a] * q->ldk;
            for (int b = 0; b < q->l; b++)
                r[b] = (float)(q->y0[a] * q->y0[b]) * kr[q->rows[b]];   /* sign flip is exact */
        } else {
            for (int b = 0; b < q->l; b++)
                r[b] = (float)(q->y0[a] * q->y0[b] * kernel_val(q, a, b));
        }
        q->qrow[a] = r;
    }
    return q->qrow[a];
}

typedef struct {
    int l, active;
    int *orig;             /* active_set: position -> original element (svm.cpp:708-713) */
    signed char *y;
    char *st;
    double *alpha, *G, *Gbar, *QD, *Cv;
    double eps;
    int unshrink;
    qmat_t *q;
} smo_t;

#define QP(s, rowptr, pos) ((rowptr)[(s)->orig[pos]])   /* Q entry by POSITION through the permutation */

static void set_status(smo_t *s, int i)                 /* svm.cpp:593-600 */
{
    if (s->alpha[i] >= s->Cv[i]) s->st[i] = ST_UPPER;
    else if (s->alpha[i] <= 0) s->st[i] = ST_LOWER;
    else s->st[i] = ST_FREE;
}

static void swap_pos(smo_t *s, int i, int j)            /* svm.cpp:616-627 (p is constant -1) */
{
#define SW(T, a) do { T t_ = s->a[i]; s->a[i] = s->a[j]; s->a[j] = t_; } while (0)
    SW(int, orig); SW(signed char, y); SW(char, st); SW(double, alpha);
    SW(double, G); SW(double, Gbar); SW(double, QD); SW(double, Cv);
#undef SW
}

/* svm.cpp:629-668.  Both libsvm branches add, for each inactive k, the terms
 * alpha_f*Q_fk over the free active f in ascending position order, so one loop nest
 * reproduces either branch bit for bit (Q is symmetric bitwise). */
static void rebuild_gradient(smo_t *s)
{
    if (s->active == s->l) return;
    for (int k = s->active; k < s->l; k++) s->G[k] = s->Gbar[k] + (-1.0);
    for (int f = 0; f < s->active; f++)
        if (s->st[f] == ST_FREE) {
            const float *Qf = q_row(s->q, s->orig[f]);
            double af = s->alpha[f];
            for (int k = s->active; k < s->l; k++) s->G[k] += af * QP(s, Qf, k);
        }
}

/* svm.cpp:946-1047 second-order working-set selection; returns 1 when optimal */
static int pick_pair(smo_t *s, int *oi, int *oj)
{
    double Gmax = -INFINITY, Gmax2 = -INFINITY, best = INFINITY;
    int gi = -1, gj = -1;
    for (int t = 0; t < s->active; t++) {
        if (s->y[t] == +1) { if (s->st[t] != ST_UPPER && -s->G[t] >= Gmax) { Gmax = -s->G[t]; gi = t; } }
        else               { if (s->st[t] != ST_LOWER &&  s->G[t] >= Gmax) { Gmax =  s->G[t]; gi = t; } }
    }
    const float *Qi = gi >= 0 ? q_row(s->q, s->orig[gi]) : NULL;
    for (int j = 0; j < s->active; j++) {
        double gd, quad, od;
        if (s->y[j] == +1) {
            if (s->st[j] == ST_LOWER) continue;
            gd = Gmax + s->G[j];
            if (s->G[j] >= Gmax2) Gmax2 = s->G[j];
            if (!(gd > 0)) continue;
            quad = s->QD[gi] + s->QD[j] - 2.0 * s->y[gi] * QP(s, Qi, j);
        } else {
            if (s->st[j] == ST_UPPER) continue;
            gd = Gmax - s->G[j];
            if (-s->G[j] >= Gmax2) Gmax2 = -s->G[j];
            if (!(gd > 0)) continue;
            quad = s->QD[gi] + s->QD[j] + 2.0 * s->y[gi] * QP(s, Qi, j);
        }
        od = quad > 0 ? -(gd * gd) / quad : -(gd * gd) / TAU_;
        if (od <= best) { gj = j; best = od; }
    }
    if (Gmax + Gmax2 < s->eps || gj == -1) return 1;
    *oi = gi; *oj = gj;
    return 0;
}

static int shrinkable(const smo_t *s, int i, double g1, double g2)   /* svm.cpp:1049-1068 */
{
    if (s->st[i] == ST_UPPER) return s->y[i] == +1 ? (-s->G[i] > g1) : (-s->G[i] > g2);
    if (s->st[i] == ST_LOWER) return s->y[i] == +1 ? ( s->G[i] > g2) : ( s->G[i] > g1);
    return 0;
}

static void shrink(smo_t *s)                                          /* svm.cpp:1070-1129 */
{
    double g1 = -INFINITY, g2 = -INFINITY;
    for (int i = 0; i < s->active; i++) {
        if (s->y[i] == +1) {
            if (s->st[i] != ST_UPPER && -s->G[i] >= g1) g1 = -s->G[i];
            if (s->st[i] != ST_LOWER &&  s->G[i] >= g2) g2 =  s->G[i];
        } else {
            if (s->st[i] != ST_UPPER && -s->G[i] >= g2) g2 = -s->G[i];
            if (s->st[i] != ST_LOWER &&  s->G[i] >= g1) g1 =  s->G[i];
        }
    }
    if (!s->unshrink && g1 + g2 <= s->eps * 10) {
        s->unshrink = 1;
        rebuild_gradient(s);
        s->active = s->l;
    }
    for (int i = 0; i < s->active; i++)
        if (shrinkable(s, i, g1, g2)) {
            s->active--;
            while (s->active > i) {
                if (!shrinkable(s, s->active, g1, g2)) { swap_pos(s, i, s->active); break; }
                s->active--;
            }
        }
}

static double bias_term(const smo_t *s)                               /* svm.cpp:1131-1168 */
{
    int nfree = 0;
    double ub = INFINITY, lb = -INFINITY, sum = 0;
    for (int i = 0; i < s->active; i++) {
        double yG = s->y[i] * s->G[i];
        if (s->st[i] == ST_UPPER) { if (s->y[i] == -1) { if (yG < ub) ub = yG; } else { if (yG > lb) lb = yG; } }
        else if (s->st[i] == ST_LOWER) { if (s->y[i] == +1) { if (yG < ub) ub = yG; } else { if (yG > lb) lb = yG; } }
        else { nfree++; sum += yG; }
    }
    return nfree > 0 ? sum / nfree : (ub + lb) / 2;
}

/*
 * Solve one binary C-SVC sub-problem (svm.cpp:1600-1647 solve_c_svc + :670-944 Solver::Solve).
 *   X[n][d] float64; rows[l] = dataset rows in sub-problem order, first n_pos labelled +1
 *   (svm.cpp:2484-2510 puts class i first as +1, class j after as -1).
 *   Kpre (may be NULL): float32 kernel matrix [n][ldk] indexed by dataset row -- lets a test
 *   hand the solver the very matrix a GPU kernel produced.
 * Outputs: coef[l] = alpha_k*y_k in sub-problem order, *rho, *n_iter, *obj; returns 1 if max_iter hit.
 */
/* Optional working-set trace (development aid for cache studies): pairs of sub-problem row ids per iteration. */
static int *g_trace; static long g_trace_cap, g_trace_len;
void oracle_svc_set_trace(int *buf, long cap) { g_trace = buf; g_trace_cap = cap; g_trace_len = 0; }
long oracle_svc_trace_len(void) { return g_trace_len; }

int oracle_svc_solve(const double *X, int n, int d, const int *rows, int l, int n_pos,
                     int kernel, double gamma, double C, double eps, int shrinking, int max_iter,
                     const float *Kpre, long ldk,
                     double *coef, double *rho, int *n_iter, double *obj)
{
    (void)n;
    qmat_t q; smo_t s;
    memset(&q, 0, sizeof q); memset(&s, 0, sizeof s);
    q.l = l; q.X = X; q.d = d; q.rows = rows; q.Kpre = Kpre; q.ldk = ldk; q.kernel = kernel; q.gamma = gamma;
    q.xsq = (double *)malloc(sizeof(double) * l);
    q.y0 = (signed char *)malloc(l);
    q.qrow = (float **)calloc(l, sizeof(float *));
    s.l = s.active = l; s.eps = eps; s.q = &q;
    s.orig = (int *)malloc(sizeof(int) * l); s.y = (signed char *)malloc(l); s.st = (char *)malloc(l);
    s.alpha = (double *)calloc(l, sizeof(double)); s.G = (double *)malloc(sizeof(double) * l);
    s.Gbar = (double *)calloc(l, sizeof(double)); s.QD = (double *)malloc(sizeof(double) * l);
    s.Cv = (double *)malloc(sizeof(double) * l);
    for (int i = 0; i < l; i++) {
        q.y0[i] = s.y[i] = i < n_pos ? +1 : -1;
        s.orig[i] = i; s.Cv[i] = C; s.G[i] = -1.0;           /* p = -1, alpha = 0 (svm.cpp:1611-1626) */
        if (!Kpre || kernel == KERNEL_LINEAR) q.xsq[i] = dot_rows(&q, i, i);
        set_status(&s, i);
    }
    for (int i = 0; i < l; i++)                               /* QD (svm.cpp:1436-1437), float64 */
        s.QD[i] = kernel == KERNEL_RBF ? 1.0 /* exp(-g*(x+x-2x)) == exp(0) */ : q.xsq[i];

    int iter = 0, timed_out = 0;
    int counter = (l < 1000 ? l : 1000) + 1;
    for (;;) {
        if (max_iter != -1 && iter >= max_iter) { timed_out = 1; break; }
        if (--counter == 0) { counter = l < 1000 ? l : 1000; if (shrinking) shrink(&s); }
        int i, j;
        if (pick_pair(&s, &i, &j)) {
            rebuild_gradient(&s);
            s.active = l;
            if (pick_pair(&s, &i, &j)) break;
            counter = 1;
        }
        iter++;
        if (g_trace && g_trace_len + 2 <= g_trace_cap) { g_trace[g_trace_len++] = s.orig[i]; g_trace[g_trace_len++] = s.orig[j]; }
        const float *Qi = q_row(&q, s.orig[i]), *Qj = q_row(&q, s.orig[j]);
        double Ci = s.Cv[i], Cj = s.Cv[j], oai = s.alpha[i], oaj = s.alpha[j];
        if (s.y[i] != s.y[j]) {                                /* svm.cpp:772-815 */
            double quad = s.QD[i] + s.QD[j] + 2 * QP(&s, Qi, j);
            if (quad <= 0) quad = TAU_;
            double delta = (-s.G[i] - s.G[j]) / quad, diff = s.alpha[i] - s.alpha[j];
            s.alpha[i] += delta; s.alpha[j] += delta;
            if (diff > 0) { if (s.alpha[j] < 0) { s.alpha[j] = 0; s.alpha[i] = diff; } }
            else          { if (s.alpha[i] < 0) { s.alpha[i] = 0; s.alpha[j] = -diff; } }
            if (diff > Ci - Cj) { if (s.alpha[i] > Ci) { s.alpha[i] = Ci; s.alpha[j] = Ci - diff; } }
            else                { if (s.alpha[j] > Cj) { s.alpha[j] = Cj; s.alpha[i] = Cj + diff; } }
        } else {                                               /* svm.cpp:816-862 */
            double quad = s.QD[i] + s.QD[j] - 2 * QP(&s, Qi, j);
            if (quad <= 0) quad = TAU_;
            double delta = (s.G[i] - s.G[j]) / quad, sum = s.alpha[i] + s.alpha[j];
            s.alpha[i] -= delta; s.alpha[j] += delta;
            if (sum > Ci) { if (s.alpha[i] > Ci) { s.alpha[i] = Ci; s.alpha[j] = sum - Ci; } }
            else          { if (s.alpha[j] < 0)  { s.alpha[j] = 0;  s.alpha[i] = sum; } }
            if (sum > Cj) { if (s.alpha[j] > Cj) { s.alpha[j] = Cj; s.alpha[i] = sum - Cj; } }
            else          { if (s.alpha[i] < 0)  { s.alpha[i] = 0;  s.alpha[j] = sum; } }
        }
        double dai = s.alpha[i] - oai, daj = s.alpha[j] - oaj;  /* svm.cpp:866-872 */
        for (int k = 0; k < s.active; k++) s.G[k] += QP(&s, Qi, k) * dai + QP(&s, Qj, k) * daj;
        int ui = s.st[i] == ST_UPPER, uj = s.st[j] == ST_UPPER; /* svm.cpp:876-905 */
        set_status(&s, i); set_status(&s, j);
        if (ui != (s.st[i] == ST_UPPER)) {
            if (ui) for (int k = 0; k < l; k++) s.Gbar[k] -= Ci * QP(&s, Qi, k);
            else    for (int k = 0; k < l; k++) s.Gbar[k] += Ci * QP(&s, Qi, k);
        }
        if (uj != (s.st[j] == ST_UPPER)) {
            if (uj) for (int k = 0; k < l; k++) s.Gbar[k] -= Cj * QP(&s, Qj, k);
            else    for (int k = 0; k < l; k++) s.Gbar[k] += Cj * QP(&s, Qj, k);
        }
    }
    *rho = bias_term(&s);
    double v = 0;                                               /* svm.cpp:913-919 */
    for (int i = 0; i < l; i++) v += s.alpha[i] * (s.G[i] + (-1.0));
    *obj = v / 2;
    for (int i = 0; i < l; i++)                                 /* svm.cpp:922-925 + :1641-1642 */
        coef[s.orig[i]] = s.alpha[i] * s.y[i];
    *n_iter = iter;
    for (int i = 0; i < l; i++) free(q.qrow[i]);
    free(q.qrow); free(q.xsq); free(q.y0);
    free(s.orig); free(s.y); free(s.st); free(s.alpha); free(s.G); free(s.Gbar); free(s.QD); free(s.Cv);
    return timed_out;
}

/*
 * Decision values of one binary sub-model on arbitrary dataset rows
 * (svm.cpp:2821-2904 svm_predict_values, k_function :436-516: float64 kernel, NOT float32-rounded;
 * rbf distance formed from the difference vector).  out[t] = sum_k coef_k K(x_t, x_rows[k]) - rho.
 */
void oracle_svc_decision(const double *X, int d, const int *rows, int l, const double *coef, double rho,
                         int kernel, double gamma, const int *trows, int nt, double *out)
{
    for (int t = 0; t < nt; t++) {
        const double *xt = X + (size_t)trows[t] * d;
        double sum = 0;
        for (int k = 0; k < l; k++) {
            if (coef[k] == 0) continue;                          /* non-SVs are dropped from the model */
            const double *xs = X + (size_t)rows[k] * d;
            double kv = 0;
            if (kernel == KERNEL_LINEAR) { for (int c = 0; c < d; c++) kv += xt[c] * xs[c]; }
            else { for (int c = 0; c < d; c++) { double df = xt[c] - xs[c]; kv += df * df; } kv = exp(-gamma * kv); }
            sum += coef[k] * kv;
        }
        out[t] = sum - rho;
    }
}

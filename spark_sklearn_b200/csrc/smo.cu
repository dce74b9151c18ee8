// smo.cu -- batched C-SVC dual solver: one CTA per (candidate, fold, class-pair) sub-problem.
//
// Restates scikit-learn's libsvm Solver (svm.cpp:670-944 Solve, :946-1047 select_working_set,
// :1049-1129 do_shrinking, :629-668 reconstruct_gradient, :1131-1168 calculate_rho) as a
// block-parallel kernel that reproduces libsvm's ITERATE SEQUENCE, not just its fixed point:
//   * same WSS2 pair selection incl. tie-breaking ("last index wins" -> (value, index) reductions),
//   * same shrinking schedule and the same swap permutation (parallel two-pointer partition),
//   * same float64 arithmetic op-for-op: every multiply/add is individually rounded
//     (__dmul_rn/__dadd_rn: nvcc may not contract them into FMAs; libsvm's x86-64 build has none),
//   * Q entries are the float32-rounded kernel values (Qfloat), read from the K matrix built by
//     gram.cu; the sign y_i*y_j is applied on the fly (exact).
// Libsvm stops at a KKT gap of 1e-3; two solvers that merely agree on the optimum differ by ~1e-3 in
// decision values and flip test points near the margin, which is more than the 1e-4 budget on
// mean_test_score.  Following the same trajectory removes that.
//
// State per CTA, resident for the whole solve:
//   shared:  G[l] float64 gradient, col[l] int32 dataset row of each position (= active_set
//            composed with the row list), fl[l] uint8 (bit2: y=+1, bits0-1: alpha status)
//   registers: the float32 Q_i row of the positions a thread owns (position t = k*NT + tid)
//   global (L2-resident): alpha[l], Gbar[l]
// Per iteration: two dependent gathers of a K row (HBM/L2), two block-wide arg-reductions.
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr double TAU = 1e-12;
constexpr int ST_LOWER = 0, ST_UPPER = 1, ST_FREE = 2;
constexpr int YPOS = 4;
constexpr int SAFETY_MAX_ITER = 10000000;   // max_iter=-1 means "no limit" in libsvm; bound a runaway solve

struct ArgD {          // (value, packed index) pair; packed = (position << 3) | flags, -1 = none
    double v;
    int p;
};

__device__ __forceinline__ ArgD shfl_xor(ArgD a, int m)
{
    ArgD r;
    r.v = __shfl_xor_sync(0xffffffffu, a.v, m);
    r.p = __shfl_xor_sync(0xffffffffu, a.p, m);
    return r;
}
// "largest value, then largest index" == sequential scan with `>=` (svm.cpp:964-978)
__device__ __forceinline__ ArgD amax(ArgD a, ArgD b) { return (b.v > a.v || (b.v == a.v && b.p > a.p)) ? b : a; }

struct SelB {          // phase-B payload: objective decrease, packed index, G_j, signed Q_ij
    double od;
    int p;
    double g;
    float q;
};
__device__ __forceinline__ SelB shfl_xor(SelB a, int m)
{
    SelB r;
    r.od = __shfl_xor_sync(0xffffffffu, a.od, m);
    r.p = __shfl_xor_sync(0xffffffffu, a.p, m);
    r.g = __shfl_xor_sync(0xffffffffu, a.g, m);
    r.q = __shfl_xor_sync(0xffffffffu, a.q, m);
    return r;
}
// "smallest value, then largest index" == sequential scan with `<=` (svm.cpp:1003-1007)
__device__ __forceinline__ SelB bmin(SelB a, SelB b) { return (b.od < a.od || (b.od == a.od && b.p > a.p)) ? b : a; }

template <int NT>
struct Red {           // static shared scratch for the block reductions (double-buffered by phase)
    ArgD a[NT / 32];
    SelB b[NT / 32];
    double m[NT / 32];
    double m2[NT / 32];
    int cnt[NT / 32];
    int bcast[4];
};

template <int NT>
__device__ __forceinline__ double block_max(double v, double *buf)
{
#pragma unroll
    for (int m = 16; m; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    __syncthreads();                               // buf free (previous readers done)
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = -CUDART_INF;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) r = fmax(r, buf[w]);
    return r;
}

// exclusive block scan of a 0/1 predicate over threads; returns this thread's exclusive rank and the total
template <int NT>
__device__ __forceinline__ int block_rank(bool pred, int *cnt, int &total)
{
    const unsigned b = __ballot_sync(0xffffffffu, pred);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) cnt[w] = __popc(b);
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NT / 32; i++) {
        const int c = cnt[i];
        if (i < w) base += c;
        tot += c;
    }
    total = tot;
    return base + __popc(b & ((1u << lane) - 1u));
}

template <int NT, int KPT>
__global__ void __launch_bounds__(NT, 1)
smo_kernel(const SmoProblem *__restrict__ probs, const int *__restrict__ order, int lcap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Red<NT> red;

    const SmoProblem P = probs[order[blockIdx.x]];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = P.l;
    double *G = reinterpret_cast<double *>(smem_raw);
    int *col = reinterpret_cast<int *>(G + lcap);
    unsigned char *fl = reinterpret_cast<unsigned char *>(col + lcap);
    unsigned char *mark = fl + lcap;                 // shrink marks
    const float *__restrict__ K = P.K;
    const int64_t ldk = P.ldk;
    const double C = P.C, eps = P.eps;
    double *__restrict__ alpha = P.alpha;
    double *__restrict__ Gbar = P.Gbar;
    const bool use_gbar = P.shrinking != 0;

    unsigned long long t_start = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

    // ---- initial point: alpha = 0, G = p = -1 (svm.cpp:1611-1626, :716-736) ----
    for (int t = tid; t < l; t += NT) {
        G[t] = -1.0;
        col[t] = P.rows[t];
        fl[t] = (unsigned char)((t < P.n_pos ? YPOS : 0) | ST_LOWER);
        alpha[t] = 0.0;
        if (use_gbar) Gbar[t] = 0.0;
    }
    __syncthreads();

    int active = l, iter = 0, timed_out = 0;
    int counter = (l < 1000 ? l : 1000) + 1;
    bool unshrink = false;
    const int max_iter = P.max_iter == -1 ? SAFETY_MAX_ITER : P.max_iter;

    float qi[KPT];     // signed float32 Q_i row at the owned active positions

    // QD of a position (svm.cpp:1436-1437): rbf -> exp(0) == 1; linear -> float64 |x|^2
    auto QD = [&](int t) -> double { return P.qd ? P.qd[col[t]] : 1.0; };

    // ---------------- reconstruct_gradient (svm.cpp:629-668) ----------------
    auto rebuild_gradient = [&]() {
        if (active == l) return;
        // compact the free active positions (ascending) into scratch: [0..nf) = position
        int nf = 0;
        for (int base = 0; base < active; base += NT) {
            const int t = base + tid;
            const bool isf = t < active && (fl[t] & 3) == ST_FREE;
            int tot;
            const int r = block_rank<NT>(isf, red.cnt, tot);
            if (isf) P.scratch[nf + r] = t;
            nf += tot;
        }
        __syncthreads();
        double g[KPT];
        int ck[KPT];           // dataset row of the owned inactive positions, sign bit in bit 31; -1 = not owned
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            const bool in = t >= active && t < l;
            g[k] = in ? __dadd_rn(Gbar[t], -1.0) : 0.0;
            ck[k] = in ? (col[t] | ((fl[t] & YPOS) ? 0 : 0x40000000)) : -1;
        }
#pragma unroll 2
        for (int r = 0; r < nf; r++) {
            const int f = P.scratch[r];
            const float *__restrict__ Kf = K + (size_t)col[f] * ldk;
            const double af = __ldcg(alpha + f);
            const bool yf = (fl[f] & YPOS) != 0;
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                if (ck[k] >= 0) {
                    const float kv = __ldg(Kf + (ck[k] & 0x3fffffff));
                    const bool yt = (ck[k] & 0x40000000) == 0;
                    const float q = (yt == yf) ? kv : -kv;
                    g[k] = __dadd_rn(g[k], __dmul_rn(af, (double)q));      // G[j] += alpha_i * Q_i[j]
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            if (t >= active && t < l) G[t] = g[k];
        }
        __syncthreads();
    };

    // ---------------- select_working_set (svm.cpp:946-1047) ----------------
    // returns true when optimal; otherwise i/j hold packed (position<<3|flags), and gmax, gj, qij are set
    int pi = -1, pj = -1;
    double gmax = 0, g_j = 0;
    float q_ij = 0.f;
    auto select = [&]() -> bool {
        ArgD a;
        a.v = -CUDART_INF; a.p = -1;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            if (t < active) {
                const int f = fl[t];
                const double g = G[t];
                if (f & YPOS) { if ((f & 3) != ST_UPPER && -g >= a.v) { a.v = -g; a.p = (t << 3) | f; } }
                else          { if ((f & 3) != ST_LOWER &&  g >= a.v) { a.v =  g; a.p = (t << 3) | f; } }
            }
        }
#pragma unroll
        for (int m = 16; m; m >>= 1) a = amax(a, shfl_xor(a, m));
        if (lane == 0) red.a[warp] = a;
        __syncthreads();
        a = red.a[0];
#pragma unroll
        for (int w = 1; w < NT / 32; w++) a = amax(a, red.a[w]);
        pi = a.p; gmax = a.v;

        SelB b;
        b.od = CUDART_INF; b.p = -1; b.g = 0; b.q = 0.f;
        double gmax2 = -CUDART_INF;
        if (pi >= 0) {
            const int i = pi >> 3;
            const bool yi = (pi & YPOS) != 0;
            const double yi2 = yi ? 2.0 : -2.0;            // 2.0*y[i]
            const double QDi = QD(i);
            const float *__restrict__ Ki = K + (size_t)col[i] * ldk;
            float kv[KPT];
#pragma unroll
            for (int k = 0; k < KPT; k++) {                 // issue the whole gather first (MLP)
                const int t = k * NT + tid;
                kv[k] = t < active ? __ldg(Ki + col[t]) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int t = k * NT + tid;
                if (t < active) {
                    const int f = fl[t];
                    const bool yt = (f & YPOS) != 0;
                    const float q = (yt == yi) ? kv[k] : -kv[k];
                    qi[k] = q;
                    const double g = G[t];
                    double gd;
                    bool cand = false;
                    double quad = 0;
                    if (yt) {
                        if ((f & 3) != ST_LOWER) {
                            gd = __dadd_rn(gmax, g);
                            if (g >= gmax2) gmax2 = g;
                            if (gd > 0) { cand = true; quad = __dsub_rn(__dadd_rn(QDi, QD(t)), __dmul_rn(yi2, (double)q)); }
                        }
                    } else {
                        if ((f & 3) != ST_UPPER) {
                            gd = __dsub_rn(gmax, g);
                            if (-g >= gmax2) gmax2 = -g;
                            if (gd > 0) { cand = true; quad = __dadd_rn(__dadd_rn(QDi, QD(t)), __dmul_rn(yi2, (double)q)); }
                        }
                    }
                    if (cand) {
                        const double num = -__dmul_rn(gd, gd);
                        const double od = quad > 0 ? __ddiv_rn(num, quad) : __ddiv_rn(num, TAU);
                        if (od <= b.od) { b.od = od; b.p = (t << 3) | f; b.g = g; b.q = q; }
                    }
                }
            }
        } else {
            // no candidate i: Gmax = -inf; libsvm's second loop still forms Gmax2 (unused: -inf+x < eps)
        }
#pragma unroll
        for (int m = 16; m; m >>= 1) {
            b = bmin(b, shfl_xor(b, m));
            gmax2 = fmax(gmax2, __shfl_xor_sync(0xffffffffu, gmax2, m));
        }
        if (lane == 0) { red.b[warp] = b; red.m[warp] = gmax2; }
        __syncthreads();
        b = red.b[0]; gmax2 = red.m[0];
#pragma unroll
        for (int w = 1; w < NT / 32; w++) { b = bmin(b, red.b[w]); gmax2 = fmax(gmax2, red.m[w]); }
        pj = b.p; g_j = b.g; q_ij = b.q;
        if (pi < 0) return true;
        return (__dadd_rn(gmax, gmax2) < eps) || pj < 0;
    };

    // ---------------- do_shrinking (svm.cpp:1070-1129) ----------------
    auto do_shrink = [&]() {
        double g1 = -CUDART_INF, g2 = -CUDART_INF;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            if (t < active) {
                const int f = fl[t];
                const double g = G[t];
                if (f & YPOS) {
                    if ((f & 3) != ST_UPPER) g1 = fmax(g1, -g);
                    if ((f & 3) != ST_LOWER) g2 = fmax(g2, g);
                } else {
                    if ((f & 3) != ST_UPPER) g2 = fmax(g2, -g);
                    if ((f & 3) != ST_LOWER) g1 = fmax(g1, g);
                }
            }
        }
        g1 = block_max<NT>(g1, red.m);
        g2 = block_max<NT>(g2, red.m2);
        if (!unshrink && __dadd_rn(g1, g2) <= __dmul_rn(eps, 10.0)) {
            unshrink = true;
            rebuild_gradient();
            active = l;
        }
        // be_shrunk marks (svm.cpp:1049-1068)
        int keep_local = 0;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            if (t < active) {
                const int f = fl[t];
                const double g = G[t];
                bool s = false;
                if ((f & 3) == ST_UPPER) s = (f & YPOS) ? (-g > g1) : (-g > g2);
                else if ((f & 3) == ST_LOWER) s = (f & YPOS) ? (g > g2) : (g > g1);
                mark[t] = s ? 1 : 0;
                keep_local += s ? 0 : 1;
            }
        }
        // new active size = number of kept positions
#pragma unroll
        for (int m = 16; m; m >>= 1) keep_local += __shfl_xor_sync(0xffffffffu, keep_local, m);
        __syncthreads();
        if (lane == 0) red.cnt[warp] = keep_local;
        __syncthreads();
        int na = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) na += red.cnt[w];
        if (na == active) return;                         // uniform: nothing shrunk
        // Two-pointer partition == pair the k-th shrunk position below na (ascending) with the k-th
        // kept position at/above na (descending).  plist -> scratch[0..), qlist -> scratch[l..)
        int *plist = P.scratch, *qlist = P.scratch + l;
        int np = 0, nq = 0;
        for (int base = 0; base < na; base += NT) {
            const int t = base + tid;
            const bool pr = t < na && mark[t];
            int tot;
            const int r = block_rank<NT>(pr, red.cnt, tot);
            if (pr) plist[np + r] = t;
            np += tot;
        }
        for (int base = na; base < active; base += NT) {
            const int t = base + tid;
            const bool pr = t < active && !mark[t];
            int tot;
            const int r = block_rank<NT>(pr, red.cnt, tot);
            if (pr) qlist[nq + r] = t;                    // ascending; pair index = np-1-rank
            nq += tot;
        }
        __syncthreads();
        for (int r = tid; r < np; r += NT) {              // np == nq
            const int p = plist[r], q = qlist[np - 1 - r];
            const double gp = G[p]; G[p] = G[q]; G[q] = gp;
            const int cp = col[p]; col[p] = col[q]; col[q] = cp;
            const unsigned char fp = fl[p]; fl[p] = fl[q]; fl[q] = fp;
            const double ap = __ldcg(alpha + p), aq = __ldcg(alpha + q);
            alpha[p] = aq; alpha[q] = ap;
            const double bp = __ldcg(Gbar + p), bq = __ldcg(Gbar + q);
            Gbar[p] = bq; Gbar[q] = bp;
        }
        active = na;
        __syncthreads();
    };

    // ---------------- main loop (svm.cpp:742-907) ----------------
    for (;;) {
        if (iter >= max_iter) { timed_out = 1; break; }
        if (--counter == 0) {
            counter = l < 1000 ? l : 1000;
            if (P.shrinking) do_shrink();
        }
        if (select()) {
            rebuild_gradient();
            active = l;
            if (select()) break;
            counter = 1;
        }
        ++iter;

        const int i = pi >> 3, j = pj >> 3;
        const bool yi = (pi & YPOS) != 0, yj = (pj & YPOS) != 0;
        const double Gi = yi ? -gmax : gmax;               // gmax = -y_i G_i, negation exact
        const double Gj = g_j;
        const float *__restrict__ Ki = K + (size_t)col[i] * ldk;
        const float *__restrict__ Kj = K + (size_t)col[j] * ldk;
        float kvj[KPT];
#pragma unroll
        for (int k = 0; k < KPT; k++) {                     // issue the Q_j gather before the scalar update
            const int t = k * NT + tid;
            kvj[k] = t < active ? __ldg(Kj + col[t]) : 0.f;
        }
        const double old_ai = __ldcg(alpha + i), old_aj = __ldcg(alpha + j);
        __syncthreads();   // every thread holds the old alpha_i/alpha_j before their owners publish new ones
        const double QDi = QD(i), QDj = QD(j);
        double ai = old_ai, aj = old_aj;
        const double Qij = (double)q_ij;                     // signed Q_i[j]
        if (yi != yj) {                                      // svm.cpp:772-815
            double quad = __dadd_rn(__dadd_rn(QDi, QDj), __dmul_rn(2.0, Qij));
            if (quad <= 0) quad = TAU;
            const double delta = __ddiv_rn(__dsub_rn(-Gi, Gj), quad);
            const double diff = __dsub_rn(ai, aj);
            ai = __dadd_rn(ai, delta); aj = __dadd_rn(aj, delta);
            if (diff > 0) { if (aj < 0) { aj = 0; ai = diff; } }
            else          { if (ai < 0) { ai = 0; aj = -diff; } }
            if (diff > __dsub_rn(C, C)) { if (ai > C) { ai = C; aj = __dsub_rn(C, diff); } }
            else                        { if (aj > C) { aj = C; ai = __dadd_rn(C, diff); } }
        } else {                                             // svm.cpp:816-862
            double quad = __dsub_rn(__dadd_rn(QDi, QDj), __dmul_rn(2.0, Qij));
            if (quad <= 0) quad = TAU;
            const double delta = __ddiv_rn(__dsub_rn(Gi, Gj), quad);
            const double sum = __dadd_rn(ai, aj);
            ai = __dsub_rn(ai, delta); aj = __dadd_rn(aj, delta);
            if (sum > C) { if (ai > C) { ai = C; aj = __dsub_rn(sum, C); } }
            else         { if (aj < 0) { aj = 0; ai = sum; } }
            if (sum > C) { if (aj > C) { aj = C; ai = __dsub_rn(sum, C); } }
            else         { if (ai < 0) { ai = 0; aj = sum; } }
        }
        const double dai = __dsub_rn(ai, old_ai), daj = __dsub_rn(aj, old_aj);
        const int sti = ai >= C ? ST_UPPER : (ai <= 0 ? ST_LOWER : ST_FREE);
        const int stj = aj >= C ? ST_UPPER : (aj <= 0 ? ST_LOWER : ST_FREE);
        const bool ui = (pi & 3) == ST_UPPER, uj = (pj & 3) == ST_UPPER;
        const bool need_i = use_gbar && (ui != (sti == ST_UPPER));
        const bool need_j = use_gbar && (uj != (stj == ST_UPPER));

        // G update over the active set (svm.cpp:866-872); G_bar over all l when a bound status flips (:876-905)
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            if (t < l) {
                const bool act = t < active;
                if (act || need_i || need_j) {
                    const bool yt = (fl[t] & YPOS) != 0;
                    float q_i, q_j;
                    if (act) {
                        q_i = qi[k];
                        q_j = (yt == yj) ? kvj[k] : -kvj[k];
                        G[t] = __dadd_rn(G[t], __dadd_rn(__dmul_rn((double)q_i, dai), __dmul_rn((double)q_j, daj)));
                    } else {
                        const int c = col[t];
                        q_i = 0.f; q_j = 0.f;
                        if (need_i) { const float v = __ldg(Ki + c); q_i = (yt == yi) ? v : -v; }
                        if (need_j) { const float v = __ldg(Kj + c); q_j = (yt == yj) ? v : -v; }
                    }
                    if (need_i || need_j) {
                        double gb = __ldcg(Gbar + t);
                        if (need_i) { const double x = __dmul_rn(C, (double)q_i); gb = ui ? __dsub_rn(gb, x) : __dadd_rn(gb, x); }
                        if (need_j) { const double x = __dmul_rn(C, (double)q_j); gb = uj ? __dsub_rn(gb, x) : __dadd_rn(gb, x); }
                        Gbar[t] = gb;
                    }
                }
            }
        }
        // the owners of i and j publish alpha and status (read by everyone only after later barriers)
        if (tid == i % NT) { alpha[i] = ai; fl[i] = (unsigned char)((pi & YPOS) | sti); }
        if (tid == j % NT) { alpha[j] = aj; fl[j] = (unsigned char)((pj & YPOS) | stj); }
    }

    // ---------------- calculate_rho (svm.cpp:1131-1168): sequential float64 sum, libsvm's order ----
    __syncthreads();
    if (tid == 0) {
        int nfree = 0;
        double ub = CUDART_INF, lb = -CUDART_INF, sum = 0;
        for (int t = 0; t < active; t++) {
            const int f = fl[t];
            const double yG = (f & YPOS) ? G[t] : -G[t];
            if ((f & 3) == ST_UPPER) { if (!(f & YPOS)) ub = fmin(ub, yG); else lb = fmax(lb, yG); }
            else if ((f & 3) == ST_LOWER) { if (f & YPOS) ub = fmin(ub, yG); else lb = fmax(lb, yG); }
            else { ++nfree; sum = __dadd_rn(sum, yG); }
        }
        *P.out_rho = nfree > 0 ? __ddiv_rn(sum, (double)nfree) : __ddiv_rn(__dadd_rn(ub, lb), 2.0);
    }
    // coefficients alpha_k*y_k scattered by dataset row (svm.cpp:922-925, :1641-1642); SV counts
    int nsv = 0, nbsv = 0;
    for (int t = tid; t < l; t += NT) {
        const double a = __ldcg(alpha + t);
        P.coef[col[t]] = (fl[t] & YPOS) ? a : -a;
        nsv += a > 0;
        nbsv += a >= C;
    }
#pragma unroll
    for (int m = 16; m; m >>= 1) {
        nsv += __shfl_xor_sync(0xffffffffu, nsv, m);
        nbsv += __shfl_xor_sync(0xffffffffu, nbsv, m);
    }
    if (lane == 0) { red.cnt[warp] = nsv; red.a[warp].p = nbsv; }
    __syncthreads();
    if (tid == 0) {
        int s = 0, b = 0;
        for (int w = 0; w < NT / 32; w++) { s += red.cnt[w]; b += red.a[w].p; }
        P.out_info[0] = iter; P.out_info[1] = timed_out; P.out_info[2] = s; P.out_info[3] = b;
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        P.out_ns[0] = t_start; P.out_ns[1] = t_end;
    }
}

template <int NT, int KPT>
cudaError_t launch_one(const SmoProblem *probs, const int *order, int n_prob, int lmax, cudaStream_t st)
{
    const int lcap = (lmax + 15) & ~15;
    const size_t smem = (size_t)lcap * (8 + 4 + 1 + 1);
    cudaError_t e = cudaFuncSetAttribute(smo_kernel<NT, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smo_kernel<NT, KPT><<<n_prob, NT, smem, st>>>(probs, order, lcap);
    return cudaGetLastError();
}

}  // namespace

int smo_max_rows() { return 1024 * 16; }

cudaError_t launch_smo(const SmoProblem *d_probs, const int *d_order, int n_prob, int lmax, cudaStream_t st,
                       std::string *why)
{
    if (n_prob <= 0) return cudaSuccess;
    if (lmax <= 256) return launch_one<64, 4>(d_probs, d_order, n_prob, lmax, st);
    if (lmax <= 1024) return launch_one<256, 4>(d_probs, d_order, n_prob, lmax, st);
    if (lmax <= 4096) return launch_one<512, 8>(d_probs, d_order, n_prob, lmax, st);
    if (lmax <= 8192) return launch_one<512, 16>(d_probs, d_order, n_prob, lmax, st);
    if (lmax <= 16384) return launch_one<1024, 16>(d_probs, d_order, n_prob, lmax, st);
    if (why) *why = "SVC sub-problem larger than 16384 rows is not supported by the resident-state SMO kernel";
    return cudaErrorInvalidValue;
}

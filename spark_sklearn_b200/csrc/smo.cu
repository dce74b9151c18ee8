// smo.cu -- batched C-SVC dual solver: one CTA per (candidate, fold, class-pair) sub-problem.
//
// Restates scikit-learn's libsvm Solver (svm.cpp:670-944 Solve, :946-1047 select_working_set,
// :1049-1129 do_shrinking, :629-668 reconstruct_gradient, :1131-1168 calculate_rho) as a
// block-parallel kernel that reproduces libsvm's ITERATE SEQUENCE, not just its fixed point:
//   * same WSS2 pair selection incl. tie-breaking ("last index wins" == (value, index) arg-reductions),
//   * same shrinking schedule and the same swap permutation (parallel two-pointer partition),
//   * same float64 arithmetic op-for-op: every multiply/add/divide is individually rounded
//     (__dmul_rn/__dadd_rn/__ddiv_rn are never contracted; libsvm's x86-64 build has no FMA),
//   * Q entries are the float32-rounded kernel values (Qfloat) read from the K matrix of gram.cu.
// Libsvm stops at a KKT gap of 1e-3; two solvers that merely agree on the optimum differ by ~1e-3 in
// decision values and flip test points near the margin -- more than the 1e-4 budget on
// mean_test_score.  Following the same trajectory removes that.
//
// Sign-free state ("m-domain").  With m_t = -y_t*G_t and mbar_t = -y_t*Gbar_t (negation is exact, so
// every rounding is mirrored bit for bit) all per-element sign handling of libsvm disappears:
//   select i:  argmax { m_t : t in I_up }                      (svm.cpp:964-978, both y branches)
//   select j:  gd = Gmax - m_t, quad = (QD_i+QD_t) - 2*K_it     (svm.cpp:986-1037, both y branches;
//              -2*y_i*Q_it == -2*K_it for y_t=+1 and +2*y_i*Q_it == -2*K_it for y_t=-1)
//   update:    m_t += fl(fl(K_it*a) + fl(K_jt*b)),  a = -y_i*dalpha_i, b = -y_j*dalpha_j   (:866-872)
//   G_bar:     mbar_t += fl(c*K_it),  c = -/+ y_i*C_i                                       (:876-905)
// where K is the UNSIGNED float32 kernel row.  Set membership is two precomputed flag bits.
//
// Throughput design (B200: 64 FP64 lanes/SM, 4 issue slots/clk/SM -- this kernel is FP64-pipe and
// issue bound, then latency bound):
//   * the exact float64 division of WSS2 is evaluated only for elements that survive a 20-bit
//     reciprocal filter (rcp.approx.ftz.f64) with a provably safe band; the winner is still chosen
//     from exactly-rounded libsvm values, so the selection is bit-identical;
//   * float32->float64 widening of the K entries uses integer bit operations (ALU pipe) instead of
//     F2F (quarter-rate on the FP64 pipe);
//   * block arg-reductions run on REDUX.MAX over order-preserving 64-bit integer keys, not on
//     FP64 compares + shuffles;
//   * all solver state (m, mbar, alpha, column map, flags) of a sub-problem with <= 8192 rows is
//     resident in shared memory (27 B/row, 221 KB), the Q_i row of the owned positions in registers.
// Per iteration: two dependent gathers of a K row (HBM/L2), three CTA barriers.
#include "common.cuh"
#include <math_constants.h>
#include <cstdlib>

namespace {

constexpr double TAU = 1e-12;
constexpr int ST_LOWER = 0, ST_UPPER = 1, ST_FREE = 2;
constexpr int F_YPOS = 4, F_UP = 8, F_LOW = 16, F_MARK = 32;
constexpr int IDX_SHIFT = 5;                 // packed index = (position << 5) | (flags & 31)
constexpr int SAFETY_MAX_ITER = 10000000;    // max_iter=-1 is "no limit" in libsvm; bound a runaway solve
constexpr double BAND = 1.0 - 1.0 / 4096.0;  // filter band 2^-12 >> 2 * (rcp.approx error ~2^-20 + two roundings)

__device__ __forceinline__ int mkflags(bool ypos, int st)
{
    const bool up = ypos ? st != ST_UPPER : st != ST_LOWER;     // I_up  membership (svm.cpp:964-978)
    const bool low = ypos ? st != ST_LOWER : st != ST_UPPER;    // I_low membership (svm.cpp:986-1037)
    return st | (ypos ? F_YPOS : 0) | (up ? F_UP : 0) | (low ? F_LOW : 0);
}

// order-preserving map double -> uint64 (larger double <=> larger key) and back
__device__ __forceinline__ unsigned long long dkey(double x)
{
    const long long u = __double_as_longlong(x);
    return (unsigned long long)u ^ ((unsigned long long)(u >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k)
{
    return __longlong_as_double((long long)(k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull)));
}

// exact float -> double widening on the integer pipe; zero/denormal/inf/nan take the F2F path
__device__ __forceinline__ double f2d(float x)
{
    const unsigned u = __float_as_uint(x);
    const unsigned e = u & 0x7f800000u;
    if (__builtin_expect(e == 0u || e == 0x7f800000u, 0)) return (double)x;
    const unsigned hi = (u & 0x80000000u) | (((u & 0x7fffffffu) >> 3) + 0x38000000u);
    return __hiloint2double((int)hi, (int)(u << 29));
}

__device__ __forceinline__ double rcp_approx(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// K-row gather flavours: 0 = ld.global.nc (L1-allocating), 1 = ld.global.cg (L2 only), 2 = nc + L1::no_allocate
template <int LD>
__device__ __forceinline__ float load_k(const float *p)
{
    if constexpr (LD == 1) return __ldcg(p);
    else if constexpr (LD == 2) {
        float v;
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
        return v;
    } else return __ldg(p);
}

struct KArg { unsigned hi, lo; int idx; };

// warp arg-max over (64-bit key, index): largest key, ties -> largest index.  3 REDUX.
__device__ __forceinline__ KArg warp_argmax(unsigned hi, unsigned lo, int idx)
{
    KArg r;
    r.hi = __reduce_max_sync(0xffffffffu, hi);
    r.lo = __reduce_max_sync(0xffffffffu, hi == r.hi ? lo : 0u);
    r.idx = __reduce_max_sync(0xffffffffu, (hi == r.hi && lo == r.lo) ? idx : -1);
    return r;
}
__device__ __forceinline__ unsigned long long warp_keymax(unsigned long long k)
{
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

struct Red {            // static shared scratch; NW <= 32 warps
    unsigned a_hi[32], a_lo[32]; int a_idx[32];                 // phase A partials
    unsigned b_hi[32], b_lo[32]; int b_idx[32];                 // phase B partials
    unsigned m_hi[32], m_lo[32];                                // Gmax2 partials
    double pl_mg[32], pl_kv[32], pl_alpha[32];                  // phase B payload of each warp's winner
    double bc_d[4]; int bc_i[4];                                // scalars broadcast by warp 0
    double dm[32], dm2[32]; int cnt[32];                        // cold-path reductions
};

template <int NT>
__device__ __forceinline__ double block_max(double v, double *buf)
{
#pragma unroll
    for (int m = 16; m; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = -CUDART_INF;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) r = fmax(r, buf[w]);
    return r;
}

// exclusive block scan of a predicate over the threads (position order); returns rank and total
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

template <int NT, int KPT, bool SMEM_STATE, int LD, bool PROF>
__global__ void __launch_bounds__(NT, 1)
smo_kernel(const SmoProblem *__restrict__ probs, const int *__restrict__ order, int lcap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Red red;
    constexpr int NW = NT / 32;

    const SmoProblem P = probs[order[blockIdx.x]];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = P.l;
    // ---- resident state ----
    double *mG = reinterpret_cast<double *>(smem_raw);                      // m_t = -y_t G_t
    double *mGbar, *alpha;
    unsigned char *after;
    if constexpr (SMEM_STATE) {
        mGbar = mG + lcap; alpha = mGbar + lcap;
        after = reinterpret_cast<unsigned char *>(alpha + lcap);
    } else {
        mGbar = P.Gbar; alpha = P.alpha;
        after = reinterpret_cast<unsigned char *>(mG + lcap);
    }
    unsigned short *col = reinterpret_cast<unsigned short *>(after);       // dataset row of each position
    unsigned char *fl = reinterpret_cast<unsigned char *>(col + lcap);
    const float *__restrict__ K = P.K;
    const int64_t ldk = P.ldk;
    const double C = P.C, eps = P.eps;
    const bool use_gbar = P.shrinking != 0;
    const double *__restrict__ qd = P.qd;

    unsigned long long t_start = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

    // ---- initial point: alpha = 0, G = p = -1  =>  m_t = y_t (svm.cpp:1611-1626, :716-736) ----
    for (int t = tid; t < l; t += NT) {
        const bool yp = t < P.n_pos;
        mG[t] = yp ? 1.0 : -1.0;
        col[t] = (unsigned short)P.rows[t];
        fl[t] = (unsigned char)mkflags(yp, ST_LOWER);
        alpha[t] = 0.0;
        if (use_gbar) mGbar[t] = 0.0;
    }
    __syncthreads();

    int active = l, iter = 0, timed_out = 0;
    int counter = (l < 1000 ? l : 1000) + 1;
    bool unshrink = false;
    const int max_iter = P.max_iter == -1 ? SAFETY_MAX_ITER : P.max_iter;

    long long prof[6] = {0, 0, 0, 0, 0, 0};
    long long tprev = PROF ? clock64() : 0;
    auto tick = [&](int slot) {
        if constexpr (PROF) {
            const long long now = clock64();
            prof[slot] += now - tprev;
            tprev = now;
        }
    };
    double qi[KPT];      // unsigned K_i row (widened) at the owned active positions t = k*NT + tid

    auto QD = [&](int t) -> double { return qd ? qd[col[t]] : 1.0; };   // svm.cpp:1436-1437

    // ---------------- reconstruct_gradient (svm.cpp:629-668), m-domain ----------------
    // G_k = (Gbar_k + p_k) + sum_{free f, ascending} alpha_f Q_fk   <=>
    // m_k = (mbar_k + y_k) + sum_f fl((-y_f alpha_f) K_fk), same roundings mirrored.
    auto rebuild_gradient = [&]() {
        if (active == l) return;
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
        int ck[KPT];
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            const bool in = t >= active && t < l;
            g[k] = in ? __dadd_rn(mGbar[t], (fl[t] & F_YPOS) ? 1.0 : -1.0) : 0.0;
            ck[k] = in ? (int)col[t] : -1;
        }
#pragma unroll 2
        for (int r = 0; r < nf; r++) {
            const int f = P.scratch[r];
            const float *__restrict__ Kf = K + (size_t)col[f] * ldk;
            const double af = (fl[f] & F_YPOS) ? -alpha[f] : alpha[f];      // -y_f alpha_f
#pragma unroll
            for (int k = 0; k < KPT; k++)
                if (ck[k] >= 0) g[k] = __dadd_rn(g[k], __dmul_rn(af, f2d(load_k<LD>(Kf + ck[k]))));
        }
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            if (t >= active && t < l) mG[t] = g[k];
        }
        __syncthreads();
    };

    // ---------------- select_working_set (svm.cpp:946-1047) ----------------
    int pi = -1, pj = -1;            // packed (position << 5 | flags)
    double gmax = 0, mg_j = 0, k_ij = 0, alpha_i = 0, alpha_j = 0;
    auto select = [&]() -> bool {
        // ---- phase A: i = argmax m_t over I_up ----
        {
            double best = -CUDART_INF;
            int bidx = -1;
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int t = k * NT + tid;
                if (t < active) {
                    const int f = fl[t];
                    const double m = mG[t];
                    if ((f & F_UP) && m >= best) { best = m; bidx = (t << IDX_SHIFT) | (f & 31); }
                }
            }
            const unsigned long long key = dkey(best);
            const KArg w = warp_argmax((unsigned)(key >> 32), (unsigned)key, bidx);
            if (lane == 0) { red.a_hi[warp] = w.hi; red.a_lo[warp] = w.lo; red.a_idx[warp] = w.idx; }
            tick(0);
            __syncthreads();                                                      // barrier 1
            tick(1);
            const bool v = lane < NW;
            const KArg a = warp_argmax(v ? red.a_hi[lane] : 0u, v ? red.a_lo[lane] : 0u, v ? red.a_idx[lane] : -1);
            pi = a.idx;
            gmax = dkey_inv(((unsigned long long)a.hi << 32) | a.lo);
        }
        // ---- phase B: j = argmin -(gd^2)/quad over I_low with gd > 0; Gmax2 = max -m_t over I_low ----
        double bestn = -CUDART_INF;          // best NEGATED objective change (positive), exact value
        int bidx = -1;
        double b_mg = 0, b_kv = 0;
        double mgmin = CUDART_INF;
        if (pi >= 0) {
            const int i = pi >> IDX_SHIFT;
            alpha_i = alpha[i];
            const double QDi = QD(i);
            const float *__restrict__ Ki = K + (size_t)col[i] * ldk;
            float kv[KPT];
#pragma unroll
            for (int k = 0; k < KPT; k++) {                                  // issue the whole gather first
                const int t = k * NT + tid;
                kv[k] = t < active ? load_k<LD>(Ki + col[t]) : 0.f;
            }
            double runmax = -CUDART_INF, thr = -CUDART_INF;   // running max of approx NEGATED od, and its band
            unsigned mask = 0;
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int t = k * NT + tid;
                if (t < active) {
                    const double dq = f2d(kv[k]);
                    qi[k] = dq;
                    if (fl[t] & F_LOW) {
                        const double m = mG[t];
                        mgmin = fmin(mgmin, m);
                        const double gd = __dsub_rn(gmax, m);
                        if (gd > 0) {
                            const double quad = __dsub_rn(__dadd_rn(QDi, QD(t)), __dmul_rn(2.0, dq));
                            const double g2 = __dmul_rn(gd, gd);
                            const double ap = quad > 0 ? g2 * rcp_approx(quad) : g2 * 1e12;
                            if (ap >= thr) {                                  // may still be the exact winner
                                mask |= 1u << k;
                                if (ap > runmax) { runmax = ap; thr = ap * BAND; }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < KPT; k++) {                                   // exact libsvm values for the survivors
                if (mask & (1u << k)) {
                    const int t = k * NT + tid;
                    const double m = mG[t];
                    const double gd = __dsub_rn(gmax, m);
                    const double quad = __dsub_rn(__dadd_rn(QDi, QD(t)), __dmul_rn(2.0, qi[k]));
                    const double g2 = __dmul_rn(gd, gd);
                    const double nod = quad > 0 ? __ddiv_rn(g2, quad) : __ddiv_rn(g2, TAU);   // == -obj_diff
                    if (nod >= bestn) { bestn = nod; bidx = (t << IDX_SHIFT) | (fl[t] & 31); b_mg = m; b_kv = qi[k]; }
                }
            }
        }
        {
            const unsigned long long key = dkey(bestn);
            const KArg w = warp_argmax((unsigned)(key >> 32), (unsigned)key, bidx);
            const unsigned long long km = warp_keymax(dkey(-mgmin));
            if (bidx >= 0 && bidx == w.idx) {                                  // this lane owns the warp's winner
                red.pl_mg[warp] = b_mg; red.pl_kv[warp] = b_kv; red.pl_alpha[warp] = alpha[bidx >> IDX_SHIFT];
            }
            if (lane == 0) {
                red.b_hi[warp] = w.hi; red.b_lo[warp] = w.lo; red.b_idx[warp] = w.idx;
                red.m_hi[warp] = (unsigned)(km >> 32); red.m_lo[warp] = (unsigned)km;
            }
            tick(2);
            __syncthreads();                                                      // barrier 2
            tick(3);
            const bool v = lane < NW;
            const KArg b = warp_argmax(v ? red.b_hi[lane] : 0u, v ? red.b_lo[lane] : 0u, v ? red.b_idx[lane] : -1);
            const unsigned long long km2 =
                warp_keymax(v ? (((unsigned long long)red.m_hi[lane] << 32) | red.m_lo[lane]) : 0ull);
            pj = b.idx;
            if (pi < 0) return true;
            const double gmax2 = dkey_inv(km2);
            if (pj >= 0) {
                const int wj = ((pj >> IDX_SHIFT) % NT) >> 5;
                mg_j = red.pl_mg[wj]; k_ij = red.pl_kv[wj]; alpha_j = red.pl_alpha[wj];
            }
            return (__dadd_rn(gmax, gmax2) < eps) || pj < 0;
        }
    };

    // ---------------- do_shrinking (svm.cpp:1070-1129), m-domain ----------------
    // Gmax1 = max{m_t : I_up}, Gmax2 = max{-m_t : I_low}; be_shrunk(t) = (!up && m_t > Gmax1) || (!low && -m_t > Gmax2)
    auto do_shrink = [&]() {
        double g1 = -CUDART_INF, g2 = -CUDART_INF;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            if (t < active) {
                const int f = fl[t];
                const double m = mG[t];
                if (f & F_UP) g1 = fmax(g1, m);
                if (f & F_LOW) g2 = fmax(g2, -m);
            }
        }
        g1 = block_max<NT>(g1, red.dm);
        g2 = block_max<NT>(g2, red.dm2);
        if (!unshrink && __dadd_rn(g1, g2) <= __dmul_rn(eps, 10.0)) {
            unshrink = true;
            rebuild_gradient();
            active = l;
        }
        int keep_local = 0;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            if (t < active) {
                const int f = fl[t] & 31;
                const double m = mG[t];
                const bool s = (!(f & F_UP) && m > g1) || (!(f & F_LOW) && -m > g2);
                fl[t] = (unsigned char)(f | (s ? F_MARK : 0));
                keep_local += s ? 0 : 1;
            }
        }
#pragma unroll
        for (int m = 16; m; m >>= 1) keep_local += __shfl_xor_sync(0xffffffffu, keep_local, m);
        __syncthreads();
        if (lane == 0) red.cnt[warp] = keep_local;
        __syncthreads();
        int na = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) na += red.cnt[w];
        if (na != active) {
            // Two-pointer partition == pair the k-th marked position below na (ascending) with the k-th
            // unmarked position at/above na (descending).
            int *plist = P.scratch, *qlist = P.scratch + l;
            int np = 0, nq = 0;
            for (int base = 0; base < na; base += NT) {
                const int t = base + tid;
                const bool pr = t < na && (fl[t] & F_MARK);
                int tot;
                const int r = block_rank<NT>(pr, red.cnt, tot);
                if (pr) plist[np + r] = t;
                np += tot;
            }
            for (int base = na; base < active; base += NT) {
                const int t = base + tid;
                const bool pr = t < active && !(fl[t] & F_MARK);
                int tot;
                const int r = block_rank<NT>(pr, red.cnt, tot);
                if (pr) qlist[nq + r] = t;
                nq += tot;
            }
            __syncthreads();
            for (int r = tid; r < np; r += NT) {                 // np == nq; disjoint pairs
                const int p = plist[r], q = qlist[np - 1 - r];
                const double gp = mG[p]; mG[p] = mG[q]; mG[q] = gp;
                const unsigned short cp = col[p]; col[p] = col[q]; col[q] = cp;
                const unsigned char fp = fl[p]; fl[p] = fl[q]; fl[q] = fp;
                const double ap = alpha[p], aq = alpha[q]; alpha[p] = aq; alpha[q] = ap;
                const double bp = mGbar[p], bq = mGbar[q]; mGbar[p] = bq; mGbar[q] = bp;
            }
            active = na;
        }
        __syncthreads();
        for (int t = tid; t < l; t += NT) fl[t] &= 31;         // clear marks (owner-mapped)
        __syncthreads();
    };

    // ---------------- main loop (svm.cpp:742-907) ----------------
    for (;;) {
        if (iter >= max_iter) { timed_out = 1; break; }
        if (--counter == 0) {
            counter = l < 1000 ? l : 1000;
            if (P.shrinking) do_shrink();
            if constexpr (PROF) tprev = clock64();
        }
        if (select()) {
            rebuild_gradient();
            active = l;
            if (select()) break;
            counter = 1;
        }
        ++iter;

        const int i = pi >> IDX_SHIFT, j = pj >> IDX_SHIFT;
        const float *__restrict__ Ki = K + (size_t)col[i] * ldk;
        const float *__restrict__ Kj = K + (size_t)col[j] * ldk;
        float kvj[KPT];
#pragma unroll
        for (int k = 0; k < KPT; k++) {                          // issue the Q_j gather before the scalar update
            const int t = k * NT + tid;
            kvj[k] = t < active ? load_k<LD>(Kj + col[t]) : 0.f;
        }
        if (warp == 0) {                                         // analytic 2-variable update, once per CTA
            const bool yi = (pi & F_YPOS) != 0, yj = (pj & F_YPOS) != 0;
            const double Gi = yi ? -gmax : gmax;                 // G = -y m (exact)
            const double Gj = yj ? -mg_j : mg_j;
            const double QDi = QD(i), QDj = QD(j);
            const double Qij = (yi == yj) ? k_ij : -k_ij;        // signed Q_i[j]
            double ai = alpha_i, aj = alpha_j;
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
            if (lane == 0) {
                const double dai = __dsub_rn(ai, alpha_i), daj = __dsub_rn(aj, alpha_j);
                red.bc_d[0] = yi ? -dai : dai;                   // a = -y_i dalpha_i
                red.bc_d[1] = yj ? -daj : daj;                   // b = -y_j dalpha_j
                red.bc_d[2] = ai; red.bc_d[3] = aj;
                red.bc_i[0] = ai >= C ? ST_UPPER : (ai <= 0 ? ST_LOWER : ST_FREE);
                red.bc_i[1] = aj >= C ? ST_UPPER : (aj <= 0 ? ST_LOWER : ST_FREE);
            }
        }
        tick(4);
        __syncthreads();                                                          // barrier 3
        tick(1);
        const double a = red.bc_d[0], b = red.bc_d[1];
        const int sti = red.bc_i[0], stj = red.bc_i[1];
        const bool need_i = use_gbar && (((pi & 3) == ST_UPPER) != (sti == ST_UPPER));
        const bool need_j = use_gbar && (((pj & 3) == ST_UPPER) != (stj == ST_UPPER));

        // m update over the active set (svm.cpp:866-872)
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = k * NT + tid;
            if (t < active)
                mG[t] = __dadd_rn(mG[t], __dadd_rn(__dmul_rn(qi[k], a), __dmul_rn(f2d(kvj[k]), b)));
        }
        // G_bar over all l when a bound status flips (svm.cpp:876-905): i first, then j
        if (need_i || need_j) {
            // Gbar -= C Q_i (was upper) / += C Q_i (became upper)  <=>  mbar += fl(c K_i), c = +/- y_i C
            const double ci = (((pi & 3) == ST_UPPER) == ((pi & F_YPOS) != 0)) ? C : -C;
            const double cj = (((pj & 3) == ST_UPPER) == ((pj & F_YPOS) != 0)) ? C : -C;
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int t = k * NT + tid;
                if (t < l) {
                    const bool act = t < active;
                    double gb = mGbar[t];
                    if (need_i) gb = __dadd_rn(gb, __dmul_rn(ci, act ? qi[k] : f2d(load_k<LD>(Ki + col[t]))));
                    if (need_j) gb = __dadd_rn(gb, __dmul_rn(cj, act ? f2d(kvj[k]) : f2d(load_k<LD>(Kj + col[t]))));
                    mGbar[t] = gb;
                }
            }
        }
        // the owners of i and j publish alpha and status (owner-only data until the next barriers)
        if (tid == i % NT) { alpha[i] = red.bc_d[2]; fl[i] = (unsigned char)mkflags((pi & F_YPOS) != 0, sti); }
        if (tid == j % NT) { alpha[j] = red.bc_d[3]; fl[j] = (unsigned char)mkflags((pj & F_YPOS) != 0, stj); }
        tick(5);
    }

    // ---------------- calculate_rho (svm.cpp:1131-1168): sequential float64 sum in libsvm's order ----
    __syncthreads();
    if (tid == 0) {
        int nfree = 0;
        double ub = CUDART_INF, lb = -CUDART_INF, sum = 0;
        for (int t = 0; t < active; t++) {
            const int f = fl[t];
            const double yG = -mG[t];                            // y_t G_t
            if ((f & 3) == ST_UPPER) { if (!(f & F_YPOS)) ub = fmin(ub, yG); else lb = fmax(lb, yG); }
            else if ((f & 3) == ST_LOWER) { if (f & F_YPOS) ub = fmin(ub, yG); else lb = fmax(lb, yG); }
            else { ++nfree; sum = __dadd_rn(sum, yG); }
        }
        *P.out_rho = nfree > 0 ? __ddiv_rn(sum, (double)nfree) : __ddiv_rn(__dadd_rn(ub, lb), 2.0);
    }
    // coefficients alpha_k*y_k scattered by dataset row (svm.cpp:922-925, :1641-1642); SV counts
    int nsv = 0, nbsv = 0;
    for (int t = tid; t < l; t += NT) {
        const double av = alpha[t];
        P.coef[col[t]] = (fl[t] & F_YPOS) ? av : -av;
        nsv += av > 0;
        nbsv += av >= C;
    }
#pragma unroll
    for (int m = 16; m; m >>= 1) {
        nsv += __shfl_xor_sync(0xffffffffu, nsv, m);
        nbsv += __shfl_xor_sync(0xffffffffu, nbsv, m);
    }
    if (lane == 0) { red.cnt[warp] = nsv; red.a_idx[warp] = nbsv; }
    __syncthreads();
    if (tid == 0) {
        int s = 0, bs = 0;
        for (int w = 0; w < NW; w++) { s += red.cnt[w]; bs += red.a_idx[w]; }
        P.out_info[0] = iter; P.out_info[1] = timed_out; P.out_info[2] = s; P.out_info[3] = bs;
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        P.out_ns[0] = t_start; P.out_ns[1] = t_end;
        if constexpr (PROF)
            for (int q = 0; q < 6; q++) P.out_ns[2 + q] = (unsigned long long)prof[q];
    }
}

template <int NT, int KPT, bool SMEM_STATE, int LD, bool PROF>
cudaError_t launch_one(const SmoProblem *probs, const int *order, int n_prob, int lmax, cudaStream_t st)
{
    const int lcap = (lmax + 15) & ~15;
    const size_t smem = (size_t)lcap * (SMEM_STATE ? (8 + 8 + 8 + 2 + 1) : (8 + 2 + 1));
    auto kern = smo_kernel<NT, KPT, SMEM_STATE, LD, PROF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<n_prob, NT, smem, st>>>(probs, order, lcap);
    return cudaGetLastError();
}

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

}  // namespace

int smo_max_rows() { return 1024 * 16; }

cudaError_t launch_smo(const SmoProblem *d_probs, const int *d_order, int n_prob, int lmax, cudaStream_t st,
                       std::string *why)
{
    if (n_prob <= 0) return cudaSuccess;
    if (lmax <= 512) return launch_one<128, 4, true, 0, false>(d_probs, d_order, n_prob, lmax, st);
    if (lmax <= 2048) return launch_one<256, 8, true, 0, false>(d_probs, d_order, n_prob, lmax, st);
    if (lmax <= 4096) return launch_one<512, 8, true, 0, false>(d_probs, d_order, n_prob, lmax, st);
    if (lmax <= 8192) {
        // tuning switches (development only): gather flavour / state placement / phase profile
        const int ld = env_int("B200GS_SMO_LD", 0), state = env_int("B200GS_SMO_STATE", 1), prof = env_int("B200GS_SMO_PROF", 0);
        if (prof) return state ? launch_one<1024, 8, true, 0, true>(d_probs, d_order, n_prob, lmax, st)
                               : launch_one<1024, 8, false, 0, true>(d_probs, d_order, n_prob, lmax, st);
        if (!state) return ld == 1 ? launch_one<1024, 8, false, 1, false>(d_probs, d_order, n_prob, lmax, st)
                                   : launch_one<1024, 8, false, 0, false>(d_probs, d_order, n_prob, lmax, st);
        if (ld == 1) return launch_one<1024, 8, true, 1, false>(d_probs, d_order, n_prob, lmax, st);
        if (ld == 2) return launch_one<1024, 8, true, 2, false>(d_probs, d_order, n_prob, lmax, st);
        return launch_one<1024, 8, true, 0, false>(d_probs, d_order, n_prob, lmax, st);
    }
    if (lmax <= 16384) return launch_one<1024, 16, false, 0, false>(d_probs, d_order, n_prob, lmax, st);
    if (why) *why = "SVC sub-problem larger than 16384 rows is not supported by the resident-state SMO kernel";
    return cudaErrorInvalidValue;
}

// smo_cluster.cu -- the SMO solver of smo.cu with ONE sub-problem spread over a thread-block cluster of CL CTAs.
//
// Why: a search is bounded by its longest sub-problem (config 2: 68,716 dependent iterations), and one iteration of
// the single-CTA kernel is issue-bound on the rows a CTA owns.  Here CL CTAs (CL SMs) each own 1/CL of the rows --
// state in their own shared memory -- and only the two arg-reductions of working-set selection cross CTAs: every CTA
// reduces its rows and publishes one 64-byte record into every peer's shared memory with st.async (DSMEM stores that
// complete_tx on the RECEIVER's mbarrier); a CTA waits only on its own mbarrier, so the all-gather costs one DSMEM
// latency instead of a cluster-wide barrier.  All CTAs then combine the CL records identically and redundantly run
// the scalar two-variable update.  The two K rows of an iteration are brought into EVERY CTA's shared memory by
// multicast bulk copies (cp.async.bulk ... .multicast::cluster): each CTA requests 1/CL of the row from L2/HBM and the
// copy engine delivers it to all CL shared memories (block-cyclic ownership over shrunk positions means every CTA
// needs columns from all over the row; per-thread gathers made every SM pull the whole row through its own L1 miss
// path).  Arithmetic, tie-breaking, shrinking schedule and swap permutation are those of smo.cu (bit-identical
// results; see that file's header for the restatement of libsvm svm.cpp:629-1168).
//
// Ordering argument for the barrier-free exchange (records double-buffered by parity, one row buffer): a CTA can
// finish exchange n only after every peer SENT record n, and a peer sends record n only after all its warps passed the
// CTA barrier that precedes the send -- i.e. after they finished reading exchange n-1's records and the row buffer
// contents of the previous phase.  So a peer's record n+1 never overwrites a slot still being read (other parity), and a
// multicast row write issued after exchange n never lands in a buffer a peer still gathers from.
//
// Row ownership is block-cyclic: position t lives in CTA (t / NT) % CL, local slot (t / (NT*CL)) * NT + t % NT, so the
// active prefix [0, active) stays balanced over the CTAs as shrinking proceeds.
#include "smo_common.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace {

using namespace smo;

constexpr int XW = 16;                       // words per exchange record

template <int CL>
struct __align__(16) Xch {                   // double-buffered all-gather slots: [parity][source rank][word]
    unsigned w[2][CL][XW];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned cta)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();                                    // a lost signal must not hang the GPU
    }
}

__device__ __forceinline__ unsigned lo32(double x) { return (unsigned)__double_as_longlong(x); }
__device__ __forceinline__ unsigned hi32(double x) { return (unsigned)((unsigned long long)__double_as_longlong(x) >> 32); }
__device__ __forceinline__ double mk64(unsigned lo, unsigned hi) { return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo)); }

// ROWBUF: K rows by multicast bulk copy into shared memory (row length rowcap floats); otherwise per-thread gathers.
template <int NT, int KPT, int CL, bool FAST, bool ROWBUF, bool PROF>
__global__ void __launch_bounds__(NT, (NT >= 1024 ? 1 : (NT >= 512 ? 2 : 3)))
smo_cluster_kernel(const SmoProblem *__restrict__ probs, const int *__restrict__ order, int rowcap, int fmode)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Red red;
    __shared__ Xch<CL> xch;
    __shared__ __align__(8) unsigned long long xbar[2];                      // one mbarrier per exchange parity
    __shared__ __align__(8) unsigned long long rowbar;                       // row-buffer fill
    __shared__ int chunk_cnt[2][KPT * CL];                                   // shrink / rebuild chunk counts (all chunks)
    constexpr int NW = NT / 32;
    constexpr int LCAP = NT * KPT;                                          // rows owned by this CTA

    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const SmoProblem *__restrict__ Pp = probs + order[blockIdx.x / CL];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = Pp->l;
    double *const mG = reinterpret_cast<double *>(smem_raw);
    double *const mGbar = mG + LCAP;
    double *const alpha = mG + 2 * LCAP;
    unsigned short *const col = reinterpret_cast<unsigned short *>(mG + 3 * LCAP);
    unsigned char *const fl = reinterpret_cast<unsigned char *>(col + LCAP);
    float *const rowbuf = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(fl + LCAP) + 127) & ~(uintptr_t)127);
    const float *__restrict__ const K = Pp->K;
    const int64_t ldk = Pp->ldk;
    const double eps = Pp->eps;
    const bool use_gbar = Pp->shrinking != 0;
    const double *__restrict__ const qd = FAST ? nullptr : Pp->qd;
    int *const scratch = Pp->scratch;                                        // global: plist/qlist or free-row list
    double *const gscratch = Pp->Gbar;                                       // global doubles (unused Gbar workspace)

    // position <-> (owner, slot)
    auto gpos = [&](int k) -> int { return (k * CL + (int)rank) * NT + tid; };
    auto owner_of = [&](int t) -> unsigned { return (unsigned)((t / NT) % CL); };
    auto slot_of = [&](int t) -> int { return (t / (NT * CL)) * NT + (t % NT); };

    unsigned long long t_start = 0;
    if (tid == 0 && rank == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

    // ---- initial point: alpha = 0, G = -1  =>  m_t = y_t ----
    {
        const int n_pos = Pp->n_pos;
        const int *__restrict__ rows = Pp->rows;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = gpos(k), s = k * NT + tid;
            if (t < l) {
                const bool yp = t < n_pos;
                mG[s] = yp ? 1.0 : -1.0;
                col[s] = (unsigned short)rows[t];
                fl[s] = (unsigned char)mkflags(yp, ST_LOWER);
                alpha[s] = 0.0;
                mGbar[s] = 0.0;
            }
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&xbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&xbar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&rowbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();                                                         // every CTA's barriers exist before any remote signal

    long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = PROF ? clock64() : 0;
    auto tick = [&](int slot) {
        if constexpr (PROF) {
            const long long now = clock64();
            prof[slot] += now - tprev;
            tprev = now;
        }
    };

    int active = l, iter = 0, timed_out = 0;
    int counter = (l < 1000 ? l : 1000) + 1;
    bool unshrink = false;
    const int max_iter = Pp->max_iter == -1 ? SAFETY_MAX_ITER : Pp->max_iter;
    int par = 0;                                                            // exchange parity

    double qi[KPT];

    auto QDc = [&](int c) -> double {                                       // by dataset row
        if constexpr (FAST) return 1.0;
        else return qd ? qd[c] : 1.0;
    };
    auto widen = [&](float x) -> double {
        if constexpr (FAST) {
            const unsigned u = __float_as_uint(x);
            return __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29));
        } else return f2d(x);
    };

    // publish this CTA's record into every CTA of the cluster, then make all records visible
    // (v must be warp-uniform in warp 0: it is built from REDUX results and broadcast shared-memory reads)
    unsigned xphase = 0;                                                    // bit p: parity the next wait on xbar[p] uses
    auto exchange = [&](const unsigned (&v)[XW]) -> const unsigned (*)[XW] {
        const unsigned bar = smem_u32(&xbar[par]);
        if (warp == 0) {
            if (lane == 0) mbar_expect_tx(bar, CL * XW * 4);
            if (lane < CL * 4) {                                             // lane -> (destination CTA, 16-byte chunk)
                const int c = lane & 3;
                const unsigned dst = (unsigned)lane >> 2;
                const unsigned a0 = c == 0 ? v[0] : (c == 1 ? v[4] : (c == 2 ? v[8] : v[12]));
                const unsigned a1 = c == 0 ? v[1] : (c == 1 ? v[5] : (c == 2 ? v[9] : v[13]));
                const unsigned a2 = c == 0 ? v[2] : (c == 1 ? v[6] : (c == 2 ? v[10] : v[14]));
                const unsigned a3 = c == 0 ? v[3] : (c == 1 ? v[7] : (c == 2 ? v[11] : v[15]));
                const unsigned raddr = mapa_u32(smem_u32(&xch.w[par][rank][c * 4]), dst);
                const unsigned rbar = mapa_u32(bar, dst);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                             ::"r"(raddr), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(rbar) : "memory");
            }
        }
        mbar_wait(bar, (xphase >> par) & 1u);
        xphase ^= 1u << par;
        const unsigned (*r)[XW] = xch.w[par];
        par ^= 1;
        return r;
    };

    // K row `r` (dataset row) into every CTA's row buffer: this CTA requests slice `rank`, multicast to all CL CTAs
    unsigned rowphase = 0;
    auto fetch_row = [&](int r) {
        if constexpr (ROWBUF) {
            const unsigned bar = smem_u32(&rowbar);
            if (fmode == 1) {                                                // experiment: unicast, every CTA pulls the whole row
                if (tid == 0) {
                    mbar_expect_tx(bar, (unsigned)rowcap * 4u);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(rowbuf)), "l"(K + (size_t)r * ldk), "r"((unsigned)rowcap * 4u), "r"(bar) : "memory");
                }
            } else if (fmode == 2) {                                         // experiment: slice split over 4 copies
                if (tid == 0) mbar_expect_tx(bar, (unsigned)rowcap * 4u);
                if (tid < 4) {
                    const unsigned slice = (unsigned)rowcap / CL, sub = slice / 4;
                    const unsigned off = rank * slice + tid * sub;
                    const unsigned short mask = (unsigned short)((1u << CL) - 1u);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                                 ::"r"(smem_u32(rowbuf + off)), "l"(K + (size_t)r * ldk + off), "r"(sub * 4u), "r"(bar), "h"(mask) : "memory");
                }
            } else if (tid == 0) {
                mbar_expect_tx(bar, (unsigned)rowcap * 4u);                  // all CL slices land here
                const unsigned slice = (unsigned)rowcap / CL;               // rowcap % 32 == 0: 16-byte multiples
                const unsigned off = rank * slice;
                const unsigned short mask = (unsigned short)((1u << CL) - 1u);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                             ::"r"(smem_u32(rowbuf + off)), "l"(K + (size_t)r * ldk + off), "r"(slice * 4u), "r"(bar), "h"(mask) : "memory");
            }
        }
    };
    auto wait_row = [&]() {
        if constexpr (ROWBUF) {
            mbar_wait(smem_u32(&rowbar), rowphase);
            rowphase ^= 1u;
        }
    };

    // ---------------- local scan (normally fused into the update loop) ----------------
    double la = -CUDART_INF, lm = CUDART_INF;
    int la_idx = -1;
    auto local_scan = [&]() {
        la = -CUDART_INF; lm = CUDART_INF; la_idx = -1;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = gpos(k), s = k * NT + tid;
            if (t < active) {
                const int f = fl[s];
                const double m = mG[s];
                if ((f & F_UP) && m >= la) { la = m; la_idx = (t << IDX_SHIFT) | f; }
                if (f & F_LOW) lm = fmin(lm, m);
            }
        }
    };

    // ---------------- reconstruct_gradient (svm.cpp:629-668) over the cluster ----------------
    auto rebuild_gradient = [&]() {
        if (active == l) return;
        // free active rows in position order: per-chunk compaction + all-gather of the chunk counts
        int rankF[KPT];
        bool isF[KPT];
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = gpos(k), s = k * NT + tid;
            isF[k] = t < active && (fl[s] & 3) == ST_FREE;
            int tot;
            rankF[k] = block_rank<NT>(isF[k], red.cnt, tot);
            if (tid == 0) chunk_cnt[0][k * CL + rank] = tot;
        }
        __syncthreads();
        {
            unsigned v[XW] = {0};
#pragma unroll
            for (int k = 0; k < KPT; k++) v[k] = (unsigned)chunk_cnt[0][k * CL + rank];
            const unsigned (*r)[XW] = exchange(v);
            if (tid < KPT * CL) { const int c = tid; chunk_cnt[0][c] = (int)r[c % CL][c / CL]; }
            __syncthreads();
        }
        int nf = 0;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int c = k * CL + (int)rank;
            int base = 0;
            for (int cc = 0; cc < c; cc++) base += chunk_cnt[0][cc];
            if (isF[k]) {
                const int s = k * NT + tid;
                scratch[base + rankF[k]] = (int)col[s];
                gscratch[base + rankF[k]] = (fl[s] & F_YPOS) ? -alpha[s] : alpha[s];      // -y_f alpha_f
            }
        }
        for (int cc = 0; cc < KPT * CL; cc++) nf += chunk_cnt[0][cc];
        __threadfence();
        cluster.sync();
        double g[KPT];
        int ck[KPT];
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = gpos(k), s = k * NT + tid;
            const bool in = t >= active && t < l;
            g[k] = in ? __dadd_rn(mGbar[s], (fl[s] & F_YPOS) ? 1.0 : -1.0) : 0.0;
            ck[k] = in ? (int)col[s] : -1;
        }
#pragma unroll 2
        for (int r = 0; r < nf; r++) {
            const float *__restrict__ Kf = K + (size_t)__ldcg(scratch + r) * ldk;
            const double af = __ldcg(gscratch + r);
#pragma unroll
            for (int k = 0; k < KPT; k++)
                if (ck[k] >= 0) g[k] = __dadd_rn(g[k], __dmul_rn(af, widen(__ldg(Kf + ck[k]))));
        }
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = gpos(k), s = k * NT + tid;
            if (t >= active && t < l) mG[s] = g[k];
        }
        cluster.sync();                                                      // scratch is reused by the next phase
    };

    // ---------------- select_working_set (svm.cpp:946-1047) ----------------
    int pi = -1, pj = -1, col_i = 0, col_j = 0;
    double gmax = 0, mg_j = 0, k_ij = 0, alpha_i = 0, alpha_j = 0;
    auto select = [&]() -> bool {
        double gmax2;
        {   // phase A: local arg-max, then all-gather over the cluster
            const unsigned long long key = dkey(la);
            const KArg w = warp_argmax((unsigned)(key >> 32), (unsigned)key, la_idx);
            const unsigned long long km = warp_keymax(dkey(-lm));
            if (lane == 0) {
                red.a_hi[warp] = w.hi; red.a_lo[warp] = w.lo; red.a_idx[warp] = w.idx;
                red.m_hi[warp] = (unsigned)(km >> 32); red.m_lo[warp] = (unsigned)km;
            }
            __syncthreads();
            const bool v = lane < NW;
            const KArg a = warp_argmax(v ? red.a_hi[lane] : 0u, v ? red.a_lo[lane] : 0u, v ? red.a_idx[lane] : -1);
            const unsigned long long km2 =
                warp_keymax(v ? (((unsigned long long)red.m_hi[lane] << 32) | red.m_lo[lane]) : 0ull);
            unsigned rec[XW] = {0};
            rec[0] = a.hi; rec[1] = a.lo; rec[2] = (unsigned)a.idx; rec[3] = (unsigned)(km2 >> 32); rec[4] = (unsigned)km2;
            if (a.idx >= 0) {
                const int s = slot_of(a.idx >> IDX_SHIFT);
                const double ai = alpha[s];
                rec[5] = lo32(ai); rec[6] = hi32(ai); rec[7] = (unsigned)col[s];
            }
            tick(0);
            const unsigned (*r)[XW] = exchange(rec);
            tick(1);
            int best = 0;
            unsigned long long kmax = ((unsigned long long)r[0][3] << 32) | r[0][4];
#pragma unroll
            for (int c = 1; c < CL; c++) {
                const unsigned long long kb = ((unsigned long long)r[best][0] << 32) | r[best][1];
                const unsigned long long kc = ((unsigned long long)r[c][0] << 32) | r[c][1];
                if (kc > kb || (kc == kb && (int)r[c][2] > (int)r[best][2])) best = c;
                const unsigned long long k2 = ((unsigned long long)r[c][3] << 32) | r[c][4];
                kmax = k2 > kmax ? k2 : kmax;
            }
            pi = (int)r[best][2];
            gmax = dkey_inv(((unsigned long long)r[best][0] << 32) | r[best][1]);
            gmax2 = dkey_inv(kmax);
            alpha_i = mk64(r[best][5], r[best][6]);
            col_i = (int)r[best][7];
        }
        if (pi < 0 || __dadd_rn(gmax, gmax2) < eps) return true;
        // phase B
        const double QDi = QDc(col_i);
        const float *__restrict__ Ki = K + (size_t)col_i * ldk;
        {
            float kv[KPT];
            if constexpr (ROWBUF) { fetch_row(col_i); wait_row(); }
            tick(2);
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int t = gpos(k), s = k * NT + tid;
                if constexpr (ROWBUF) kv[k] = t < active ? rowbuf[col[s]] : 0.f;
                else kv[k] = t < active ? __ldg(Ki + col[s]) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < KPT; k++) qi[k] = widen(kv[k]);
        }
        // Approximate gd^2/quad is tracked by the HIGH WORD of the (non-negative) double only: 32-bit compares and
        // moves instead of 64-bit ones.  High words order like the doubles to 2^-20; the band test below (514 units
        // >= 2^-12 relative) sends every near-tie to the exact libsvm quotients, so the choice stays bit-identical.
        unsigned b1k = 0u, b2k = 0u;                    // keys of the best and second-best candidate (0 = none)
        int k1 = -1;
        double m1 = 0, q1 = 0;
        int idx1 = -1;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = gpos(k), s = k * NT + tid;
            if (t < active) {
                const int f = fl[s];
                const double m = mG[s];
                const double gd = __dsub_rn(gmax, m);
                if ((f & F_LOW) && gd > 0) {
                    const double quad = FAST ? __dsub_rn(2.0, __dadd_rn(qi[k], qi[k]))
                                             : __dsub_rn(__dadd_rn(QDi, QDc(col[s])), __dmul_rn(2.0, qi[k]));
                    const double g2 = __dmul_rn(gd, gd);
                    const double ap = quad > 0 ? g2 * rcp_approx(quad) : g2 * 1e12;
                    const unsigned key = (unsigned)__double2hiint(ap) + 1u;      // +1: a valid candidate is never 0
                    const bool gt = key > b1k;
                    b2k = gt ? b1k : max(b2k, key);
                    b1k = gt ? key : b1k;
                    k1 = gt ? k : k1;
                }
            }
        }
        if (k1 >= 0) {
            const int t1 = (k1 * CL + (int)rank) * NT + tid, s1 = k1 * NT + tid;
            idx1 = (t1 << IDX_SHIFT) | fl[s1];
            m1 = mG[s1];
#pragma unroll
            for (int k = 0; k < KPT; k++) q1 = k == k1 ? qi[k] : q1;
        }
        unsigned top1k, top2k;
        {
            const unsigned w1 = __reduce_max_sync(0xffffffffu, b1k);
            const int widx = __reduce_max_sync(0xffffffffu, (b1k == w1) ? idx1 : -1);
            const unsigned w2 = __reduce_max_sync(0xffffffffu, (idx1 == widx) ? b2k : b1k);
            if (idx1 >= 0 && idx1 == widx) {
                const int s = slot_of(idx1 >> IDX_SHIFT);
                red.pl_mg[warp] = m1; red.pl_kv[warp] = q1; red.pl_alpha[warp] = alpha[s];
                red.cnt[warp] = (int)col[s];
            }
            if (lane == 0) { red.b_hi[warp] = w1; red.b_idx[warp] = widx; red.t_hi[warp] = w2; }
            __syncthreads();
            const bool v = lane < NW;
            const unsigned bk = v ? red.b_hi[lane] : 0u;
            const int bi = v ? red.b_idx[lane] : -1;
            const unsigned c1 = __reduce_max_sync(0xffffffffu, bk);
            const int cidx = __reduce_max_sync(0xffffffffu, (bk == c1) ? bi : -1);
            const unsigned c2 = __reduce_max_sync(0xffffffffu, (v && bi == cidx) ? red.t_hi[lane] : bk);
            unsigned rec[XW] = {0};
            rec[0] = c1; rec[2] = (unsigned)cidx; rec[3] = c2;
            if (cidx >= 0) {
                const int wj = ((cidx >> IDX_SHIFT) % NT) >> 5;
                const double a = red.pl_mg[wj], bb = red.pl_kv[wj], c = red.pl_alpha[wj];
                rec[5] = lo32(a); rec[6] = hi32(a); rec[7] = lo32(bb); rec[8] = hi32(bb); rec[9] = lo32(c); rec[10] = hi32(c);
                rec[11] = (unsigned)red.cnt[wj];
            }
            tick(3);
            const unsigned (*r)[XW] = exchange(rec);
            tick(4);
            int best = 0;
#pragma unroll
            for (int c = 1; c < CL; c++)
                if (r[c][0] > r[best][0] || (r[c][0] == r[best][0] && (int)r[c][2] > (int)r[best][2])) best = c;
            top2k = 0u;
#pragma unroll
            for (int c = 0; c < CL; c++) { const unsigned kc = c == best ? r[c][3] : r[c][0]; top2k = kc > top2k ? kc : top2k; }
            top1k = r[best][0];
            pj = (int)r[best][2];
            if (pj < 0) return true;
            mg_j = mk64(r[best][5], r[best][6]); k_ij = mk64(r[best][7], r[best][8]); alpha_j = mk64(r[best][9], r[best][10]);
            col_j = (int)r[best][11];
        }
        if (top1k - top2k <= 514u) {
            // exact tie-break among the elements of the band (rare)
            __syncthreads();                                 // slower warps may still be reading red.pl_* / red.cnt above
            const unsigned thrk = top1k > 514u ? top1k - 514u : 1u;
            double bestn = -CUDART_INF;
            int bidx = -1;
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int t = gpos(k), s = k * NT + tid;
                if (t < active) {
                    const int f = fl[s];
                    const double m = mG[s];
                    const double gd = __dsub_rn(gmax, m);
                    if ((f & F_LOW) && gd > 0) {
                        const double quad = __dsub_rn(__dadd_rn(QDi, QDc(col[s])), __dmul_rn(2.0, qi[k]));
                        const double g2 = __dmul_rn(gd, gd);
                        const double ap = quad > 0 ? g2 * rcp_approx(quad) : g2 * 1e12;
                        if ((unsigned)__double2hiint(ap) + 1u >= thrk) {
                            const double nod = quad > 0 ? __ddiv_rn(g2, quad) : __ddiv_rn(g2, TAU);
                            if (nod >= bestn) { bestn = nod; bidx = (t << IDX_SHIFT) | f; m1 = m; q1 = qi[k]; }
                        }
                    }
                }
            }
            const unsigned long long key = dkey(bestn);
            const KArg w = warp_argmax((unsigned)(key >> 32), (unsigned)key, bidx);
            if (bidx >= 0 && bidx == w.idx) {
                const int s = slot_of(bidx >> IDX_SHIFT);
                red.pl_mg[warp] = m1; red.pl_kv[warp] = q1; red.pl_alpha[warp] = alpha[s];
                red.cnt[warp] = (int)col[s];
            }
            if (lane == 0) { red.x_hi[warp] = w.hi; red.x_lo[warp] = w.lo; red.x_idx[warp] = w.idx; }
            __syncthreads();
            const bool v = lane < NW;
            const KArg b = warp_argmax(v ? red.x_hi[lane] : 0u, v ? red.x_lo[lane] : 0u, v ? red.x_idx[lane] : -1);
            unsigned rec[XW] = {0};
            rec[0] = b.hi; rec[1] = b.lo; rec[2] = (unsigned)b.idx;
            if (b.idx >= 0) {
                const int wj = ((b.idx >> IDX_SHIFT) % NT) >> 5;
                const double a = red.pl_mg[wj], bb = red.pl_kv[wj], c = red.pl_alpha[wj];
                rec[5] = lo32(a); rec[6] = hi32(a); rec[7] = lo32(bb); rec[8] = hi32(bb); rec[9] = lo32(c); rec[10] = hi32(c);
                rec[11] = (unsigned)red.cnt[wj];
            }
            const unsigned (*r)[XW] = exchange(rec);
            int best = 0;
#pragma unroll
            for (int c = 1; c < CL; c++) {
                const unsigned long long kb = ((unsigned long long)r[best][0] << 32) | r[best][1];
                const unsigned long long kc = ((unsigned long long)r[c][0] << 32) | r[c][1];
                if (kc > kb || (kc == kb && (int)r[c][2] > (int)r[best][2])) best = c;
            }
            pj = (int)r[best][2];
            mg_j = mk64(r[best][5], r[best][6]); k_ij = mk64(r[best][7], r[best][8]); alpha_j = mk64(r[best][9], r[best][10]);
            col_j = (int)r[best][11];
        }
        return false;
    };

    // ---------------- do_shrinking (svm.cpp:1070-1129) over the cluster ----------------
    auto do_shrink = [&]() {
        double g1 = -CUDART_INF, g2 = -CUDART_INF;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = gpos(k), s = k * NT + tid;
            if (t < active) {
                const int f = fl[s];
                const double m = mG[s];
                if (f & F_UP) g1 = fmax(g1, m);
                if (f & F_LOW) g2 = fmax(g2, -m);
            }
        }
        g1 = block_max<NT>(g1, red.dm);
        g2 = block_max<NT>(g2, red.dm2);
        {
            unsigned v[XW] = {0};
            v[0] = lo32(g1); v[1] = hi32(g1); v[2] = lo32(g2); v[3] = hi32(g2);
            const unsigned (*r)[XW] = exchange(v);
#pragma unroll
            for (int c = 0; c < CL; c++) { g1 = fmax(g1, mk64(r[c][0], r[c][1])); g2 = fmax(g2, mk64(r[c][2], r[c][3])); }
        }
        if (!unshrink && __dadd_rn(g1, g2) <= __dmul_rn(eps, 10.0)) {
            unshrink = true;
            rebuild_gradient();
            active = l;
        }
        int keep_local = 0;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = gpos(k), s = k * NT + tid;
            if (t < active) {
                const int f = fl[s] & 31;
                const double m = mG[s];
                const bool sh = (!(f & F_UP) && m > g1) || (!(f & F_LOW) && -m > g2);
                fl[s] = (unsigned char)(f | (sh ? F_MARK : 0));
                keep_local += sh ? 0 : 1;
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
        {
            unsigned v[XW] = {0};
            v[0] = (unsigned)na;
            const unsigned (*r)[XW] = exchange(v);
            na = 0;
#pragma unroll
            for (int c = 0; c < CL; c++) na += (int)r[c][0];
        }
        if (na != active) {
            // marked positions below na (ascending) pair with unmarked positions at/above na (descending)
            int rankP[KPT], rankQ[KPT];
            bool isP[KPT], isQ[KPT];
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int t = gpos(k), s = k * NT + tid;
                const bool mk = t < active && (fl[s] & F_MARK);
                isP[k] = t < na && mk;
                isQ[k] = t >= na && t < active && !mk;
                int totP, totQ;
                rankP[k] = block_rank<NT>(isP[k], red.cnt, totP);
                rankQ[k] = block_rank<NT>(isQ[k], red.cnt, totQ);
                if (tid == 0) { chunk_cnt[0][k * CL + rank] = totP; chunk_cnt[1][k * CL + rank] = totQ; }
            }
            __syncthreads();
            {
                unsigned v[XW] = {0};
#pragma unroll
                for (int k = 0; k < KPT; k++) { v[k] = (unsigned)chunk_cnt[0][k * CL + rank]; v[KPT + k] = (unsigned)chunk_cnt[1][k * CL + rank]; }
                const unsigned (*r)[XW] = exchange(v);
                if (tid < KPT * CL) {
                    const int c = tid;
                    chunk_cnt[0][c] = (int)r[c % CL][c / CL];
                    chunk_cnt[1][c] = (int)r[c % CL][KPT + c / CL];
                }
                __syncthreads();
            }
            int *plist = scratch, *qlist = scratch + l;
            int np = 0;
            for (int cc = 0; cc < KPT * CL; cc++) np += chunk_cnt[0][cc];
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int c = k * CL + (int)rank;
                int bp = 0, bq = 0;
                for (int cc = 0; cc < c; cc++) { bp += chunk_cnt[0][cc]; bq += chunk_cnt[1][cc]; }
                if (isP[k]) plist[bp + rankP[k]] = gpos(k);
                if (isQ[k]) qlist[bq + rankQ[k]] = gpos(k);
            }
            __threadfence();
            cluster.sync();
            // swaps through distributed shared memory: pair r handled by one thread of the cluster
            for (int r = (int)rank * NT + tid; r < np; r += CL * NT) {
                const int p = __ldcg(plist + r), q = __ldcg(qlist + np - 1 - r);
                const unsigned op = owner_of(p), oq = owner_of(q);
                const int sp = slot_of(p), sq = slot_of(q);
                double *mGp = cluster.map_shared_rank(mG, op), *mGq = cluster.map_shared_rank(mG, oq);
                double *bp_ = cluster.map_shared_rank(mGbar, op), *bq_ = cluster.map_shared_rank(mGbar, oq);
                double *ap_ = cluster.map_shared_rank(alpha, op), *aq_ = cluster.map_shared_rank(alpha, oq);
                unsigned short *cp_ = cluster.map_shared_rank(col, op), *cq_ = cluster.map_shared_rank(col, oq);
                unsigned char *fp_ = cluster.map_shared_rank(fl, op), *fq_ = cluster.map_shared_rank(fl, oq);
                const double g_p = mGp[sp], g_q = mGq[sq]; mGp[sp] = g_q; mGq[sq] = g_p;
                const double b_p = bp_[sp], b_q = bq_[sq]; bp_[sp] = b_q; bq_[sq] = b_p;
                const double a_p = ap_[sp], a_q = aq_[sq]; ap_[sp] = a_q; aq_[sq] = a_p;
                const unsigned short c_p = cp_[sp], c_q = cq_[sq]; cp_[sp] = c_q; cq_[sq] = c_p;
                const unsigned char f_p = fp_[sp], f_q = fq_[sq]; fp_[sp] = f_q; fq_[sq] = f_p;
            }
            cluster.sync();
            active = na;
        }
#pragma unroll
        for (int k = 0; k < KPT; k++) { const int t = gpos(k), s = k * NT + tid; if (t < l) fl[s] &= 31; }
        __syncthreads();
    };

    // ---------------- main loop (svm.cpp:742-907) ----------------
    bool scan_valid = false;
    for (;;) {
        if (iter >= max_iter) { timed_out = 1; break; }
        if (--counter == 0) {
            counter = l < 1000 ? l : 1000;
            if (use_gbar) { do_shrink(); scan_valid = false; }
            if constexpr (PROF) tprev = clock64();
        }
        if (!scan_valid) local_scan();
        if (select()) {
            rebuild_gradient();
            active = l;
            __syncthreads();
            local_scan();
            if (select()) break;
            counter = 1;
        }
        ++iter;

        const int i = pi >> IDX_SHIFT, j = pj >> IDX_SHIFT;
        const float *__restrict__ Kj = K + (size_t)col_j * ldk;
        float kvj[KPT];
        if constexpr (ROWBUF) fetch_row(col_j);                  // every CTA is past exchange 2: row i is in registers everywhere
        else {
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int t = gpos(k), s = k * NT + tid;
                kvj[k] = t < active ? __ldg(Kj + col[s]) : 0.f;
            }
        }
        if (warp == 0) {                                         // every CTA runs the identical scalar update
            const double C = Pp->C;
            const bool yi = (pi & F_YPOS) != 0, yj = (pj & F_YPOS) != 0;
            const double Gi = yi ? -gmax : gmax;
            const double Gj = yj ? -mg_j : mg_j;
            const double QDi = QDc(col_i), QDj = QDc(col_j);
            const double Qij = (yi == yj) ? k_ij : -k_ij;
            double ai = alpha_i, aj = alpha_j;
            if (yi != yj) {
                double quad = __dadd_rn(__dadd_rn(QDi, QDj), __dmul_rn(2.0, Qij));
                if (quad <= 0) quad = TAU;
                const double delta = __ddiv_rn(__dsub_rn(-Gi, Gj), quad);
                const double diff = __dsub_rn(ai, aj);
                ai = __dadd_rn(ai, delta); aj = __dadd_rn(aj, delta);
                if (diff > 0) { if (aj < 0) { aj = 0; ai = diff; } }
                else          { if (ai < 0) { ai = 0; aj = -diff; } }
                if (diff > __dsub_rn(C, C)) { if (ai > C) { ai = C; aj = __dsub_rn(C, diff); } }
                else                        { if (aj > C) { aj = C; ai = __dadd_rn(C, diff); } }
            } else {
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
                red.bc_d[0] = yi ? -dai : dai;
                red.bc_d[1] = yj ? -daj : daj;
                red.bc_d[2] = ai; red.bc_d[3] = aj;
                red.bc_i[0] = ai >= C ? ST_UPPER : (ai <= 0 ? ST_LOWER : ST_FREE);
                red.bc_i[1] = aj >= C ? ST_UPPER : (aj <= 0 ? ST_LOWER : ST_FREE);
            }
        }
        tick(5);
        __syncthreads();
        tick(6);
        if constexpr (ROWBUF) {
            wait_row();
            tick(7);
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int t = gpos(k), s = k * NT + tid;
                kvj[k] = t < active ? rowbuf[col[s]] : 0.f;
            }
        }
        const double a = red.bc_d[0], b = red.bc_d[1];
        const int sti = red.bc_i[0], stj = red.bc_i[1];
        if (owner_of(i) == rank && tid == i % NT) { const int s = slot_of(i); alpha[s] = red.bc_d[2]; fl[s] = (unsigned char)mkflags((pi & F_YPOS) != 0, sti); }
        if (owner_of(j) == rank && tid == j % NT) { const int s = slot_of(j); alpha[s] = red.bc_d[3]; fl[s] = (unsigned char)mkflags((pj & F_YPOS) != 0, stj); }

        la = -CUDART_INF; lm = CUDART_INF; la_idx = -1;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = gpos(k), s = k * NT + tid;
            if (t < active) {
                const double m = __dadd_rn(mG[s], __dadd_rn(__dmul_rn(qi[k], a), __dmul_rn(widen(kvj[k]), b)));
                mG[s] = m;
                const int f = fl[s];
                if ((f & F_UP) && m >= la) { la = m; la_idx = (t << IDX_SHIFT) | f; }
                if (f & F_LOW) lm = fmin(lm, m);
            }
        }
        scan_valid = true;
        const bool need_i = use_gbar && (((pi & 3) == ST_UPPER) != (sti == ST_UPPER));
        const bool need_j = use_gbar && (((pj & 3) == ST_UPPER) != (stj == ST_UPPER));
        if (need_i || need_j) {
            const double C = Pp->C;
            const float *__restrict__ Ki = K + (size_t)col_i * ldk;
            const double ci = (((pi & 3) == ST_UPPER) == ((pi & F_YPOS) != 0)) ? C : -C;
            const double cj = (((pj & 3) == ST_UPPER) == ((pj & F_YPOS) != 0)) ? C : -C;
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const int t = gpos(k), s = k * NT + tid;
                if (t < l) {
                    const bool act = t < active;
                    double gb = mGbar[s];
                    if (need_i) gb = __dadd_rn(gb, __dmul_rn(ci, act ? qi[k] : widen(__ldg(Ki + col[s]))));
                    if (need_j) gb = __dadd_rn(gb, __dmul_rn(cj, widen(act ? kvj[k] : __ldg(Kj + col[s]))));
                    mGbar[s] = gb;
                }
            }
        }
        tick(8);
    }

    // ---------------- calculate_rho: sequential float64 sum in libsvm's (position) order, via DSMEM ----------------
    cluster.sync();
    const double C = Pp->C;
    if (rank == 0 && tid == 0) {
        const double *mGr[CL];
        const unsigned char *flr[CL];
#pragma unroll
        for (int c = 0; c < CL; c++) { mGr[c] = cluster.map_shared_rank(mG, c); flr[c] = cluster.map_shared_rank(fl, c); }
        int nfree = 0;
        double ub = CUDART_INF, lb = -CUDART_INF, sum = 0;
        for (int t0 = 0; t0 < active; t0 += NT) {                 // chunk by chunk: one owner per chunk
            const int c = (t0 / NT) % CL, sbase = (t0 / (NT * CL)) * NT;
            const int cnt = min(NT, active - t0);
            for (int u = 0; u < cnt; u++) {
                const int f = flr[c][sbase + u];
                const double yG = -mGr[c][sbase + u];
                if ((f & 3) == ST_UPPER) { if (!(f & F_YPOS)) ub = fmin(ub, yG); else lb = fmax(lb, yG); }
                else if ((f & 3) == ST_LOWER) { if (f & F_YPOS) ub = fmin(ub, yG); else lb = fmax(lb, yG); }
                else { ++nfree; sum = __dadd_rn(sum, yG); }
            }
        }
        *Pp->out_rho = nfree > 0 ? __ddiv_rn(sum, (double)nfree) : __ddiv_rn(__dadd_rn(ub, lb), 2.0);
    }
    int nsv = 0, nbsv = 0;
    {
        double *__restrict__ coef = Pp->coef;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int t = gpos(k), s = k * NT + tid;
            if (t < l) {
                const double av = alpha[s];
                coef[col[s]] = (fl[s] & F_YPOS) ? av : -av;
                nsv += av > 0;
                nbsv += av >= C;
            }
        }
    }
#pragma unroll
    for (int m = 16; m; m >>= 1) {
        nsv += __shfl_xor_sync(0xffffffffu, nsv, m);
        nbsv += __shfl_xor_sync(0xffffffffu, nbsv, m);
    }
    if (lane == 0) { red.cnt[warp] = nsv; red.a_idx[warp] = nbsv; }
    __syncthreads();
    {
        int s = 0, bs = 0;
        for (int w = 0; w < NW; w++) { s += red.cnt[w]; bs += red.a_idx[w]; }
        unsigned v[XW] = {0};
        v[0] = (unsigned)s; v[1] = (unsigned)bs;
        const unsigned (*r)[XW] = exchange(v);
        if (rank == 0 && tid == 0) {
            int ts = 0, tb = 0;
#pragma unroll
            for (int c = 0; c < CL; c++) { ts += (int)r[c][0]; tb += (int)r[c][1]; }
            int *info = Pp->out_info;
            info[0] = iter; info[1] = timed_out; info[2] = ts; info[3] = tb;
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            unsigned long long *ns = Pp->out_ns;
            ns[0] = t_start; ns[1] = t_end;
            if constexpr (PROF)
                for (int q = 0; q < 10; q++) ns[2 + q] = (unsigned long long)prof[q];
        }
    }
    cluster.sync();                                              // no CTA may exit while a peer still reads its shared memory
}

template <int NT, int KPT, int CL, bool FAST, bool ROWBUF, bool PROF>
cudaError_t launch_cluster(const SmoProblem *probs, const int *order, int n_prob, int rowcap, cudaStream_t st)
{
    constexpr int LCAP = NT * KPT;
    const size_t smem = (size_t)LCAP * (8 + 8 + 8 + 2 + 1) + (ROWBUF ? 128 + (size_t)rowcap * 4 : 0);
    auto kern = smo_cluster_kernel<NT, KPT, CL, FAST, ROWBUF, PROF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_prob * CL));
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int fmode = 0;
    if (const char *e = getenv("B200GS_SMO_FMODE")) fmode = atoi(e);           // development switch
    return cudaLaunchKernelEx(&cfg, kern, probs, order, rowcap, fmode);
}

int env_int_c(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

}  // namespace

// Largest sub-problem a cluster launch of size cl supports (0: unsupported cluster size)
int smo_cluster_max_rows(int cl) { return cl == 2 ? 8192 : ((cl == 4 || cl == 8) ? 16384 : 0); }

template <int NT, int KPT, int CL>
static cudaError_t launch_cluster_f(const SmoProblem *p, const int *o, int n, bool fast, int rowcap, cudaStream_t st)
{
    // the row buffer shares the 227 KB of shared memory with the state (27 B per owned row) and ~3 KB of static scratch
    const bool fits = rowcap > 0 && rowcap % 32 == 0 &&
                      (size_t)NT * KPT * 27 + 128 + (size_t)rowcap * 4 + 4096 <= 227 * 1024;
    const bool rowbuf = fits && env_int_c("B200GS_SMO_ROWBUF", 1);
    const bool prof = env_int_c("B200GS_SMO_PROF", 0) != 0;
    if (prof && rowbuf) return fast ? launch_cluster<NT, KPT, CL, true, true, true>(p, o, n, rowcap, st)
                                    : launch_cluster<NT, KPT, CL, false, true, true>(p, o, n, rowcap, st);
    if (rowbuf) return fast ? launch_cluster<NT, KPT, CL, true, true, false>(p, o, n, rowcap, st)
                            : launch_cluster<NT, KPT, CL, false, true, false>(p, o, n, rowcap, st);
    return fast ? launch_cluster<NT, KPT, CL, true, false, false>(p, o, n, 0, st)
                : launch_cluster<NT, KPT, CL, false, false, false>(p, o, n, 0, st);
}

// Shape = (threads per CTA) x (rows per thread) x (CTAs per problem).
cudaError_t launch_smo_cluster(const SmoProblem *d_probs, const int *d_order, int n_prob, int lmax, int cl, bool fast, int rowcap,
                               cudaStream_t st)
{
    if (n_prob <= 0) return cudaSuccess;
    if (lmax > smo_cluster_max_rows(cl)) return cudaErrorInvalidValue;
    int nt = 0;
    if (const char *e = getenv("B200GS_SMO_NT")) nt = atoi(e);                  // development switch
    if (cl == 2) {
        if (nt == 512) return launch_cluster_f<512, 8, 2>(d_probs, d_order, n_prob, fast, rowcap, st);
        return launch_cluster_f<1024, 4, 2>(d_probs, d_order, n_prob, fast, rowcap, st);
    }
    if (cl == 4) {
        if (lmax > 8192) return launch_cluster_f<1024, 4, 4>(d_probs, d_order, n_prob, fast, rowcap, st);
        if (nt == 256) return launch_cluster_f<256, 8, 4>(d_probs, d_order, n_prob, fast, rowcap, st);
        if (nt == 512) return launch_cluster_f<512, 4, 4>(d_probs, d_order, n_prob, fast, rowcap, st);
        return launch_cluster_f<1024, 2, 4>(d_probs, d_order, n_prob, fast, rowcap, st);      // one SM per CTA: co-exists with the single-CTA kernel
    }
    if (cl == 8) {
        if (lmax > 8192) return launch_cluster_f<512, 4, 8>(d_probs, d_order, n_prob, fast, rowcap, st);
        if (nt == 128) return launch_cluster_f<128, 8, 8>(d_probs, d_order, n_prob, fast, rowcap, st);
        if (nt == 1024) return launch_cluster_f<1024, 1, 8>(d_probs, d_order, n_prob, fast, rowcap, st);
        return launch_cluster_f<256, 4, 8>(d_probs, d_order, n_prob, fast, rowcap, st);
    }
    return cudaErrorInvalidValue;
}

// smo_colown.cu -- cluster SMO solver with STATIC element ownership ("column owner") and register-resident state.
//
// Same algorithm and the same bit-exact trajectory as smo.cu (libsvm svm.cpp:629-1168 restated there), one sub-problem
// per thread-block cluster of CL CTAs.  What differs from a position-owned layout is WHERE an element lives:
//
//   * libsvm's shrinking permutes positions (swap_index, svm.cpp:616-627) and every tie-break and summation order is
//     defined on positions.  Here an element (one training row) never moves: thread (rank, tid, k) owns element
//     e = (k*CL + rank)*NT + tid for the whole solve and carries its CURRENT POSITION as state.  A libsvm swap of
//     positions p and q is an exchange of two integers; `pos < active` is the active-set test; arg-reductions compare
//     (value, position) exactly as libsvm's "last index wins" scans do.
//   * Because the element -> dataset-row map is fixed and the sub-problem's rows are ascending runs of the (class-sorted)
//     dataset, a K row is gathered with COALESCED loads: each CTA reads only its own 1/CL of the row, once, straight
//     into registers -- one DRAM round trip, no column indirection, no row staging.  (Position-owned layouts scatter a
//     CTA's columns over the whole row after the first shrink: every CTA then pulls all of it.)
//   * m = -y*G, the position, the flags and the K_i values of the owned elements live in REGISTERS; alpha and
//     mbar = -y*G_bar (touched by two elements per iteration / on status flips) in shared memory.
//
// The hot loop has NO barrier of any kind -- neither __syncthreads nor a cluster barrier.  Each of the two arg-reductions
// of an iteration is a WARP-level all-gather: a warp reduces its own elements (REDUX), its winner lane st.async-writes
// the warp's 32/48-byte record into every CTA's shared memory (bytes counted on each RECEIVER's mbarrier), and every
// warp of every CTA waits on its own CTA's mbarrier and reduces the same CL*NW records.  The scalar two-variable update
// is then computed redundantly by every thread.  Ordering: a warp can finish stage s only after every warp of the
// cluster SENT its stage-s record, and a warp sends its next record only after it has read all records of the current
// stage; stages alternate A, B (, X), each with its own slots and mbarrier, so no slot is overwritten while readable.
// Cold paths keep CTA-level records (double-buffered by parity) behind __syncthreads.
//
// Cold paths (every 1000 iterations / at unshrink): do_shrinking builds libsvm's two-pointer partition from a
// position-indexed mark array (all-gathered through DSMEM), each CTA deriving the identical swap map; the gradient
// reconstruction and calculate_rho walk positions in ascending order through position-indexed global scratch.
#include "smo_common.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace {

using namespace smo;

constexpr int XW = 16;                       // words per exchange record
constexpr int NOPOS = 0x3fffffff;            // position of a slot that holds no element

template <int CL>
struct __align__(16) Xch {                   // double-buffered all-gather slots: [parity][source rank][word]
    unsigned w[2][CL][XW];
};

__device__ __forceinline__ unsigned lo32(double x) { return (unsigned)__double_as_longlong(x); }
__device__ __forceinline__ unsigned hi32(double x) { return (unsigned)((unsigned long long)__double_as_longlong(x) >> 32); }
__device__ __forceinline__ double mk64(unsigned lo, unsigned hi) { return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo)); }

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

template <int NT, int KPT, int CL, bool FAST, bool PROF>
__global__ void __launch_bounds__(NT, 1)
smo_colown_kernel(const SmoProblem *__restrict__ probs, const int *__restrict__ order)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Red red;
    __shared__ Xch<CL> xch;
    __shared__ __align__(8) unsigned long long xbar[2];                      // cold paths: one mbarrier per exchange parity
    constexpr int NW = NT / 32;
    constexpr int R = CL * NW;                                              // warp records per all-gather
    constexpr int RPL = (R + 31) / 32;                                      // records per lane
    __shared__ __align__(16) unsigned wxA[R][8];                            // stage A records: key hi lo | idx | km hi lo | alpha lo hi | col
    __shared__ __align__(16) unsigned wxB[R][12];                           // stage B: b1 | idx | b2 | col | mg | K_ij | alpha | pad
    __shared__ __align__(16) unsigned wxX[R][12];                           // exact tie-break: hi | idx | lo | col | mg | K_ij | alpha | pad
    __shared__ __align__(8) unsigned long long wbar[3];                      // one mbarrier per stage
    constexpr int LCAP = NT * KPT;                                          // elements owned by this CTA

    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const SmoProblem *__restrict__ Pp = probs + order[blockIdx.x / CL];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = Pp->l;
    const int lhalf = (l + 1) / 2;
    // ---- shared memory: per-element alpha / mbar, then the cold-path arrays indexed by POSITION ----
    double *const alpha = reinterpret_cast<double *>(smem_raw);             // [k*NT + tid]
    double *const mGbar = alpha + LCAP;
    int *const lists = reinterpret_cast<int *>(mGbar + LCAP);               // l ints: plist | qlist, or the free list
    unsigned short *const swapmap = reinterpret_cast<unsigned short *>(lists + l);   // l
    unsigned char *const mk = reinterpret_cast<unsigned char *>(swapmap + l);        // l
    const float *__restrict__ const K = Pp->K;
    const int64_t ldk = Pp->ldk;
    const double eps = Pp->eps;
    const double Cc = Pp->C;
    const bool use_gbar = Pp->shrinking != 0;
    const double *__restrict__ const qd = FAST ? nullptr : Pp->qd;
    int *const scratch = Pp->scratch;                                        // global ints, >= 2*l
    double *const gscratch = Pp->Gbar;                                       // global doubles, l (G_bar lives in shared memory here)

    unsigned long long t_start = 0;
    if (tid == 0 && rank == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

    // ---- element state in registers ----
    double m[KPT];            // m = -y*G
    int pos[KPT];             // current libsvm position (NOPOS: empty slot)
    int fl[KPT];              // status | F_YPOS | F_UP | F_LOW
    int colr[KPT];            // dataset row == column of K
    // initial point: alpha = 0, G = -1  =>  m = y (svm.cpp:1611-1626, :716-736)
    {
        const int n_pos = Pp->n_pos;
        const int *__restrict__ rows = Pp->rows;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int e = (k * CL + (int)rank) * NT + tid, s = k * NT + tid;
            alpha[s] = 0.0; mGbar[s] = 0.0;
            if (e < l) {
                const bool yp = e < n_pos;
                m[k] = yp ? 1.0 : -1.0; pos[k] = e; fl[k] = mkflags(yp, ST_LOWER); colr[k] = rows[e];
            } else { m[k] = 0.0; pos[k] = NOPOS; fl[k] = 0; colr[k] = 0; }
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&xbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&xbar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&wbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&wbar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&wbar[2])));
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
    unsigned xphase = 0;                                                    // bit p: parity the next wait on xbar[p] uses

    float kvi[KPT];                                                         // unsigned K_i at the owned elements (float32 as stored)

    auto QDc = [&](int c) -> double {                                       // by dataset row (svm.cpp:1436-1437)
        if constexpr (FAST) return 1.0;
        else return qd ? qd[c] : 1.0;
    };
    auto widen = [&](float x) -> double {
        if constexpr (FAST) {
            const unsigned u = __float_as_uint(x);
            return __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29));
        } else return f2d(x);
    };

    // all-gather of one record per CTA (v must be warp-uniform in warp 0: REDUX results and broadcast shared reads)
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

    // ---------------- local scan (normally fused into the update loop) ----------------
    // la: max m over the owned I_up elements, ties -> larger position (libsvm's ascending ">=" scan); lm: min m over I_low
    double la = -CUDART_INF, lm = CUDART_INF;
    int la_pos = -1, la_k = 0;
    auto scan_elem = [&](int k) {
        const int f = fl[k];
        const double mv = m[k];
        if ((f & F_UP) && (mv > la || (mv == la && pos[k] > la_pos))) { la = mv; la_pos = pos[k]; la_k = k; }
        if (f & F_LOW) lm = fmin(lm, mv);
    };
    auto local_scan = [&]() {
        la = -CUDART_INF; lm = CUDART_INF; la_pos = -1; la_k = 0;
#pragma unroll
        for (int k = 0; k < KPT; k++)
            if (pos[k] < active) scan_elem(k);
    };

    // ---------------- reconstruct_gradient (svm.cpp:629-668) ----------------
    // m_k = (mbar_k + y_k) + sum over free active f in ASCENDING POSITION of fl((-y_f alpha_f) K_fk), for inactive k
    auto rebuild_gradient = [&]() {
        if (active == l) return;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            if (pos[k] < active) {
                const bool fr = (fl[k] & 3) == ST_FREE;
                scratch[pos[k]] = fr ? colr[k] : -1;
                if (fr) { const double av = alpha[k * NT + tid]; gscratch[pos[k]] = (fl[k] & F_YPOS) ? -av : av; }
            }
        }
        __threadfence();
        cluster.sync();
        int nf = 0;
        for (int base = 0; base < active; base += NT) {                     // every CTA compacts the same list
            const int t = base + tid;
            const bool isf = t < active && __ldcg(scratch + t) >= 0;
            int tot;
            const int r = block_rank<NT>(isf, red.cnt, tot);
            if (isf) lists[nf + r] = t;
            nf += tot;
        }
        __syncthreads();
        double g[KPT];
        bool in[KPT];
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            in[k] = pos[k] >= active && pos[k] != NOPOS;
            g[k] = in[k] ? __dadd_rn(mGbar[k * NT + tid], (fl[k] & F_YPOS) ? 1.0 : -1.0) : 0.0;
        }
#pragma unroll 2
        for (int r = 0; r < nf; r++) {
            const int p = lists[r];
            const float *__restrict__ Kf = K + (size_t)__ldcg(scratch + p) * ldk;
            const double af = __ldcg(gscratch + p);
#pragma unroll
            for (int k = 0; k < KPT; k++)
                if (in[k]) g[k] = __dadd_rn(g[k], __dmul_rn(af, widen(__ldg(Kf + colr[k]))));
        }
#pragma unroll
        for (int k = 0; k < KPT; k++)
            if (in[k]) m[k] = g[k];
        cluster.sync();                                                      // scratch is reused by the next cold path
    };

    // ---------------- select_working_set (svm.cpp:946-1047) ----------------
    // Warp-level all-gather, no CTA barrier: every warp reduces its own elements, its winner lane writes the warp's record
    // into slot (rank*NW + warp) of EVERY CTA of the cluster (st.async; bytes counted on each receiver's mbarrier), and
    // every warp of every CTA then reduces the same R = CL*NW records.  Stage A, stage B and the rare exact tie-break
    // each own a record array and an mbarrier; a warp sends its stage-(s+1) record only after it has read all stage-s
    // records, so a slot is never overwritten while any warp can still read it (stages strictly alternate).
    auto send_record = [&](unsigned slot_base, unsigned bar, int words, const unsigned (&v)[12]) {
        const unsigned slot = slot_base + (unsigned)((int)rank * NW + warp) * (unsigned)words * 4u;
#pragma unroll
        for (int c = 0; c < CL; c++) {
            const unsigned ra = mapa_u32(slot, (unsigned)c), rb = mapa_u32(bar, (unsigned)c);
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                         ::"r"(ra), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(rb) : "memory");
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                         ::"r"(ra + 16u), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(rb) : "memory");
            if (words == 12)
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                             ::"r"(ra + 32u), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(rb) : "memory");
        }
    };
    unsigned wphase = 0;                                 // bit s: parity the next wait on wbar[s] uses
    auto stage_wait = [&](int sidx) {
        mbar_wait(smem_u32(&wbar[sidx]), (wphase >> sidx) & 1u);
        wphase ^= 1u << sidx;
    };

    int pi = -1, pj = -1, col_i = 0, col_j = 0;          // packed (position << 5 | flags), dataset rows
    double gmax = 0, mg_j = 0, k_ij = 0, alpha_i = 0, alpha_j = 0;
    auto select = [&]() -> bool {
        double gmax2;
        {   // ---- stage A: i = argmax m over I_up, Gmax2 = max -m over I_low ----
            const unsigned long long key = dkey(la);
            const int la_idx = la_pos >= 0 ? (la_pos << IDX_SHIFT) : -1;             // the sender adds the flags
            const KArg w = warp_argmax((unsigned)(key >> 32), (unsigned)key, la_idx);
            const unsigned long long km = warp_keymax(dkey(-lm));
            if (tid == 0) mbar_expect_tx(smem_u32(&wbar[0]), R * 32);
            if (w.idx >= 0 ? la_idx == w.idx : lane == 0) {                          // exactly one lane per warp
                int c = colr[0], f = fl[0];
#pragma unroll
                for (int k = 1; k < KPT; k++) { c = k == la_k ? colr[k] : c; f = k == la_k ? fl[k] : f; }
                const double av = alpha[la_k * NT + tid];
                unsigned v[12];
                v[0] = w.hi; v[1] = w.lo; v[2] = w.idx >= 0 ? (unsigned)(w.idx | f) : 0xffffffffu;
                v[3] = (unsigned)(km >> 32); v[4] = (unsigned)km; v[5] = lo32(av); v[6] = hi32(av); v[7] = (unsigned)c;
                v[8] = v[9] = v[10] = v[11] = 0u;
                send_record(smem_u32(&wxA[0][0]), smem_u32(&wbar[0]), 8, v);
            }
            tick(0);
            stage_wait(0);
            // every warp reduces all R records: lane-local best of its RPL records, then REDUX
            unsigned bh = 0u, bl = 0u; int bi = -1, br = 0;
            unsigned long long kml = 0ull;
#pragma unroll
            for (int q = 0; q < RPL; q++) {
                const int r = lane + 32 * q;
                if (r < R) {
                    const uint4 x = *reinterpret_cast<const uint4 *>(&wxA[r][0]);
                    const unsigned k4 = wxA[r][4];
                    const int xi = (int)x.z;
                    if (xi >= 0 && (x.x > bh || (x.x == bh && (x.y > bl || (x.y == bl && xi > bi))))) { bh = x.x; bl = x.y; bi = xi; br = r; }
                    const unsigned long long k2 = ((unsigned long long)x.w << 32) | k4;
                    kml = k2 > kml ? k2 : kml;
                }
            }
            const KArg a = warp_argmax(bh, bl, bi);
            const unsigned long long km2 = warp_keymax(kml);
            pi = a.idx;
            if (pi >= 0) {
                const int wl = __ffs(__ballot_sync(0xffffffffu, bi == a.idx)) - 1;
                const int rw = __shfl_sync(0xffffffffu, br, wl);
                const uint4 y = *reinterpret_cast<const uint4 *>(&wxA[rw][4]);           // km lo | alpha lo | alpha hi | col
                alpha_i = mk64(y.y, y.z);
                col_i = (int)y.w;
            }
            gmax = dkey_inv(((unsigned long long)a.hi << 32) | a.lo);
            gmax2 = dkey_inv(km2);
            tick(1);
        }
        if (pi < 0 || __dadd_rn(gmax, gmax2) < eps) return true;              // svm.cpp:1040-1041
        // ---- stage B: j = argmin -(gd^2)/quad over I_low with gd > 0 (svm.cpp:980-1037) ----
        const double QDi = QDc(col_i);
        const float *__restrict__ Ki = K + (size_t)col_i * ldk;
#pragma unroll
        for (int k = 0; k < KPT; k++) kvi[k] = pos[k] < active ? __ldg(Ki + colr[k]) : 0.f;    // coalesced: fixed columns
        tick(2);
        // Approximate gd^2/quad tracked by a 32-bit order-preserving key; every near-tie (keys within BAND units) is decided
        // by the exact libsvm quotients below, so the choice stays bit-identical (error analysis: smo.cu, same filter).
        constexpr unsigned BAND = FAST ? 64u : 514u;
        constexpr unsigned KEY_TINY = 0x0D800000u;                          // float bits of 2^-100
        auto approx_key = [&](double gd, float kvf, int c) -> unsigned {
            if constexpr (FAST) {
                const float gdf = __double2float_rn(gd);
                const float quadf = __fmaf_rn(-2.f, kvf, 2.f);              // == fl32(2 - 2K): 2K is exact
                const float g2f = __fmul_rn(gdf, gdf);
                float r;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(quadf));
                const float apf = quadf > 0.f ? __fmul_rn(g2f, r) : __fmul_rn(g2f, 1e12f);
                return __float_as_uint(apf) + 1u;                           // +1: a valid candidate is never 0
            } else {
                const double quad = __dsub_rn(__dadd_rn(QDi, QDc(c)), __dmul_rn(2.0, widen(kvf)));
                const double g2 = __dmul_rn(gd, gd);
                const double ap = quad > 0 ? g2 * rcp_approx(quad) : g2 * 1e12;
                return (unsigned)__double2hiint(ap) + 1u;
            }
        };
        unsigned b1k = 0u, b2k = 0u;                    // keys of the best and second-best candidate (0 = none)
        int k1 = -1;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            if (pos[k] < active) {
                const int f = fl[k];
                const double gd = __dsub_rn(gmax, m[k]);
                if ((f & F_LOW) && gd > 0) {
                    const unsigned key = approx_key(gd, kvi[k], colr[k]);
                    const bool gt = key > b1k;
                    b2k = gt ? b1k : max(b2k, key);
                    b1k = gt ? key : b1k;
                    k1 = gt ? k : k1;
                }
            }
        }
        // winner payload of this thread (selected lazily: only the sender lane needs it)
        auto payload = [&](int kk, unsigned (&v)[12]) {
            int cc = colr[0];
            float kq = kvi[0];
            double mm = m[0];
#pragma unroll
            for (int k = 1; k < KPT; k++) { const bool s_ = k == kk; cc = s_ ? colr[k] : cc; kq = s_ ? kvi[k] : kq; mm = s_ ? m[k] : mm; }
            const double qq = widen(kq), av = alpha[kk * NT + tid];
            v[3] = (unsigned)cc; v[4] = lo32(mm); v[5] = hi32(mm); v[6] = lo32(qq); v[7] = hi32(qq); v[8] = lo32(av); v[9] = hi32(av);
            v[10] = v[11] = 0u;
        };
        int idx1 = -1;
        if (k1 >= 0) {
            int p1 = pos[0], f1 = fl[0];
#pragma unroll
            for (int k = 1; k < KPT; k++) { const bool s_ = k == k1; p1 = s_ ? pos[k] : p1; f1 = s_ ? fl[k] : f1; }
            idx1 = (p1 << IDX_SHIFT) | f1;
        }
        unsigned top1k, top2k;
        {
            const unsigned w1 = __reduce_max_sync(0xffffffffu, b1k);
            const int widx = __reduce_max_sync(0xffffffffu, (b1k == w1) ? idx1 : -1);
            const unsigned w2 = __reduce_max_sync(0xffffffffu, (idx1 == widx) ? b2k : b1k);
            if (tid == 0) mbar_expect_tx(smem_u32(&wbar[1]), R * 48);
            if (widx >= 0 ? idx1 == widx : lane == 0) {
                unsigned v[12];
                if (widx >= 0) payload(k1, v);
                else { v[3] = v[4] = v[5] = v[6] = v[7] = v[8] = v[9] = v[10] = v[11] = 0u; }
                v[0] = w1; v[1] = (unsigned)widx; v[2] = w2;
                send_record(smem_u32(&wxB[0][0]), smem_u32(&wbar[1]), 12, v);
            }
            tick(3);
            stage_wait(1);
            unsigned bk = 0u; int bi = -1, br = 0;
#pragma unroll
            for (int q = 0; q < RPL; q++) {
                const int r = lane + 32 * q;
                if (r < R) {
                    const uint2 x = *reinterpret_cast<const uint2 *>(&wxB[r][0]);
                    const int xi = (int)x.y;
                    if (xi >= 0 && (x.x > bk || (x.x == bk && xi > bi))) { bk = x.x; bi = xi; br = r; }
                }
            }
            top1k = __reduce_max_sync(0xffffffffu, bk);
            pj = __reduce_max_sync(0xffffffffu, (bk == top1k) ? bi : -1);
            if (pj < 0) return true;                                               // Gmin_idx == -1
            unsigned sk = 0u;                                                       // runner-up: best b1 of the others, b2 of the winner
#pragma unroll
            for (int q = 0; q < RPL; q++) {
                const int r = lane + 32 * q;
                if (r < R) {
                    const unsigned kk = (int)wxB[r][1] == pj ? wxB[r][2] : ((int)wxB[r][1] >= 0 ? wxB[r][0] : 0u);
                    sk = kk > sk ? kk : sk;
                }
            }
            top2k = __reduce_max_sync(0xffffffffu, sk);
            const int wl = __ffs(__ballot_sync(0xffffffffu, bi == pj)) - 1;
            const int rw = __shfl_sync(0xffffffffu, br, wl);
            const uint4 y = *reinterpret_cast<const uint4 *>(&wxB[rw][4]);             // mg lo hi | kv lo hi
            const uint2 z = *reinterpret_cast<const uint2 *>(&wxB[rw][8]);             // alpha lo hi
            col_j = (int)wxB[rw][3];
            mg_j = mk64(y.x, y.y); k_ij = mk64(y.z, y.w); alpha_j = mk64(z.x, z.y);
            tick(4);
        }
        if (top1k - top2k <= BAND || (FAST && top1k <= KEY_TINY)) {
            // ---- exact tie-break: libsvm's correctly rounded quotients for every element in the band (rare) ----
            const unsigned thrk = (top1k > BAND && !(FAST && top1k <= KEY_TINY)) ? top1k - BAND : 1u;
            double bestn = -CUDART_INF;
            int bidx = -1, bk_ = 0;
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                if (pos[k] < active) {
                    const int f = fl[k];
                    const double gd = __dsub_rn(gmax, m[k]);
                    if ((f & F_LOW) && gd > 0) {
                        if (approx_key(gd, kvi[k], colr[k]) >= thrk) {
                            const double quad = __dsub_rn(__dadd_rn(QDi, QDc(colr[k])), __dmul_rn(2.0, widen(kvi[k])));
                            const double g2 = __dmul_rn(gd, gd);
                            const double nod = quad > 0 ? __ddiv_rn(g2, quad) : __ddiv_rn(g2, TAU);   // == -obj_diff
                            const int cand = (pos[k] << IDX_SHIFT) | f;
                            if (nod > bestn || (nod == bestn && cand > bidx)) { bestn = nod; bidx = cand; bk_ = k; }
                        }
                    }
                }
            }
            const unsigned long long key = dkey(bestn);
            const KArg w = warp_argmax((unsigned)(key >> 32), (unsigned)key, bidx);
            if (tid == 0) mbar_expect_tx(smem_u32(&wbar[2]), R * 48);
            if (w.idx >= 0 ? bidx == w.idx : lane == 0) {
                unsigned v[12];
                if (w.idx >= 0) payload(bk_, v);
                else { v[3] = v[4] = v[5] = v[6] = v[7] = v[8] = v[9] = v[10] = v[11] = 0u; }
                v[0] = w.hi; v[1] = (unsigned)w.idx; v[2] = w.lo;
                send_record(smem_u32(&wxX[0][0]), smem_u32(&wbar[2]), 12, v);
            }
            stage_wait(2);
            unsigned bh = 0u, bl = 0u; int bi = -1, br = 0;
#pragma unroll
            for (int q = 0; q < RPL; q++) {
                const int r = lane + 32 * q;
                if (r < R) {
                    const uint4 x = *reinterpret_cast<const uint4 *>(&wxX[r][0]);       // hi | idx | lo | col
                    const int xi = (int)x.y;
                    if (xi >= 0 && (x.x > bh || (x.x == bh && (x.z > bl || (x.z == bl && xi > bi))))) { bh = x.x; bl = x.z; bi = xi; br = r; }
                }
            }
            const KArg b = warp_argmax(bh, bl, bi);
            pj = b.idx;                                                         // >= 0: the approximate winner is in the band
            const int wl = __ffs(__ballot_sync(0xffffffffu, bi == pj)) - 1;
            const int rw = __shfl_sync(0xffffffffu, br, wl);
            const uint4 y = *reinterpret_cast<const uint4 *>(&wxX[rw][4]);
            const uint2 z = *reinterpret_cast<const uint2 *>(&wxX[rw][8]);
            col_j = (int)wxX[rw][3];
            mg_j = mk64(y.x, y.y); k_ij = mk64(y.z, y.w); alpha_j = mk64(z.x, z.y);
        }
        return false;
    };

    // ---------------- do_shrinking (svm.cpp:1070-1129) ----------------
    // Gmax1 = max{m : I_up}, Gmax2 = max{-m : I_low}; be_shrunk = (!up && m > Gmax1) || (!low && -m > Gmax2)
    auto do_shrink = [&]() {
        double g1 = -CUDART_INF, g2 = -CUDART_INF;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            if (pos[k] < active) {
                if (fl[k] & F_UP) g1 = fmax(g1, m[k]);
                if (fl[k] & F_LOW) g2 = fmax(g2, -m[k]);
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
        bool mark[KPT];
        int keep_local = 0;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            mark[k] = false;
            if (pos[k] < active) {
                const int f = fl[k];
                mark[k] = (!(f & F_UP) && m[k] > g1) || (!(f & F_LOW) && -m[k] > g2);
                keep_local += mark[k] ? 0 : 1;
            }
        }
#pragma unroll
        for (int s = 16; s; s >>= 1) keep_local += __shfl_xor_sync(0xffffffffu, keep_local, s);
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
            // libsvm's two-pointer sweep pairs the k-th marked position below na (ascending) with the k-th unmarked
            // position at/above na (descending).  Marks are all-gathered by position; every CTA derives the same map.
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                if (pos[k] < active) {
#pragma unroll
                    for (int c = 0; c < CL; c++) *cluster.map_shared_rank(mk + pos[k], c) = mark[k] ? 1 : 0;
                }
            }
            cluster.sync();
            int *plist = lists, *qlist = lists + lhalf;
            int np = 0, nq = 0;
            for (int base = 0; base < na; base += NT) {
                const int t = base + tid;
                const bool pr = t < na && mk[t];
                int tot;
                const int r = block_rank<NT>(pr, red.cnt, tot);
                if (pr) plist[np + r] = t;
                np += tot;
            }
            for (int base = na; base < active; base += NT) {
                const int t = base + tid;
                const bool pr = t < active && !mk[t];
                int tot;
                const int r = block_rank<NT>(pr, red.cnt, tot);
                if (pr) qlist[nq + r] = t;
                nq += tot;
            }
            __syncthreads();
            for (int r = tid; r < np; r += NT) {                 // np == nq; disjoint pairs
                const int p = plist[r], q = qlist[np - 1 - r];
                swapmap[p] = (unsigned short)q; swapmap[q] = (unsigned short)p;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                if (pos[k] < active) {
                    const bool moved = pos[k] < na ? mark[k] : !mark[k];
                    if (moved) pos[k] = (int)swapmap[pos[k]];
                }
            }
            active = na;
            // mk / lists / swapmap are next written 1000 iterations (thousands of exchanges) later: no barrier needed here
        }
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
            local_scan();
            if (select()) break;
            counter = 1;
        }
        ++iter;

        const int i = pi >> IDX_SHIFT, j = pj >> IDX_SHIFT;
        const float *__restrict__ Kj = K + (size_t)col_j * ldk;
        float kvj[KPT];
#pragma unroll
        for (int k = 0; k < KPT; k++) kvj[k] = pos[k] < active ? __ldg(Kj + colr[k]) : 0.f;     // in flight during the scalar update
        // analytic two-variable update, computed redundantly (and identically) by every thread: no broadcast, no barrier
        double a, b, ai = alpha_i, aj = alpha_j;
        int sti, stj;
        {
            const double C = Cc;
            const bool yi = (pi & F_YPOS) != 0, yj = (pj & F_YPOS) != 0;
            const double Gi = yi ? -gmax : gmax;                 // G = -y m (exact)
            const double Gj = yj ? -mg_j : mg_j;
            const double QDi = QDc(col_i), QDj = QDc(col_j);
            const double Qij = (yi == yj) ? k_ij : -k_ij;        // signed Q_i[j]
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
            const double dai = __dsub_rn(ai, alpha_i), daj = __dsub_rn(aj, alpha_j);
            a = yi ? -dai : dai;                                 // a = -y_i dalpha_i
            b = yj ? -daj : daj;                                 // b = -y_j dalpha_j
            sti = ai >= C ? ST_UPPER : (ai <= 0 ? ST_LOWER : ST_FREE);
            stj = aj >= C ? ST_UPPER : (aj <= 0 ? ST_LOWER : ST_FREE);
        }
        tick(5);
        // the owners of i and j take the new alpha and status FIRST: the fused scan below must see the new sets
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            if (pos[k] == i) { alpha[k * NT + tid] = ai; fl[k] = mkflags((pi & F_YPOS) != 0, sti); }
            if (pos[k] == j) { alpha[k * NT + tid] = aj; fl[k] = mkflags((pj & F_YPOS) != 0, stj); }
        }
        // m update over the active set (svm.cpp:866-872), fused with the next iteration's local scan
        la = -CUDART_INF; lm = CUDART_INF; la_pos = -1; la_k = 0;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            if (pos[k] < active) {
                m[k] = __dadd_rn(m[k], __dadd_rn(__dmul_rn(widen(kvi[k]), a), __dmul_rn(widen(kvj[k]), b)));
                scan_elem(k);
            }
        }
        scan_valid = true;
        tick(7);
        // G_bar over all l when a bound status flips (svm.cpp:876-905): i first, then j
        const bool need_i = use_gbar && (((pi & 3) == ST_UPPER) != (sti == ST_UPPER));
        const bool need_j = use_gbar && (((pj & 3) == ST_UPPER) != (stj == ST_UPPER));
        if (need_i || need_j) {
            // Gbar -= C Q_i (was upper) / += C Q_i (became upper)  <=>  mbar += fl(c K_i), c = +/- y_i C
            const float *__restrict__ Ki = K + (size_t)col_i * ldk;
            const double ci = (((pi & 3) == ST_UPPER) == ((pi & F_YPOS) != 0)) ? Cc : -Cc;
            const double cj = (((pj & 3) == ST_UPPER) == ((pj & F_YPOS) != 0)) ? Cc : -Cc;
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                if (pos[k] != NOPOS) {
                    const bool act = pos[k] < active;
                    double gb = mGbar[k * NT + tid];
                    if (need_i) gb = __dadd_rn(gb, __dmul_rn(ci, widen(act ? kvi[k] : __ldg(Ki + colr[k]))));
                    if (need_j) gb = __dadd_rn(gb, __dmul_rn(cj, widen(act ? kvj[k] : __ldg(Kj + colr[k]))));
                    mGbar[k * NT + tid] = gb;
                }
            }
        }
        tick(8);
    }


    // ---------------- calculate_rho (svm.cpp:1131-1168): sequential float64 sum in ascending position ----------------
    cluster.sync();
#pragma unroll
    for (int k = 0; k < KPT; k++) {
        if (pos[k] < active) { gscratch[pos[k]] = -m[k]; scratch[pos[k]] = fl[k]; }     // y*G and flags, by position
    }
    __threadfence();
    cluster.sync();
    if (rank == 0 && tid == 0) {
        int nfree = 0;
        double ub = CUDART_INF, lb = -CUDART_INF, sum = 0;
#pragma unroll 8
        for (int t = 0; t < active; t++) {
            const int f = __ldcg(scratch + t);
            const double yG = __ldcg(gscratch + t);
            if ((f & 3) == ST_UPPER) { if (!(f & F_YPOS)) ub = fmin(ub, yG); else lb = fmax(lb, yG); }
            else if ((f & 3) == ST_LOWER) { if (f & F_YPOS) ub = fmin(ub, yG); else lb = fmax(lb, yG); }
            else { ++nfree; sum = __dadd_rn(sum, yG); }
        }
        *Pp->out_rho = nfree > 0 ? __ddiv_rn(sum, (double)nfree) : __ddiv_rn(__dadd_rn(ub, lb), 2.0);
    }
    // coefficients alpha*y scattered by dataset row (svm.cpp:922-925, :1641-1642); SV counts
    int nsv = 0, nbsv = 0;
    {
        double *__restrict__ coef = Pp->coef;
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            if (pos[k] != NOPOS) {
                const double av = alpha[k * NT + tid];
                coef[colr[k]] = (fl[k] & F_YPOS) ? av : -av;
                nsv += av > 0;
                nbsv += av >= Cc;
            }
        }
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        nsv += __shfl_xor_sync(0xffffffffu, nsv, s);
        nbsv += __shfl_xor_sync(0xffffffffu, nbsv, s);
    }
    __syncthreads();
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
    cluster.sync();                                              // no CTA may exit while a peer can still write its shared memory
}

template <int NT, int KPT, int CL, bool FAST, bool PROF>
cudaError_t launch_co(const SmoProblem *probs, const int *order, int n_prob, int lmax, cudaStream_t st)
{
    constexpr int LCAP = NT * KPT;
    // alpha + mbar per owned element, then 7 bytes per POSITION of cold-path scratch; padded to a whole SM's worth so a
    // cluster CTA never shares its SM (a resident small CTA would keep a full-SM single-CTA solver from being scheduled)
    size_t smem = (size_t)LCAP * 16 + (size_t)lmax * 7 + 64;
    const char *sh = getenv("B200GS_SMO_CO_SHARE");                         // development switch: let cluster CTAs share an SM
    if (!(sh && atoi(sh)) && smem < 160 * 1024) smem = 160 * 1024;
    auto kern = smo_colown_kernel<NT, KPT, CL, FAST, PROF>;
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
    return cudaLaunchKernelEx(&cfg, kern, probs, order);
}

template <int NT, int KPT, int CL>
cudaError_t launch_co_f(const SmoProblem *p, const int *o, int n, int lmax, bool fast, cudaStream_t st)
{
    const char *e = getenv("B200GS_SMO_PROF");
    const bool prof = e && atoi(e) != 0;                                    // development switch: per-phase cycle counters
    if (prof) return fast ? launch_co<NT, KPT, CL, true, true>(p, o, n, lmax, st) : launch_co<NT, KPT, CL, false, true>(p, o, n, lmax, st);
    return fast ? launch_co<NT, KPT, CL, true, false>(p, o, n, lmax, st) : launch_co<NT, KPT, CL, false, false>(p, o, n, lmax, st);
}

}  // namespace

// Largest sub-problem a column-owner cluster launch of size cl supports (0: unsupported cluster size)
int smo_colown_max_rows(int cl) { return cl == 2 ? 8192 : ((cl == 4 || cl == 8) ? 16384 : 0); }

// Shape = (threads per CTA) x (elements per thread) x (CTAs per problem).
cudaError_t launch_smo_colown(const SmoProblem *d_probs, const int *d_order, int n_prob, int lmax, int cl, bool fast, cudaStream_t st)
{
    if (n_prob <= 0) return cudaSuccess;
    if (lmax > smo_colown_max_rows(cl)) return cudaErrorInvalidValue;
    int nt = 0;
    if (const char *e = getenv("B200GS_SMO_NT")) nt = atoi(e);                  // development switch
    // 512 threads at most: the register-resident state needs more than the 64 registers a 1024-thread CTA leaves a thread
    if (cl == 2) return launch_co_f<512, 8, 2>(d_probs, d_order, n_prob, lmax, fast, st);
    if (cl == 4) {
        if (lmax > 8192) return launch_co_f<512, 8, 4>(d_probs, d_order, n_prob, lmax, fast, st);
        if (nt == 256) return launch_co_f<256, 8, 4>(d_probs, d_order, n_prob, lmax, fast, st);
        return launch_co_f<512, 4, 4>(d_probs, d_order, n_prob, lmax, fast, st);
    }
    if (cl == 8) {
        if (lmax > 8192) return launch_co_f<512, 4, 8>(d_probs, d_order, n_prob, lmax, fast, st);
        if (nt == 512) return launch_co_f<512, 2, 8>(d_probs, d_order, n_prob, lmax, fast, st);
        return launch_co_f<256, 4, 8>(d_probs, d_order, n_prob, lmax, fast, st);
    }
    return cudaErrorInvalidValue;
}

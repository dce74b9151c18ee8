// gemm_tc.cu -- C[M][N] = sum_k A[M][k] * B[N][k]  ("NT", both operands K-major) on the 5th-generation
// tensor cores: tcgen05.mma kind::tf32 with the accumulator in TMEM, operands staged by TMA
// (cp.async.bulk.tensor, SWIZZLE_128B) through a 3-stage mbarrier pipeline.  sm_100a only.
//
// fp32-faithful on TF32 tensor cores ("3xTF32"): every fp32 operand is pre-split into
//   hi = x with the low 13 mantissa bits cleared (exactly a TF32 value),  lo = x - hi (exact in fp32)
// and each K-step issues hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator.  The dropped
// lo*lo term and the truncation of lo to 11 bits are ~2^-22 relative, i.e. at the level of the fp32
// accumulation itself -- the precision scikit-learn's own fp32 BLAS calls have on this path
// (Ridge: sklearn/linear_model/_ridge.py:215-227 keeps float32 throughout).
//
// Replaces (reference path -> scikit-learn): safe_sparse_dot(X.T, X) of _solve_cholesky (_ridge.py:219),
// X @ w and X.T @ grad of the logistic objective (_linear_loss.py), and -- opt-in -- the libsvm kernel rows.
//
// The TMEM accumulator adds with truncation, so a long contraction drifts by ~sqrt(K)*2^-24 of sum|a||b| -- harmless
// for a Gram diagonal, ruinous for a heavily cancelled sum such as a gradient near its zero.  Callers therefore
// bound every accumulation chain to TC_KCHUNK terms (one TcBatch per K-chunk) and add the partial tiles in float64
// on the CUDA cores (launch_sum_partials).
//
// Tile: 128 x 128 x 32 (fp32) per stage.  PERSISTENT, warp-specialised CTAs (one per SM, 6 warps), tiles dealt round-robin:
//   warp 0 / lane 0: TMA producer (3-stage ring over ALL its tiles -- the ring never drains between tiles)
//   warp 1 / lane 0: MMA issuer (12 MMAs per stage) into one of TWO 128-column TMEM accumulators
//   warps 2-5: epilogue of the other accumulator: tcgen05.ld 32x32b.x32 -> padded shared-memory transpose -> COALESCED
//              global stores (a TMEM lane is a tile row: storing straight from registers touches 32 cache lines per request).
// The contractions of this package are short (K <= TC_KCHUNK = 512 per accumulation chain, see above), so a tile is
// ~16 k-steps: with one tile per CTA the barrier/TMEM set-up and the epilogue were half of its life (31 % tensor pipe,
// profiles/r01_gram_tc_ncu_summary.txt); now they overlap the next tile's main loop.
// symmetric (A == B, M == N): only tiles on or above the diagonal are computed and their transposes stored too.
// Batched: per-batch row offsets into A and B, K range and output pointer.
#include "common.cuh"
#include <cuda.h>
#include <cstdio>
#include <algorithm>

namespace {

constexpr int BM = 128, BN = 128, BK = 32;            // BK fp32 = 128 B = one SWIZZLE_128B atom row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;               // 16 KB per operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;           // A_hi, A_lo, B_hi, B_lo
constexpr int EPI_WARPS = 4, NTHREADS = 64 + EPI_WARPS * 32;
constexpr int STG_LD = 33;                            // padded row of the epilogue transpose buffers
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr uint32_t TMEM_COLS = 256;                   // two 128-column fp32 accumulators

// instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c_format F32 (1) @4, a/b format
// TF32 (2) @7/@10, K-major both, N>>3 @17, M>>4 @24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps the kernel instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute SmemDescriptor): start>>4 @0, LBO=1 @16,
// SBO = 8 rows * 128 B = 1024 B (>>4 = 64) @32, version 1 @46, layout_type 2 (SWIZZLE_128B) @61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// tile t of the launch -> (batch, m-tile, n-tile); symmetric: n-tile >= m-tile only
__device__ __forceinline__ void decode_tile(int t, int tiles_m, int tiles_n, int symmetric, int &z, int &mi, int &ni)
{
    const int per = symmetric ? tiles_m * (tiles_m + 1) / 2 : tiles_m * tiles_n;
    z = t / per;
    int r = t - z * per;
    if (!symmetric) { mi = r / tiles_n; ni = r - mi * tiles_n; return; }
    mi = 0;
    while (r >= tiles_m - mi) { r -= tiles_m - mi; mi++; }
    ni = mi + r;
}

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_nt_tf32x3_kernel(const __grid_constant__ CUtensorMap a_hi, const __grid_constant__ CUtensorMap a_lo,
                      const __grid_constant__ CUtensorMap b_hi, const __grid_constant__ CUtensorMap b_lo,
                      const TcBatch *__restrict__ batches, int n_batches, int M, int N, float alpha, int accumulate_c, int symmetric)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    float *stg_all = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES + STG_BYTES);   // full[S], empty[S], tfull[2], tempty[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
    const int n_tiles = n_batches * (symmetric ? tiles_m * (tiles_m + 1) / 2 : tiles_m * tiles_n);
    const uint32_t b_full = smem_u32(&bars[0]), b_empty = smem_u32(&bars[STAGES]), b_tfull = smem_u32(&bars[2 * STAGES]),
                   b_tempty = smem_u32(&bars[2 * STAGES + 2]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
        for (int q = 0; q < 2; q++) { mbar_init(b_tfull + 8 * q, 1); mbar_init(b_tempty + 8 * q, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer: one ring over all tiles of this CTA ----------------
            int it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int z, mi, ni;
                decode_tile(t, tiles_m, tiles_n, symmetric, z, mi, ni);
                const TcBatch bt = batches[z];
                const int nkb = (bt.k1 - bt.k0 + BK - 1) / BK;
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES;
                    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                    mbar_wait(b_empty + 8 * s, ph ^ 1u);                        // slot free
                    const uint32_t full = b_full + 8 * s;
                    mbar_expect_tx(full, STAGE_BYTES);
                    const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
                    const int kc = bt.k0 + kb * BK;
                    tma_load_2d(base, &a_hi, kc, bt.a_row0 + mi * BM, full);
                    tma_load_2d(base + TILE_BYTES, &a_lo, kc, bt.a_row0 + mi * BM, full);
                    tma_load_2d(base + 2 * TILE_BYTES, &b_hi, kc, bt.b_row0 + ni * BN, full);
                    tma_load_2d(base + 3 * TILE_BYTES, &b_lo, kc, bt.b_row0 + ni * BN, full);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer: accumulator buffers alternate between tiles ----------------
            int it = 0, ti = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ti++) {
                int z, mi, ni;
                decode_tile(t, tiles_m, tiles_n, symmetric, z, mi, ni);
                const TcBatch bt = batches[z];
                const int nkb = (bt.k1 - bt.k0 + BK - 1) / BK;
                const int buf = ti & 1;
                mbar_wait(b_tempty + 8 * buf, (((uint32_t)ti >> 1) & 1u) ^ 1u);      // the epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES;
                    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                    mbar_wait(b_full + 8 * s, ph);                               // TMA bytes landed
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t dah = make_desc(base), dal = make_desc(base + TILE_BYTES);
                    const uint64_t dbh = make_desc(base + 2 * TILE_BYTES), dbl = make_desc(base + 3 * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; k++) {                           // UMMA_K = 8 for tf32: 32 B per step
                        const uint64_t off = (uint64_t)(k * 32 >> 4);            // advance inside the swizzle atom
                        umma_tf32(d_tmem, dah + off, dbh + off, (kb | k) != 0);
                        umma_tf32(d_tmem, dah + off, dbl + off, 1);
                        umma_tf32(d_tmem, dal + off, dbh + off, 1);
                    }
                    umma_commit(b_empty + 8 * s);                                // frees the smem slot when the MMAs retire
                }
                umma_commit(b_tfull + 8 * buf);                                  // accumulator complete
            }
        }
    } else {
        // ---------------- epilogue warps: TMEM -> registers -> smem transpose -> coalesced global ----------------
        const int q = warp & 3;                                                  // the TMEM lane quarter this warp may read
        float *stg = stg_all + (size_t)(warp - 2) * 32 * STG_LD;
        int ti = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ti++) {
            int z, mi, ni;
            decode_tile(t, tiles_m, tiles_n, symmetric, z, mi, ni);
            const TcBatch bt = batches[z];
            const int nkb = (bt.k1 - bt.k0 + BK - 1) / BK;
            const int buf = ti & 1;
            mbar_wait(b_tfull + 8 * buf, ((uint32_t)ti >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m0 = mi * BM + q * 32, n0 = ni * BN;
            const bool diag = symmetric && mi == ni;       // diagonal tile: store j >= i, mirror the strict upper part (bitwise symmetric:
                                                           // the hi*lo and lo*hi products of C[i][j] and C[j][i] accumulate in swapped order)
            const bool mirror = symmetric != 0;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                if (nkb > 0) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                                 : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) r[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 32; j++) stg[lane * STG_LD + j] = alpha * __uint_as_float(r[j]);   // lane == tile row
                __syncwarp();
                const int col = n0 + c0 + lane;
                if (col < N) {
#pragma unroll 4
                    for (int rr = 0; rr < 32; rr++) {                            // a warp request = 32 consecutive columns of one row
                        const int row = m0 + rr;
                        if (row < M && !(diag && col < row)) {
                            float *dst = bt.c + (size_t)row * bt.ldc + col;
                            const float v = stg[rr * STG_LD + lane];
                            *dst = accumulate_c ? *dst + v : v;
                        }
                    }
                }
                if (mirror) {                                                    // C[n][m] = C[m][n]: columns become rows
                    const int mcol = m0 + lane;
                    if (mcol < M) {
#pragma unroll 4
                        for (int jj = 0; jj < 32; jj++) {
                            const int mrow = n0 + c0 + jj;
                            if (mrow < N && !(diag && mrow <= mcol)) {
                                float *dst = bt.c + (size_t)mrow * bt.ldc + mcol;
                                const float v = stg[lane * STG_LD + jj];
                                *dst = accumulate_c ? *dst + v : v;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) mbar_arrive(b_tempty + 8 * buf);                      // this warp has drained its quarter
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// hi = x with the low 13 mantissa bits cleared (a TF32 value), lo = x - hi (exact)
__global__ void split_tf32_kernel(const float *__restrict__ x, float *__restrict__ hi, float *__restrict__ lo, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        hi[i] = h;
        lo[i] = v - h;
    }
}

// out[i] = (float) sum_c partial[c][i] accumulated in float64 (round-to-nearest), i < per
__global__ void sum_partials_kernel(const float *__restrict__ partial, int n_chunks, int64_t per, float *__restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0;
        for (int c = 0; c < n_chunks; c++) s += (double)partial[(size_t)c * per + i];
        out[i] = (float)s;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace

// 2-D tensor map over a row-major fp32 matrix [rows][ld] exposing `cols` columns, box = BK x 128 rows, 128B swizzle
cudaError_t tc_make_map(TcMap *out, const float *base, int64_t rows, int64_t cols, int64_t ld)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    if ((ld * 4) % 16 != 0 || ((uintptr_t)base & 15)) return cudaErrorInvalidValue;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_sum_partials(const float *partial, int n_chunks, int64_t per, float *out, cudaStream_t st)
{
    sum_partials_kernel<<<(unsigned)std::min<int64_t>((per + 255) / 256, 1184), 256, 0, st>>>(partial, n_chunks, per, out);
    return cudaGetLastError();
}

cudaError_t launch_split_tf32(const float *x, float *hi, float *lo, size_t n, cudaStream_t st)
{
    split_tf32_kernel<<<592, 256, 0, st>>>(x, hi, lo, n);
    return cudaGetLastError();
}

cudaError_t launch_gemm_nt_tf32x3(const TcMap &a_hi, const TcMap &a_lo, const TcMap &b_hi, const TcMap &b_lo,
                                  const TcBatch *d_batches, int n_batches, int M, int N, float alpha, bool accumulate,
                                  cudaStream_t st, bool symmetric)
{
    // per device and per launch (cheap): a handle per GPU may launch this kernel from its own host thread
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (symmetric && M != N) return cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tm = (M + BM - 1) / BM, tn = (N + BN - 1) / BN;
    const long long tiles = (long long)n_batches * (symmetric ? (long long)tm * (tm + 1) / 2 : (long long)tm * tn);
    if (tiles <= 0) return cudaSuccess;
    const int grid = (int)std::min<long long>(tiles, sms);
    gemm_nt_tf32x3_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(*reinterpret_cast<const CUtensorMap *>(&a_hi),
                                                              *reinterpret_cast<const CUtensorMap *>(&a_lo),
                                                              *reinterpret_cast<const CUtensorMap *>(&b_hi),
                                                              *reinterpret_cast<const CUtensorMap *>(&b_lo), d_batches, n_batches, M, N, alpha,
                                                              accumulate ? 1 : 0, symmetric ? 1 : 0);
    return cudaGetLastError();
}

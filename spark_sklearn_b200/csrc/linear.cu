// linear.cu -- Ridge / LogisticRegression searches (C ABI entry points).
#include "common.cuh"
#include <vector>

extern "C" {

int gs_ridge(gs_handle *h, int32_t, const double *, int32_t, uint32_t, double *, double *, float *, float *)
{
    gs_set_error(h, "gs_ridge: not implemented in this build"); return GS_ERR_UNSUPPORTED;
}
int gs_ridge_refit(gs_handle *h, double, int32_t, double *)
{
    gs_set_error(h, "gs_ridge_refit: not implemented in this build"); return GS_ERR_UNSUPPORTED;
}
int gs_logreg(gs_handle *h, int32_t, const double *, double, int32_t, int32_t, uint32_t, double *, double *, int32_t *, float *, float *)
{
    gs_set_error(h, "gs_logreg: not implemented in this build"); return GS_ERR_UNSUPPORTED;
}
int gs_logreg_refit(gs_handle *h, double, double, int32_t, int32_t, double *, int32_t *)
{
    gs_set_error(h, "gs_logreg_refit: not implemented in this build"); return GS_ERR_UNSUPPORTED;
}

}

// ---- test hook: one tensor-core GEMM with host buffers ----
extern "C" int gs_debug_gemm_nt(gs_handle *h, const float *A, int32_t M, const float *B, int32_t N, int32_t K, float *C)
{
    if (!h || !A || !B || !C || M <= 0 || N <= 0 || K <= 0) return GS_ERR_ARG;
    GS_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int Kp = (K + 31) & ~31;                                   // zero-padded contraction length
    DevBuf a, ah, al, b, bh, bl, c, bt;
    GS_CUDA(a.reserve((size_t)M * Kp * 4)); GS_CUDA(ah.reserve((size_t)M * Kp * 4)); GS_CUDA(al.reserve((size_t)M * Kp * 4));
    GS_CUDA(b.reserve((size_t)N * Kp * 4)); GS_CUDA(bh.reserve((size_t)N * Kp * 4)); GS_CUDA(bl.reserve((size_t)N * Kp * 4));
    GS_CUDA(c.reserve((size_t)M * N * 4)); GS_CUDA(bt.reserve(sizeof(TcBatch)));
    GS_CUDA(cudaMemsetAsync(a.p, 0, (size_t)M * Kp * 4, st));
    GS_CUDA(cudaMemsetAsync(b.p, 0, (size_t)N * Kp * 4, st));
    GS_CUDA(cudaMemcpy2DAsync(a.p, (size_t)Kp * 4, A, (size_t)K * 4, (size_t)K * 4, M, cudaMemcpyHostToDevice, st));
    GS_CUDA(cudaMemcpy2DAsync(b.p, (size_t)Kp * 4, B, (size_t)K * 4, (size_t)K * 4, N, cudaMemcpyHostToDevice, st));
    GS_CUDA(launch_split_tf32(a.as<float>(), ah.as<float>(), al.as<float>(), (size_t)M * Kp, st));
    GS_CUDA(launch_split_tf32(b.as<float>(), bh.as<float>(), bl.as<float>(), (size_t)N * Kp, st));
    TcMap mah, mal, mbh, mbl;
    GS_CUDA(tc_make_map(&mah, ah.as<float>(), M, Kp, Kp)); GS_CUDA(tc_make_map(&mal, al.as<float>(), M, Kp, Kp));
    GS_CUDA(tc_make_map(&mbh, bh.as<float>(), N, Kp, Kp)); GS_CUDA(tc_make_map(&mbl, bl.as<float>(), N, Kp, Kp));
    TcBatch hb{0, 0, 0, Kp, c.as<float>(), (int64_t)N};
    GS_CUDA(cudaMemcpyAsync(bt.p, &hb, sizeof hb, cudaMemcpyHostToDevice, st));
    GS_CUDA(launch_gemm_nt_tf32x3(mah, mal, mbh, mbl, bt.as<TcBatch>(), 1, M, N, 1.0f, false, st));
    GS_CUDA(cudaMemcpyAsync(C, c.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost, st));
    GS_CUDA(cudaStreamSynchronize(st));
    a.release(); ah.release(); al.release(); b.release(); bh.release(); bl.release(); c.release(); bt.release();
    return GS_OK;
}

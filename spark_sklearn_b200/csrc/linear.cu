// linear.cu -- Ridge / LogisticRegression searches (C ABI entry points).
#include "common.cuh"

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

// fused_ridge.cu -- engine instantiations for the FRidge device functor (heavy geometry).
#include "dispatch.h"

#define LIST_(F, f) VB_CASE_DX(F, f, 8) VB_CASE_D(F, f, 2) VB_CASE_D(F, f, 4) VB_CASE_D(F, f, 6) VB_CASE_D(F, f, 8) \
    VB_CASE_D(F, f, 10) VB_CASE_D(F, f, 12) VB_CASE_D(F, f, 16) VB_CASE_D(F, f, 20)

int launch_fused_ridge_heavy(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st)
{
    const FRidge& f = *(const FRidge*)functor;
    VB_DISPATCH_D(FRidge, f, LIST_);
}

int eval_ridge(const void* functor, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st)
{
    const FRidge& f = *(const FRidge*)functor;
    VB_EVAL_D(FRidge, f, 4) VB_EVAL_D(FRidge, f, 8) VB_EVAL_D(FRidge, f, 12) VB_EVAL_D(FRidge, f, 20)
    return -22;
}

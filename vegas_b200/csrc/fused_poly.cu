// fused_poly.cu -- engine instantiations for the FPoly device functor (heavy geometry).
#include "dispatch.h"

#define LIST_(F, f) VB_CASE_D(F, f, 2) VB_CASE_D(F, f, 4) VB_CASE_D(F, f, 6) VB_CASE_D(F, f, 8) \
    VB_CASE_D(F, f, 10) VB_CASE_D(F, f, 12) VB_CASE_D(F, f, 16) VB_CASE_D(F, f, 20)

int launch_fused_poly_heavy(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st)
{
    const FPoly& f = *(const FPoly*)functor;
    VB_DISPATCH_D(FPoly, f, LIST_);
}

int eval_poly(const void* functor, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st)
{
    const FPoly& f = *(const FPoly*)functor;
    VB_EVAL_D(FPoly, f, 4) VB_EVAL_D(FPoly, f, 8) VB_EVAL_D(FPoly, f, 12) VB_EVAL_D(FPoly, f, 20)
    return -22;
}

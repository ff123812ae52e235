// fused_genz.cu -- engine instantiations for the FGenz device functor (heavy geometry).
#include "dispatch.h"

#define LIST_(F, f) VB_CASE_D(F, f, 2) VB_CASE_D(F, f, 4) VB_CASE_D(F, f, 6) VB_CASE_D(F, f, 8) \
    VB_CASE_D(F, f, 10) VB_CASE_D(F, f, 12) VB_CASE_D(F, f, 16) VB_CASE_D(F, f, 20)

int launch_fused_genz_heavy(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st)
{
    const FGenz& f = *(const FGenz*)functor;
    VB_DISPATCH_D(FGenz, f, LIST_);
}

int eval_genz(const void* functor, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st)
{
    const FGenz& f = *(const FGenz*)functor;
    VB_EVAL_D(FGenz, f, 4) VB_EVAL_D(FGenz, f, 8) VB_EVAL_D(FGenz, f, 12) VB_EVAL_D(FGenz, f, 20)
    return -22;
}

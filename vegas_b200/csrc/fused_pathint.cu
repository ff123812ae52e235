// fused_pathint.cu -- engine instantiations for the lattice path-integral functor (vector-valued).
#include "dispatch.h"

#define LIST_(F, f) VB_CASE_D(F, f, 8) VB_CASE_D(F, f, 10) VB_CASE_D(F, f, 12) VB_CASE_D(F, f, 16)

template <int NX0>
static int launch_nx0(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st)
{
    const FPathInt<NX0>& f = *(const FPathInt<NX0>*)functor;
    VB_DISPATCH_D(FPathInt<NX0>, f, LIST_);
}

int launch_fused_pathint(const EngineP& p, const void* functor, int nx0, LaunchCfg& cfg, cudaStream_t st)
{
    switch (nx0) {
    case 0: return launch_nx0<0>(p, functor, cfg, st);
    case 6: return launch_nx0<6>(p, functor, cfg, st);
    default: return -22;
    }
}

template <int NX0>
static int eval_nx0(const void* functor, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st)
{
    const FPathInt<NX0>& f = *(const FPathInt<NX0>*)functor;
    VB_EVAL_D(FPathInt<NX0>, f, 8) VB_EVAL_D(FPathInt<NX0>, f, 10) VB_EVAL_D(FPathInt<NX0>, f, 16)
    return -22;
}

int eval_pathint(const void* functor, int nx0, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st)
{
    switch (nx0) {
    case 0: return eval_nx0<0>(functor, dim, x, rows, out, sm_count, st);
    case 6: return eval_nx0<6>(functor, dim, x, rows, out, sm_count, st);
    default: return -22;
    }
}

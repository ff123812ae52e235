// fused_light.cu -- engine instantiations of the light geometry (two 256-thread CTAs per SM, grid and
// histogram windows shared by all its warps) for the cheap built-in integrands, and the
// light-or-heavy choice of every family.
// one copy of the exp table here: the 15 KB the conflict-free 16-copy layout costs are worth more as
// histogram / grid windows to these sampler-bound kernels (B200, N = 1 ridge: 5.96 ms against 6.20 ms)
#define VB_EXP_COPIES 1
#include "dispatch.h"

int launch_fused_poly_heavy(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st);
int launch_fused_gaussmix_heavy(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st);
int launch_fused_ridge_heavy(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st);
int launch_fused_genz_heavy(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st);

#define LIGHT_(F, f) VB_CASE_LX(F, f, 4) VB_CASE_LX(F, f, 8) VB_CASE_LX(F, f, 10) VB_CASE_L(F, f, 4) VB_CASE_L(F, f, 8) VB_CASE_L(F, f, 10) VB_CASE_L(F, f, 16) VB_CASE_L(F, f, 20)

template <class F>
static int light_or(const EngineP& p, const F& f, LaunchCfg& cfg, cudaStream_t st)
{
    VB_DISPATCH_D(F, f, LIGHT_);
}

#define FAMILY_(name, F, FL)                                                                       \
    int launch_fused_##name(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st)  \
    {                                                                                              \
        if (cfg.light) {                                                                           \
            FL fl{*(const F*)functor};                                                             \
            int g = light_or<FL>(p, fl, cfg, st);                                                  \
            if (g != -22 && g != -23 && g != -24) return g;                                        \
            cfg.light = false;                     /* not compiled / not applicable: heavy */      \
        }                                                                                          \
        return launch_fused_##name##_heavy(p, functor, cfg, st);                                   \
    }

struct FPolyL : FPoly {};
struct FGaussMixL : FGaussMix {};
struct FGenzL : FGenz {};
FAMILY_(poly, FPoly, FPolyL)
FAMILY_(gaussmix, FGaussMix, FGaussMixL)
FAMILY_(ridge, FRidge, FRidgeLight)
FAMILY_(genz, FGenz, FGenzL)

// reduce_buffer.cu -- engine instantiations for the unfused path: w and f come from HBM buffers.
#include "dispatch.h"

int launch_buffer(const EngineP& p, int nf, LaunchCfg& cfg, cudaStream_t st)
{
    switch (nf) {
#define C_(N) case N: { BufferSrc<N> s_; return launch_engine(p, s_, cfg, st); }
    C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8)
#undef C_
    default: return -22;
    }
}

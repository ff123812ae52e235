// ctx.h -- the library's context object and the helpers its translation units share.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/vegas_b200.h"
#include "dispatch.h"

// error reporting: message kept per thread for vb200_last_error(); returns `code`
int vb_fail(int code, const char* fmt, ...);
#define fail vb_fail
#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return fail(-2, "%s: %s", #call, cudaGetErrorString(e_));       \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct vb200_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_per_sm = 0, smem_per_block_optin = 0;
    int last_grid = 0, last_bps = 0, last_wtot = 0, last_nt = 0, last_ch = 0;
    bool very_light = false;                              // ... and only a handful of flops per sample (launch_engine: staging capacity)
    bool light_hint = false;                              // the integrand is cheap: prefer the light engine geometry      // geometry of the most recent engine launch
    int64_t last_smem = 0;
    uint64_t seed = 0;
    PhiloxKey key;
    // map
    bool have_map = false;
    MapP map;
    DevBuf grid;
    // strata
    bool have_strata = false;
    StrataP st;
    int64_t cstride[VB_MAXD];
    int64_t nchunks = 0;
    // plan
    bool have_plan = false;
    AllocP al;
    int64_t plan_total = 0, plan_min = 0, plan_max = 0;
    int64_t plan_max_chunk = 0, plan_items = 0;
    bool plan_light = false;                              // the launched pre-pass also planned the light geometry's chunks
    int64_t nsuper = 0, plan_super_items = -1;            // light geometry: VB_LCH-cube chunks and their items (-1: not planned)
    DevBuf super_items, super_item_off;
    DevBuf chunk_tot, chunk_off, chunk_items, item_off, stats;
    // the exclusive scans are made when first used: the fused path needs none of them unless a chunk was split
    bool chunk_off_valid = false, item_off_valid = false, super_off_valid = false;
    std::vector<long long> chunk_off_host;   // fetched lazily by the unfused path
    std::vector<long long> item_off_host;    // same (only when some chunk was split: plan_items != nchunks)
    // integrand
    int fid = -1, nf = 0, nx0 = 0;
    std::vector<char> functor;        // host copy of the functor struct
    DevBuf fparams;                   // device arrays the functor points to
    // scratch
    DevBuf partials, scratch, counter;
    DevBuf ecounter;                  // the engine's work-item counter (zero between launches: k_finalize resets it)
    DevBuf sigf_shadow;               // engine output of sigf while chunks are split into items (see run_engine)
    // vb200_iteration_begin/_end: pinned, device-mapped landing zone of the iteration head
    void* head_host = nullptr;
    void* head_dev = nullptr;
    int64_t head_pending = 0;         // words the iteration in flight will deliver (0: none)
    int64_t launches = 0;
};
#define VB_HEAD_WORDS 128

// work items of a local chunk range, per geometry (see set_items in vegas_b200.cu)
struct ItemsSel { const int64_t* off[2]; int64_t begin[2], end[2]; };
int vb_set_items(vb200_ctx* c, int64_t chunk_begin, int64_t chunk_end, ItemsSel& it, cudaStream_t st);
// row offsets of the chunks on the device (made on first use after a plan) / fetched to the host
int vb_ensure_chunk_off(vb200_ctx* c, cudaStream_t st);
int vb_fetch_chunk_off(vb200_ctx* c, cudaStream_t st = 0);

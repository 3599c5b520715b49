// blend_bwd.cuh -- declarations shared by the two implementations of loop A of the backward
// (blend_bwd.cu: butterfly reduction per (warp, splat); blend_bwd_transposed.cu: splat-per-lane accumulation).
#pragma once
#include "common.cuh"

namespace gsb {

struct BlendBwdParams {
    int H, W, tiles_x;
    const int *tile_start;
    const int *tile_end;
    const int *sorted_vals;
    const float4 *records;
    const float *grad_image;
    const float *acc_alpha;
    const int *last_effective;
    float *accum;      // rows of 12 floats
    float *mag_image;  // (H,W,2)
};

#ifdef GSB_HOST_EMU  // tests/simt: the kernels compiled as host C++ under a lock-step SIMT emulator
__device__ __forceinline__ float ex2_approx_b(float x) { return exp2f(x); }
__device__ __forceinline__ float rcp_approx(float x) { return 1.0f / x; }
__device__ __forceinline__ float sqrt_approx(float x) { return sqrtf(x); }
#else
__device__ __forceinline__ float ex2_approx_b(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {  // MUFU.RCP, <= 1 ulp
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {  // MUFU.RSQ based, ~1 ulp
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif

// Work counters for the CPU-side emulation only (tests/simt, scripts/emu_work_stats.py); nothing in a device build.
#ifdef GSB_HOST_EMU
#define GSB_EMU_COUNT(slot, n) (simt_emu::counters()[slot] += (long long)(n))
#else
#define GSB_EMU_COUNT(slot, n) ((void)0)
#endif
enum EmuCounter {
    EC_BF_VISITS = 0,       // butterfly kernel: (warp, splat) visits
    EC_BF_VISITS_ANY = 1,   //   ... with at least one contributing pixel (these pay the butterfly + RED)
    EC_BF_PAIRS = 2,        //   contributing (pixel, splat) pairs
    EC_TB_SPLATS = 3,       // transposed kernel: (warp, splat) list entries
    EC_TB_CHUNKS = 4,       //   chunks processed
    EC_TB_ROWS = 5,         //   accumulator rows flushed (splats with a contributing pixel)
    EC_BATCHES = 6,         // staging batches (per CTA)
};

int launch_blend_backward_transposed(const BlendBwdParams &p, int tiles, bool exact_exp, bool stats,
                                     cudaStream_t stream);

}  // namespace gsb

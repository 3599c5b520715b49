// blend_fwd.cu -- per-tile front-to-back alpha blending (replaces gaussian_point_rasterisation,
// GPCR:318-485, with get_point_probability_density_from_conic_and_rescale, UT:275-284).
//
// One CTA per 16x16 tile, one pixel per thread; a warp owns an 8x4 pixel patch so that the
// alpha < 1/255 rejection and the saturation exit stay warp-coherent.  The tile's splat list is
// streamed through shared memory in batches of 256 packed 48-byte records (3 x float4, gathered by
// the sorted in-camera offsets); the inner loop reads them as broadcasts.  The CTA leaves the list as
// soon as every pixel has saturated (__syncthreads_and) -- the reference walks the whole list
// (GPCR:387-394).  Compute-bound (FP32 + MUFU.EX2), not HBM-bound: 48 B per (tile, splat) are reused
// by 256 pixels.
#include "common.cuh"

namespace gsb {

struct BlendFwdParams {
    int H, W, tiles_x;
    const int *tile_start;
    const int *tile_end;
    const int *sorted_vals;
    const float4 *records;
    float *image;
    float *depth;
    float *acc_alpha;
    int *last_effective;
    int *valid_count;
};

#ifdef GSB_HOST_EMU  // tests/simt: host build under the SIMT emulator
__device__ __forceinline__ float ex2_approx(float x) { return exp2f(x); }
#else
__device__ __forceinline__ float ex2_approx(float x) {
    // one MUFU.EX2; rel. error ~2^-22
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif

#ifndef GSB_FWD_MIN_BLOCKS
#define GSB_FWD_MIN_BLOCKS 5
#endif
template <bool RGB_ONLY, bool EXACT_EXP>
__global__ void __launch_bounds__(GSB_TILE_PIXELS, GSB_FWD_MIN_BLOCKS)
blend_forward_kernel(const BlendFwdParams p) {
    // double-buffered staging area: [buf][plane][splat]; planes: u v a b | c rescale opacity depth | r g b radius
    __shared__ float4 s_rec[2 * 3 * GSB_TILE_PIXELS];
    constexpr int PLANE = GSB_TILE_PIXELS * 16;  // bytes between the three record planes
    __shared__ unsigned int s_bits[2][8][8];  // [buf][consumer warp patch][loader warp] -> splats that can reach it

    const int tile = blockIdx.x;
    const int tu = tile % p.tiles_x, tv = tile / p.tiles_x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // warp w covers the 8x4 patch at ((w & 1) * 8, (w >> 1) * 4)
    const int pu = tu * GSB_TILE_WIDTH + (warp & 1) * 8 + (lane & 7);
    const int pv = tv * GSB_TILE_HEIGHT + (warp >> 1) * 4 + (lane >> 3);
    const float px = (float)pu + 0.5f, py = (float)pv + 0.5f;  // GPCR:442 pixel centre
    const float tile_x0 = (float)(tu * GSB_TILE_WIDTH), tile_y0 = (float)(tv * GSB_TILE_HEIGHT);
    const int start = p.tile_start[tile], end = p.tile_end[tile];

    // T is the working transmittance: it is forced to 0 once the pixel has saturated, which makes every
    // later splat fail the T(1-a) >= 1e-4 test without a separate flag; Tlive keeps the value to output.
    float T = 1.0f, Tlive = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, D = 0.0f, Wt = 0.0f;
    int last = start, cnt = 0;
    const unsigned int sa = smem_u32(s_rec);

    // One barrier per batch: batch k is staged into buffer k&1 while slower warps may still be blending
    // batch k-1 from the other buffer; passing barrier k implies everybody is done with batch k-1.
    int buf = 0;
    for (int base = start; base < end; base += GSB_TILE_PIXELS, buf ^= 1) {
        float4 *const s_r0 = s_rec + buf * 3 * GSB_TILE_PIXELS;
        float4 *const s_r1 = s_r0 + GSB_TILE_PIXELS, *const s_r2 = s_r0 + 2 * GSB_TILE_PIXELS;
        const int idx = base + tid;
        unsigned int mask = 0;
        if (idx < end) {
            const int o = __ldg(&p.sorted_vals[idx]);
            const float4 *rec = p.records + 3 * (size_t)o;
            const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1);
            if (EXACT_EXP) {
                s_r0[tid] = r0;
                s_r1[tid] = r1;
            } else {
                // fast path: fold -1/2, log2(e) and rescale*opacity into the staged record, so that the
                // inner loop is  alpha = ex2(A dx^2 + C dy^2 + B dx dy) * ro
                constexpr float L2E = 1.4426950408889634f;
                s_r0[tid] = make_float4(r0.x, r0.y, -0.5f * L2E * r0.z, -L2E * r0.w);
                s_r1[tid] = make_float4(-0.5f * L2E * r1.x, r1.y * r1.z, 0.0f, r1.w);
            }
            s_r2[tid] = __ldg(rec + 2);
            mask = splat_patch_mask(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y * r1.z, tile_x0, tile_y0);
        }
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const unsigned int bits = __ballot_sync(0xffffffffu, (mask >> w) & 1u);
            if (lane == 0) s_bits[buf][w][warp] = bits;
        }
        if (__syncthreads_and(T == 0.0f)) break;  // staging visible + tile-level early exit
        if (__all_sync(0xffffffffu, T == 0.0f)) continue;  // whole patch saturated: only help with loads
        const unsigned int sb = sa + buf * (3 * PLANE);
#pragma unroll 1
        for (int lw = 0; lw < 8; ++lw) {
            unsigned int bits = s_bits[buf][warp][lw];
            while (bits) {
                const int j = lw * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                if (lane == 0) GSB_EMU_COUNT(EC_FW_VISITS, 1);
                const unsigned int ja = sb + j * 16;
                const float4 r0 = lds128<0>(ja);      // u v a b      (fast: u v A B)
                const float4 r1 = lds128<PLANE>(ja);  // c rescale opacity depth (fast: C ro - depth)
                const float dx = px - r0.x, dy = py - r0.y;
                float alpha;
                if (EXACT_EXP) {  // the reference's op order (UT:275-284)
                    const float power = -0.5f * (dx * dx * r0.z + dy * dy * r1.x) - dx * dy * r0.w;
                    alpha = expf(power) * r1.y * r1.z;
                } else {
                    alpha = ex2_approx(dx * (r0.z * dx + r0.w * dy) + r1.x * dy * dy) * r1.y;
                }
                if (!(alpha < 1.0f / 255.0f)) {             // GPCR:451 (same comparison as the reference)
                    GSB_EMU_COUNT(EC_FW_PAIRS, 1);
                    alpha = fminf(alpha, 0.99f);            // GPCR:453
                    const float nT = T * (1.0f - alpha);
                    if (nT >= 0.0001f) {
                        last = base + j + 1;
                        const float4 r2 = lds128<2 * PLANE>(ja);
                        const float wgt = alpha * T;
                        if (EXACT_EXP) {
                            C0 += r2.x * alpha * T;
                            C1 += r2.y * alpha * T;
                            C2 += r2.z * alpha * T;
                            if (!RGB_ONLY) D += r1.w * alpha * T;
                        } else {
                            C0 = fmaf(r2.x, wgt, C0);
                            C1 = fmaf(r2.y, wgt, C1);
                            C2 = fmaf(r2.z, wgt, C2);
                            if (!RGB_ONLY) D = fmaf(r1.w, wgt, D);
                        }
                        if (!RGB_ONLY) {
                            Wt += wgt;
                            cnt += 1;
                        }
                        T = nT;
                        Tlive = nT;
                    } else {
                        T = 0.0f;  // GPCR:457-460: saturated; this splat is NOT blended
                    }
                }
            }
        }
    }
    const size_t pix = (size_t)pv * p.W + pu;
    p.image[3 * pix] = C0;
    p.image[3 * pix + 1] = C1;
    p.image[3 * pix + 2] = C2;
    if (!RGB_ONLY) {
        p.depth[pix] = D / fmaxf(Wt, 1e-6f);  // GPCR:479-480
        p.acc_alpha[pix] = 1.0f - Tlive;
        p.last_effective[pix] = last;
        p.valid_count[pix] = cnt;
    }
}

#ifndef GSB_HOST_EMU
int launch_blend_forward(const GsbForwardArgs &a, const Workspace &ws, cudaStream_t stream) {
    const GsbWorkspaceLayout &L = ws.layout;
    BlendFwdParams p;
    p.H = a.camera_height;
    p.W = a.camera_width;
    p.tiles_x = a.camera_width / GSB_TILE_WIDTH;
    p.tile_start = ws.tile_start;
    p.tile_end = ws.tile_end;
    p.sorted_vals = (L.sort_passes % 2) == 1 ? ws.vals_b : ws.vals_a;
    p.records = ws.records;
    p.image = a.rasterized_image;
    p.depth = a.rasterized_depth;
    p.acc_alpha = a.pixel_accumulated_alpha;
    p.last_effective = a.pixel_offset_of_last_effective_point;
    p.valid_count = a.pixel_valid_point_count;
    const int tiles = p.tiles_x * (a.camera_height / GSB_TILE_HEIGHT);
    if (tiles <= 0) return GSB_OK;
    const bool exact = (a.flags & GSB_FLAG_EXACT_EXP) != 0;
    if (a.rgb_only) {
        if (exact) blend_forward_kernel<true, true><<<tiles, GSB_TILE_PIXELS, 0, stream>>>(p);
        else blend_forward_kernel<true, false><<<tiles, GSB_TILE_PIXELS, 0, stream>>>(p);
    } else {
        if (exact) blend_forward_kernel<false, true><<<tiles, GSB_TILE_PIXELS, 0, stream>>>(p);
        else blend_forward_kernel<false, false><<<tiles, GSB_TILE_PIXELS, 0, stream>>>(p);
    }
    GSB_CUDA_CHECK(cudaGetLastError());
    return GSB_OK;
}
#endif  // GSB_HOST_EMU

}  // namespace gsb

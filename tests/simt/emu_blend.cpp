// emu_blend.cpp -- the forward blend and loop A of the backward (both implementations) compiled as host C++ under simt_emu.h.
// TEST INFRASTRUCTURE, see simt_emu.h; built by tests/test_simt_blend_cpu.py with g++.
#include "simt_emu.h"
// the kernel sources, unmodified (their launchers are compiled out under GSB_HOST_EMU)
#include "../../taichi_3d_gaussian_splatting_b200/csrc/blend_fwd.cu"
#include "../../taichi_3d_gaussian_splatting_b200/csrc/blend_bwd.cu"
#include "../../taichi_3d_gaussian_splatting_b200/csrc/blend_bwd_transposed.cu"

namespace gsb {
void set_error(const char *, ...) {}
}  // namespace gsb

extern "C" long long emu_blend_backward(int transposed, int exact_exp, int stats, int H, int W, const int *tile_start,
                                        const int *tile_end, const int *sorted_vals, const float *records,
                                        const float *grad_image, const float *acc_alpha, const int *last_effective,
                                        float *accum, float *mag_image) {
    using namespace gsb;
    BlendBwdParams p;
    p.H = H;
    p.W = W;
    p.tiles_x = W / GSB_TILE_WIDTH;
    p.tile_start = tile_start;
    p.tile_end = tile_end;
    p.sorted_vals = sorted_vals;
    p.records = reinterpret_cast<const float4 *>(records);
    p.grad_image = grad_image;
    p.acc_alpha = acc_alpha;
    p.last_effective = last_effective;
    p.accum = accum;
    p.mag_image = mag_image;
    const int tiles = p.tiles_x * (H / GSB_TILE_HEIGHT);
    simt_emu::M().switches = 0;
    if (!transposed) {
        if (exact_exp && stats) simt_emu::launch(blend_backward_kernel<true, true>, tiles, GSB_TILE_PIXELS, p);
        else if (exact_exp) simt_emu::launch(blend_backward_kernel<true, false>, tiles, GSB_TILE_PIXELS, p);
        else if (stats) simt_emu::launch(blend_backward_kernel<false, true>, tiles, GSB_TILE_PIXELS, p);
        else simt_emu::launch(blend_backward_kernel<false, false>, tiles, GSB_TILE_PIXELS, p);
    } else if (exact_exp) {
        if (stats) simt_emu::launch(blend_backward_transposed_kernel<true, true>, tiles, GSB_TILE_PIXELS, p);
        else simt_emu::launch(blend_backward_transposed_kernel<true, false>, tiles, GSB_TILE_PIXELS, p);
    } else {
        if (stats) simt_emu::launch(blend_backward_transposed_kernel<false, true>, tiles, GSB_TILE_PIXELS, p);
        else simt_emu::launch(blend_backward_transposed_kernel<false, false>, tiles, GSB_TILE_PIXELS, p);
    }
    return simt_emu::M().switches;
}

extern "C" long long emu_blend_forward(int rgb_only, int exact_exp, int H, int W, const int *tile_start, const int *tile_end,
                                       const int *sorted_vals, const float *records, float *image, float *depth,
                                       float *acc_alpha, int *last_effective, int *valid_count) {
    using namespace gsb;
    BlendFwdParams p;
    p.H = H;
    p.W = W;
    p.tiles_x = W / GSB_TILE_WIDTH;
    p.tile_start = tile_start;
    p.tile_end = tile_end;
    p.sorted_vals = sorted_vals;
    p.records = reinterpret_cast<const float4 *>(records);
    p.image = image;
    p.depth = depth;
    p.acc_alpha = acc_alpha;
    p.last_effective = last_effective;
    p.valid_count = valid_count;
    const int tiles = p.tiles_x * (H / GSB_TILE_HEIGHT);
    simt_emu::M().switches = 0;
    if (rgb_only) {
        if (exact_exp) simt_emu::launch(blend_forward_kernel<true, true>, tiles, GSB_TILE_PIXELS, p);
        else simt_emu::launch(blend_forward_kernel<true, false>, tiles, GSB_TILE_PIXELS, p);
    } else {
        if (exact_exp) simt_emu::launch(blend_forward_kernel<false, true>, tiles, GSB_TILE_PIXELS, p);
        else simt_emu::launch(blend_forward_kernel<false, false>, tiles, GSB_TILE_PIXELS, p);
    }
    return simt_emu::M().switches;
}

// loop B + P4 (per-point chain rule, SH-band masking, gradient factors): backward_points_kernel of blend_bwd.cu; with
// grad_sum / grad_col non-null the COMPACT instantiation (view-parallel exchange) instead of the dense one
extern "C" long long emu_backward_points(long long N, const int *point_offset, const float *records,
                                         const float *point_in_camera, const float *accum, const float *poses,
                                         const float *xyz, const float *features, const int *obj_id,
                                         const float *t_pc_cam, const float *K, int color_max_sh_band, float q_f, float s_f,
                                         float a_f, float c_f, float h_f, float *grad_xyz, float *grad_feat, float *grad_sum,
                                         float *grad_col, int *ctl_num_in_camera, int *ctl_num_pixels, float *ctl_vs_grad,
                                         float *ctl_vs_grad_avg, float *ctl_pos_grad, float *ctl_pos_grad_norm) {
    using namespace gsb;
    PointsBwdParams p;
    p.N = N;
    p.point_offset = point_offset;
    p.records = reinterpret_cast<const float4 *>(records);
    p.point_in_camera = point_in_camera;
    p.accum = accum;
    p.poses = reinterpret_cast<const PoseBlock *>(poses);
    p.xyz = xyz;
    p.features = features;
    p.obj_id = obj_id;
    p.t_pc_cam = t_pc_cam;
    p.K = K;
    const int band = color_max_sh_band;
    p.first_cleared = band <= 0 ? 1 : band == 1 ? 4 : band == 2 ? 9 : 16;  // as launch_backward_points
    p.q_f = q_f;
    p.s_f = s_f;
    p.a_f = a_f;
    p.c_f = c_f;
    p.h_f = h_f;
    p.grad_xyz = grad_xyz;
    p.grad_feat = grad_feat;
    p.grad_sum_compact = grad_sum;
    p.grad_color_compact = grad_col;
    p.ctl_num_in_camera = ctl_num_in_camera;
    p.ctl_num_pixels = ctl_num_pixels;
    p.ctl_vs_grad = ctl_vs_grad;
    p.ctl_vs_grad_avg = ctl_vs_grad_avg;
    p.ctl_pos_grad = ctl_pos_grad;
    p.ctl_pos_grad_norm = ctl_pos_grad_norm;
    p.skip_flag = nullptr;
    simt_emu::M().switches = 0;
    const int blocks = (int)std::min<long long>((N + GSB_POINTS_THREADS - 1) / GSB_POINTS_THREADS, 16 * 148);
    if (N > 0) {
        if (grad_sum) simt_emu::launch(backward_points_kernel<true>, blocks, GSB_POINTS_THREADS, p);
        else simt_emu::launch(backward_points_kernel<false>, blocks, GSB_POINTS_THREADS, p);
    }
    return simt_emu::M().switches;
}

// gsb200_expand_view_gradients: dense gradients of a batch of views from the exchanged compact rows
extern "C" long long emu_expand_view_gradients(long long N, int R, const float *grad_sum, const float *grad_color_views,
                                               long long view_stride, const float *xyz, const int *obj_id,
                                               int color_max_sh_band, float c_f, float h_f, float *grad_xyz, float *grad_feat,
                                               int part) {
    using namespace gsb;
    ExpandParams p;
    p.N = N;
    p.R = R;
    p.grad_sum = grad_sum;
    p.grad_color_views = grad_color_views;
    p.view_stride = view_stride;
    p.xyz = xyz;
    p.obj_id = obj_id;
    const int band = color_max_sh_band;
    p.first_cleared = band <= 0 ? 1 : band == 1 ? 4 : band == 2 ? 9 : 16;
    p.c_f = c_f;
    p.h_f = h_f;
    p.grad_xyz = grad_xyz;
    p.grad_feat = grad_feat;
    simt_emu::M().switches = 0;
    const int blocks = (int)std::min<long long>((N + GSB_POINTS_THREADS - 1) / GSB_POINTS_THREADS, 16 * 148);
    if (N > 0) {
        if (part == 1) simt_emu::launch(expand_view_gradients_kernel<1>, blocks, GSB_POINTS_THREADS, p);
        else if (part == 2) simt_emu::launch(expand_view_gradients_kernel<2>, blocks, GSB_POINTS_THREADS, p);
        else simt_emu::launch(expand_view_gradients_kernel<0>, blocks, GSB_POINTS_THREADS, p);
    }
    return simt_emu::M().switches;
}

// Loop A of the backward on the tiles [tile_first, tile_last) only, with the work counters of blend_bwd.cuh (EmuCounter)
// returned in counters_out[8]: lets scripts/emu_work_stats.py shard a full-size frame over processes.
extern "C" void emu_blend_backward_stats(int transposed, int stats, int H, int W, int tile_first, int tile_last,
                                         const int *tile_start, const int *tile_end, const int *sorted_vals,
                                         const float *records, const float *grad_image, const float *acc_alpha,
                                         const int *last_effective, float *accum, float *mag_image, long long *counters_out) {
    using namespace gsb;
    BlendBwdParams p;
    p.H = H;
    p.W = W;
    p.tiles_x = W / GSB_TILE_WIDTH;
    p.tile_start = tile_start;
    p.tile_end = tile_end;
    p.sorted_vals = sorted_vals;
    p.records = reinterpret_cast<const float4 *>(records);
    p.grad_image = grad_image;
    p.acc_alpha = acc_alpha;
    p.last_effective = last_effective;
    p.accum = accum;
    p.mag_image = mag_image;
    const int tiles = p.tiles_x * (H / GSB_TILE_HEIGHT);
    for (int k = 0; k < 16; ++k) simt_emu::counters()[k] = 0;
    if (!transposed) {
        if (stats) simt_emu::launch_range(blend_backward_kernel<false, true>, tiles, tile_first, tile_last, GSB_TILE_PIXELS, p);
        else simt_emu::launch_range(blend_backward_kernel<false, false>, tiles, tile_first, tile_last, GSB_TILE_PIXELS, p);
    } else {
        if (stats) simt_emu::launch_range(blend_backward_transposed_kernel<false, true>, tiles, tile_first, tile_last, GSB_TILE_PIXELS, p);
        else simt_emu::launch_range(blend_backward_transposed_kernel<false, false>, tiles, tile_first, tile_last, GSB_TILE_PIXELS, p);
    }
    for (int k = 0; k < 8; ++k) counters_out[k] = simt_emu::counters()[k];
}

// The forward blend on the tiles [tile_first, tile_last) only, with the work counters (EmuCounter) in counters_out[16].
extern "C" void emu_blend_forward_stats(int H, int W, int tile_first, int tile_last, const int *tile_start, const int *tile_end,
                                        const int *sorted_vals, const float *records, float *image, float *depth,
                                        float *acc_alpha, int *last_effective, int *valid_count, long long *counters_out) {
    using namespace gsb;
    BlendFwdParams p;
    p.H = H;
    p.W = W;
    p.tiles_x = W / GSB_TILE_WIDTH;
    p.tile_start = tile_start;
    p.tile_end = tile_end;
    p.sorted_vals = sorted_vals;
    p.records = reinterpret_cast<const float4 *>(records);
    p.image = image;
    p.depth = depth;
    p.acc_alpha = acc_alpha;
    p.last_effective = last_effective;
    p.valid_count = valid_count;
    const int tiles = p.tiles_x * (H / GSB_TILE_HEIGHT);
    for (int k = 0; k < 16; ++k) simt_emu::counters()[k] = 0;
    simt_emu::launch_range(blend_forward_kernel<false, false>, tiles, tile_first, tile_last, GSB_TILE_PIXELS, p);
    for (int k = 0; k < 16; ++k) counters_out[k] = simt_emu::counters()[k];
}

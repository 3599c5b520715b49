// emu_image_loss.cpp -- the fused image loss (csrc/image_loss.cu: clamp + L1 + D-SSIM, forward and gradient) compiled as
// host C++ under simt_emu.h.  TEST INFRASTRUCTURE, see simt_emu.h.
#include "simt_emu.h"
#include "../../taichi_3d_gaussian_splatting_b200/csrc/image_loss.cu"

extern "C" long long emu_image_loss_temp_bytes(int H, int W) { return gsb::image_loss_layout(H, W).total; }

extern "C" long long emu_image_loss(const float *pred_hwc, const float *gt_chw, int H, int W, float lambda_value,
                                    float upstream, float *loss_out3, float *grad_hwc, void *temp) {
    using namespace gsb;
    ImageLossParams p;
    ImageLossLayout L;
    image_loss_params(pred_hwc, gt_chw, H, W, lambda_value, upstream, loss_out3, grad_hwc, temp, &p, &L);
    simt_emu::M().switches = 0;
    p.tiles_x = L.tiles_mx;
    p.tiles_y = L.tiles_my;
    simt_emu::launch(ssim_map_kernel, 3 * L.tiles_mx * L.tiles_my, IL_THREADS, p);
    p.tiles_x = L.tiles_ix;
    p.tiles_y = L.tiles_iy;
    simt_emu::launch(image_loss_grad_kernel, 3 * L.tiles_ix * L.tiles_iy, IL_THREADS, p);
    return simt_emu::M().switches;
}

// the fused clamp + L1 loss + gradient kernel (csrc/loss.cu, gsb200_l1_loss)
#include "../../taichi_3d_gaussian_splatting_b200/csrc/loss.cu"

extern "C" long long emu_l1_loss_temp_bytes() { return (long long)(4 + gsb::L1_MAX_BLOCKS) * 4; }

extern "C" void emu_l1_loss(const float *pred, const float *target, long long n, int clamp01, float upstream, float *loss,
                            float *grad, void *temp) {
    using namespace gsb;
    long long blocks = 1;
    const L1Params p = l1_params(pred, target, n, clamp01, upstream, loss, grad, temp, &blocks);
    simt_emu::launch(l1_loss_kernel, (int)blocks, L1_THREADS, p);
}

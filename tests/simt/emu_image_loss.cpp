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

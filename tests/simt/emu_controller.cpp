// emu_controller.cpp -- the controller's fused accumulator update (csrc/controller.cu) under simt_emu.h.  TEST INFRASTRUCTURE.
#include "simt_emu.h"
#include "../../taichi_3d_gaussian_splatting_b200/csrc/controller.cu"

extern "C" void emu_controller_update(const int *ids, long long M, const int *num_pixels, const float *magnitude,
                                      const float *grad_xyz, int *acc_num_in_camera, int *acc_num_pixels, float *acc_vs_grad,
                                      float *acc_vs_grad_avg, float *acc_pos_grad, float *acc_pos_grad_norm, int blocks) {
    using namespace gsb;
    ControllerUpdateParams p{ids, M, num_pixels, magnitude, grad_xyz, acc_num_in_camera, acc_num_pixels, acc_vs_grad,
                             acc_vs_grad_avg, acc_pos_grad, acc_pos_grad_norm};
    if (M > 0) simt_emu::launch(controller_update_kernel, blocks, CU_THREADS, p);
}

// emu_adam.cpp -- the fused Adam step (csrc/adam.cu) compiled as host C++ under simt_emu.h.  TEST INFRASTRUCTURE.
#include "simt_emu.h"
#include "../../taichi_3d_gaussian_splatting_b200/csrc/adam.cu"

extern "C" void emu_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, double lr,
                              double beta1, double beta2, double eps, int step, int blocks) {
    using namespace gsb;
    const AdamParams p = adam_params(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step);
    simt_emu::launch(adam_step_kernel, blocks, ADAM_THREADS, p);
}

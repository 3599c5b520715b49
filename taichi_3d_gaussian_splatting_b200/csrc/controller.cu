// controller.cu -- the per-iteration accumulator update of the densification controller in ONE kernel:
// GaussianPointAdaptiveController.update, GaussianPointAdaptiveController.py:130-143, runs inside every backward (it is
// the rasteriser's backward hook) as six indexed read-modify-writes plus a division, a NaN fix-up and a row norm -- ~15
// torch launches over the M in-camera points.  The in-camera ids are unique (GPCR:861-864), so the scatter needs no atomics.
#include "common.cuh"

namespace gsb {

struct ControllerUpdateParams {
    const int *ids;           // point_id_in_camera_list (M)
    long long M;
    const int *num_pixels;    // num_affected_pixels (M)
    const float *magnitude;   // magnitude_grad_viewspace (M)
    const float *grad_xyz;    // grad_point_in_camera (M,3)
    int *acc_num_in_camera;   // (N)
    int *acc_num_pixels;      // (N)
    float *acc_vs_grad;       // accumulated_view_space_position_gradients (N)
    float *acc_vs_grad_avg;   // accumulated_view_space_position_gradients_avg (N)
    float *acc_pos_grad;      // accumulated_position_gradients (N,3)
    float *acc_pos_grad_norm; // accumulated_position_gradients_norm (N)
};

constexpr int CU_THREADS = 256;
__global__ void __launch_bounds__(CU_THREADS) controller_update_kernel(const ControllerUpdateParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.M; i += stride) {
        const long long id = p.ids[i];
        const int npix = p.num_pixels[i];
        const float mag = p.magnitude[i];
        const float gx = p.grad_xyz[3 * i], gy = p.grad_xyz[3 * i + 1], gz = p.grad_xyz[3 * i + 2];
        p.acc_num_in_camera[id] += 1;                 // :133
        p.acc_num_pixels[id] += npix;                 // :134
        p.acc_vs_grad[id] += mag;                     // :135-136
        float avg = mag / (float)npix;                // :137-139  (0/0 -> NaN -> 0; x/0 stays inf like the reference)
        if (avg != avg) avg = 0.0f;
        p.acc_vs_grad_avg[id] += avg;
        p.acc_pos_grad[3 * id] += gx;                 // :140-141
        p.acc_pos_grad[3 * id + 1] += gy;
        p.acc_pos_grad[3 * id + 2] += gz;
        p.acc_pos_grad_norm[id] += sqrtf(gx * gx + gy * gy + gz * gz);  // :142-143
    }
}

}  // namespace gsb

#ifndef GSB_HOST_EMU
extern "C" int gsb200_controller_update(const int32_t *point_id_in_camera_list, int64_t num_points_in_camera,
                                        const int32_t *num_affected_pixels, const float *magnitude_grad_viewspace,
                                        const float *grad_point_in_camera, int32_t *accumulated_num_in_camera,
                                        int32_t *accumulated_num_pixels, float *accumulated_view_space_position_gradients,
                                        float *accumulated_view_space_position_gradients_avg,
                                        float *accumulated_position_gradients, float *accumulated_position_gradients_norm,
                                        void *stream) {
    using namespace gsb;
    if (num_points_in_camera < 0) {
        set_error("controller_update: negative point count");
        return GSB_EINVAL;
    }
    if (num_points_in_camera == 0) return GSB_OK;
    if (!point_id_in_camera_list || !num_affected_pixels || !magnitude_grad_viewspace || !grad_point_in_camera ||
        !accumulated_num_in_camera || !accumulated_num_pixels || !accumulated_view_space_position_gradients ||
        !accumulated_view_space_position_gradients_avg || !accumulated_position_gradients ||
        !accumulated_position_gradients_norm) {
        set_error("controller_update: null pointer argument");
        return GSB_EINVAL;
    }
    ControllerUpdateParams p;
    p.ids = point_id_in_camera_list;
    p.M = num_points_in_camera;
    p.num_pixels = num_affected_pixels;
    p.magnitude = magnitude_grad_viewspace;
    p.grad_xyz = grad_point_in_camera;
    p.acc_num_in_camera = accumulated_num_in_camera;
    p.acc_num_pixels = accumulated_num_pixels;
    p.acc_vs_grad = accumulated_view_space_position_gradients;
    p.acc_vs_grad_avg = accumulated_view_space_position_gradients_avg;
    p.acc_pos_grad = accumulated_position_gradients;
    p.acc_pos_grad_norm = accumulated_position_gradients_norm;
    long long blocks = (num_points_in_camera + CU_THREADS - 1) / CU_THREADS;
    const long long cap = 16LL * num_sms();
    if (blocks > cap) blocks = cap;
    controller_update_kernel<<<(int)blocks, CU_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
    GSB_CUDA_CHECK(cudaGetLastError());
    return GSB_OK;
}
#endif  // GSB_HOST_EMU

// adam.cu -- one-kernel Adam step for the scene's two parameter tensors: the reference trainer keeps two
// torch.optim.Adam instances, features (N,56) and positions (N,3), betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad
// (GaussianPointTrainer.py:126-129, stepped at :176-177).  Same arithmetic as torch's single-tensor Adam:
//   m += (g - m) (1 - b1);  v = v b2 + (1 - b2) g g;  p += -(lr / (1 - b1^t)) * (m / (sqrt(v) / sqrt(1 - b2^t) + eps))
// HBM-bound: 16 B read + 12 B written per element, float4 accesses; the bias corrections are computed on the host in double
// like torch does.
#include "common.cuh"

namespace gsb {

struct AdamParams {
    float *param;
    const float *grad;
    float *exp_avg;
    float *exp_avg_sq;
    long long n;
    float one_minus_beta1, beta2, one_minus_beta2, eps;
    float neg_step_size;         // -lr / (1 - beta1^t)
    float bias_correction2_sqrt; // sqrt(1 - beta2^t)
    const long long *skip_flag;  // optional device flag: non-zero = no-op (fused train step after a key-capacity overflow)
};

__device__ __forceinline__ void adam_one(const AdamParams &p, float &w, float g, float &m, float &v) {
    m = m + (g - m) * p.one_minus_beta1;
    v = v * p.beta2 + p.one_minus_beta2 * g * g;
    const float denom = sqrtf(v) / p.bias_correction2_sqrt + p.eps;
    w = w + p.neg_step_size * (m / denom);
}

constexpr int ADAM_THREADS = 256;
__global__ void __launch_bounds__(ADAM_THREADS) adam_step_kernel(const AdamParams p) {
    if (p.skip_flag != nullptr && *p.skip_flag != 0) return;
    const long long n4 = p.n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float4 *w4 = reinterpret_cast<float4 *>(p.param);
    const float4 *g4 = reinterpret_cast<const float4 *>(p.grad);
    float4 *m4 = reinterpret_cast<float4 *>(p.exp_avg);
    float4 *v4 = reinterpret_cast<float4 *>(p.exp_avg_sq);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 w = w4[i], m = m4[i], v = v4[i];
        const float4 g = __ldg(g4 + i);
        adam_one(p, w.x, g.x, m.x, v.x);
        adam_one(p, w.y, g.y, m.y, v.y);
        adam_one(p, w.z, g.z, m.z, v.z);
        adam_one(p, w.w, g.w, m.w, v.w);
        w4[i] = w;
        m4[i] = m;
        v4[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {  // tail (n not a multiple of 4)
        const long long i = (n4 << 2) + threadIdx.x;
        adam_one(p, p.param[i], p.grad[i], p.exp_avg[i], p.exp_avg_sq[i]);
    }
}

// lr / betas / eps arrive as doubles, like the Python floats torch's Adam works with: 1 - beta is formed in double and
// rounded once (1 - 0.999 -> 0.001f; forming it from a float beta would be off by 1e-5 relative).
static inline AdamParams adam_params(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, double lr,
                                     double beta1, double beta2, double eps, int step) {
    AdamParams p;
    p.param = param;
    p.grad = grad;
    p.exp_avg = exp_avg;
    p.exp_avg_sq = exp_avg_sq;
    p.n = n;
    p.one_minus_beta1 = (float)(1.0 - beta1);
    p.beta2 = (float)beta2;
    p.one_minus_beta2 = (float)(1.0 - beta2);
    p.eps = (float)eps;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    p.neg_step_size = (float)(-lr / bc1);
    p.bias_correction2_sqrt = (float)sqrt(bc2);
    p.skip_flag = nullptr;
    return p;
}

#ifndef GSB_HOST_EMU
int launch_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, double lr, double beta1,
                     double beta2, double eps, int step, const long long *skip_flag, cudaStream_t stream) {
    if (n <= 0) return GSB_OK;
    AdamParams p = adam_params(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step);
    p.skip_flag = skip_flag;
    long long blocks = ((n >> 2) + ADAM_THREADS - 1) / ADAM_THREADS;
    const long long cap = 16LL * num_sms();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_step_kernel<<<(int)blocks, ADAM_THREADS, 0, stream>>>(p);
    GSB_CUDA_CHECK(cudaGetLastError());
    return GSB_OK;
}
#endif

}  // namespace gsb

#ifndef GSB_HOST_EMU
extern "C" int gsb200_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t num_elements,
                                double lr, double beta1, double beta2, double eps, int32_t step, void *stream) {
    using namespace gsb;
    if (num_elements < 0 || step < 1 || (num_elements > 0 && (!param || !grad || !exp_avg || !exp_avg_sq))) {
        set_error("adam_step: bad arguments (n=%lld, step=%d)", (long long)num_elements, step);
        return GSB_EINVAL;
    }
    if (num_elements == 0) return GSB_OK;
    if (reinterpret_cast<uintptr_t>(param) % 16 || reinterpret_cast<uintptr_t>(grad) % 16 ||
        reinterpret_cast<uintptr_t>(exp_avg) % 16 || reinterpret_cast<uintptr_t>(exp_avg_sq) % 16) {
        set_error("adam_step: pointers must be 16-byte aligned");
        return GSB_EINVAL;
    }
    return launch_adam_step(param, grad, exp_avg, exp_avg_sq, num_elements, lr, beta1, beta2, eps, step, nullptr,
                            static_cast<cudaStream_t>(stream));
}
#endif  // GSB_HOST_EMU

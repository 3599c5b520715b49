// image_loss.cu -- the trainer's whole image loss and its gradient in two kernels:
//   pred = clamp(rasterised image, 0, 1)                                  GaussianPointTrainer.py:168-170
//   L    = (1 - lambda) * mean|pred - gt| + lambda * (1 - SSIM(pred, gt))  LossFunction.py:20-38
// SSIM is the published pytorch_msssim algorithm the reference calls (LossFunction.py:4,31; the package is absent here,
// loss.py restates it in torch and is the parity target): 11-tap Gaussian window, sigma 1.5, applied separably with VALID
// padding, K1 = 0.01, K2 = 0.03, data_range 1, mean over the (H-10) x (W-10) x 3 map.  torch autograd runs ~60 small
// kernels for this (5 grouped convolutions forward, their transposes backward, ~25 elementwise); here
//   kernel 1: one CTA per 16 x 16 tile of the SSIM map and channel: 26 x 26 input patches -> shared memory, separable
//             filter of {x, y, x^2, y^2, xy}, SSIM value (block-reduced), and the three derivative maps
//             d ssim/d{mu_x, E[x^2], E[xy]} (the filter outputs the prediction enters linearly);
//   kernel 2: one CTA per 16 x 16 tile of the image and channel: the transposed separable filter of the derivative maps,
//             + the L1 term, through the clamp, written straight into the (H, W, 3) gradient of the rasterised image;
//             its last CTA adds the partial sums in block order (deterministic) and writes {L, L1, 1 - SSIM}.
// Image layouts are the trainer's own: prediction (H, W, 3) as the rasteriser returns it, ground truth (3, H, W) as the
// dataset yields it (no permute / contiguous copies).
#include "common.cuh"

namespace gsb {

constexpr int IL_TILE = 16;
constexpr int IL_WIN = 11;
constexpr int IL_HALO = IL_WIN - 1;          // 10
constexpr int IL_PATCH = IL_TILE + IL_HALO;  // 26
constexpr int IL_THREADS = IL_TILE * IL_TILE;

struct ImageLossParams {
    const float *pred_hwc;  // (H, W, 3) rasterised image, unclamped
    const float *gt_chw;    // (3, H, W)
    int H, W, Hm, Wm;       // image and SSIM-map sizes (Hm = H - 10, Wm = W - 10)
    int tiles_x, tiles_y;   // of the current kernel's grid (per channel)
    float win[IL_WIN];
    float c1, c2;
    float l1_scale;         // upstream * (1 - lambda) / (3 H W)
    float ssim_scale;       // -upstream * lambda / (3 Hm Wm)
    float lambda_value;
    float *dmaps;           // [3 maps][3 channels][Hm][Wm]: d ssim / d mu_x, d E[x^2], d E[xy]
    double *ssim_partials;  // one per CTA of kernel 1
    double *l1_partials;    // one per CTA of kernel 2
    int n_ssim_partials;
    unsigned int *ticket;
    float *grad_hwc;        // may be null (loss only)
    float *loss_out;        // {L, L1, 1 - SSIM}
};

__device__ __forceinline__ double il_block_sum(double v, double *s_part) {
    // fixed-order sum over the 256 threads: warp butterflies, then the 8 warp totals in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < IL_THREADS / 32; ++w) t += s_part[w];
    return t;
}

__global__ void __launch_bounds__(IL_THREADS) ssim_map_kernel(const ImageLossParams p) {
    __shared__ float s_x[IL_PATCH][IL_PATCH + 1];
    __shared__ float s_y[IL_PATCH][IL_PATCH + 1];
    __shared__ float s_h[5][IL_PATCH][IL_TILE + 1];  // horizontally filtered x, y, x^2, y^2, xy
    __shared__ double s_part[IL_THREADS / 32];
    const int tid = threadIdx.x;
    const int per_channel = p.tiles_x * p.tiles_y;
    const int ch = blockIdx.x / per_channel, t = blockIdx.x - ch * per_channel;
    const int i0 = (t / p.tiles_x) * IL_TILE, j0 = (t % p.tiles_x) * IL_TILE;  // map = image coordinates of the window's corner
    for (int e = tid; e < IL_PATCH * IL_PATCH; e += IL_THREADS) {
        const int r = e / IL_PATCH, c = e - r * IL_PATCH;
        const int gr = i0 + r, gc = j0 + c;
        float x = 0.0f, y = 0.0f;
        if (gr < p.H && gc < p.W) {
            x = fminf(fmaxf(__ldg(&p.pred_hwc[((size_t)gr * p.W + gc) * 3 + ch]), 0.0f), 1.0f);
            y = __ldg(&p.gt_chw[((size_t)ch * p.H + gr) * p.W + gc]);
        }
        s_x[r][c] = x;
        s_y[r][c] = y;
    }
    __syncthreads();
    for (int e = tid; e < IL_PATCH * IL_TILE; e += IL_THREADS) {
        const int r = e / IL_TILE, c = e - r * IL_TILE;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f, a4 = 0.0f;
#pragma unroll
        for (int k = 0; k < IL_WIN; ++k) {
            const float w = p.win[k], x = s_x[r][c + k], y = s_y[r][c + k];
            a0 = fmaf(w, x, a0);
            a1 = fmaf(w, y, a1);
            a2 = fmaf(w, x * x, a2);
            a3 = fmaf(w, y * y, a3);
            a4 = fmaf(w, x * y, a4);
        }
        s_h[0][r][c] = a0;
        s_h[1][r][c] = a1;
        s_h[2][r][c] = a2;
        s_h[3][r][c] = a3;
        s_h[4][r][c] = a4;
    }
    __syncthreads();
    const int li = tid / IL_TILE, lj = tid - li * IL_TILE;
    const int i = i0 + li, j = j0 + lj;
    double mine = 0.0;
    if (i < p.Hm && j < p.Wm) {
        float mu1 = 0.0f, mu2 = 0.0f, exx = 0.0f, eyy = 0.0f, exy = 0.0f;
#pragma unroll
        for (int k = 0; k < IL_WIN; ++k) {
            const float w = p.win[k];
            mu1 = fmaf(w, s_h[0][li + k][lj], mu1);
            mu2 = fmaf(w, s_h[1][li + k][lj], mu2);
            exx = fmaf(w, s_h[2][li + k][lj], exx);
            eyy = fmaf(w, s_h[3][li + k][lj], eyy);
            exy = fmaf(w, s_h[4][li + k][lj], exy);
        }
        const float s1 = exx - mu1 * mu1, s2 = eyy - mu2 * mu2, s12 = exy - mu1 * mu2;
        const float A1 = 2.0f * mu1 * mu2 + p.c1, B1 = mu1 * mu1 + mu2 * mu2 + p.c1;
        const float A2 = 2.0f * s12 + p.c2, B2 = s1 + s2 + p.c2;
        const float l = A1 / B1, cs = A2 / B2;
        mine = (double)(l * cs);
        if (p.dmaps) {
            const float inv_b2 = 1.0f / B2;
            const float d_exx = -l * cs * inv_b2;
            const float d_exy = 2.0f * l * inv_b2;
            const float d_mu = 2.0f * cs * (mu2 - l * mu1) / B1 + 2.0f * l * inv_b2 * (mu1 * cs - mu2);
            const size_t plane = (size_t)p.Hm * p.Wm, at = ((size_t)ch * p.Hm + i) * p.Wm + j;
            p.dmaps[at] = d_mu;
            p.dmaps[3 * plane + at] = d_exx;
            p.dmaps[6 * plane + at] = d_exy;
        }
    }
    const double total = il_block_sum(mine, s_part);
    if (tid == 0) p.ssim_partials[blockIdx.x] = total;
}

__global__ void __launch_bounds__(IL_THREADS) image_loss_grad_kernel(const ImageLossParams p) {
    __shared__ float s_d[3][IL_PATCH][IL_PATCH + 1];    // derivative maps around the tile (zero outside the map)
    __shared__ float s_h[3][IL_PATCH][IL_TILE + 1];
    __shared__ double s_part[IL_THREADS / 32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const int per_channel = p.tiles_x * p.tiles_y;
    const int ch = blockIdx.x / per_channel, t = blockIdx.x - ch * per_channel;
    const int r0 = (t / p.tiles_x) * IL_TILE, c0 = (t % p.tiles_x) * IL_TILE;
    const int li = tid / IL_TILE, lj = tid - li * IL_TILE;
    const int r = r0 + li, c = c0 + lj;
    const bool inside = r < p.H && c < p.W;
    float x = 0.0f, y = 0.0f;
    if (inside) {
        x = __ldg(&p.pred_hwc[((size_t)r * p.W + c) * 3 + ch]);
        y = __ldg(&p.gt_chw[((size_t)ch * p.H + r) * p.W + c]);
    }
    const bool pass = x >= 0.0f && x <= 1.0f;  // torch.clamp hands the gradient through on [min, max]
    const float xc = fminf(fmaxf(x, 0.0f), 1.0f);
    const float e = xc - y;
    float g = 0.0f;
    if (p.grad_hwc) {
        // pixel (r, c) sits at tap (a, b) of the windows whose corner is (r - a, c - b): patch row u <-> map row r0 - 10 + u
        const size_t plane = (size_t)p.Hm * p.Wm;
        for (int q = tid; q < IL_PATCH * IL_PATCH; q += IL_THREADS) {
            const int u = q / IL_PATCH, v = q - u * IL_PATCH;
            const int mi = r0 - IL_HALO + u, mj = c0 - IL_HALO + v;
            const bool ok = mi >= 0 && mi < p.Hm && mj >= 0 && mj < p.Wm;
            const size_t at = ok ? ((size_t)ch * p.Hm + mi) * p.Wm + mj : 0;
            s_d[0][u][v] = ok ? p.dmaps[at] : 0.0f;
            s_d[1][u][v] = ok ? p.dmaps[3 * plane + at] : 0.0f;
            s_d[2][u][v] = ok ? p.dmaps[6 * plane + at] : 0.0f;
        }
        __syncthreads();
        for (int q = tid; q < IL_PATCH * IL_TILE; q += IL_THREADS) {
            const int u = q / IL_TILE, v = q - u * IL_TILE;
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
            for (int b = 0; b < IL_WIN; ++b) {  // map column (c0 + v) - b  <->  patch column v + 10 - b
                const float w = p.win[b];
                a0 = fmaf(w, s_d[0][u][v + IL_HALO - b], a0);
                a1 = fmaf(w, s_d[1][u][v + IL_HALO - b], a1);
                a2 = fmaf(w, s_d[2][u][v + IL_HALO - b], a2);
            }
            s_h[0][u][v] = a0;
            s_h[1][u][v] = a1;
            s_h[2][u][v] = a2;
        }
        __syncthreads();
        float f_mu = 0.0f, f_xx = 0.0f, f_xy = 0.0f;
#pragma unroll
        for (int a = 0; a < IL_WIN; ++a) {
            const float w = p.win[a];
            f_mu = fmaf(w, s_h[0][li + IL_HALO - a][lj], f_mu);
            f_xx = fmaf(w, s_h[1][li + IL_HALO - a][lj], f_xx);
            f_xy = fmaf(w, s_h[2][li + IL_HALO - a][lj], f_xy);
        }
        const float d_ssim = f_mu + 2.0f * xc * f_xx + y * f_xy;
        const float d_l1 = e > 0.0f ? p.l1_scale : (e < 0.0f ? -p.l1_scale : 0.0f);
        g = pass ? fmaf(p.ssim_scale, d_ssim, d_l1) : 0.0f;
        if (inside) p.grad_hwc[((size_t)r * p.W + c) * 3 + ch] = g;
    }
    const double total = il_block_sum(inside ? (double)fabsf(e) : 0.0, s_part);
    if (tid == 0) {
        p.l1_partials[blockIdx.x] = total;
        __threadfence();
        s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && tid == 0) {  // the last CTA to finish: both partial arrays are complete; fixed summation order
        __threadfence();
        double l1 = 0.0, ss = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b) l1 += ((volatile double *)p.l1_partials)[b];
        for (int b = 0; b < p.n_ssim_partials; ++b) ss += ((volatile double *)p.ssim_partials)[b];
        const double l1_mean = l1 / (3.0 * (double)p.H * (double)p.W);
        const double d_ssim = 1.0 - ss / (3.0 * (double)p.Hm * (double)p.Wm);
        p.loss_out[0] = (float)((1.0 - (double)p.lambda_value) * l1_mean + (double)p.lambda_value * d_ssim);
        p.loss_out[1] = (float)l1_mean;
        p.loss_out[2] = (float)d_ssim;
        *p.ticket = 0u;  // ready for the next call on this temp buffer
    }
}

// temp layout: [ticket: 16 B][ssim partials][l1 partials][derivative maps]
struct ImageLossLayout {
    int tiles_mx, tiles_my, tiles_ix, tiles_iy;
    long long off_ssim, off_l1, off_maps, total;
};
static inline ImageLossLayout image_loss_layout(int H, int W) {
    ImageLossLayout L;
    const int Hm = H - IL_HALO, Wm = W - IL_HALO;
    L.tiles_mx = (Wm + IL_TILE - 1) / IL_TILE;
    L.tiles_my = (Hm + IL_TILE - 1) / IL_TILE;
    L.tiles_ix = (W + IL_TILE - 1) / IL_TILE;
    L.tiles_iy = (H + IL_TILE - 1) / IL_TILE;
    L.off_ssim = 16;
    L.off_l1 = L.off_ssim + 8LL * 3 * L.tiles_mx * L.tiles_my;
    L.off_maps = (L.off_l1 + 8LL * 3 * L.tiles_ix * L.tiles_iy + 255) / 256 * 256;
    L.total = L.off_maps + 4LL * 9 * Hm * Wm;
    return L;
}

// The window of loss.py::_gaussian_window in float32: exp(-(k - 5)^2 / (2 sigma^2)), normalised.
static inline void image_loss_window(float *win) {
    float sum = 0.0f;
    for (int k = 0; k < IL_WIN; ++k) {
        const float d = (float)(k - IL_WIN / 2);
        win[k] = expf(-(d * d) / (2.0f * 1.5f * 1.5f));
        sum += win[k];
    }
    for (int k = 0; k < IL_WIN; ++k) win[k] /= sum;
}

static inline bool image_loss_params(const float *pred, const float *gt, int H, int W, float lambda_value, float upstream,
                                     float *loss_out, float *grad, void *temp, ImageLossParams *p, ImageLossLayout *L) {
    *L = image_loss_layout(H, W);
    p->pred_hwc = pred;
    p->gt_chw = gt;
    p->H = H;
    p->W = W;
    p->Hm = H - IL_HALO;
    p->Wm = W - IL_HALO;
    image_loss_window(p->win);
    p->c1 = 0.01f * 0.01f;
    p->c2 = 0.03f * 0.03f;
    p->l1_scale = upstream * (1.0f - lambda_value) / (3.0f * (float)H * (float)W);
    p->ssim_scale = -upstream * lambda_value / (3.0f * (float)p->Hm * (float)p->Wm);
    p->lambda_value = lambda_value;
    char *base = static_cast<char *>(temp);
    p->ticket = reinterpret_cast<unsigned int *>(base);
    p->ssim_partials = reinterpret_cast<double *>(base + L->off_ssim);
    p->l1_partials = reinterpret_cast<double *>(base + L->off_l1);
    p->dmaps = grad ? reinterpret_cast<float *>(base + L->off_maps) : nullptr;
    p->n_ssim_partials = 3 * L->tiles_mx * L->tiles_my;
    p->grad_hwc = grad;
    p->loss_out = loss_out;
    return true;
}

}  // namespace gsb

#ifndef GSB_HOST_EMU
extern "C" {

int64_t gsb200_image_loss_temp_bytes(int32_t camera_height, int32_t camera_width) {
    if (camera_height <= gsb::IL_HALO || camera_width <= gsb::IL_HALO) return 0;
    return gsb::image_loss_layout(camera_height, camera_width).total;
}

int gsb200_image_loss(const float *rasterized_image, const float *ground_truth_image, int32_t camera_height,
                      int32_t camera_width, float lambda_value, float upstream_grad, float *loss_out3,
                      float *grad_rasterized_image, void *temp, int64_t temp_bytes, void *stream) {
    using namespace gsb;
    if (!rasterized_image || !ground_truth_image || !loss_out3 || !temp || camera_height <= IL_HALO ||
        camera_width <= IL_HALO || temp_bytes < gsb200_image_loss_temp_bytes(camera_height, camera_width) ||
        reinterpret_cast<uintptr_t>(temp) % 16) {
        set_error("image_loss: bad arguments (H=%d W=%d temp_bytes=%lld; images must exceed the 11-tap window)",
                  camera_height, camera_width, (long long)temp_bytes);
        return GSB_EINVAL;
    }
    ImageLossParams p;
    ImageLossLayout L;
    image_loss_params(rasterized_image, ground_truth_image, camera_height, camera_width, lambda_value, upstream_grad,
                      loss_out3, grad_rasterized_image, temp, &p, &L);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    p.tiles_x = L.tiles_mx;
    p.tiles_y = L.tiles_my;
    ssim_map_kernel<<<3 * L.tiles_mx * L.tiles_my, IL_THREADS, 0, st>>>(p);
    GSB_CUDA_CHECK(cudaGetLastError());
    p.tiles_x = L.tiles_ix;
    p.tiles_y = L.tiles_iy;
    image_loss_grad_kernel<<<3 * L.tiles_ix * L.tiles_iy, IL_THREADS, 0, st>>>(p);
    GSB_CUDA_CHECK(cudaGetLastError());
    return GSB_OK;
}

}  // extern "C"
#endif  // GSB_HOST_EMU

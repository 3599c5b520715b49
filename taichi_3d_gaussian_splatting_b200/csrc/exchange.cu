// exchange.cu -- the collectives of the compact view-parallel gradient exchange (parallel.py) as ONE hand-written kernel over
// NVLink / NVSwitch multicast memory (NVLS), instead of an NCCL all-reduce followed by an NCCL all-gather:
//   * the (N,12) summable columns: every rank owns 1/R of the rows, pulls their sum over all ranks with
//     multimem.ld_reduce (the switch adds the R copies in flight) and pushes the result back to all ranks with multimem.st
//     (the switch replicates the store) -- a two-shot all-reduce without staging buffers or protocol flags;
//   * the per-view block [3N colour-argument gradients | 3 n_obj camera centres]: every rank pushes its own block into
//     slot `rank` of every rank's buffer with multimem.st -- an all-gather with one store stream per rank.
// Per GPU 48 MB/R + 12 MB leave and 48 MB + 96 MB arrive at N = 1e6, R = 8 (the NVLS minimum for this exchange); nothing is
// copied through intermediate buffers.  The two halves can be launched separately (`phases`): a rank pushes its block as soon as its
// own per-point kernel has finished -- BEFORE the barrier, into a buffer alternating with the step parity so that no peer can
// still be reading it -- which lets the early ranks' pushes run under the slowest rank's compute; only the all-reduce needs
// every rank's rows and follows the barrier.  The buffers live in ONE symmetric allocation (torch.distributed._symmetric_memory:
// same offset on every rank, mapped into a multicast address); the caller brackets the kernel with two cross-rank barriers
// (all compact rows written / all multicast stores landed).  sm_90+ PTX (multimem.*); SASS shows them as multimem ops.
#include "common.cuh"

namespace gsb {

struct MultimemExchangeParams {
    float *mc_sum;             // multicast address of the (N,12) rows
    long long sum_float4;      // number of float4 in them (3 N)
    float *mc_blocks;          // multicast address of the (R, block_stride) blocks
    const float *local_block;  // this rank's own block (local address)
    long long block_stride;    // floats between blocks
    long long block_float4;    // float4 per block
    int rank, world;
};

#ifndef GSB_HOST_EMU
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float *mc, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

constexpr int MX_THREADS = 512;
constexpr int MX_U = 4;  // independent requests in flight per thread (the reduction is a round trip through the switch)

// all-gather: this rank's block -> slot `rank` on every rank (the switch replicates each store)
__device__ __forceinline__ void push_block(const MultimemExchangeParams &p, long long tid, long long stride) {
    float *dst = p.mc_blocks + (size_t)p.rank * p.block_stride;
    const float4 *src = reinterpret_cast<const float4 *>(p.local_block);
    for (long long i = tid; i < p.block_float4; i += MX_U * stride) {
        float4 v[MX_U];
#pragma unroll
        for (int u = 0; u < MX_U; ++u)
            if (i + u * stride < p.block_float4) v[u] = src[i + u * stride];
#pragma unroll
        for (int u = 0; u < MX_U; ++u)
            if (i + u * stride < p.block_float4) multimem_st(dst + 4 * (i + u * stride), v[u]);
    }
}

// two-shot all-reduce of this rank's slice of the summable rows
__device__ __forceinline__ void reduce_slice(const MultimemExchangeParams &p, long long tid, long long stride) {
    const long long per = (p.sum_float4 + p.world - 1) / p.world;
    const long long lo = per * p.rank;
    const long long hi = lo + per < p.sum_float4 ? lo + per : p.sum_float4;
    for (long long i = lo + tid; i < hi; i += MX_U * stride) {
        float4 v[MX_U];
#pragma unroll
        for (int u = 0; u < MX_U; ++u)
            if (i + u * stride < hi) v[u] = multimem_ld_reduce_add(p.mc_sum + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < MX_U; ++u)
            if (i + u * stride < hi) multimem_st(p.mc_sum + 4 * (i + u * stride), v[u]);
    }
}

template <bool PUSH, bool REDUCE>
__global__ void __launch_bounds__(MX_THREADS) multimem_exchange_kernel(const MultimemExchangeParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (PUSH) push_block(p, tid, stride);  // first: its stores are fire-and-forget and overlap the reduction's round trips
    if (REDUCE) reduce_slice(p, tid, stride);
}
#endif

}  // namespace gsb

#ifndef GSB_HOST_EMU
extern "C" int gsb200_exchange_multimem(const GsbMultimemExchangeArgs *a) {
    using namespace gsb;
    if (!a || !a->multicast_grad_sum || !a->multicast_blocks || !a->local_block || a->num_points < 0 || a->world_size < 1 ||
        a->rank < 0 || a->rank >= a->world_size || a->block_stride < 3 * a->num_points + 3 * (int64_t)a->num_objects ||
        a->block_stride % 4 != 0) {
        set_error("exchange_multimem: bad arguments");
        return GSB_EINVAL;
    }
    if (reinterpret_cast<uintptr_t>(a->multicast_grad_sum) % 16 || reinterpret_cast<uintptr_t>(a->multicast_blocks) % 16 ||
        reinterpret_cast<uintptr_t>(a->local_block) % 16) {
        set_error("exchange_multimem: pointers must be 16-byte aligned");
        return GSB_EINVAL;
    }
    if (a->num_points == 0) return GSB_OK;
    MultimemExchangeParams p;
    p.mc_sum = a->multicast_grad_sum;
    p.sum_float4 = 3 * a->num_points;
    p.mc_blocks = a->multicast_blocks;
    p.local_block = a->local_block;
    p.block_stride = a->block_stride;
    p.block_float4 = (3 * a->num_points + 3 * (int64_t)a->num_objects + 3) / 4;
    p.rank = a->rank;
    p.world = a->world_size;
    const int blocks = a->num_blocks > 0 ? a->num_blocks : 2 * num_sms();
    cudaStream_t st = static_cast<cudaStream_t>(a->stream);
    const int phases = a->phases == 0 ? 3 : a->phases;
    if (phases == 3) multimem_exchange_kernel<true, true><<<blocks, MX_THREADS, 0, st>>>(p);
    else if (phases == 1) multimem_exchange_kernel<true, false><<<blocks, MX_THREADS, 0, st>>>(p);
    else if (phases == 2) multimem_exchange_kernel<false, true><<<blocks, MX_THREADS, 0, st>>>(p);
    else {
        set_error("exchange_multimem: phases must be 0 (both), 1 (gather), 2 (all-reduce) or 3");
        return GSB_EINVAL;
    }
    GSB_CUDA_CHECK(cudaGetLastError());
    return GSB_OK;
}
#endif

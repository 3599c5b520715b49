// sort.cu -- device LSD radix sort of (tile | depth) keys with the in-camera offset as payload
// (replaces torch.sort + gather, GPCR:947-950) and the per-tile range detection (GPCR:175-193).
//
// One-sweep organisation: one histogram kernel builds the digit histograms of every pass; each
// pass is a single kernel in which a CTA (a) pulls its 4096-key tile into shared memory with one
// TMA bulk copy (cp.async.bulk + mbarrier -> SASS UBLKCP), (b) ranks the keys stably with
// warp-level match_any, (c) obtains the global digit offsets by a per-digit decoupled look-back
// over the preceding CTAs, and (d) scatters keys and payloads from a block-sorted shared-memory
// staging area so that global stores go out in runs.  Keys are only as wide as the live bits:
// ceil(log2 T) tile bits + the bits of int(far*scale); 32-bit keys whenever that is <= 32 bits.
// Stability (ties keep ascending in-camera offset = ascending point id) is what reproduces the
// reference's blend order.  HBM-bound: 2*(key+4) B per key per pass + one key read for histograms.
#include "common.cuh"

namespace gsb {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr unsigned int SS_AGGREGATE = 1u << 30;
constexpr unsigned int SS_INCLUSIVE = 2u << 30;
constexpr unsigned int SS_VALUE_MASK = (1u << 30) - 1;

__device__ __forceinline__ unsigned int ld_u32_volatile(const unsigned int *p) {
    return *reinterpret_cast<const volatile unsigned int *>(p);
}
__device__ __forceinline__ void st_u32_volatile(unsigned int *p, unsigned int v) {
    *reinterpret_cast<volatile unsigned int *>(p) = v;
}

// ---- mbarrier / bulk-copy helpers (TMA 1-D bulk copy, global -> shared)
__device__ __forceinline__ unsigned int smem_addr(const void *p) {
    return (unsigned int)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, unsigned int bytes,
                                              unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------------------ histograms of all passes
template <typename KeyT>
__global__ void __launch_bounds__(256)
sort_histogram_kernel(const KeyT *__restrict__ keys, const long long *__restrict__ n_dev,
                      long long capacity, int passes, unsigned int *__restrict__ hist) {
    __shared__ unsigned int s_hist[8 * RADIX];
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    long long n = *n_dev;
    if (n > capacity) n = capacity;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const KeyT k = keys[i];
        for (int p = 0; p < passes; ++p)
            atomicAdd(&s_hist[p * RADIX + (int)((k >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) {
        const unsigned int c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ------------------------------------------------------------------ one radix pass
template <typename KeyT>
struct PassSmem {
    alignas(128) KeyT keys[SORT_TILE];  // TMA destination, later the block-sorted key staging area
    int vals[SORT_TILE];                // block-sorted payload staging area
    unsigned int warp_cnt[SORT_BLOCK_THREADS / 32][RADIX];
    unsigned int digit_start[RADIX];    // exclusive start of each digit inside the block-sorted tile
    unsigned int global_base[RADIX];    // destination of local position p with digit d: global_base[d] + p
    unsigned int scan_tmp[SORT_BLOCK_THREADS / 32];
    unsigned long long mbar;
    unsigned int ticket;
};

template <typename KeyT>
__global__ void __launch_bounds__(SORT_BLOCK_THREADS, GSB_SORT_MIN_BLOCKS)
onesweep_pass_kernel(const KeyT *__restrict__ keys_in, const int *__restrict__ vals_in,
                     KeyT *__restrict__ keys_out, int *__restrict__ vals_out,
                     const long long *__restrict__ n_dev, long long capacity, int shift,
                     const unsigned int *__restrict__ hist /* this pass, 256 bins */,
                     unsigned int *__restrict__ state /* this pass: [blocks][256] */,
                     unsigned int *__restrict__ ticket_ctr) {
    extern __shared__ unsigned char smem_raw[];
    PassSmem<KeyT> &s = *reinterpret_cast<PassSmem<KeyT> *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        s.ticket = atomicAdd(ticket_ctr, 1u);
        mbar_init(&s.mbar, 1);
    }
    for (int i = tid; i < (SORT_BLOCK_THREADS / 32) * RADIX; i += SORT_BLOCK_THREADS)
        (&s.warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const unsigned int blk = s.ticket;
    long long n = *n_dev;
    if (n > capacity) n = capacity;
    const long long tile_base = (long long)blk * SORT_TILE;
    if (tile_base >= n) return;
    const int count = (int)min((long long)SORT_TILE, n - tile_base);

    // (a) key tile -> shared memory with one TMA bulk copy (16-byte granules); the < 16-byte tail of a
    //     partial last tile is fetched with ordinary loads.
    const unsigned int bulk_bytes = ((unsigned int)count * (unsigned int)sizeof(KeyT)) & ~15u;
    const int bulk_elems = (int)(bulk_bytes / sizeof(KeyT));
    if (tid == 0 && bulk_bytes) {
        mbar_arrive_expect_tx(&s.mbar, bulk_bytes);
        bulk_copy_g2s(s.keys, keys_in + tile_base, bulk_bytes, &s.mbar);
    }
    if (bulk_elems + tid < count) s.keys[bulk_elems + tid] = keys_in[tile_base + bulk_elems + tid];
    // payloads straight to registers, warp-striped (coalesced)
    int vals[SORT_ITEMS_PER_THREAD];
    const int wbase = warp * (32 * SORT_ITEMS_PER_THREAD);
#pragma unroll
    for (int j = 0; j < SORT_ITEMS_PER_THREAD; ++j) {
        const int idx = wbase + j * 32 + lane;
        vals[j] = idx < count ? __ldg(&vals_in[tile_base + idx]) : 0;
    }
    if (bulk_bytes) mbar_wait(&s.mbar, 0);
    __syncthreads();  // tail keys written by other threads

    // (b) stable ranking: per-warp digit counters, match_any groups equal digits in lane order
    KeyT keys[SORT_ITEMS_PER_THREAD];
    unsigned short ranks[SORT_ITEMS_PER_THREAD];
    const unsigned int lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < SORT_ITEMS_PER_THREAD; ++j) {
        const int idx = wbase + j * 32 + lane;
        const bool valid = idx < count;
        keys[j] = s.keys[idx];
        const int d = valid ? (int)((keys[j] >> shift) & (RADIX - 1)) : RADIX;
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        unsigned int prev = 0;
        if (valid) prev = s.warp_cnt[warp][d];
        ranks[j] = (unsigned short)(prev + __popc(peers & lt_mask));
        __syncwarp();
        if (valid && (peers & lt_mask) == 0) s.warp_cnt[warp][d] = prev + __popc(peers);
        __syncwarp();
    }
    __syncthreads();

    // per-digit totals, warp-exclusive bases, block-exclusive digit starts
    unsigned int my_count = 0;
    {
#pragma unroll
        for (int w = 0; w < SORT_BLOCK_THREADS / 32; ++w) {
            const unsigned int c = s.warp_cnt[w][tid];
            s.warp_cnt[w][tid] = my_count;
            my_count += c;
        }
        unsigned int incl = my_count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s.scan_tmp[warp] = incl;
        __syncthreads();
        unsigned int wprefix = 0;
#pragma unroll
        for (int w = 0; w < SORT_BLOCK_THREADS / 32; ++w)
            if (w < warp) wprefix += s.scan_tmp[w];
        const unsigned int dstart = wprefix + incl - my_count;
        s.digit_start[tid] = dstart;

        // global exclusive prefix of this digit over all digits (from the pass histogram)
        unsigned int h = hist[tid];
        unsigned int hincl = h;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int o = __shfl_up_sync(0xffffffffu, hincl, d);
            if (lane >= d) hincl += o;
        }
        __syncthreads();  // scan_tmp reuse
        if (lane == 31) s.scan_tmp[warp] = hincl;
        __syncthreads();
        unsigned int hprefix = 0;
#pragma unroll
        for (int w = 0; w < SORT_BLOCK_THREADS / 32; ++w)
            if (w < warp) hprefix += s.scan_tmp[w];
        const unsigned int digit_global = hprefix + hincl - h;

        // (c) decoupled look-back over preceding CTAs for digit `tid`
        unsigned int *my_state = state + (size_t)blk * RADIX + tid;
        unsigned int exclusive = 0;
        if (blk == 0) {
            st_u32_volatile(my_state, SS_INCLUSIVE | my_count);
        } else {
            st_u32_volatile(my_state, SS_AGGREGATE | my_count);
            long long look = (long long)blk - 1;
            while (true) {
                const unsigned int *ps = state + (size_t)look * RADIX + tid;
                unsigned int w = ld_u32_volatile(ps);
                while ((w >> 30) == 0) w = ld_u32_volatile(ps);
                exclusive += w & SS_VALUE_MASK;
                if ((w >> 30) == 2) break;
                --look;
            }
            st_u32_volatile(my_state, SS_INCLUSIVE | (exclusive + my_count));
        }
        s.global_base[tid] = digit_global + exclusive - dstart;
    }
    __syncthreads();

    // block-sorted staging in shared memory (the TMA buffer is dead: all keys are in registers)
#pragma unroll
    for (int j = 0; j < SORT_ITEMS_PER_THREAD; ++j) {
        const int idx = wbase + j * 32 + lane;
        if (idx < count) {
            const int d = (int)((keys[j] >> shift) & (RADIX - 1));
            const unsigned int pos = s.digit_start[d] + s.warp_cnt[warp][d] + ranks[j];
            s.keys[pos] = keys[j];
            s.vals[pos] = vals[j];
        }
    }
    __syncthreads();
    // (d) scatter: consecutive local positions of one digit go to consecutive global addresses
#pragma unroll
    for (int j = 0; j < SORT_ITEMS_PER_THREAD; ++j) {
        const int pidx = j * SORT_BLOCK_THREADS + tid;
        if (pidx < count) {
            const KeyT k = s.keys[pidx];
            const int d = (int)((k >> shift) & (RADIX - 1));
            const unsigned int dst = s.global_base[d] + (unsigned int)pidx;
            keys_out[dst] = k;
            vals_out[dst] = s.vals[pidx];
        }
    }
}

template <typename KeyT>
static int sort_pairs_typed(const KeyT *keys_in, const int *vals_in, KeyT *keys_out, int *vals_out,
                            const long long *n_dev, int64_t capacity, int end_bit, unsigned int *hist,
                            unsigned int *state, unsigned int *tickets, KeyT *tmp_keys, int *tmp_vals,
                            long long *sel_out, cudaStream_t stream) {
    const int passes = (end_bit + RADIX_BITS - 1) / RADIX_BITS;
    const int blocks = (int)((capacity + SORT_TILE - 1) / SORT_TILE);
    if (blocks == 0 || passes == 0) return GSB_OK;
    const size_t smem = sizeof(PassSmem<KeyT>) + 128;
    static bool attr_set = false;
    if (!attr_set) {
        GSB_CUDA_CHECK(cudaFuncSetAttribute(onesweep_pass_kernel<KeyT>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int hist_blocks = blocks < 4 * num_sms() ? blocks : 4 * num_sms();
    sort_histogram_kernel<KeyT><<<hist_blocks, 256, 0, stream>>>(keys_in, n_dev, capacity, passes, hist);
    GSB_CUDA_CHECK(cudaGetLastError());
    // ping-pong: pass p reads src, writes dst.  We arrange that the LAST pass writes keys_out.
    const KeyT *src_k = keys_in;
    const int *src_v = vals_in;
    for (int p = 0; p < passes; ++p) {
        const bool last_to_out = ((passes - 1 - p) % 2) == 0;
        KeyT *dst_k = last_to_out ? keys_out : tmp_keys;
        int *dst_v = last_to_out ? vals_out : tmp_vals;
        onesweep_pass_kernel<KeyT><<<blocks, SORT_BLOCK_THREADS, smem, stream>>>(
            src_k, src_v, dst_k, dst_v, n_dev, capacity, p * RADIX_BITS, hist + p * RADIX,
            state + (size_t)p * blocks * RADIX, tickets + p);
        GSB_CUDA_CHECK(cudaGetLastError());
        src_k = dst_k;
        src_v = dst_v;
    }
    (void)sel_out;
    return GSB_OK;
}

// keys_in may alias tmp_keys' partner: the caller provides (in, out, tmp) with in != out != tmp unless
// passes is such that `in` is never written; we require three distinct buffers only when passes >= 2
// and in must stay intact; the frame pipeline passes in = keys_a, out = keys_b, tmp = keys_a
// (the emitted keys are dead after the first pass has consumed them).
int sort_pairs_device(const void *keys_in, const int *vals_in, void *keys_out, int *vals_out,
                      const long long *n_dev, int64_t n_capacity, int key_bytes, int end_bit,
                      unsigned int *hist, unsigned int *state, unsigned int *tickets, void *tmp_keys,
                      int *tmp_vals, long long *sel_out, cudaStream_t stream) {
    if (key_bytes == 4)
        return sort_pairs_typed<unsigned int>((const unsigned int *)keys_in, vals_in,
                                              (unsigned int *)keys_out, vals_out, n_dev, n_capacity,
                                              end_bit, hist, state, tickets, (unsigned int *)tmp_keys,
                                              tmp_vals, sel_out, stream);
    if (key_bytes == 8)
        return sort_pairs_typed<unsigned long long>(
            (const unsigned long long *)keys_in, vals_in, (unsigned long long *)keys_out, vals_out, n_dev,
            n_capacity, end_bit, hist, state, tickets, (unsigned long long *)tmp_keys, tmp_vals, sel_out,
            stream);
    set_error("sort: key_bytes must be 4 or 8, got %d", key_bytes);
    return GSB_EINVAL;
}

// ------------------------------------------------------------------ tile ranges (GPCR:175-193)
template <typename KeyT>
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const KeyT *__restrict__ keys, const long long *__restrict__ n_dev, long long capacity,
                   int depth_bits, int num_tiles, int *__restrict__ tile_start,
                   int *__restrict__ tile_end) {
    long long n = n_dev ? *n_dev : capacity;
    if (n > capacity) n = capacity;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int t = (int)(keys[i] >> depth_bits);
        const int tp = i > 0 ? (int)(keys[i - 1] >> depth_bits) : -1;
        if (t != tp && t < num_tiles) {
            if (i > 0) {
                tile_start[t] = (int)i;
                if (tp < num_tiles) tile_end[tp] = (int)i;
            }
        }
        if (i == n - 1 && t < num_tiles) tile_end[t] = (int)n;
    }
}

int launch_sort(const Workspace &ws, int64_t key_capacity, cudaStream_t stream) {
    const GsbWorkspaceLayout &L = ws.layout;
    // in = keys_a (emitted), out = keys_b, tmp = keys_a: pass p alternates b/a so that the last pass
    // lands in keys_b; the emitted keys in a are dead once pass 0 has read them -- but pass 0 must not
    // write a.  With an odd number of passes pass 0 writes b (fine); with an even number pass 0 would
    // write a (= its own input), so in that case we sort into a with b as scratch.
    const int passes = L.sort_passes;
    const bool out_is_b = (passes % 2) == 1;
    void *out_k = out_is_b ? ws.keys_b : ws.keys_a;
    int *out_v = out_is_b ? ws.vals_b : ws.vals_a;
    void *tmp_k = out_is_b ? ws.keys_a : ws.keys_b;
    int *tmp_v = out_is_b ? ws.vals_a : ws.vals_b;
    // even pass count: pass 0 writes tmp (= b), pass 1 writes out (= a, input already consumed) ...
    (void)key_capacity;
    return sort_pairs_device(ws.keys_a, ws.vals_a, out_k, out_v, ws.counters + CNT_K,
                             L.key_capacity_padded, L.key_bytes, L.tile_bits + L.depth_bits,
                             ws.sort_hist, ws.sort_state, ws.tickets + TICKET_SORT0, tmp_k, tmp_v,
                             nullptr, stream);
}

int launch_tile_ranges_raw(const long long *keys_i64, int64_t n, int *tile_start, int *tile_end,
                           int num_tiles, cudaStream_t stream) {
    if (n <= 0) return GSB_OK;
    long long blocks = (n + 255) / 256;
    const long long cap_blocks = 8LL * num_sms();
    if (blocks > cap_blocks) blocks = cap_blocks;
    tile_ranges_kernel<unsigned long long><<<(int)blocks, 256, 0, stream>>>(
        (const unsigned long long *)keys_i64, nullptr, n, 32, num_tiles, tile_start, tile_end);
    GSB_CUDA_CHECK(cudaGetLastError());
    return GSB_OK;
}

int launch_tile_ranges(const Workspace &ws, int64_t key_capacity, int num_tiles, cudaStream_t stream) {
    const GsbWorkspaceLayout &L = ws.layout;
    const bool out_is_b = (L.sort_passes % 2) == 1;
    const void *keys = out_is_b ? ws.keys_b : ws.keys_a;
    long long blocks = (key_capacity + 255) / 256;
    const long long cap_blocks = 8LL * num_sms();
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (blocks <= 0) return GSB_OK;
    if (L.key_bytes == 4)
        tile_ranges_kernel<unsigned int><<<(int)blocks, 256, 0, stream>>>(
            (const unsigned int *)keys, ws.counters + CNT_K, key_capacity, L.depth_bits, num_tiles,
            ws.tile_start, ws.tile_end);
    else
        tile_ranges_kernel<unsigned long long><<<(int)blocks, 256, 0, stream>>>(
            (const unsigned long long *)keys, ws.counters + CNT_K, key_capacity, L.depth_bits, num_tiles,
            ws.tile_start, ws.tile_end);
    GSB_CUDA_CHECK(cudaGetLastError());
    return GSB_OK;
}

}  // namespace gsb

// sort.cu -- device LSD radix sort of (tile | depth) keys with the in-camera offset as payload
// (replaces torch.sort + gather, GPCR:947-950) and the per-tile range detection (GPCR:175-193).
//
// One-sweep organisation: one histogram kernel builds the digit histograms of every pass; each
// pass is a single kernel in which a CTA (a) pulls its 3072-key tile into shared memory with one
// TMA bulk copy (cp.async.bulk + mbarrier -> SASS UBLKCP), (b) ranks the keys stably with
// warp-level match_any, (c) obtains the global digit offsets by a per-digit decoupled look-back
// over the preceding CTAs, and (d) scatters keys and payloads from a block-sorted shared-memory
// staging area so that global stores go out in runs.  Keys are only as wide as the live bits:
// ceil(log2 T) tile bits + the bits of int(far*scale); 32-bit keys whenever that is <= 32 bits.
// Stability (ties keep ascending in-camera offset = ascending point id) is what reproduces the
// reference's blend order.  HBM-bound: 2*(key+4) B per key per pass + one key read for histograms.
#include "common.cuh"

namespace gsb {

constexpr unsigned int SS_AGGREGATE = 1u << 30;
constexpr unsigned int SS_INCLUSIVE = 2u << 30;
constexpr unsigned int SS_VALUE_MASK = (1u << 30) - 1;
#ifndef GSB_SORT_LOOKBACK
#define GSB_SORT_LOOKBACK 4
#endif
constexpr int LOOKBACK = GSB_SORT_LOOKBACK;  // predecessors whose state is fetched per look-back round

// Digit width for a key of `bits` live bits.  The kernels are templated on 8- and 10-bit digits; measured on
// B200 (C3, 30-bit keys, K = 4.0e6) three 10-bit passes cost 244 us against 199 us for four 8-bit passes
// (1024-bin ranking + 4 KB of look-back state per CTA outweigh the saved pass), so 8 bits are always used.
int sort_radix_bits(int bits) {
    (void)bits;
    return 8;
}

__device__ __forceinline__ unsigned int ld_u32_volatile(const unsigned int *p) {
    return *reinterpret_cast<const volatile unsigned int *>(p);
}
__device__ __forceinline__ void st_u32_volatile(unsigned int *p, unsigned int v) {
    *reinterpret_cast<volatile unsigned int *>(p) = v;
}

// ------------------------------------------------------------------ histograms of all passes
template <typename KeyT, int RBITS>
__global__ void __launch_bounds__(256)
sort_histogram_kernel(const KeyT *__restrict__ keys, const long long *__restrict__ n_dev,
                      long long capacity, int passes, unsigned int *__restrict__ hist) {
    constexpr int RADIX = 1 << RBITS;
    __shared__ unsigned int s_hist[8 * RADIX];
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    long long n = *n_dev;
    if (n > capacity) n = capacity;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const KeyT k = keys[i];
        for (int p = 0; p < passes; ++p)
            atomicAdd(&s_hist[p * RADIX + (int)((k >> (p * RBITS)) & (RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) {
        const unsigned int c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ------------------------------------------------------------------ one radix pass
template <typename KeyT, int RBITS>
struct PassSmem {
    static constexpr int RADIX = 1 << RBITS;
    alignas(128) KeyT keys[SORT_TILE];  // TMA destination, later the block-sorted key staging area
    int vals[SORT_TILE];                // block-sorted payload staging area
    unsigned short warp_cnt[SORT_BLOCK_THREADS / 32][RADIX];  // per-warp digit counters (<= 512 each)
    unsigned int digit_start[RADIX];    // exclusive start of each digit inside the block-sorted tile
    unsigned int global_base[RADIX];    // destination of local position p with digit d: global_base[d] + p
    unsigned int scan_tmp[SORT_BLOCK_THREADS / 32];
    unsigned long long mbar;
    unsigned int ticket;
};

// exclusive block scan of one value per thread (256 threads); returns the exclusive prefix of `v`
template <typename S>
__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, S &s, int lane, int warp) {
    unsigned int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    __syncthreads();  // scan_tmp free
    if (lane == 31) s.scan_tmp[warp] = incl;
    __syncthreads();
    unsigned int wprefix = 0;
#pragma unroll
    for (int w = 0; w < SORT_BLOCK_THREADS / 32; ++w)
        if (w < warp) wprefix += s.scan_tmp[w];
    return wprefix + incl - v;
}

template <typename KeyT, int RBITS>
__global__ void __launch_bounds__(SORT_BLOCK_THREADS, GSB_SORT_MIN_BLOCKS)
onesweep_pass_kernel(const KeyT *__restrict__ keys_in, const int *__restrict__ vals_in,
                     KeyT *__restrict__ keys_out, int *__restrict__ vals_out,
                     const long long *__restrict__ n_dev, long long capacity, int shift,
                     const unsigned int *__restrict__ hist /* this pass, RADIX bins */,
                     unsigned int *__restrict__ state /* this pass: [blocks][RADIX] */,
                     unsigned int *__restrict__ ticket_ctr) {
    constexpr int RADIX = 1 << RBITS;
    constexpr int DPT = RADIX / SORT_BLOCK_THREADS;  // digits owned by a thread: [tid*DPT, tid*DPT+DPT)
#ifdef GSB_HOST_EMU
    unsigned char *const smem_raw = simt_emu::dynamic_smem();
#else
    extern __shared__ unsigned char smem_raw[];
#endif
    PassSmem<KeyT, RBITS> &s = *reinterpret_cast<PassSmem<KeyT, RBITS> *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        s.ticket = atomicAdd(ticket_ctr, 1u);
        mbar_init(&s.mbar, 1);
    }
    {
        unsigned int *z = reinterpret_cast<unsigned int *>(&s.warp_cnt[0][0]);
        for (int i = tid; i < (SORT_BLOCK_THREADS / 32) * RADIX / 2; i += SORT_BLOCK_THREADS) z[i] = 0;
    }
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
        if (valid && (peers & lt_mask) == 0) s.warp_cnt[warp][d] = (unsigned short)(prev + __popc(peers));
        __syncwarp();
    }
    __syncthreads();

    // per-digit totals (thread t owns digits t*DPT..), warp-exclusive bases; publish the aggregates at once
    unsigned int cnt[DPT], excl[DPT];
    unsigned int *my_state = state + (size_t)blk * RADIX + tid * DPT;
    unsigned int tsum = 0;
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const int d = tid * DPT + k;
        unsigned int c = 0;
#pragma unroll
        for (int w = 0; w < SORT_BLOCK_THREADS / 32; ++w) {
            const unsigned int x = s.warp_cnt[w][d];
            s.warp_cnt[w][d] = (unsigned short)c;
            c += x;
        }
        cnt[k] = c;
        tsum += c;
        st_u32_volatile(my_state + k, (blk == 0 ? SS_INCLUSIVE : SS_AGGREGATE) | c);
    }
    {   // block-exclusive digit starts
        unsigned int run = block_exclusive_scan(tsum, s, lane, warp);
#pragma unroll
        for (int k = 0; k < DPT; ++k) {
            s.digit_start[tid * DPT + k] = run;
            run += cnt[k];
        }
    }
    unsigned int dglobal[DPT];
    {   // global exclusive prefix of each digit over all digits (from the pass histogram)
        unsigned int h[DPT], hs = 0;
#pragma unroll
        for (int k = 0; k < DPT; ++k) {
            h[k] = hist[tid * DPT + k];
            hs += h[k];
        }
        unsigned int run = block_exclusive_scan(hs, s, lane, warp);
#pragma unroll
        for (int k = 0; k < DPT; ++k) {
            dglobal[k] = run;
            run += h[k];
        }
    }
    __syncthreads();  // digit_start and the warp bases are visible

    // block-sorted staging in shared memory (the TMA buffer is dead: all keys are in registers).  This needs
    // only block-local offsets, so it runs BEFORE the look-back and gives the predecessors time to publish.
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

    // (c) decoupled look-back over preceding CTAs for this thread's digits, four predecessors in flight per round
#pragma unroll
    for (int k = 0; k < DPT; ++k) excl[k] = 0;
    if (blk != 0) {
        bool done[DPT];
#pragma unroll
        for (int k = 0; k < DPT; ++k) done[k] = false;
        long long look = (long long)blk - 1;
        bool all_done = false;
        while (!all_done) {
            unsigned int w[LOOKBACK][DPT];
#pragma unroll
            for (int r = 0; r < LOOKBACK; ++r)
#pragma unroll
                for (int k = 0; k < DPT; ++k)
                    w[r][k] = (look - r >= 0) ? ld_u32_volatile(state + (size_t)(look - r) * RADIX + tid * DPT + k)
                                              : SS_INCLUSIVE;
#pragma unroll
            for (int r = 0; r < LOOKBACK; ++r) {
#pragma unroll
                for (int k = 0; k < DPT; ++k) {
                    if (done[k]) continue;
                    while ((w[r][k] >> 30) == 0)
                        w[r][k] = ld_u32_volatile(state + (size_t)(look - r) * RADIX + tid * DPT + k);
                    excl[k] += w[r][k] & SS_VALUE_MASK;
                    if ((w[r][k] >> 30) == 2) done[k] = true;
                }
            }
            all_done = true;
#pragma unroll
            for (int k = 0; k < DPT; ++k) all_done = all_done && done[k];
            look -= LOOKBACK;
        }
#pragma unroll
        for (int k = 0; k < DPT; ++k) st_u32_volatile(my_state + k, SS_INCLUSIVE | (excl[k] + cnt[k]));
    }
#pragma unroll
    for (int k = 0; k < DPT; ++k)
        s.global_base[tid * DPT + k] = dglobal[k] + excl[k] - s.digit_start[tid * DPT + k];
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

#ifndef GSB_HOST_EMU
template <typename KeyT, int RBITS>
static int sort_pairs_typed(const KeyT *keys_in, const int *vals_in, KeyT *keys_out, int *vals_out,
                            const long long *n_dev, int64_t capacity, int end_bit, unsigned int *hist,
                            unsigned int *state, unsigned int *tickets, KeyT *tmp_keys, int *tmp_vals,
                            cudaStream_t stream) {
    constexpr int RADIX = 1 << RBITS;
    const int passes = (end_bit + RBITS - 1) / RBITS;
    const int blocks = (int)((capacity + SORT_TILE - 1) / SORT_TILE);
    if (blocks == 0 || passes == 0) return GSB_OK;
    const size_t smem = sizeof(PassSmem<KeyT, RBITS>) + 128;
    static bool attr_set = false;
    if (!attr_set) {
        GSB_CUDA_CHECK(cudaFuncSetAttribute(onesweep_pass_kernel<KeyT, RBITS>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int hist_blocks = blocks < 4 * num_sms() ? blocks : 4 * num_sms();
    sort_histogram_kernel<KeyT, RBITS><<<hist_blocks, 256, 0, stream>>>(keys_in, n_dev, capacity, passes, hist);
    GSB_CUDA_CHECK(cudaGetLastError());
    // ping-pong: pass p reads src, writes dst.  We arrange that the LAST pass writes keys_out.
    const KeyT *src_k = keys_in;
    const int *src_v = vals_in;
    for (int p = 0; p < passes; ++p) {
        const bool last_to_out = ((passes - 1 - p) % 2) == 0;
        KeyT *dst_k = last_to_out ? keys_out : tmp_keys;
        int *dst_v = last_to_out ? vals_out : tmp_vals;
        onesweep_pass_kernel<KeyT, RBITS><<<blocks, SORT_BLOCK_THREADS, smem, stream>>>(
            src_k, src_v, dst_k, dst_v, n_dev, capacity, p * RBITS, hist + p * RADIX,
            state + (size_t)p * blocks * RADIX, tickets + p);
        GSB_CUDA_CHECK(cudaGetLastError());
        src_k = dst_k;
        src_v = dst_v;
    }
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
    (void)sel_out;
    const int rbits = sort_radix_bits(end_bit);
    typedef unsigned int u32;
    typedef unsigned long long u64;
    if (key_bytes == 4 && rbits == 8)
        return sort_pairs_typed<u32, 8>((const u32 *)keys_in, vals_in, (u32 *)keys_out, vals_out, n_dev, n_capacity,
                                        end_bit, hist, state, tickets, (u32 *)tmp_keys, tmp_vals, stream);
    if (key_bytes == 4)
        return sort_pairs_typed<u32, 10>((const u32 *)keys_in, vals_in, (u32 *)keys_out, vals_out, n_dev, n_capacity,
                                         end_bit, hist, state, tickets, (u32 *)tmp_keys, tmp_vals, stream);
    if (key_bytes == 8 && rbits == 8)
        return sort_pairs_typed<u64, 8>((const u64 *)keys_in, vals_in, (u64 *)keys_out, vals_out, n_dev, n_capacity,
                                        end_bit, hist, state, tickets, (u64 *)tmp_keys, tmp_vals, stream);
    if (key_bytes == 8)
        return sort_pairs_typed<u64, 10>((const u64 *)keys_in, vals_in, (u64 *)keys_out, vals_out, n_dev, n_capacity,
                                         end_bit, hist, state, tickets, (u64 *)tmp_keys, tmp_vals, stream);
    set_error("sort: key_bytes must be 4 or 8, got %d", key_bytes);
    return GSB_EINVAL;
}

#endif  // GSB_HOST_EMU

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

#ifndef GSB_HOST_EMU
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

#endif  // GSB_HOST_EMU

}  // namespace gsb

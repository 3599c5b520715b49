// emu_sort.cpp -- the one-sweep LSD radix sort and the tile-range kernel (csrc/sort.cu) compiled as host C++ under
// simt_emu.h.  TEST INFRASTRUCTURE, see simt_emu.h.  The pass loop below restates sort_pairs_typed() of sort.cu (histogram
// kernel, then one kernel per 8-bit digit, ping-pong so that the last pass lands in the output); the CTAs of a pass run in
// launch = ticket order, so the per-digit look-back always finds its predecessors complete.
#include "simt_emu.h"
#include "../../taichi_3d_gaussian_splatting_b200/csrc/sort.cu"

namespace {
template <class KeyT>
struct HistArgs {
    const KeyT *keys;
    const long long *n;
    long long cap;
    int passes;
    unsigned int *hist;
};
template <class KeyT>
struct PassArgs {
    const KeyT *ki;
    const int *vi;
    KeyT *ko;
    int *vo;
    const long long *n;
    long long cap;
    int shift;
    const unsigned int *hist;
    unsigned int *state, *ticket;
};
template <class KeyT>
long long sort_typed(const KeyT *keys_in, const int *vals_in, KeyT *keys_out, int *vals_out, long long n, int end_bit) {
    using namespace gsb;
    constexpr int RB = 8, RADIX = 1 << RB;
    const int passes = (end_bit + RB - 1) / RB;
    const long long cap = (n + SORT_TILE - 1) / SORT_TILE * SORT_TILE;
    const int blocks = (int)(cap / SORT_TILE);
    if (blocks == 0 || passes == 0) return 0;
    std::vector<KeyT> in(cap, KeyT(0)), tmp(cap), out(cap);
    std::vector<int> vin(cap, 0), vtmp(cap), vout(cap);
    std::copy(keys_in, keys_in + n, in.begin());
    std::copy(vals_in, vals_in + n, vin.begin());
    std::vector<unsigned int> hist(8 * 1024, 0u), state((size_t)passes * blocks * RADIX, 0u), tickets(8, 0u);
    simt_emu::M().switches = 0;
    HistArgs<KeyT> ha{in.data(), &n, cap, passes, hist.data()};
    simt_emu::launch([](const HistArgs<KeyT> &a) { sort_histogram_kernel<KeyT, RB>(a.keys, a.n, a.cap, a.passes, a.hist); },
                     std::min(blocks, 4 * 148), 256, ha);
    const KeyT *src_k = in.data();
    const int *src_v = vin.data();
    for (int p = 0; p < passes; ++p) {
        const bool last_to_out = ((passes - 1 - p) % 2) == 0;
        KeyT *dst_k = last_to_out ? out.data() : tmp.data();
        int *dst_v = last_to_out ? vout.data() : vtmp.data();
        PassArgs<KeyT> pa{src_k, src_v, dst_k, dst_v, &n, cap, p * RB, hist.data() + p * RADIX,
                          state.data() + (size_t)p * blocks * RADIX, tickets.data() + p};
        simt_emu::launch(
            [](const PassArgs<KeyT> &a) {
                onesweep_pass_kernel<KeyT, RB>(a.ki, a.vi, a.ko, a.vo, a.n, a.cap, a.shift, a.hist, a.state, a.ticket);
            },
            blocks, SORT_BLOCK_THREADS, pa);
        src_k = dst_k;
        src_v = dst_v;
    }
    std::copy(out.begin(), out.begin() + n, keys_out);
    std::copy(vout.begin(), vout.begin() + n, vals_out);
    return simt_emu::M().switches;
}
template <class KeyT>
struct RangeArgs {
    const KeyT *keys;
    const long long *n;
    long long cap;
    int depth_bits, tiles;
    int *start, *end;
};
}  // namespace

extern "C" long long emu_sort_pairs(const void *keys_in, const int *vals_in, void *keys_out, int *vals_out, long long n,
                                    int key_bytes, int end_bit) {
    if (key_bytes == 4)
        return sort_typed<unsigned int>((const unsigned int *)keys_in, vals_in, (unsigned int *)keys_out, vals_out, n, end_bit);
    return sort_typed<unsigned long long>((const unsigned long long *)keys_in, vals_in, (unsigned long long *)keys_out,
                                          vals_out, n, end_bit);
}

// tile_start / tile_end must be zero-initialised (GPCR:954-957)
extern "C" void emu_tile_ranges(const void *sorted_keys, long long n, int key_bytes, int depth_bits, int num_tiles,
                                int *tile_start, int *tile_end) {
    using namespace gsb;
    if (n <= 0) return;
    const int blocks = (int)std::min<long long>((n + 255) / 256, 8 * 148);
    if (key_bytes == 4) {
        RangeArgs<unsigned int> a{(const unsigned int *)sorted_keys, &n, n, depth_bits, num_tiles, tile_start, tile_end};
        simt_emu::launch([](const RangeArgs<unsigned int> &r) {
            tile_ranges_kernel<unsigned int>(r.keys, r.n, r.cap, r.depth_bits, r.tiles, r.start, r.end);
        }, blocks, 256, a);
    } else {
        RangeArgs<unsigned long long> a{(const unsigned long long *)sorted_keys, &n, n, depth_bits, num_tiles, tile_start,
                                        tile_end};
        simt_emu::launch([](const RangeArgs<unsigned long long> &r) {
            tile_ranges_kernel<unsigned long long>(r.keys, r.n, r.cap, r.depth_bits, r.tiles, r.start, r.end);
        }, blocks, 256, a);
    }
}

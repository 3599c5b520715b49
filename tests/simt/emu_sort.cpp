// emu_sort.cpp -- the one-sweep LSD radix sort and the tile-range kernel (csrc/sort.cu) compiled as host C++ under
// simt_emu.h.  TEST INFRASTRUCTURE, see simt_emu.h.  The pass loop below restates sort_pairs_typed() of sort.cu (histogram
// kernel, then one launch per 8-bit digit of the widest possible key; the kernels themselves decide which passes run and
// rotate through the three buffers so that the sorted list ends in the output); the CTAs of a pass run in launch = ticket
// order, so the per-digit look-back always finds its predecessors complete.
#include "simt_emu.h"
#include "../../taichi_3d_gaussian_splatting_b200/csrc/sort.cu"

namespace {
template <class KeyT>
struct HistArgs {
    const KeyT *keys;
    const long long *n;
    long long cap;
    int depth_bits, end_bit;
    const int *max_depth_key;
    unsigned int *hist, *done;
};
// depth_bits / max_depth_key: the live-bit compaction of sort.cu (max_depth_key == nullptr: every bit is live)
template <class KeyT>
long long sort_typed(const KeyT *keys_in, const int *vals_in, KeyT *keys_out, int *vals_out, long long n, int depth_bits,
                     int end_bit, const int *max_depth_key) {
    using namespace gsb;
    const int passes = (end_bit + RBITS - 1) / RBITS;  // launches: the worst case, as sort_pairs_typed() does
    const long long cap = (n + SORT_TILE - 1) / SORT_TILE * SORT_TILE;
    const int blocks = (int)(cap / SORT_TILE);
    if (blocks == 0 || passes == 0) return 0;
    std::vector<KeyT> in(cap, KeyT(0)), tmp(cap), out(cap);
    std::vector<int> vin(cap, 0), vtmp(cap), vout(cap);
    std::copy(keys_in, keys_in + n, in.begin());
    std::copy(vals_in, vals_in + n, vin.begin());
    const std::vector<KeyT> in_before(in);
    std::vector<unsigned int> hist(8 * 1024, 0u), state((size_t)passes * blocks * RADIX, 0u), tickets(16, 0u);
    simt_emu::M().switches = 0;
    HistArgs<KeyT> ha{in.data(), &n, cap, depth_bits, end_bit, max_depth_key, hist.data(), tickets.data() + 8};
    simt_emu::launch(
        [](const HistArgs<KeyT> &a) {
            sort_histogram_kernel<KeyT>(a.keys, a.n, a.cap, a.depth_bits, a.end_bit, a.max_depth_key, a.hist, a.done);
        },
        std::min(blocks, 4 * 148), SORT_BLOCK_THREADS, ha);
    PassParams<KeyT> P;
    P.keys_a = in.data();
    P.vals_a = vin.data();
    P.keys_b = out.data();
    P.vals_b = vout.data();
    P.keys_c = tmp.data();
    P.vals_c = vtmp.data();
    P.n_dev = &n;
    P.capacity = cap;
    P.depth_bits = depth_bits;
    P.end_bit = end_bit;
    P.blocks = blocks;
    P.max_depth_key = max_depth_key;
    P.hist = hist.data();
    P.state = state.data();
    P.tickets = tickets.data();
    for (int p = 0; p < passes; ++p) {
        P.pass = p;
        simt_emu::launch([](const PassParams<KeyT> &a) { onesweep_pass_kernel<KeyT>(a); }, blocks, SORT_BLOCK_THREADS, P);
    }
    if (in != in_before) return -1;  // the input buffer must never be written
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
        return sort_typed<unsigned int>((const unsigned int *)keys_in, vals_in, (unsigned int *)keys_out, vals_out, n, 0,
                                        end_bit, nullptr);
    return sort_typed<unsigned long long>((const unsigned long long *)keys_in, vals_in, (unsigned long long *)keys_out,
                                          vals_out, n, 0, end_bit, nullptr);
}

// the frame pipeline's call: keys = tile << depth_bits | depth key, *max_depth_key = the frame's largest depth key
extern "C" long long emu_sort_pairs_compacted(const void *keys_in, const int *vals_in, void *keys_out, int *vals_out,
                                              long long n, int key_bytes, int depth_bits, int end_bit,
                                              const int *max_depth_key) {
    if (key_bytes == 4)
        return sort_typed<unsigned int>((const unsigned int *)keys_in, vals_in, (unsigned int *)keys_out, vals_out, n,
                                        depth_bits, end_bit, max_depth_key);
    return sort_typed<unsigned long long>((const unsigned long long *)keys_in, vals_in, (unsigned long long *)keys_out,
                                          vals_out, n, depth_bits, end_bit, max_depth_key);
}

// the digit selector of one pass on its own (csrc/sort.cu make_digit_sel / digit_of / active_passes): digit `pass` of the
// compacted key  tile << live | depth  cut out of the stored key  tile << depth_bits | depth; returns -1 for a pass beyond
// the last active one
extern "C" int emu_sort_digit(unsigned long long key, int key_bytes, int pass, int depth_bits, int end_bit, int live) {
    using namespace gsb;
    if (pass >= active_passes(end_bit, depth_bits, live)) return -1;
    if (key_bytes == 4) return digit_of<unsigned int>((unsigned int)key, make_digit_sel<unsigned int>(pass, depth_bits, live));
    return digit_of<unsigned long long>(key, make_digit_sel<unsigned long long>(pass, depth_bits, live));
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

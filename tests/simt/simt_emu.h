// simt_emu.h -- a minimal lock-step SIMT emulator for CPU-side logic tests of the CUDA kernels (TEST INFRASTRUCTURE:
// nothing under taichi_3d_gaussian_splatting_b200/ uses it, and it is not a fallback -- it runs one CTA at a time on
// one host thread, ~10^5 times slower than the GPU).
//
// A kernel source file is compiled as host C++ with GSB_HOST_EMU defined.  Every CUDA thread of a CTA is a ucontext
// fiber; the fibers run round-robin and switch only inside the collectives (__syncthreads, __syncwarp, shuffles,
// ballots, votes), where a thread deposits its operand and waits until its whole warp / CTA has arrived.  That is
// exactly the convergence contract of the *_sync intrinsics with a full mask, so a kernel that deadlocks or reads a
// partner's stale value here is also wrong on the GPU; what the emulator does NOT model is timing, bank conflicts,
// memory-ordering races between warps (fibers switch only at collectives) and the approximate MUFU functions
// (ex2/rcp/sqrt approximations are libm calls).  __shared__ variables become function-local statics (one CTA at a time).
#pragma once
#define GSB_HOST_EMU 1
#define __host__
#define __device__
#define __global__
#define __shared__ static
#include <cuda_runtime.h>  // vector types and cudaStream_t only
#include <ucontext.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

namespace simt_emu {

struct Dim3 {
    unsigned int x = 1, y = 1, z = 1;
};
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
};
struct WarpState {
    unsigned int gen = 0;
    int arrived = 0;
    unsigned int vals[2][32];
};
struct BlockState {
    unsigned int gen = 0;
    int arrived = 0;
    int conj[2] = {1, 1};
};
struct Machine {
    Dim3 tid, bid, bdim, gdim;
    int cur = 0, nthreads = 0;
    std::vector<Fiber> fibers;
    ucontext_t main_ctx;
    std::function<void()> body;
    WarpState warps[32];
    BlockState block;
    long long switches = 0;
};
inline Machine &M() {
    static Machine m;
    return m;
}
inline char *smem_anchor() {
    static char anchor[16];
    return anchor;
}
inline unsigned char *dynamic_smem() {
    alignas(16) static unsigned char buf[232448];
    return buf;
}
inline long long *counters() {  // work counters a kernel source may bump through GSB_EMU_COUNT (emulation only)
    static long long c[16];
    return c;
}
inline void yield() {
    Machine &m = M();
    ++m.switches;
    swapcontext(&m.fibers[m.cur].ctx, &m.main_ctx);
}
inline void trampoline() {
    Machine &m = M();
    m.body();
    m.fibers[m.cur].done = true;
    swapcontext(&m.fibers[m.cur].ctx, &m.main_ctx);
}
// Deposit v, wait for the 32 lanes of the warp, return the round's 32 operands (valid until the lane's next collective:
// the buffers alternate, and round g+2 cannot start before every lane has arrived at round g+1, i.e. finished reading g).
inline const unsigned int *warp_exchange(unsigned int v) {
    Machine &m = M();
    WarpState &w = m.warps[m.tid.x >> 5];
    const unsigned int g = w.gen;
    w.vals[g & 1][m.tid.x & 31] = v;
    const int width = std::min(32, m.nthreads - (int)(m.tid.x & ~31u));
    if (++w.arrived == width) {
        w.arrived = 0;
        ++w.gen;
    } else {
        while (w.gen == g) yield();
    }
    return w.vals[g & 1];
}
inline int block_barrier(int pred) {
    Machine &m = M();
    BlockState &b = m.block;
    const unsigned int g = b.gen;
    b.conj[g & 1] &= (pred != 0);
    if (++b.arrived == m.nthreads) {
        b.arrived = 0;
        b.conj[(g + 1) & 1] = 1;
        ++b.gen;
    } else {
        while (b.gen == g) yield();
    }
    return b.conj[g & 1];
}

// Run kernel(p) for the CTAs [first, last) of a 1-D grid of `grid` 1-D blocks (launch(): all of them).
template <class Kernel, class Params>
void launch_range(Kernel kernel, int grid, int first, int last, int block, const Params &p) {
    Machine &m = M();
    assert(block % 32 == 0 && block <= 1024);
    m.nthreads = block;
    m.bdim.x = block;
    m.gdim.x = grid;
    if ((int)m.fibers.size() < block) m.fibers.resize(block);
    for (auto &f : m.fibers)
        if (f.stack.empty()) f.stack.resize(256 * 1024);
    m.body = [&]() { kernel(p); };
    for (int b = first; b < last; ++b) {
        m.bid.x = b;
        for (auto &w : m.warps) w = WarpState();
        m.block = BlockState();
        for (int t = 0; t < block; ++t) {
            Fiber &f = m.fibers[t];
            f.done = false;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.data();
            f.ctx.uc_stack.ss_size = f.stack.size();
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, trampoline, 0);
        }
        int alive = block;
        while (alive > 0) {
            int progressed = 0;
            for (int t = 0; t < block; ++t) {
                Fiber &f = m.fibers[t];
                if (f.done) continue;
                m.cur = t;
                m.tid.x = t;
                swapcontext(&m.main_ctx, &f.ctx);
                ++progressed;
                if (f.done) --alive;
            }
            assert(progressed > 0);
        }
    }
}

template <class Kernel, class Params>
void launch(Kernel kernel, int grid, int block, const Params &p) {
    launch_range(kernel, grid, 0, grid, block, p);
}

}  // namespace simt_emu

#define threadIdx (simt_emu::M().tid)
#define blockIdx (simt_emu::M().bid)
#define blockDim (simt_emu::M().bdim)
#define gridDim (simt_emu::M().gdim)

using std::max;
using std::min;

inline void __syncthreads() { simt_emu::block_barrier(1); }
inline int __syncthreads_and(int pred) { return simt_emu::block_barrier(pred); }
inline void __syncwarp(unsigned int mask = 0xffffffffu) {
    assert(mask == 0xffffffffu);
    simt_emu::warp_exchange(0u);
}
namespace simt_emu {
// value of lane `src(lane)` (the caller's own value when src is out of range), 4- or 8-byte operands
template <class T, class Src>
inline T shuffle(unsigned int mask, T v, Src src) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- and 64-bit shuffles only");
    assert(mask == 0xffffffffu);
    (void)mask;
    unsigned int w[2] = {0u, 0u}, got[2] = {0u, 0u};
    std::memcpy(w, &v, sizeof(T));
    const int lane = (int)(threadIdx.x & 31);
    const int from = src(lane);
    for (unsigned int k = 0; k < sizeof(T) / 4; ++k) {
        const unsigned int *r = warp_exchange(w[k]);
        got[k] = (from >= 0 && from < 32) ? r[from] : w[k];
    }
    T out;
    std::memcpy(&out, got, sizeof(T));
    return out;
}
}  // namespace simt_emu
template <class T>
inline T __shfl_xor_sync(unsigned int mask, T v, int lane_mask) {
    return simt_emu::shuffle(mask, v, [=](int lane) { return lane ^ lane_mask; });
}
template <class T>
inline T __shfl_sync(unsigned int mask, T v, int src_lane) {
    return simt_emu::shuffle(mask, v, [=](int) { return src_lane & 31; });
}
template <class T>
inline T __shfl_up_sync(unsigned int mask, T v, unsigned int delta) {
    return simt_emu::shuffle(mask, v, [=](int lane) { return lane - (int)delta; });
}
template <class T>
inline T __shfl_down_sync(unsigned int mask, T v, unsigned int delta) {
    return simt_emu::shuffle(mask, v, [=](int lane) { return lane + (int)delta; });
}
inline unsigned int __ballot_sync(unsigned int mask, int pred) {
    assert(mask == 0xffffffffu);
    const unsigned int *r = simt_emu::warp_exchange(pred ? 1u : 0u);
    unsigned int bits = 0;
    for (int l = 0; l < 32; ++l) bits |= (r[l] & 1u) << l;
    return bits;
}
template <class T>
inline unsigned int __match_any_sync(unsigned int mask, T v) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- and 64-bit operands only");
    assert(mask == 0xffffffffu);
    (void)mask;
    unsigned int w[2] = {0u, 0u};
    std::memcpy(w, &v, sizeof(T));
    unsigned int peers = 0xffffffffu;
    for (unsigned int k = 0; k < sizeof(T) / 4; ++k) {
        const unsigned int *r = simt_emu::warp_exchange(w[k]);
        unsigned int same = 0;
        for (int l = 0; l < 32; ++l) same |= (r[l] == w[k] ? 1u : 0u) << l;
        peers &= same;
    }
    return peers;
}
inline int __any_sync(unsigned int mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __all_sync(unsigned int mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }

inline float atomicAdd(float *a, float v) {
    const float old = *a;
    *a = old + v;
    return old;
}
inline unsigned int atomicAdd(unsigned int *a, unsigned int v) {
    const unsigned int old = *a;
    *a = old + v;
    return old;
}
inline unsigned long long atomicAdd(unsigned long long *a, unsigned long long v) {
    const unsigned long long old = *a;
    *a = old + v;
    return old;
}
inline unsigned int atomicOr(unsigned int *a, unsigned int v) {
    const unsigned int old = *a;
    *a = old | v;
    return old;
}
inline void __threadfence() {}
inline int atomicMax(int *a, int v) {
    const int old = *a;
    *a = std::max(old, v);
    return old;
}
template <class T>
inline T __ldg(const T *p) { return *p; }
inline int __popc(unsigned int x) { return __builtin_popcount(x); }
inline int __ffs(unsigned int x) { return __builtin_ffs((int)x); }
inline unsigned int __float_as_uint(float f) {
    unsigned int u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float __int_as_float(int i) {
    float f;
    std::memcpy(&f, &i, 4);
    return f;
}
inline int __float_as_int(float f) {
    int i;
    std::memcpy(&i, &f, 4);
    return i;
}
inline float __logf(float x) { return logf(x); }
inline int __float2int_rn(float x) { return (int)nearbyintf(x); }

// common.cuh -- shared declarations of libgsb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gsb200.h"

#define GSB_TILE_PIXELS (GSB_TILE_WIDTH * GSB_TILE_HEIGHT)

namespace gsb {

// ---- error plumbing (thread-local message, C ABI returns a code)
void set_error(const char *fmt, ...);
#define GSB_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            gsb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                           __LINE__);                                                          \
            return GSB_ECUDA;                                                                  \
        }                                                                                      \
    } while (0)

// ---- counters living at the head of the workspace
enum Counter { CNT_M = 0, CNT_K = 1, CNT_OVERFLOW = 2, CNT_SORTED_SEL = 3 };
enum Ticket { TICKET_SCAN = 0, TICKET_SORT0 = 1 /* ..+7 */ };

// Per-object pose block in the workspace (20 floats):
//   [0..11]  T_camera_pointcloud 3x4 row-major (R | t)     GP3D:51-62
//   [12..14] camera centre in the pointcloud frame (-R^T t)  UT:495-510
struct PoseBlock {
    float T[12];
    float centre[3];
    float pad[5];
};
static_assert(sizeof(PoseBlock) == 80, "PoseBlock layout");

// Resolved device pointers of one frame's workspace.
struct Workspace {
    long long *counters;
    unsigned int *tickets;
    unsigned long long *scan_state;
    unsigned int *sort_hist;
    unsigned int *sort_state;
    int *tile_start;
    int *tile_end;
    PoseBlock *poses;
    int *point_id;
    int *num_tiles;
    float4 *records;        // 3 float4 per in-camera point
    float *point_in_camera; // 3 floats per in-camera point
    void *keys_a, *keys_b;
    int *vals_a, *vals_b;
    GsbWorkspaceLayout layout;
};

int resolve_workspace(void *base, int64_t bytes, int64_t N, int32_t n_obj, int64_t key_capacity,
                      int32_t H, int32_t W, float far_plane, float depth_scale, uint32_t flags,
                      Workspace *ws);

// ---- stage launchers (each enqueues on `stream`, returns GSB_* code)
int launch_preprocess(const GsbForwardArgs &a, const Workspace &ws, cudaStream_t stream);
int launch_sort(const Workspace &ws, int64_t key_capacity, cudaStream_t stream);
int launch_tile_ranges(const Workspace &ws, int64_t key_capacity, int num_tiles, cudaStream_t stream);
int launch_tile_ranges_raw(const long long *keys_i64, int64_t n, int *tile_start, int *tile_end,
                           int num_tiles, cudaStream_t stream);
int launch_blend_forward(const GsbForwardArgs &a, const Workspace &ws, cudaStream_t stream);
int launch_blend_backward(const GsbBackwardArgs &a, const Workspace &ws, cudaStream_t stream);
int launch_backward_points(const GsbBackwardArgs &a, const Workspace &ws, cudaStream_t stream);

int sort_pairs_device(const void *keys_in, const int *vals_in, void *keys_out, int *vals_out,
                      const long long *n_dev, int64_t n_capacity, int key_bytes, int end_bit,
                      unsigned int *hist /*8*256, zeroed*/, unsigned int *state /*zeroed*/,
                      unsigned int *tickets /*8, zeroed*/, void *tmp_keys, int *tmp_vals,
                      long long *sel_out, cudaStream_t stream);

constexpr int SORT_BLOCK_THREADS = 256;
constexpr int SORT_ITEMS_PER_THREAD = 16;
constexpr int SORT_TILE = SORT_BLOCK_THREADS * SORT_ITEMS_PER_THREAD;  // 4096 keys per CTA
constexpr int SCAN_BLOCK_THREADS = 256;

static inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace gsb

// emu_preprocess.cpp -- the fused per-point stage (pose kernel + K1/P1/K2/K3/P2/K4, csrc/preprocess.cu) compiled as host
// C++ under simt_emu.h.  TEST INFRASTRUCTURE, see simt_emu.h.  The CTAs run one after the other in launch order, which is
// also their ticket order, so the decoupled look-back always finds its predecessors complete.
#include "simt_emu.h"
#include "../../taichi_3d_gaussian_splatting_b200/csrc/preprocess.cu"

extern "C" long long emu_preprocess(long long N, const float *xyz, float *features, const signed char *invalid,
                                    const int *obj_id, int n_obj, const float *q_pc, const float *t_pc, const float *K,
                                    int W, int H, float near_plane, float far_plane, float depth_scale, int depth_bits,
                                    int key_bytes, int filter_tiles, int skip_q_normalise, long long key_capacity,
                                    long long *counters /*8*/, int *point_id, int *point_offset, int *num_tiles,
                                    float *records /*12 N*/, float *point_in_camera /*3 N*/, void *keys, int *vals) {
    using namespace gsb;
    std::vector<PoseBlock> poses(n_obj > 0 ? n_obj : 1);
    struct PoseArgs {
        const float *q, *t;
        int n;
        PoseBlock *out;
    } pa{q_pc, t_pc, n_obj, poses.data()};
    simt_emu::M().switches = 0;
    if (n_obj > 0)
        simt_emu::launch([](const PoseArgs &a) { pose_kernel(a.q, a.t, a.n, a.out); }, (n_obj + 63) / 64, 64, pa);
    const int blocks = (int)((N + SCAN_BLOCK_THREADS - 1) / SCAN_BLOCK_THREADS);
    std::vector<unsigned int> tickets(16, 0u);
    std::vector<unsigned long long> scan_state(blocks + 1, 0ull);
    PreParams p;
    p.N = N;
    p.xyz = xyz;
    p.features = features;
    p.invalid = invalid;
    p.obj_id = obj_id;
    p.poses = poses.data();
    p.K = K;
    p.W = W;
    p.H = H;
    p.near_plane = near_plane;
    p.far_plane = far_plane;
    p.depth_scale = depth_scale;
    p.depth_bits = depth_bits;
    p.skip_q_normalise = skip_q_normalise;
    p.filter_tiles = filter_tiles;
    p.key_capacity = key_capacity;
    p.key_store_limit = key_capacity;
    p.num_blocks = blocks;
    p.counters = counters;
    p.tickets = tickets.data();
    p.scan_state = scan_state.data();
    p.point_id = point_id;
    p.point_offset = point_offset;
    p.num_tiles = num_tiles;
    p.records = reinterpret_cast<float4 *>(records);
    p.point_in_camera = point_in_camera;
    p.keys = keys;
    p.vals = vals;
    if (N > 0) {
        if (key_bytes == 4) simt_emu::launch(preprocess_kernel<unsigned int>, blocks, SCAN_BLOCK_THREADS, p);
        else simt_emu::launch(preprocess_kernel<unsigned long long>, blocks, SCAN_BLOCK_THREADS, p);
    }
    return simt_emu::M().switches;
}

// pose blocks alone (20 floats per object), for the emulated per-point backward
extern "C" void emu_pose(int n_obj, const float *q_pc, const float *t_pc, float *poses_out) {
    using namespace gsb;
    struct PoseArgs {
        const float *q, *t;
        int n;
        PoseBlock *out;
    } pa{q_pc, t_pc, n_obj, reinterpret_cast<PoseBlock *>(poses_out)};
    if (n_obj > 0)
        simt_emu::launch([](const PoseArgs &a) { pose_kernel(a.q, a.t, a.n, a.out); }, (n_obj + 63) / 64, 64, pa);
}

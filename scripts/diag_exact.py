"""GPU diagnostic: which per-point record columns are not bit-identical to the oracle, under which scene factors."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_forward
from gpu_helpers import cuda_scene, make_op, n, run_forward
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene

def check(label, scene):
    o, fwd, feats = oracle_forward(scene)
    sc = cuda_scene(scene)
    op = make_op(exact_exp=True)
    run_forward(op, sc)
    fr = op.last_frame
    rec = n(fr.records)
    exp = np.concatenate([fwd.point_uv, fwd.point_uv_conic_and_rescale[:, :2], fwd.point_uv_conic_and_rescale[:, 2:],
                          fwd.point_alpha_after_activation[:, None], fwd.point_in_camera[:, 2:3], fwd.point_color,
                          fwd.point_radii[:, None]], axis=1)
    names = ["u", "v", "a", "b", "c", "rescale", "opacity", "depth", "r", "g", "b_", "radius"]
    bad = {nm: int((rec[:, i] != exp[:, i]).sum()) for i, nm in enumerate(names)}
    qbad = int((n(sc.point_cloud_features)[:, :4] != feats[:, :4]).sum())
    print(f"{label:40s} M={rec.shape[0]} mismatching entries per column: {bad} q_written_mismatch={qbad}")
    i = np.argmax(np.abs(rec[:, 2] - exp[:, 2]))
    if bad["a"]:
        print("   worst a:", rec[i, :6], exp[i, :6], "feat", feats[fwd.point_id_in_camera_list[i], :8])

base = dict(num_points=3000, height=64, width=96, sigma_med=0.06, seed=21)
check("identity, unit q, sh0", make_scene(**base, sh_degree=0))
check("identity, unit q, sh3", make_scene(**base, sh_degree=3))
s = make_scene(**base, sh_degree=0); s.point_cloud_features[:, :4] *= 1.7
check("identity, q*1.7", s)
check("yaw7, unit q", make_scene(**base, sh_degree=0, yaw_degrees=7.0))
s = make_scene(**base, sh_degree=0); s.point_cloud[:, 2] *= 0.6
check("identity, z*0.6", s)
s = make_scene(**base, sh_degree=0); s.t_pointcloud_camera = torch.tensor([[0.3, -0.2, 0.1]])
check("translated camera", s)

# exact replica of tests/test_gpu_parity._small_scene(21) with near_plane 0.4
def small(seed):
    sc = make_scene(3000, 64, 96, 0.06, seed, sh_degree=3, yaw_degrees=7.0)
    sc.point_cloud[:, 2] = sc.point_cloud[:, 2] * 0.6
    sc.point_cloud_features[:, 7] += 1.0
    sc.point_cloud_features[:, :4] *= 1.7
    sc.point_invalid_mask[::11] = 1
    return sc

def check2(label, scene, **cfg):
    o, fwd, feats = oracle_forward(scene, **cfg)
    sc = cuda_scene(scene)
    op = make_op(exact_exp=True, **cfg)
    run_forward(op, sc)
    fr = op.last_frame
    rec = n(fr.records)
    exp = np.concatenate([fwd.point_uv, fwd.point_uv_conic_and_rescale[:, :2], fwd.point_uv_conic_and_rescale[:, 2:],
                          fwd.point_alpha_after_activation[:, None], fwd.point_in_camera[:, 2:3], fwd.point_color,
                          fwd.point_radii[:, None]], axis=1)
    names = ["u", "v", "a", "b", "c", "rescale", "opacity", "depth", "r", "g", "b_", "radius"]
    bad = {nm: int((rec[:, i] != exp[:, i]).sum()) for i, nm in enumerate(names)}
    print(f"{label:40s} M={rec.shape[0]} mismatches: {bad}")
    rows = np.nonzero((rec != exp).any(axis=1))[0][:5]
    for i in rows:
        pid = fwd.point_id_in_camera_list[i]
        print("   row", i, "id", pid, "\n     got", rec[i], "\n     exp", exp[i], "\n     xyz", scene.point_cloud[pid].numpy(), "feat[:8]", feats[pid, :8])

check2("small(21) near .4", small(21), near_plane=0.4)
check2("small(21) near .8", small(21))
s = small(21); s.point_invalid_mask[:] = 0
check2("small(21) no invalid", s, near_plane=0.4)
s = make_scene(3000, 64, 96, 0.06, 21, sh_degree=3, yaw_degrees=7.0); s.point_cloud_features[:, :4] *= 1.7
check2("yaw7 + q*1.7", s)

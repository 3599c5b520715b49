"""CPU analysis of the blend workload (oracle data): how many (warp-patch, splat) pairs are visited,
how many have any passing lane, lane efficiency.  Samples tiles of a config."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_forward
from taichi_3d_gaussian_splatting_b200.synthetic import CONFIGS, make_scene
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = CONFIGS[name]
sc = make_scene(**cfg)
o, f, _ = oracle_forward(sc)
H, W = cfg["height"], cfg["width"]
tx = W // 16
rng = np.random.default_rng(0)
tiles = rng.choice(f.tile_points_start.shape[0], 150, replace=False)
tot = dict(list=0, pairs_all=0, pairs_before_sat=0, pairs_any_pass=0, lanes_pass=0, lanes_blend=0, pairs_rect=0, pix_evals_before_sat=0)
for t in tiles:
    s, e = f.tile_points_start[t], f.tile_points_end[t]
    if e <= s: continue
    offs = f.point_offset_with_sort_key[s:e]
    uv = f.point_uv[offs]; cr = f.point_uv_conic_and_rescale[offs]; op = f.point_alpha_after_activation[offs]
    tu, tv = t % tx, t // tx
    ys, xs = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    px = (tu * 16 + xs + 0.5).reshape(-1); py = (tv * 16 + ys + 0.5).reshape(-1)
    dx = px[None, :] - uv[:, 0:1]; dy = py[None, :] - uv[:, 1:2]
    power = -0.5 * (dx * dx * cr[:, 0:1] + dy * dy * cr[:, 2:3]) - dx * dy * cr[:, 1:2]
    alpha = np.exp(power) * cr[:, 3:4] * op[:, None]            # (L, 256)
    passing = alpha >= 1 / 255
    last = f.pixel_offset_of_last_effective_point.reshape(H // 16, 16, W // 16, 16)[tv, :, tu, :].reshape(-1) - s  # per pixel
    # saturation index per pixel: pixel stops at first splat where T(1-a)<1e-4; approximate with 'last' (blended until last-1)
    cnt = f.pixel_valid_point_count.reshape(H // 16, 16, W // 16, 16)[tv, :, tu, :].reshape(-1)
    L = e - s
    idx = np.arange(L)[:, None]
    a = np.minimum(alpha, 0.99)
    # recompute per-pixel T sequence to find saturation point
    T = np.ones(256); stop = np.full(256, L)
    for j in range(L):
        act = passing[j] & (stop == L)
        nT = T * (1 - a[j])
        sat = act & (nT < 1e-4)
        stop[sat] = j
        T = np.where(act & ~sat, nT, T)
    warp_of = ((ys // 4) * 2 + (xs // 8)).reshape(-1)
    tot["list"] += L
    for w in range(8):
        m = warp_of == w
        wstop = stop[m].max()  # warp leaves when all lanes saturated
        vis = min(L, wstop + 1)
        tot["pairs_all"] += L
        tot["pairs_before_sat"] += vis
        alive = (idx[:vis] <= stop[m][None, :])
        pas = passing[:vis][:, m] & alive
        tot["pairs_any_pass"] += int(pas.any(axis=1).sum())
        tot["lanes_pass"] += int(pas.sum())
        tot["pix_evals_before_sat"] += int(alive.sum())
    tot["lanes_blend"] += int(cnt.sum())
print(name, {k: v for k, v in tot.items()})
pa, pb, pp = tot["pairs_all"], tot["pairs_before_sat"], tot["pairs_any_pass"]
print(f"pairs before warp saturation / all = {pb/pa:.3f};  pairs with >=1 passing lane / all = {pp/pa:.3f};  "
      f"passing lanes per such pair = {tot['lanes_pass']/max(pp,1):.1f};  blended lanes total/pass = {tot['lanes_blend']/max(tot['lanes_pass'],1):.2f}")

"""Estimate blend work (warp-steps per tile-splat) for alternative patch shapes: 8x4 warp patches (current) vs
two independent 4x4 half-warp sub-patches per warp."""
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
tiles = rng.choice(f.tile_points_start.shape[0], 200, replace=False)
res = dict(L=0, w84=0, w84_lanes=0, h44_pairs=0, h44_warpsteps=0, h44_lanes=0, q88=0)
ys, xs = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
p84 = ((ys // 4) * 2 + (xs // 8)).reshape(-1)
p44 = ((ys // 4) * 4 + (xs // 4)).reshape(-1)
p88 = ((ys // 8) * 2 + (xs // 8)).reshape(-1)
for t in tiles:
    s, e = f.tile_points_start[t], f.tile_points_end[t]
    if e <= s: continue
    offs = f.point_offset_with_sort_key[s:e]
    uv = f.point_uv[offs]; cr = f.point_uv_conic_and_rescale[offs]; op = f.point_alpha_after_activation[offs]
    tu, tv = t % tx, t // tx
    px = (tu * 16 + xs + 0.5).reshape(-1); py = (tv * 16 + ys + 0.5).reshape(-1)
    dx = px[None, :] - uv[:, 0:1]; dy = py[None, :] - uv[:, 1:2]
    power = -0.5 * (dx * dx * cr[:, 0:1] + dy * dy * cr[:, 2:3]) - dx * dy * cr[:, 1:2]
    passing = (np.exp(power) * cr[:, 3:4] * op[:, None]) >= 1 / 255   # (L,256); ignores saturation (upper bound)
    L = e - s
    res["L"] += L
    for w in range(8):
        anyp = passing[:, p84 == w].any(axis=1)
        res["w84"] += int(anyp.sum()); res["w84_lanes"] += int(passing[:, p84 == w].sum())
    cnt44 = np.zeros(16, int)
    for q in range(16):
        anyp = passing[:, p44 == q].any(axis=1)
        cnt44[q] = anyp.sum()
        res["h44_lanes"] += int(passing[:, p44 == q].sum())
    res["h44_pairs"] += int(cnt44.sum())
    # warp w handles sub-patches (2w, 2w+1) in lockstep -> steps = max of the two
    res["h44_warpsteps"] += int(sum(max(cnt44[2 * w], cnt44[2 * w + 1]) for w in range(8)))
    for q in range(4):
        res["q88"] += int(passing[:, p88 == q].any(axis=1).sum())
L = res["L"]
print(name, res)
print(f"per (tile,splat): 8x4 warp visits={res['w84']/L:.2f} (lanes/visit {res['w84_lanes']/res['w84']:.1f}); "
      f"4x4 half-warp visits={res['h44_pairs']/L:.2f} -> warp steps={res['h44_warpsteps']/L:.2f} (lanes/step {res['h44_lanes']/res['h44_warpsteps']:.1f}); "
      f"8x8 visits={res['q88']/L:.2f}")

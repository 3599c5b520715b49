"""Host-side replay of the device algorithm that decides which (tile, splat) pairs get a sort key
(``csrc/common.cuh``: make_splat_reach / rect_reachable; ``csrc/preprocess.cu``: warp-cooperative filter and
emission), checked against the oracle's lists on CPU:

* with the filter off the replay emits exactly the reference's pairs in the reference's order (GPCR:131-172);
* with the filter on the emitted pairs are an order-preserving subset, only the first 64 tiles of a splat are
  ever tested, and EVERY dropped pair has alpha < 1/255 on all 256 pixel centres of its tile in the oracle's
  arithmetic -- i.e. the filter is conservative and cannot change an output.
The replay mirrors the CUDA code line by line (same masks, same prefix searches, same j-th-set-bit selection), so an
algorithmic slip there shows up here without a GPU.  It is test infrastructure, not a CPU path of the product."""
import math

import numpy as np
import pytest

from helpers import oracle_forward
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene


def _bbox(u, v, radius, W, H):  # GPCR:81-103
    radius = max(radius, 1.0)
    min_u, max_u, min_v, max_v = max(0.0, u - radius), u + radius, max(0.0, v - radius), v + radius
    tw, th = W // 16, H // 16
    a = min(int(math.floor(min_u / 16)), tw)
    b = min(max(int(math.floor(max_u / 16)) + 1, a + 1), tw)
    c = min(int(math.floor(min_v / 16)), th)
    d = min(max(int(math.floor(max_v / 16)) + 1, c + 1), th)
    return a, b, c, d


def _make_reach(a, b, c, ro):  # common.cuh make_splat_reach
    r = dict(a=a, b2=2 * b, c=c, nb_ic=-b / c if c else math.nan, nb_ia=-b / a if a else math.nan,
             t2=2 * 1.002 * math.log(max(255 * ro, 1.0)) + 1e-3)
    det = a * c - b * b
    r["mode"] = 2 if (ro != ro or not det > 0 or not a > 0 or not c > 0) else (0 if ro < 0.999 / 255 else 1)
    return r


def _rect_reachable(r, X0, X1, Y0, Y1):  # common.cuh rect_reachable
    if X0 <= 0 <= X1 and Y0 <= 0 <= Y1:
        return True
    q = lambda dx, dy: dx * (r["b2"] * dy + r["a"] * dx) + r["c"] * dy * dy  # noqa: E731
    cl = lambda x, lo, hi: min(max(x, lo), hi)  # noqa: E731
    best = min(q(X0, cl(r["nb_ic"] * X0, Y0, Y1)), q(X1, cl(r["nb_ic"] * X1, Y0, Y1)),
               q(cl(r["nb_ia"] * Y0, X0, X1), Y0), q(cl(r["nb_ia"] * Y1, X0, X1), Y1))
    return not best > r["t2"]


def _select_bit32(m, j):  # preprocess.cu select_bit32
    pos = 0
    for w in (16, 8, 4, 2, 1):
        c = bin(m & ((1 << w) - 1)).count("1")
        if j >= c:
            j, pos, m = j - c, pos + w, m >> w
    return pos


def _owner(pref, q):
    lo = 0
    for step in (16, 8, 4, 2, 1):
        if lo + step < len(pref) and pref[lo + step] <= q:
            lo += step
    return lo


def replay_emission(fwd, H, W, filter_tiles):
    tiles_x, out, tested_beyond_64 = W // 16, [], 0
    M = fwd.point_uv.shape[0]
    for w0 in range(0, M, 32):
        st = []
        for off in range(w0, min(w0 + 32, M)):
            u, v = map(float, fwd.point_uv[off])
            a, b, c, rescale = map(float, fwd.point_uv_conic_and_rescale[off])
            min_tu, max_tu, min_tv, max_tv = _bbox(u, v, float(fwd.point_radii[off]), W, H)
            nt = (max_tu - min_tu) * (max_tv - min_tv)
            assert nt == fwd.num_overlap_tiles[off]
            r = _make_reach(a, b, c, rescale * float(fwd.point_alpha_after_activation[off]))
            if not filter_tiles:
                r["mode"] = 2
            n64 = min(nt, 64)
            st.append(dict(u=u, v=v, r=r, min_tu=min_tu, min_tv=min_tv, ntv=max(max_tv - min_tv, 1), nt=nt, off=off,
                           mask=((1 << n64) - 1) if r["mode"] == 2 else 0, tcap=n64 if r["mode"] == 1 else 0))
        pref = [0]
        for s in st:
            pref.append(pref[-1] + s["tcap"])
        for q in range(pref[-1]):  # one lane per PAIR
            s = st[_owner(pref, q)]
            idx = q - pref[st.index(s)]
            du = idx // s["ntv"]
            X0 = (s["min_tu"] + du) * 16 + 0.5 - s["u"]
            Y0 = (s["min_tv"] + idx - du * s["ntv"]) * 16 + 0.5 - s["v"]
            if _rect_reachable(s["r"], X0, X0 + 15, Y0, Y0 + 15):
                s["mask"] |= 1 << idx
        pk = [0]
        for s in st:
            beyond = max(s["nt"] - 64, 0)
            s["nkeys"] = 0 if s["r"]["mode"] == 0 else bin(s["mask"]).count("1") + beyond
            s["nk64"] = s["nkeys"] - beyond
            pk.append(pk[-1] + s["nkeys"])
        for q in range(pk[-1]):  # one lane per KEY
            lo = _owner(pk, q)
            s, j = st[lo], q - pk[lo]
            if j < s["nk64"]:
                mlo = s["mask"] & 0xFFFFFFFF
                plo = bin(mlo).count("1")
                idx = _select_bit32(mlo, j) if j < plo else 32 + _select_bit32(s["mask"] >> 32, j - plo)
            else:
                idx = 64 + (j - s["nk64"])
                tested_beyond_64 += 1
            du = idx // s["ntv"]
            out.append((s["off"], (s["min_tu"] + du) + (s["min_tv"] + idx - du * s["ntv"]) * tiles_x))
    return out, tested_beyond_64


def reference_pairs(fwd, H, W):
    ref = []
    for off in range(fwd.point_uv.shape[0]):
        a, b, c, d = _bbox(float(fwd.point_uv[off, 0]), float(fwd.point_uv[off, 1]), float(fwd.point_radii[off]), W, H)
        ref += [(off, tu + tv * (W // 16)) for tu in range(a, b) for tv in range(c, d)]  # tile_u outer, tile_v inner
    return ref


@pytest.mark.parametrize("npts,H,W,sigma,logit_shift", [(1500, 128, 192, 0.3, -2.0), (1200, 128, 192, 0.6, 0.0),
                                                         (4000, 96, 128, 0.05, 0.0)])
def test_reach_filter_replay_is_conservative_and_order_preserving(npts, H, W, sigma, logit_shift):
    scene = make_scene(npts, H, W, sigma, 31, sh_degree=0, yaw_degrees=4.0)
    scene.point_cloud_features[:, 7] += logit_shift
    _, fwd, _ = oracle_forward(scene)
    ref = reference_pairs(fwd, H, W)
    assert len(ref) == fwd.point_offset_with_sort_key.shape[0]
    unfiltered, _ = replay_emission(fwd, H, W, filter_tiles=False)
    assert unfiltered == ref
    kept, beyond = replay_emission(fwd, H, W, filter_tiles=True)
    where = {pair: i for i, pair in enumerate(ref)}
    order = [where[pair] for pair in kept]  # KeyError = a pair the reference does not have
    assert order == sorted(order) and len(set(order)) == len(order)
    if sigma >= 0.3:
        assert beyond > 0 and len(kept) < 0.8 * len(ref)
    # every dropped pair is dead on all 256 pixel centres of its tile (oracle arithmetic, UT:275-284)
    dropped = sorted(set(range(len(ref))) - set(order))
    assert dropped
    ys, xs = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    worst = 0.0
    for i in dropped:
        off, tile = ref[i]
        px = (tile % (W // 16)) * 16 + xs.reshape(-1) + 0.5
        py = (tile // (W // 16)) * 16 + ys.reshape(-1) + 0.5
        a, b, c, rescale = fwd.point_uv_conic_and_rescale[off].astype(np.float64)
        dx, dy = px - fwd.point_uv[off, 0], py - fwd.point_uv[off, 1]
        alpha = np.exp(-0.5 * (dx * dx * a + dy * dy * c) - dx * dy * b) * rescale * fwd.point_alpha_after_activation[off]
        worst = max(worst, float(alpha.max()))
    assert worst < 1.0 / 255.0, worst

"""Independent dense evaluator of the rasteriser in float64 torch (test helper).

This is NOT the oracle: it is a second, structurally different restatement (brute force over
pixels x splats, differentiable, autograd for the backward) used to cross-check the C oracle
(SURVEY.md §8(c) "what is NOT pinned by the reference's tests").  The reference's omissions are
made explicit so autograd reproduces the hand-written K7 gradients:

* J inside Sigma' is detached            (GaussianPoint3D.py:237-331 differentiates only q and s)
* the SH view direction is detached      (GaussianPointCloudRasterisation.py:749-756)
* ``rescale`` is detached                (utils.py:347 "we don't intend to differentiate w.r.t. rescale")
* q is treated as already normalised     (GaussianPoint3D.py:318-330, no normalisation Jacobian)
* the 0.99 clamp is straight-through     (GaussianPointCloudRasterisation.py:635-662)
* no background colour                   (GaussianPointCloudRasterisation.py:475-477)
* it applies the reference's tile-membership mask (3-sigma square bbox, GPCR:81-103), because that
  cut-off is visible in the output.
"""
import math

import torch

SH_C = [0.28209479177387814, 0.48860251190291987, 1.0925484305920792, 0.94617469575755997,
        0.31539156525251999, 0.54627421529603959, 0.59004358992664352, 2.8906114426405538,
        0.45704579946446572, 0.3731763325901154, 1.4453057213202769]


def sh_basis(d):
    d = d / d.norm(dim=-1, keepdim=True)
    x, y, z = d.unbind(-1)
    return torch.stack([
        torch.full_like(x, SH_C[0]), -SH_C[1] * y, SH_C[1] * z, -SH_C[1] * x,
        SH_C[2] * x * y, -SH_C[2] * y * z, SH_C[3] * z * z - SH_C[4], -SH_C[2] * x * z,
        SH_C[5] * x * x - SH_C[5] * y * y,
        SH_C[6] * y * (-3.0 * x * x + y * y), SH_C[7] * x * y * z, SH_C[8] * y * (1.0 - 5.0 * z * z),
        SH_C[9] * z * (5.0 * z * z - 3.0), SH_C[8] * x * (1.0 - 5.0 * z * z),
        SH_C[10] * z * (x * x - y * y), SH_C[6] * x * (-x * x + 3.0 * y * y)], dim=-1)


def quat_to_rot(q):
    x, y, z, w = q.unbind(-1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)


def dense_render(xyz, feats, invalid_mask, K, q_cam_pc, t_cam_pc, H, W, near=0.8, far=1000.0,
                 depth_scale=100.0):
    """Single-object scene. xyz (N,3), feats (N,56) (q assumed unit). Returns image (H,W,3) f64 and
    a dict of intermediates. Differentiable w.r.t. xyz / feats as the reference defines it."""
    dt = torch.float64
    xyz = xyz.to(dt)
    feats = feats.to(dt)
    K = K.to(dt)
    Rc = quat_to_rot(q_cam_pc.to(dt).reshape(4))
    tc = t_cam_pc.to(dt).reshape(3)
    pc = xyz @ Rc.T + tc
    z = pc[:, 2]
    uv = ((pc @ K.T) / z[:, None])[:, :2]
    inside = (invalid_mask.to(torch.bool) == 0) & (z > near) & (z < far) & (uv[:, 0] >= -48) & \
        (uv[:, 0] < W + 48) & (uv[:, 1] >= -48) & (uv[:, 1] < H + 48)
    ids = torch.nonzero(inside.detach()).reshape(-1)
    pc, uv, z = pc[ids], uv[ids], z[ids]
    f = feats[ids]
    M = ids.shape[0]
    q, s, logit = f[:, 0:4], f[:, 4:7], f[:, 7]
    pcd = pc.detach()
    fx, fy = K[0, 0], K[1, 1]
    zeros = torch.zeros_like(pcd[:, 0])
    J = torch.stack([torch.stack([fx / pcd[:, 2], zeros, -fx * pcd[:, 0] / pcd[:, 2] ** 2], -1),
                     torch.stack([zeros, fy / pcd[:, 2], -fy * pcd[:, 1] / pcd[:, 2] ** 2], -1)], -2)
    R = quat_to_rot(q)
    S2 = torch.diag_embed(torch.exp(2 * s))
    Sigma = R @ S2 @ R.transpose(-1, -2)
    U = J @ Rc
    cov = U @ Sigma @ U.transpose(-1, -2)
    a0, b0, c0, d0 = cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 0], cov[:, 1, 1]
    det0 = a0 * d0 - b0 * c0
    a1, d1 = a0 + 0.3, d0 + 0.3
    det1 = a1 * d1 - b0 * c0
    rescale = torch.sqrt(torch.clamp(det0 / det1, min=0.0)).detach()
    ca, cb, cc = d1 / det1, -b0 / det1, a1 / det1
    opacity = torch.sigmoid(logit)
    cam_centre = -(Rc.T @ tc)
    basis = sh_basis((xyz[ids] - cam_centre).detach())
    sh = f[:, 8:56].reshape(M, 3, 16)
    color = torch.sigmoid((sh * basis[:, None, :]).sum(-1))
    lam = (a0 + d0 + torch.sqrt((a0 - d0) ** 2 + 4 * b0 * c0)) / 2
    radius = (3.0 * torch.sqrt(lam)).detach().to(torch.float32)  # bbox decisions in f32 like the op
    uvf = uv.detach().to(torch.float32)
    r = torch.clamp(radius, min=1.0)
    tw, th = W // 16, H // 16
    min_tu = torch.clamp(torch.floor(torch.clamp(uvf[:, 0] - r, min=0.0) / 16).to(torch.int64), max=tw)
    max_tu = torch.clamp(torch.maximum(torch.floor((uvf[:, 0] + r) / 16).to(torch.int64) + 1, min_tu + 1), max=tw)
    min_tv = torch.clamp(torch.floor(torch.clamp(uvf[:, 1] - r, min=0.0) / 16).to(torch.int64), max=th)
    max_tv = torch.clamp(torch.maximum(torch.floor((uvf[:, 1] + r) / 16).to(torch.int64) + 1, min_tv + 1), max=th)
    depth_key = (z.detach().to(torch.float32) * torch.tensor(depth_scale, dtype=torch.float32)).to(torch.int32)
    order = torch.argsort(depth_key.to(torch.int64) * (M + 1) + torch.arange(M), stable=True)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    px = xs.to(dt) + 0.5
    py = ys.to(dt) + 0.5
    ptu, ptv = xs // 16, ys // 16
    T = torch.ones((H, W), dtype=dt)
    C = torch.zeros((H, W, 3), dtype=dt)
    D = torch.zeros((H, W), dtype=dt)
    Wt = torch.zeros((H, W), dtype=dt)
    cnt = torch.zeros((H, W), dtype=torch.int32)
    stopped = torch.zeros((H, W), dtype=torch.bool)
    for m in order.tolist():
        member = (ptu >= min_tu[m]) & (ptu < max_tu[m]) & (ptv >= min_tv[m]) & (ptv < max_tv[m])
        if not bool(member.any()):
            continue
        dx = px - uv[m, 0]
        dy = py - uv[m, 1]
        G = torch.exp(-0.5 * (dx * dx * ca[m] + dy * dy * cc[m]) - dx * dy * cb[m]) * rescale[m]
        alpha = G * opacity[m]
        active = member & ~stopped & (alpha.detach() >= 1.0 / 255.0)
        alpha_c = alpha + (torch.clamp(alpha, max=0.99) - alpha).detach()
        nT = T * (1 - alpha_c)
        stop_now = active & (nT.detach() < 1e-4)
        stopped = stopped | stop_now
        blend = active & ~stop_now
        w = alpha_c * T
        C = C + torch.where(blend[..., None], color[m][None, None, :] * w[..., None], torch.zeros_like(C))
        D = D + torch.where(blend, z[m].detach() * w, torch.zeros_like(D))
        Wt = Wt + torch.where(blend, w, torch.zeros_like(Wt))
        cnt = cnt + blend.to(torch.int32)
        T = torch.where(blend, nT, T)
    aux = dict(ids=ids, uv=uv, pc=pc, conic=torch.stack([ca, cb, cc, rescale], -1), opacity=opacity,
               color=color, radius=radius, depth=D / torch.clamp(Wt, min=1e-6), acc_alpha=1 - T,
               count=cnt, ntiles=(max_tu - min_tu) * (max_tv - min_tv))
    return C, aux


def postprocess_feature_grads(g, band, q_f=1.0, s_f=0.5, a_f=20.0, c_f=5.0, h_f=1.0):
    """GaussianPointCloudRasterisation.py:1102-1125, 1167-1182."""
    g = g.clone()
    first = {0: 1, 1: 4, 2: 9}.get(int(band), 16)
    for ch in range(3):
        g[:, 8 + 16 * ch + first: 8 + 16 * (ch + 1)] = 0
    g[:, :4] *= q_f
    g[:, 4:7] *= s_f
    g[:, 7] *= a_f
    for ch in range(3):
        g[:, 8 + 16 * ch] *= c_f
        g[:, 9 + 16 * ch: 8 + 16 * (ch + 1)] *= h_f
    return g

"""-m gpu: the compact view-parallel gradient exchange (GSB_FLAG_COMPACT_GRADS + gsb200_expand_view_gradients) on ONE GPU:
R "ranks" run one after the other, a stand-in for ``parallel.ViewParallelExchange`` first records every rank's buffers and
then replays the collectives' results (sum of the (N,12) columns, gather of the per-view blocks).  The gradients the
operator returns in exchange mode must equal the SUM over the views of the dense gradients (what one all-reduce of the
dense buffers gives).  The real NCCL collectives are exercised by ``bench.py --gpus N`` (which checks the same identity
against a dense all-reduce once per run) and, on CPU, by the world-size-2 gloo test."""
import numpy as np
import pytest
import torch

from taichi_3d_gaussian_splatting_b200.synthetic import make_scene

from gpu_helpers import cuda_scene, make_op, n, run_forward
from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR

pytestmark = pytest.mark.gpu


class _LocalExchange:
    def __init__(self, world, rank, store, replay):
        self.world, self.rank, self.store, self.replay = world, rank, store, replay

    def allocate(self, num_points, num_objects, device):
        stride = (3 * num_points + 3 * num_objects + 3) // 4 * 4
        return (torch.empty((num_points, 12), device=device), torch.empty((self.world, stride), device=device))

    def rows_written(self, grad_sum, blocks):
        pass

    def run(self, grad_sum, blocks):
        if not self.replay:
            self.store[self.rank] = (grad_sum.clone(), blocks[self.rank].clone())
            return
        grad_sum.zero_()
        for r in range(self.world):  # rank order, like the expansion kernel's own sum
            grad_sum += self.store[r][0]
            blocks[r].copy_(self.store[r][1])

    def run_and_expand(self, grad_sum, blocks, expand):
        """The split expansion of parallel.MulticastViewParallelExchange: the SH columns from the gathered blocks while the
        summed columns are still unsummed (here: poisoned), the summed columns afterwards."""
        if not self.replay:
            self.run(grad_sum, blocks)
            return expand(0)
        own = grad_sum.clone()
        for r in range(self.world):
            blocks[r].copy_(self.store[r][1])
        grad_sum.fill_(float("nan"))  # part 1 must not read it
        expand(1)
        grad_sum.copy_(own)
        self.run(grad_sum, blocks)
        expand(2)


@pytest.mark.parametrize("band", [3, 1])
@pytest.mark.parametrize("with_hook", [False, True])
def test_exchange_mode_returns_the_sum_over_views(band, with_hook):
    R = 3
    base = make_scene(20000, 128, 192, 0.04, 9, sh_degree=3)
    base.point_object_id[1::2] = 1
    views = []
    for v in range(R):
        sc = make_scene(20000, 128, 192, 0.04, 9, sh_degree=3, yaw_degrees=6.0 * v - 5.0)
        sc.point_object_id = base.point_object_id.clone()
        q0 = sc.q_pointcloud_camera[0]
        q1 = torch.tensor([0.01 * v, float(q0[1]) + 0.02, -0.01, float(q0[3])])
        sc.q_pointcloud_camera = torch.stack([q0, q1 / q1.norm()])
        sc.t_pointcloud_camera = torch.tensor([[0.05 * v, 0.0, -0.1], [0.2, -0.05 * v, -0.3]])
        views.append(cuda_scene(sc))
    xyz = views[0].point_cloud.clone().requires_grad_(True)
    feat = views[0].point_cloud_features.clone().requires_grad_(True)
    grads_img = [torch.randn((128, 192, 3), generator=torch.Generator().manual_seed(50 + v)).cuda() for v in range(R)]

    def run(op, v):
        sc = views[v]
        sc.point_cloud, sc.point_cloud_features = xyz, feat
        xyz.grad = feat.grad = None
        image, _, _ = run_forward(op, sc, band=band)
        image.backward(grads_img[v])
        return xyz.grad.clone(), feat.grad.clone()

    hooks = {}
    dense_x, dense_f = torch.zeros_like(xyz), torch.zeros_like(feat)
    again_x, again_f = torch.zeros_like(xyz), torch.zeros_like(feat)
    for v in range(R):
        gx, gf = run(make_op(hook=(lambda h, v=v: hooks.setdefault(("dense", v), h)) if with_hook else None), v)
        dense_x += gx
        dense_f += gf
        gx, gf = run(make_op(), v)  # the same thing once more: the float atomics of loop A land in another order every run
        again_x += gx
        again_f += gf
    noise_x = float((again_x - dense_x).abs().max())
    noise_f = float((again_f - dense_f).abs().max())
    store = {}
    for v in range(R):  # every rank's compact buffers
        op = GPCR(GPCR.GaussianPointCloudRasterisationConfig(), gradient_exchange=_LocalExchange(R, v, store, False))
        run(op, v)
    for v in (0, R - 1):  # what ranks 0 and R-1 see after the exchange
        hook = (lambda h, v=v: hooks.setdefault(("compact", v), h)) if with_hook else None
        op = GPCR(GPCR.GaussianPointCloudRasterisationConfig(), backward_valid_point_hook=hook,
                  gradient_exchange=_LocalExchange(R, v, store, True))
        gx, gf = run(op, v)
        # equal up to the run-to-run noise of the float atomics (measured above on the dense path itself) ...
        assert float((gx - dense_x).abs().max()) <= 4.0 * noise_x + 2e-6 * float(dense_x.abs().max())
        assert float((gf - dense_f).abs().max()) <= 4.0 * noise_f + 2e-6 * float(dense_f.abs().max())
        assert noise_x <= 1e-3 * float(dense_x.abs().max()) and noise_f <= 1e-3 * float(dense_f.abs().max())
        cleared = {3: 16, 1: 4}[band]
        sh = gf[:, 8:].reshape(-1, 3, 16)
        assert float(sh[:, :, cleared:].abs().max()) == 0.0 if cleared < 16 else True
        if with_hook:  # the hook of a rank sees that rank's own view
            hd, hc = hooks[("dense", v)], hooks[("compact", v)]
            assert torch.equal(hd.point_id_in_camera_list, hc.point_id_in_camera_list)
            assert float((hd.grad_point_in_camera - hc.grad_point_in_camera).abs().max()) <= 4.0 * noise_x + 2e-6 * float(dense_x.abs().max())
            assert torch.equal(hd.num_affected_pixels, hc.num_affected_pixels)
            assert hc.grad_pointfeatures_in_camera is None

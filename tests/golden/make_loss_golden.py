"""Golden values for the loss row: the REFERENCE's LossFunction (LossFunction.py:8-51, imported from /root/reference) on
seeded inputs.  Its SSIM term comes from the third-party ``pytorch_msssim`` package, which is not installed here, so that
one call is replaced by a fixed stand-in (``placeholder_ssim`` below) on BOTH sides: what is pinned is everything the
reference itself wrote -- the L1 term, the (1 - lambda) / lambda mix, the exp(s) regulariser over the valid points and its
weight -- not the SSIM arithmetic (unpinned).

    python tests/golden/make_loss_golden.py        # build container only; writes loss_vectors.json
"""
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def placeholder_ssim(x, y, data_range=1, size_average=True):
    return (x * y).mean() / data_range


def inputs(seed):
    g = torch.Generator().manual_seed(seed)
    pred = torch.rand((3, 40, 56), generator=g)
    gt = torch.rand((3, 40, 56), generator=g)
    feats = torch.randn((50, 56), generator=g)
    mask = (torch.rand((50,), generator=g) < 0.3).to(torch.int8)
    return pred, gt, mask, feats


CASES = [dict(seed=1, config=dict()), dict(seed=2, config=dict(lambda_value=0.35, regularization_weight=0.5)),
         dict(seed=3, config=dict(enable_regularization=False)), dict(seed=4, config=dict(), batched=True)]


def main():
    sys.modules["pytorch_msssim"] = types.SimpleNamespace(ssim=placeholder_ssim)
    sys.modules.setdefault("dataclass_wizard", types.SimpleNamespace(YAMLWizard=object))
    sys.path.insert(0, "/root/reference")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_loss", "/root/reference/taichi_3d_gaussian_splatting/LossFunction.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = []
    for case in CASES:
        pred, gt, mask, feats = inputs(case["seed"])
        if case.get("batched"):
            pred, gt = pred.unsqueeze(0), gt.unsqueeze(0)
        fn = ref.LossFunction(ref.LossFunction.LossFunctionConfig(**case["config"]))
        loss, l1, ld = fn(pred, gt, point_invalid_mask=mask, pointcloud_features=feats)
        loss_nofeat, _, _ = fn(pred, gt)
        out.append(dict(loss=float(loss), l1=float(l1), ld_ssim=float(ld), loss_without_features=float(loss_nofeat)))
    with open(os.path.join(HERE, "loss_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()

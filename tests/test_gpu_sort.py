"""-m gpu: the device radix sort (gsb200_sort_pairs) against torch.sort(stable=True): bit-exact keys,
bit-exact payload order (stability), across sizes that hit the empty / partial-tile / TMA-tail paths."""
import pytest
import torch

from taichi_3d_gaussian_splatting_b200 import _lib

pytestmark = pytest.mark.gpu


def _sort(keys, vals, end_bit):
    lib = _lib.load()
    nkeys = keys.shape[0]
    kb = keys.element_size()
    temp = torch.empty(max(int(lib.gsb200_sort_temp_bytes(nkeys, kb)), 256), dtype=torch.uint8, device="cuda")
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    _lib.check(lib.gsb200_sort_pairs(keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(), nkeys, kb,
                                     end_bit, temp.data_ptr(), temp.shape[0],
                                     torch.cuda.current_stream().cuda_stream), "gsb200_sort_pairs")
    torch.cuda.synchronize()
    return ko, vo


@pytest.mark.parametrize("nkeys", [0, 1, 3, 31, 3071, 3072, 3073, 4096, 6143, 9217, 100_003, 1_000_000])
@pytest.mark.parametrize("key_bytes,end_bit", [(4, 30), (4, 13), (4, 32), (8, 45), (8, 64)])
def test_sort_pairs_matches_stable_sort(nkeys, key_bytes, end_bit):
    g = torch.Generator(device="cpu").manual_seed(nkeys * 7 + end_bit)
    hi = min(end_bit, 62)
    # few distinct values -> many ties, so stability is really exercised
    distinct = 1 << min(hi, 10 if nkeys > 1000 else hi)
    raw = torch.randint(0, distinct, (nkeys,), generator=g, dtype=torch.int64)
    spread = raw * max(1, ((1 << hi) - 1) // max(distinct - 1, 1))
    if key_bytes == 4:
        keys = (spread & 0x7FFFFFFF).to(torch.int32).cuda() if end_bit < 32 else spread.to(torch.int32).cuda()
        ref_keys = keys.to(torch.int64) & 0xFFFFFFFF
    else:
        keys = spread.cuda()
        ref_keys = keys
    vals = torch.arange(nkeys, dtype=torch.int32, device="cuda")
    ko, vo = _sort(keys, vals, end_bit)
    if nkeys == 0:
        return
    exp_k, perm = torch.sort(ref_keys, stable=True)
    got_k = ko.to(torch.int64) & 0xFFFFFFFF if key_bytes == 4 else ko
    assert torch.equal(got_k, exp_k)
    assert torch.equal(vo.to(torch.int64), perm)


def test_sort_leaves_input_intact():
    keys = torch.randint(0, 1 << 20, (50_000,), dtype=torch.int64).to(torch.int32).cuda()
    vals = torch.arange(50_000, dtype=torch.int32, device="cuda")
    k0, v0 = keys.clone(), vals.clone()
    _sort(keys, vals, 20)
    assert torch.equal(keys, k0) and torch.equal(vals, v0)

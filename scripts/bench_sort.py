"""Micro-benchmark of gsb200_sort_pairs for one or more library variants: python scripts/bench_sort.py lib1.so lib2.so"""
import ctypes, sys, os, torch
n = 4_044_076
torch.manual_seed(0)
tile = torch.randint(0, 8040, (n,), dtype=torch.int64)
depth = torch.randint(200, 1000, (n,), dtype=torch.int64)
keys = ((tile << 17) | depth).to(torch.int32).cuda()
vals = torch.arange(n, dtype=torch.int32, device="cuda")
exp_k, perm = torch.sort(keys.to(torch.int64) & 0xFFFFFFFF, stable=True)
for path in sys.argv[1:]:
    lib = ctypes.CDLL(os.path.abspath(path))
    lib.gsb200_sort_temp_bytes.restype = ctypes.c_int64
    lib.gsb200_sort_temp_bytes.argtypes = [ctypes.c_int64, ctypes.c_int32]
    lib.gsb200_sort_pairs.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    tb = lib.gsb200_sort_temp_bytes(n, 4)
    temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    st = torch.cuda.current_stream().cuda_stream
    def run():
        rc = lib.gsb200_sort_pairs(keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(), n, 4, 30, temp.data_ptr(), tb, st)
        assert rc == 0
    for _ in range(3): run()
    torch.cuda.synchronize()
    ok = torch.equal(ko.to(torch.int64) & 0xFFFFFFFF, exp_k) and torch.equal(vo.to(torch.int64), perm)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print(f"{path}: {e0.elapsed_time(e1)/20*1e3:.1f} us per sort of {n} keys (incl. temp memset), correct={ok}")

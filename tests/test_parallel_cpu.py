"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes (views shard across
ranks, one gradient-sum exchange per step; no collective on the render path itself)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taichi_3d_gaussian_splatting_b200.parallel import ViewParallelExchange, exchange_gradients, shard_views

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_views_partition():
    for world in (1, 2, 3, 8):
        for views in (0, 1, 7, 8, 30):
            shards = [shard_views(views, r, world) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(views))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1000
    g_xyz = torch.full((n, 3), float(rank + 1))
    g_feat = torch.arange(n * 56, dtype=torch.float32).reshape(n, 56) * (rank + 1)
    exchange_gradients([g_xyz, None, g_feat])
    g_avg = torch.full((4,), float(rank))
    exchange_gradients([g_avg], average=True)
    flat = torch.arange(20, dtype=torch.float32) * (rank + 1)  # two views into one allocation -> one collective
    va, vb = flat[:6].view(2, 3), flat[8:20].view(3, 4)
    exchange_gradients([va, vb], fused_buffer=flat)
    other = torch.ones(4) * (rank + 1)  # not a view of `flat`: falls back to per-tensor exchange
    exchange_gradients([va, other], fused_buffer=flat)
    handles = exchange_gradients([torch.ones(3)], async_op=True)
    for h in handles:
        h.wait()
    # the compact exchange: (N,12) summable columns all-reduced, the per-view blocks all-gathered into their rank's slot
    ex = ViewParallelExchange()
    assert (ex.world, ex.rank) == (world, rank)
    gsum = torch.full((n, 12), float(rank + 1))
    blocks = torch.full((world, 3 * n + 8), -1.0)
    blocks[rank] = torch.arange(3 * n + 8, dtype=torch.float32) + 1000.0 * rank
    ex.run(gsum, blocks)
    ex2 = ViewParallelExchange(gather_group=dist.new_group())  # all-gather on a communicator of its own
    gsum2 = torch.full((n, 12), float(rank + 1))
    blocks2 = torch.full((world, 3 * n + 8), -1.0)
    blocks2[rank] = torch.arange(3 * n + 8, dtype=torch.float32) + 1000.0 * rank
    ex2.run(gsum2, blocks2)
    assert torch.equal(gsum2, gsum) and torch.equal(blocks2, blocks)
    # run_and_expand: what the operator's backward calls -- the collectives, then ONE expansion pass (part 0); on CPU tensors
    # also when the split expansion is asked for (it needs CUDA streams)
    for overlap in (False, True):
        ex3 = ViewParallelExchange(overlap_expansion=overlap)
        gsum3 = torch.full((n, 12), float(rank + 1))
        blocks3 = torch.full((world, 3 * n + 8), -1.0)
        blocks3[rank] = torch.arange(3 * n + 8, dtype=torch.float32) + 1000.0 * rank
        parts = []
        ex3.run_and_expand(gsum3, blocks3, lambda part: parts.append((part, gsum3.clone(), blocks3.clone())))
        assert [p[0] for p in parts] == [0]                                   # one pass ...
        assert torch.equal(parts[0][1], gsum) and torch.equal(parts[0][2], blocks)  # ... that sees the exchanged buffers
    torch.save({"xyz": g_xyz, "feat": g_feat, "avg": g_avg, "views": shard_views(8, rank, world), "flat": flat, "other": other,
                "gsum": gsum, "blocks": blocks},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_gradient_exchange_world2_gloo(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    n = 1000
    for o in outs:
        assert torch.equal(o["xyz"], torch.full((n, 3), 3.0))  # 1 + 2
        assert torch.equal(o["feat"], torch.arange(n * 56, dtype=torch.float32).reshape(n, 56) * 3)
        assert torch.allclose(o["avg"], torch.full((4,), 0.5))
        expect = torch.arange(20, dtype=torch.float32) * 3
        expect[:6] *= 2  # `va` went through the second (per-tensor) exchange: both ranks held x3, sum = x6
        assert torch.equal(o["flat"], expect) and torch.equal(o["other"], torch.full((4,), 3.0))
        assert torch.equal(o["gsum"], torch.full((n, 12), 3.0))
        for r in range(world):
            assert torch.equal(o["blocks"][r], torch.arange(3 * n + 8, dtype=torch.float32) + 1000.0 * r)
    assert sorted(outs[0]["views"] + outs[1]["views"]) == list(range(8))


def test_exchange_is_noop_without_process_group():
    g = torch.ones(5)
    assert exchange_gradients([g]) is None
    assert torch.equal(g, torch.ones(5))


def test_bench_reference_arm_prints_one_json_line_on_rank0_only():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 prints the line, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="4")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--workload", "C1"]
    r1 = subprocess.run(cmd, env=dict(env, RANK="1", LOCAL_RANK="1"), capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    r0 = subprocess.run(cmd, env=dict(env, RANK="0", LOCAL_RANK="0"), capture_output=True, text=True, timeout=300)
    assert r0.returncode == 0, r0.stderr
    lines = [ln for ln in r0.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0

"""CPU (gloo, world_size 2): the multi-GPU plumbing - sharding plan, weight broadcast, waveform gather."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_lpt_balanced_and_deterministic():
    from cosyvoice_b200.parallel import shard_lpt
    costs = [200 + (i * 7) % 101 for i in range(32)]
    for world in (1, 2, 4, 8):
        plan = shard_lpt(costs, world)
        assert sorted(i for p in plan for i in p) == list(range(32))
        loads = [sum(costs[i] for i in p) for p in plan]
        assert max(loads) - min(loads) <= max(costs)
        assert plan == shard_lpt(costs, world)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cosyvoice_b200.parallel import broadcast_state_dicts, gather_waveforms
    shapes = [{"a.weight": (3, 4), "a.bias": (3,)}, {"b": (5,)}, {"c.w": (2, 2, 2)}]
    dev = torch.device("cpu")
    sds = None
    if rank == 0:
        g = torch.Generator().manual_seed(1)
        sds = tuple({k: torch.randn(s, generator=g) for k, s in sh.items()} for sh in shapes)
    got = broadcast_state_dicts(sds, dev, *shapes, dist)
    chk = float(sum(v.double().sum() for sd in got for v in sd.values()))
    wavs = [torch.full((1, 10 + 3 * rank + i), float(rank * 10 + i)) for i in range(2 + rank)]
    res = gather_waveforms(wavs, dist, dev)
    ok = True
    if rank == 0:
        for r in range(world):
            ok &= len(res[r]) == 2 + r
            for i, w in enumerate(res[r]):
                ok &= tuple(w.shape) == (1, 10 + 3 * r + i) and bool((w == r * 10 + i).all())
    # device-resident variant: flat buffers, exact sizes, point-to-point into rank 0's buffer
    from cosyvoice_b200.parallel import gather_flat, split_flat
    flat = torch.cat([w.reshape(-1) for w in wavs])
    for counts in (None, [2 + r for r in range(world)]):
        got2 = gather_flat(flat, [w.shape[-1] for w in wavs], dist, dev, counts=counts)
        if rank == 0:
            host, per_rank = got2
            res2 = split_flat(host, per_rank)
            for r in range(world):
                ok &= len(res2[r]) == 2 + r
                for i, w in enumerate(res2[r]):
                    ok &= tuple(w.shape) == (1, 10 + 3 * r + i) and bool((w == r * 10 + i).all())
        else:
            ok &= got2 is None
    q.put((rank, chk, ok))
    dist.destroy_process_group()


def test_broadcast_and_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert abs(res[0][1] - res[1][1]) < 1e-9          # both ranks hold identical weights after the broadcast
    assert res[0][2] and res[1][2]

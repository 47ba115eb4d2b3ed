"""Multi-GPU plumbing: the path shards by utterance - one process per GPU, a full weight replica each.

Only two collectives exist, both outside the compute kernels (no exchange step inside LM -> flow -> HiFT):
  * broadcast of the weights from rank 0 at load time,
  * gather of the finished waveforms to rank 0: lengths (one small all_gather), then every rank's flat DEVICE buffer straight
    into its slice of one device buffer on rank 0 (grouped send / recv, exact sizes, no padding), then ONE pinned D2H copy.
`torch.distributed` is the transport (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from collections import OrderedDict

import torch


def shard_lpt(costs, world):
    """Longest-processing-time assignment of utterances to ranks (cost = expected speech tokens).
    Returns a list of index lists, one per rank; deterministic, every rank computes the same plan."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0] * world
    plan = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        plan[r].append(i)
        loads[r] += costs[i]
    return [sorted(p) for p in plan]


def broadcast_state_dicts(sds, device, llm_shapes, flow_shapes, hift_shapes, dist):
    """rank 0 holds (llm, flow, hift) state dicts; every other rank receives them.  One flat fp32 buffer per stage."""
    if dist is None or dist.get_world_size() == 1:
        return sds
    out = []
    for k, shapes in enumerate((llm_shapes, flow_shapes, hift_shapes)):
        n = sum(int(torch.Size(s).numel()) for s in shapes.values())
        if dist.get_rank() == 0:
            flat = torch.cat([sds[k][key].reshape(-1).to(device=device, dtype=torch.float32) for key in shapes])
        else:
            flat = torch.empty(n, device=device, dtype=torch.float32)
        dist.broadcast(flat, src=0)
        sd, o = OrderedDict(), 0
        for key, s in shapes.items():
            m = int(torch.Size(s).numel())
            sd[key] = flat[o:o + m].view(s)
            o += m
        out.append(sd)
    return tuple(out)


def gather_waveforms(wavs, dist, device):
    """wavs: list of [1,N_i] tensors of this rank.  Returns on rank 0 a list (per rank) of lists of CPU waveforms, else None."""
    if dist is None or dist.get_world_size() == 1:
        return [wavs]
    world, rank = dist.get_world_size(), dist.get_rank()
    lens = torch.tensor([w.shape[-1] for w in wavs], dtype=torch.int64, device=device)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = len(wavs)
    dist.all_reduce(counts)
    maxn = int(counts.max())
    lens_p = torch.zeros(maxn, dtype=torch.int64, device=device)
    lens_p[:len(wavs)] = lens
    all_lens = [torch.zeros(maxn, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_lens, lens_p)
    total = max(int(l.sum()) for l in all_lens)
    flat = torch.zeros(total, dtype=torch.float32, device=device)
    if wavs:
        cat = torch.cat([w.reshape(-1).to(device) for w in wavs])
        flat[:cat.numel()] = cat
    bufs = [torch.zeros(total, dtype=torch.float32, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(flat, bufs, dst=0)
    if rank != 0:
        return None
    out = []
    for r in range(world):
        o, lst = 0, []
        for i in range(int(counts[r])):
            n = int(all_lens[r][i])
            lst.append(bufs[r][o:o + n].cpu().unsqueeze(0))
            o += n
        out.append(lst)
    return out


def gather_flat(flat, lens, dist, device, counts=None):
    """Device-resident gather.  flat: 1-D float32 tensor on `device` = this rank's waveforms back to back (the vocoder's own
    output buffer); lens: their lengths in samples.  counts: utterances per rank when every rank already knows them (the LPT
    plan is deterministic), else exchanged.  Rank 0 returns (host_flat, per_rank_lens): ONE host tensor (pinned when CUDA) holding
    rank 0's, rank 1's, ... samples back to back, and the list of length lists; other ranks return None.

    Traffic: world x maxn int64 lengths (all_gather) + exactly the samples (point-to-point into rank 0's buffer at the right
    offset) + one D2H copy.  Round 1 went D2H -> H2D -> zero-padded NCCL gather -> one .cpu() per utterance."""
    if dist is None or dist.get_world_size() == 1:
        host = torch.empty(flat.numel(), dtype=torch.float32, pin_memory=flat.is_cuda)
        host.copy_(flat, non_blocking=True)
        if flat.is_cuda:
            torch.cuda.current_stream().synchronize()
        return host, [list(lens)]
    world, rank = dist.get_world_size(), dist.get_rank()
    if counts is None:
        c = torch.zeros(world, dtype=torch.int64, device=device)
        c[rank] = len(lens)
        dist.all_reduce(c)
        counts = [int(x) for x in c.tolist()]
    maxn = max(max(counts), 1)
    lens_p = torch.zeros(maxn, dtype=torch.int64, device=device)
    if len(lens):
        lens_p[:len(lens)] = torch.tensor(list(lens), dtype=torch.int64)
    all_lens = torch.zeros(world * maxn, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_lens, lens_p)
    if rank != 0:
        if flat.numel():          # batched like the receiving side (an un-batched send is serialised with every other op of the group)
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, flat, 0)]):
                w.wait()
        return None
    al = all_lens.view(world, maxn).tolist()
    per_rank = [[int(x) for x in al[r][:counts[r]]] for r in range(world)]
    totals = [sum(p) for p in per_rank]
    big = torch.empty(sum(totals), dtype=torch.float32, device=device)
    big[:totals[0]].copy_(flat)
    ops, o = [], totals[0]
    for r in range(1, world):
        if totals[r]:
            ops.append(dist.P2POp(dist.irecv, big[o:o + totals[r]], r))
        o += totals[r]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    host = torch.empty(big.numel(), dtype=torch.float32, pin_memory=big.is_cuda)
    host.copy_(big, non_blocking=True)
    if big.is_cuda:
        torch.cuda.current_stream().synchronize()
    return host, per_rank


def split_flat(host, per_rank):
    """views [1,n] of the gathered host buffer, one list per rank"""
    out, o = [], 0
    for lens in per_rank:
        lst = []
        for n in lens:
            lst.append(host[o:o + n].unsqueeze(0))
            o += n
        out.append(lst)
    return out

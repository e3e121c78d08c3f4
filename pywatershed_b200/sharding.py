"""Sharding a domain over the GPUs of one box.

The path partitions with no data-path collective: HRU columns
are independent (cascades are off in the reference, prms_runoff.py:850-854)
and the segment graph is a forest, so every outlet tree -- with the HRUs that
drain into it -- can live on its own GPU.  `partition_basins` assigns whole
outlet trees to ranks (greedy, heaviest first), `shard_parameters` cuts the
parameter set for one rank (tosegment / hru_segment renumbered to the local
segments), and `gather_outputs` puts per-rank results back into the domain's
order with torch.distributed (gloo on CPU, NCCL on GPUs) -- the only
communication of a sharded run.  Ensembles shard by member
(`partition_members`).
"""
from __future__ import annotations

import numpy as np


def outlet_of(tosegment) -> np.ndarray:
    """Index of the outlet segment each segment drains to (pointer jumping)."""
    tos = np.asarray(tosegment, dtype=np.int64) - 1
    nxt = np.where(tos >= 0, tos, np.arange(tos.shape[0], dtype=np.int64))
    while True:  # pointer doubling
        nn = nxt[nxt]
        if np.array_equal(nn, nxt):
            return nxt
        nxt = nn


def partition_basins(tosegment, hru_segment, nranks: int, seg_weight: float = 24.0):
    """Assign outlet trees to ranks.  Returns (seg_rank [nseg], hru_rank [nhru]).

    Weight of a tree = its HRUs + seg_weight x its segments (a segment does 24
    hourly steps a day).  HRUs without a segment (hru_segment == 0) are dealt
    round-robin.  Deterministic: ties broken by outlet index."""
    tosegment = np.asarray(tosegment, dtype=np.int64)
    hru_segment = np.asarray(hru_segment, dtype=np.int64)
    nseg = tosegment.shape[0]
    out = outlet_of(tosegment) if nseg else np.zeros(0, dtype=np.int64)
    outlets = np.unique(out)
    nsegs = np.bincount(out, minlength=nseg)[outlets]
    attached = hru_segment > 0
    nhrus = np.bincount(out[hru_segment[attached] - 1], minlength=nseg)[outlets] if nseg else 0
    weight = nhrus + seg_weight * nsegs
    order = np.lexsort((outlets, -weight))
    load = np.zeros(nranks)
    rank_of_outlet = {}
    for j in order:
        r = int(np.argmin(load))
        rank_of_outlet[int(outlets[j])] = r
        load[r] += weight[j]
    seg_rank = np.array([rank_of_outlet[int(o)] for o in out], dtype=np.int64)
    hru_rank = np.empty(hru_segment.shape[0], dtype=np.int64)
    hru_rank[attached] = seg_rank[hru_segment[attached] - 1]
    loose = np.where(~attached)[0]
    hru_rank[loose] = np.arange(loose.size) % nranks
    return seg_rank, hru_rank


def shard_parameters(params: dict, seg_rank, hru_rank, rank: int):
    """Parameters of one rank's sub-domain + (hru_idx, seg_idx): the global
    indices of its HRUs / segments, ascending (so relative order is kept)."""
    nhru = np.asarray(params["hru_area"]).shape[0]
    nseg = np.asarray(params["tosegment"]).shape[0]
    hru_idx = np.where(np.asarray(hru_rank) == rank)[0]
    seg_idx = np.where(np.asarray(seg_rank) == rank)[0]
    new_seg = np.zeros(nseg + 1, dtype=np.int64)  # global segment number -> local number
    new_seg[seg_idx + 1] = np.arange(1, seg_idx.size + 1)
    out = {}
    for k, v in params.items():
        v = np.asarray(v)
        if k in ("tosegment", "tosegment_nhm"):
            out[k] = new_seg[np.asarray(params["tosegment"], dtype=np.int64)[seg_idx]] \
                if k == "tosegment" else v[seg_idx]
        elif k == "hru_segment":
            out[k] = new_seg[v.astype(np.int64)[hru_idx]]
        elif k == "snarea_curve":
            out[k] = v
        elif v.ndim >= 1 and v.shape[-1] == nhru and nhru != nseg:
            out[k] = np.ascontiguousarray(v[..., hru_idx])
        elif v.ndim >= 1 and v.shape[-1] == nseg and nhru != nseg:
            out[k] = np.ascontiguousarray(v[..., seg_idx])
        elif v.ndim >= 1 and v.shape[-1] == nhru:  # nhru == nseg: decide by name
            is_seg = k in ("mann_n", "seg_depth", "seg_length", "seg_slope", "x_coef",
                           "segment_flow_init", "segment_type", "nhm_seg", "seg_lat", "seg_width")
            out[k] = np.ascontiguousarray(v[..., seg_idx if is_seg else hru_idx])
        else:
            out[k] = v
    return out, hru_idx, seg_idx


def partition_members(nmember: int, nranks: int, rank: int) -> np.ndarray:
    """Ensemble members of one rank: contiguous blocks, sizes differ by <= 1."""
    base, extra = divmod(nmember, nranks)
    lo = rank * base + min(rank, extra)
    return np.arange(lo, lo + base + (1 if rank < extra else 0))


def gather_outputs(local: np.ndarray, idx: np.ndarray, n_total: int, group=None):
    """All-gather per-rank outputs [..., n_local] into the domain's order
    [..., n_total].  `idx` = the global unit indices of this rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
        else torch.device("cpu")
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([idx.size], dtype=torch.int64, device=dev), group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    lead = local.shape[:-1]
    pad = np.zeros(lead + (nmax,), dtype=np.float64)
    pad[..., :idx.size] = local
    ipad = np.full(nmax, -1, dtype=np.int64)
    ipad[:idx.size] = idx
    t = torch.from_numpy(pad).to(dev)
    ti = torch.from_numpy(ipad).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    iparts = [torch.empty_like(ti) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    dist.all_gather(iparts, ti, group=group)
    full = np.empty(lead + (n_total,), dtype=np.float64)
    for p, ip, n in zip(parts, iparts, sizes):
        full[..., ip[:n].cpu().numpy()] = p[..., :n].cpu().numpy()
    return full


def reduce_budget_violations(viol: np.ndarray, group=None) -> np.ndarray:
    """Sum of the per-day, per-process violation counts over ranks."""
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
        else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(viol, dtype=np.int64)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()

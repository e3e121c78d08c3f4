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

A segment graph that does NOT split (the real DRB: 434 of its 456 segments are
one outlet tree; a chained network) still leaves the HRU columns independent:
`partition_hrus` deals the HRUs to the ranks segment by segment,
`lateral_inflow` restates the HRU side of PRMSChannel (prms_channel.py:396-421)
on a rank's three volume fluxes, `exchange_lateral` is the ONE exchange step of
that variant (a reduce of seg_lateral_inflow [T, nsegment] per
block of days to the rank that owns the network), and `SplitColumnsOneRouter`
drives it: every rank computes its columns, the router rank routes the whole
network with PRMSChannel alone (pws_run_channel_device) beside its own columns.
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


# ------------------------------------------------------ HRUs split, one router
_SEG_NAMES = ("mann_n", "seg_depth", "seg_length", "seg_slope", "x_coef", "segment_flow_init",
              "segment_type", "tosegment", "tosegment_nhm", "nhm_seg", "seg_lat", "seg_width",
              "seg_cum_area", "seg_elev", "obsin_segment", "obsout_segment", "seg_humidity")


def partition_hrus(hru_segment, nranks: int, weights=None) -> np.ndarray:
    """hru_rank [nhru] for a domain whose segment graph stays on one rank: all HRUs
    of a segment go to the same rank (so each segment's lateral inflow is summed by
    ONE rank, in the reference's order: ascending HRU index, prms_channel.py:396-421),
    segments are dealt in runs of consecutive numbers holding about nhru / nranks HRUs
    (`weights` [nranks]: relative shares instead, e.g. a smaller one for the rank that
    also routes); HRUs without a segment are dealt round-robin."""
    hs = np.asarray(hru_segment, dtype=np.int64)
    att = hs > 0
    nseg = int(hs.max(initial=0))
    cnt = np.bincount(hs[att] - 1, minlength=nseg)
    before = np.cumsum(cnt) - cnt
    total = max(int(cnt.sum()), 1)
    w = np.ones(nranks) if weights is None else np.asarray(weights, dtype=np.float64)
    if w.shape != (nranks,) or (w < 0).any() or w.sum() <= 0:
        raise ValueError("partition_hrus: weights must be [nranks], non-negative, not all zero")
    edges = np.cumsum(w) / w.sum() * total            # rank r takes HRU counts [edges[r-1], edges[r])
    seg_rank = np.minimum(np.searchsorted(edges, before, side="right"), nranks - 1)
    hru_rank = np.empty(hs.shape[0], dtype=np.int64)
    hru_rank[att] = seg_rank[hs[att] - 1]
    loose = np.where(~att)[0]
    hru_rank[loose] = np.arange(loose.size) % nranks
    return hru_rank


def shard_hru_parameters(params: dict, hru_rank, rank: int):
    """The HRU side of one rank (no segment parameter: a domain for an Engine with
    channel=False) + hru_idx, the global indices of its HRUs, ascending."""
    nhru = np.asarray(params["hru_area"]).shape[0]
    hru_idx = np.where(np.asarray(hru_rank) == rank)[0]
    out = {}
    for k, v in params.items():
        if k in _SEG_NAMES or k == "hru_segment":
            continue
        v = np.asarray(v)
        if k != "snarea_curve" and v.ndim >= 1 and v.shape[-1] == nhru:
            out[k] = np.ascontiguousarray(v[..., hru_idx])
        else:
            out[k] = v
    return out, hru_idx


LATERAL_TERMS = ("sroff_vol", "ssres_flow_vol", "gwres_flow_vol")


def lateral_plan(hru_segment, hru_idx):
    """Passes of the lateral-inflow sum of one rank: [(local HRU positions, global
    segment indices)] -- pass k holds the k-th HRU (ascending HRU index) of every
    segment this rank feeds, so a pass touches a segment at most once and the passes
    in order add a segment's HRUs in the reference's order."""
    seg = np.asarray(hru_segment, dtype=np.int64)[np.asarray(hru_idx)] - 1
    att = np.where(seg >= 0)[0]
    order = np.lexsort((att, seg[att]))
    h, sg = att[order], seg[att][order]
    n = sg.shape[0]
    first = np.ones(n, dtype=bool)
    first[1:] = sg[1:] != sg[:-1]
    occ = np.arange(n) - np.maximum.accumulate(np.where(first, np.arange(n), 0))
    return [(h[occ == k], sg[occ == k]) for k in range(int(occ.max(initial=-1)) + 1)]


def lateral_inflow(sroff_vol, ssres_flow_vol, gwres_flow_vol, plan, nseg: int, out=None):
    """seg_lateral_inflow [T, nseg] (cfs) of this rank's HRUs, zero on segments it does
    not feed: prms_channel.py:396-421 on arrays [T, nhru_local] (numpy, or torch tensors
    on any device with `plan` as index tensors of that device).  Same operations in the
    same order as the fused kernel's HRU side (hru_column.cuh, "PRMSChannel, HRU side"):
    ((sroff_vol + ssres_flow_vol) + gwres_flow_vol) / 86400 per HRU, HRUs of a segment
    added in ascending order starting from 0."""
    q = (sroff_vol + ssres_flow_vol) + gwres_flow_vol
    if isinstance(q, np.ndarray):
        q = q / 86400.0
    else:
        # a tensor divisor: torch turns division by a Python scalar into a multiplication
        # by its reciprocal on CUDA, which is not the correctly rounded quotient
        q = q / q.new_full((1,), 86400.0)
    if out is None:
        out = np.zeros((q.shape[0], nseg)) if isinstance(q, np.ndarray) else q.new_zeros((q.shape[0], nseg))
    else:
        out[...] = 0.0
    for h, sg in plan:
        out[:, sg] = out[:, sg] + q[:, h]
    return out


def exchange_lateral(lat, router: int = 0, group=None):
    """The exchange step: every segment's lateral inflow was summed by exactly one rank
    and is zero elsewhere, so a SUM-reduce to the router (x + 0 is exact in any order)
    delivers the whole network's [T, nseg] rows bit for bit.  `lat`: a torch tensor,
    reduced in place on the router (NCCL on GPUs, gloo on CPU)."""
    import torch.distributed as dist

    dist.reduce(lat, dst=router, op=dist.ReduceOp.SUM, group=group)
    return lat


class SplitColumnsOneRouter:
    """One domain on N GPUs when its segment graph does not split: rank r computes the
    columns of its HRUs (Engine over `shard_hru_parameters`, no channel), the router
    rank also owns an Engine over the whole domain's segments and routes every block
    with PRMSChannel alone after the exchange.  The next block's columns are enqueued
    before a block's exchange, so columns, exchange and routing of neighbouring blocks
    overlap on the device."""

    def __init__(self, params: dict, start, rank: int, world: int, device: int, router: int = 0,
                 hru_vars=(), seg_vars=None, block_days: int = 256, group=None,
                 router_share: float = 1.0, engine_cls=None, **engine_kw):
        """`router_share`: the router's share of the HRUs relative to another rank's (it
        also routes: a smaller share evens the ranks out; at least a few HRUs stay).
        `engine_cls` (tests only): a stand-in for Engine with `device=None`, buffers on
        the CPU -- tests/test_sharding.py drives this class under gloo with the oracle."""
        import torch

        if engine_cls is None:
            from .engine import Engine
        else:
            Engine = engine_cls

        self.rank, self.world, self.router, self.group = rank, world, router, group
        self.dev = torch.device("cuda", device) if device is not None else torch.device("cpu")
        self.T = int(block_days)
        self.nseg = int(np.asarray(params["tosegment"]).shape[0])
        w = np.ones(world)
        w[router] = max(float(router_share), 0.02) if world > 1 else 1.0
        hru_rank = partition_hrus(params["hru_segment"], world, w)
        q, self.hru_idx = shard_hru_parameters(params, hru_rank, rank)
        self.cols = Engine(q, start, device=device, channel=False, **engine_kw)
        self.hru_vars = list(hru_vars)
        sel = self.hru_vars + [v for v in LATERAL_TERMS if v not in self.hru_vars]
        self.cols.select_outputs(sel, ())
        if self.cols.sel_i32:
            raise ValueError("SplitColumnsOneRouter: float64 HRU variables only")
        self._term = [self.cols.sel_f64.index(v) for v in LATERAL_TERMS]
        self.plan = [(torch.from_numpy(h).to(self.dev), torch.from_numpy(s).to(self.dev))
                     for h, s in lateral_plan(params["hru_segment"], self.hru_idx)]
        self.route = None
        if rank == router:
            self.route = Engine(params, start, device=device, **engine_kw)
            self.seg_vars = list(self.route.reg.seg_vars if seg_vars is None else seg_vars)
            self.route.select_outputs((), self.seg_vars)
        nv = len(self.cols.sel_f64)
        self.out = [torch.empty((self.T, nv, self.cols.nhru), dtype=torch.float64, device=self.dev)
                    for _ in range(2)]
        self.lat = []    # ring of exchange buffers, grown by run_device (LAT_RING_MAX blocks deep)

    LAT_RING_MAX = 16

    def reset(self):
        self.cols.reset()
        if self.route is not None:
            self.route.reset()

    def close(self):
        self.cols.close()
        if self.route is not None:
            self.route.close()

    def run_device(self, prcp, tmax, tmin, seg_out=None, hru_out=None):
        """Advance len(prcp) days: forcings = torch tensors [nd, nhru_local] on this
        rank's GPU; `seg_out` (router only): [nd, len(seg_vars), nseg] on the GPU;
        `hru_out` (optional): [nd, len(hru_vars), nhru_local].  Returns when the
        device has finished."""
        import torch

        nd, nloc = prcp.shape[0], self.cols.nhru
        blocks = [(d, min(self.T, nd - d)) for d in range(0, nd, self.T)]
        # one exchange buffer per block in flight: the router's launches queue up on its
        # routing stream while later blocks are reduced, and the host waits for the routing
        # only when a buffer comes round again
        R = min(len(blocks), self.LAT_RING_MAX)
        while len(self.lat) < R:
            self.lat.append(torch.empty((self.T, self.nseg), dtype=torch.float64, device=self.dev))

        def columns(b):
            d, n = blocks[b]
            o = d * nloc * 8
            self.cols.run_device(n, prcp.data_ptr() + o, tmax.data_ptr() + o, tmin.data_ptr() + o,
                                 self.out[b % 2].data_ptr(), sync=False)

        columns(0)
        for b, (d, n) in enumerate(blocks):
            self.cols.sync()                          # block b's columns are in self.out[b % 2]
            if b + 1 < len(blocks):
                columns(b + 1)                        # runs beside the exchange and the routing of b
            if self.route is not None and b >= R:
                self.route.sync()                     # the routing that read self.lat[b % R] is done
            o = self.out[b % 2][:n]
            lat = self.lat[b % R][:n]
            lateral_inflow(o[:, self._term[0]], o[:, self._term[1]], o[:, self._term[2]], self.plan,
                           self.nseg, out=lat)
            if hru_out is not None and self.hru_vars:
                hru_out[d:d + n] = o[:, :len(self.hru_vars)]
            if self.world > 1:
                exchange_lateral(lat, self.router, self.group)
            if self.dev.type == "cuda":
                torch.cuda.current_stream().synchronize()
            if self.route is not None:
                self.route.run_channel_device(n, lat.data_ptr(),
                                              seg_out[d:d + n].data_ptr() if seg_out is not None else 0,
                                              sync=False)
        if self.route is not None:
            self.route.sync()

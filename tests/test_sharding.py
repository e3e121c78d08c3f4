"""Multi-GPU host logic on CPU: world_size-2 gloo run of the basin sharding
(pywatershed_b200/sharding.py).  Each rank takes its outlet trees, computes
its shard (the oracle stands in for the GPU here -- this is a test of the
partition / renumbering / gather arithmetic, not of the kernels), and the
gathered result must equal the unsharded run bit for bit: the path shards with
no data-path collective."""
import os
import socket

import numpy as np
import pytest

import scenarios


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_partition_keeps_outlet_trees_whole(drb):
    from pywatershed_b200 import sharding

    p, f, start = drb
    pt, _ = scenarios.tile(p, {k: v[:2] for k, v in f.items()}, 5)
    for nranks in (1, 2, 3, 8):
        seg_rank, hru_rank = sharding.partition_basins(pt["tosegment"], pt["hru_segment"], nranks)
        tos = pt["tosegment"]
        down = tos > 0
        assert (seg_rank[down] == seg_rank[tos[down] - 1]).all()  # no edge crosses ranks
        att = pt["hru_segment"] > 0
        assert (hru_rank[att] == seg_rank[pt["hru_segment"][att] - 1]).all()
        assert set(np.unique(seg_rank)) <= set(range(nranks))
        # balanced to within one big tree (434 of the 456 segments of a tile are one tree)
        load = np.bincount(seg_rank, minlength=nranks)
        if nranks <= 5:
            assert load.max() - load.min() <= 460
        total = 0
        for r in range(nranks):
            q, hi, si = sharding.shard_parameters(pt, seg_rank, hru_rank, r)
            total += hi.size
            assert q["tosegment"].max(initial=0) <= si.size and q["hru_segment"].max(initial=0) <= si.size
            assert q["tmax_cbh_adj"].shape == (12, hi.size) and q["seg_length"].shape == (si.size,)
        assert total == pt["hru_area"].shape[0]
    np.testing.assert_array_equal(
        np.concatenate([sharding.partition_members(10, 4, r) for r in range(4)]), np.arange(10))


def _worker(rank, world, port, nd, ret):
    import sys
    import pathlib as pl

    root = pl.Path(__file__).resolve().parent.parent
    for q in (root, root / "tests", root / "oracle"):
        if str(q) not in sys.path:
            sys.path.insert(0, str(q))
    import torch.distributed as dist

    import prms_oracle as po
    import scenarios as sc
    from pywatershed_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, f, start = sc.load_drb()
        pt, ft = sc.tile(p, {k: v[:nd] for k, v in f.items()}, 3)
        nhru, nseg = pt["hru_area"].shape[0], pt["tosegment"].shape[0]
        seg_rank, hru_rank = sharding.partition_basins(pt["tosegment"], pt["hru_segment"], world)
        q, hi, si = sharding.shard_parameters(pt, seg_rank, hru_rank, rank)
        o = po.Oracle(q, start=start)
        names_h, names_s = ["pkwater_equiv", "sroff_vol", "soil_moist"], ["seg_outflow"]
        res, viol = o.run(ft["prcp"][:, hi], ft["tmax"][:, hi], ft["tmin"][:, hi], names_h, names_s)
        o.close()
        full = {k: sharding.gather_outputs(res[k], hi, nhru) for k in names_h}
        full["seg_outflow"] = sharding.gather_outputs(res["seg_outflow"], si, nseg)
        tot_viol = sharding.reduce_budget_violations(viol)
        if rank == 0:
            ref_o = po.Oracle(pt, start=start)
            ref, rviol = ref_o.run(ft["prcp"], ft["tmax"], ft["tmin"], names_h, names_s)
            ok = all(np.array_equal(full[k], ref[k]) for k in names_h + names_s)
            ok = ok and np.array_equal(tot_viol, rviol)
            ret["ok"] = bool(ok)
            ret["nlocal"] = int(hi.size)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_run_equals_unsharded():
    import torch.multiprocessing as mp

    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 25, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
    assert 0 < ret["nlocal"] < 3 * 765


# ---------------------------------------------------------------- HRUs split, one router
def test_partition_hrus_keeps_a_segments_hrus_together(drb):
    from pywatershed_b200 import sharding

    p, f, start = drb
    pt, _ = scenarios.tile(p, {k: v[:2] for k, v in f.items()}, 3)
    hs = pt["hru_segment"]
    for nranks in (1, 2, 3, 8):
        hru_rank = sharding.partition_hrus(hs, nranks)
        att = hs > 0
        for sgm in np.unique(hs[att]):
            assert np.unique(hru_rank[hs == sgm]).size == 1
        load = np.bincount(hru_rank, minlength=nranks)
        assert load.sum() == hs.size and load.max() - load.min() <= 8 + hs.size // 100
        seen = np.zeros(hs.size, dtype=int)
        for r in range(nranks):
            q, hi = sharding.shard_hru_parameters(pt, hru_rank, r)
            seen[hi] += 1
            assert "tosegment" not in q and "hru_segment" not in q and "seg_length" not in q
            assert q["tmax_cbh_adj"].shape == (12, hi.size) and q["snarea_curve"].shape == pt["snarea_curve"].shape
            plan = sharding.lateral_plan(hs, hi)
            touched = np.concatenate([sg for _, sg in plan]) if plan else np.zeros(0, int)
            assert np.array_equal(np.sort(np.concatenate([h for h, _ in plan])), np.where(hs[hi] > 0)[0])
            for _, sg in plan:
                assert np.unique(sg).size == sg.size          # a pass touches a segment once
            assert np.array_equal(np.unique(touched), np.unique(hs[hi][hs[hi] > 0] - 1))
        assert (seen == 1).all()
    # a lighter share for the rank that also routes
    hru_rank = sharding.partition_hrus(hs, 4, weights=[0.25, 1, 1, 1])
    load = np.bincount(hru_rank, minlength=4)
    assert load.sum() == hs.size and abs(load[0] - hs.size / 13) <= 8 and load[1:].min() > 3 * load[0]
    for sgm in np.unique(hs[hs > 0]):
        assert np.unique(hru_rank[hs == sgm]).size == 1


def test_lateral_inflow_is_the_channels_hru_side(drb):
    """lateral_inflow on the oracle's three fluxes = the oracle's seg_lateral_inflow, bit for bit."""
    import prms_oracle as po
    from pywatershed_b200 import sharding

    p, f, start = drb
    nd = 20
    o = po.Oracle(p, start=start)
    res, _ = o.run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd], sharding.LATERAL_TERMS, ["seg_lateral_inflow"])
    o.close()
    plan = sharding.lateral_plan(p["hru_segment"], np.arange(p["hru_area"].shape[0]))
    lat = sharding.lateral_inflow(*(res[k] for k in sharding.LATERAL_TERMS), plan, p["tosegment"].shape[0])
    np.testing.assert_array_equal(lat, res["seg_lateral_inflow"])


def _worker_split(rank, world, port, nd, ret):
    import sys
    import pathlib as pl

    root = pl.Path(__file__).resolve().parent.parent
    for q in (root, root / "tests", root / "oracle"):
        if str(q) not in sys.path:
            sys.path.insert(0, str(q))
    import torch
    import torch.distributed as dist

    import bench
    import prms_oracle as po
    import scenarios as sc
    from pywatershed_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, f, start = sc.load_drb()
        ntile = 3
        pt, ft = sc.tile(p, {k: v[:nd] for k, v in f.items()}, ntile)
        pt, _ = bench.chain_tiles(p, pt, ntile)           # ONE tree: basin sharding cannot split it
        nhru, nseg = pt["hru_area"].shape[0], pt["tosegment"].shape[0]
        hru_rank = sharding.partition_hrus(pt["hru_segment"], world)
        q, hi = sharding.shard_hru_parameters(pt, hru_rank, rank)
        o = po.Oracle(q, start=start)                      # the oracle stands in for the GPU
        names_h = ["pkwater_equiv", "soil_moist"] + list(sharding.LATERAL_TERMS)
        res, _ = o.run(ft["prcp"][:, hi], ft["tmax"][:, hi], ft["tmin"][:, hi], names_h, ())
        o.close()
        plan = sharding.lateral_plan(pt["hru_segment"], hi)
        lat = sharding.lateral_inflow(*(res[k] for k in sharding.LATERAL_TERMS), plan, nseg)
        lat_t = torch.from_numpy(lat)
        sharding.exchange_lateral(lat_t, 0)               # the one exchange step
        full = {k: sharding.gather_outputs(res[k], hi, nhru) for k in names_h}
        if rank == 0:
            routed = po.Oracle(pt, start=start).run_channel(lat_t.numpy())
            ref_o = po.Oracle(pt, start=start)
            ref, _ = ref_o.run(ft["prcp"], ft["tmax"], ft["tmin"], names_h, po.SEG_VARS)
            ok = all(np.array_equal(full[k], ref[k]) for k in names_h)
            ok = ok and all(np.array_equal(routed[k], ref[k]) for k in po.SEG_VARS)
            ret["ok"] = bool(ok)
            ret["nlocal"] = int(hi.size)
            ret["flow"] = float(ref["seg_outflow"][-1].sum())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_split_columns_one_router_equals_unsharded():
    """A network that is ONE tree (chained tiles): HRU columns split over two ranks,
    lateral inflow reduced to rank 0 (gloo), which routes the whole network -- all
    five segment variables equal the unsharded run bit for bit."""
    import torch.multiprocessing as mp

    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_split, args=(2, port, 25, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
    assert 0 < ret["nlocal"] < 3 * 765 and ret["flow"] > 0.0


# ---------------------------------------------------------------- the driver class itself, under gloo
class _OracleEngine:
    """The slice of Engine that sharding.SplitColumnsOneRouter uses, on the CPU oracle and
    CPU buffers (raw pointers, as the library takes them): a stand-in for tests only."""

    def __init__(self, params, start, device=None, channel=True, **kw):
        import types

        import prms_oracle as po

        self._po, self._params, self._start = po, params, start
        self.o = po.Oracle(params, start=start)
        self.nhru, self.nseg = self.o.nhru, (self.o.nseg if channel else 0)
        self.reg = types.SimpleNamespace(seg_vars=list(po.SEG_VARS))
        self.sel_f64, self.sel_i32, self.sel_seg = [], [], []
        self.day, self._launches = 0, 0

    @staticmethod
    def _arr(ptr, shape):
        import ctypes as C

        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=shape)

    def select_outputs(self, hru_vars=(), seg_vars=()):
        self.sel_f64, self.sel_seg = list(hru_vars), list(seg_vars)

    def run_device(self, n, d_prcp, d_tmax, d_tmin, d_out_f64=0, d_out_i32=0, d_seg_out=0, viol=None,
                   sync=True):
        f = [self._arr(q, (n, self.nhru)) for q in (d_prcp, d_tmax, d_tmin)]
        res, _ = self.o.run(*f, self.sel_f64, (), budgets=False)
        out = self._arr(d_out_f64, (n, len(self.sel_f64), self.nhru))
        for k, v in enumerate(self.sel_f64):
            out[:, k] = res[v]
        self.day += n
        self._launches += 1

    def run_channel_device(self, n, d_lat, d_out, sync=True):
        res = self.o.run_channel(self._arr(d_lat, (n, self.nseg)), self.sel_seg)
        if d_out:
            out = self._arr(d_out, (n, len(self.sel_seg), self.nseg))
            for k, v in enumerate(self.sel_seg):
                out[:, k] = res[v]
        self._launches += 1

    def sync(self):
        pass

    def reset(self):
        self.o.close()
        self.o = self._po.Oracle(self._params, start=self._start)
        self.day = 0

    def close(self):
        self.o.close()

    def launch_count(self):
        return self._launches


def _worker_driver(rank, world, port, nd, ret):
    import sys
    import pathlib as pl

    root = pl.Path(__file__).resolve().parent.parent
    for q in (root, root / "tests", root / "oracle"):
        if str(q) not in sys.path:
            sys.path.insert(0, str(q))
    import torch
    import torch.distributed as dist

    import bench
    import prms_oracle as po
    import scenarios as sc
    from pywatershed_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, f, start = sc.load_drb()
        ntile = 3
        pt, ft = sc.tile(p, {k: v[:nd] for k, v in f.items()}, ntile)
        pt, _ = bench.chain_tiles(p, pt, ntile)
        nhru, nseg = pt["hru_area"].shape[0], pt["tosegment"].shape[0]
        hv = ["pkwater_equiv", "sroff_vol"]
        run = sharding.SplitColumnsOneRouter(pt, start, rank, world, None, hru_vars=hv, block_days=4,
                                             router_share=0.5, engine_cls=_OracleEngine)
        run.LAT_RING_MAX = 3                               # 7 blocks: the exchange buffers come round twice
        hi = run.hru_idx
        forc = [torch.from_numpy(np.ascontiguousarray(ft[k][:, hi])) for k in ("prcp", "tmax", "tmin")]
        seg = torch.zeros((nd, 5, nseg), dtype=torch.float64) if run.route is not None else None
        hru = torch.zeros((nd, len(hv), hi.size), dtype=torch.float64)
        for _ in range(2):                                 # a second pass after reset() gives the same
            run.reset()
            run.run_device(*forc, seg_out=seg, hru_out=hru)
        full = {v: sharding.gather_outputs(hru[:, k].numpy(), hi, nhru) for k, v in enumerate(hv)}
        sizes = sharding.gather_outputs(np.full((1, hi.size), float(rank)), hi, nhru)
        if rank == 0:
            ref, _ = po.Oracle(pt, start=start).run(ft["prcp"], ft["tmax"], ft["tmin"], hv, po.SEG_VARS)
            ok = all(np.array_equal(full[v], ref[v]) for v in hv)
            ok = ok and all(np.array_equal(seg[:, k].numpy(), ref[v]) for k, v in enumerate(run.seg_vars))
            ret["ok"] = bool(ok)
            ret["share0"] = float((sizes == 0).mean())
            ret["flow"] = float(ref["seg_outflow"][-1].sum())
        run.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_split_columns_one_router_driver_under_gloo():
    """sharding.SplitColumnsOneRouter itself (blocks, double-buffered results, the ring of
    exchange buffers, the router's lighter share) on two gloo ranks with the oracle standing
    in for the Engine: HRU and segment results equal the unsharded oracle run bit for bit."""
    import torch.multiprocessing as mp

    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_driver, args=(2, port, 26, ret), nprocs=2, join=True)
    assert ret.get("ok") is True and ret["flow"] > 0.0
    assert 0.25 < ret["share0"] < 0.42                    # a third of the HRUs for the router

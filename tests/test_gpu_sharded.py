"""Basin sharding on real GPUs: two ranks (NCCL), each with its outlet trees of
ONE tiled domain on its own B200 through the Engine, results gathered with
pywatershed_b200.sharding -- must equal the unsharded GPU run bit for bit (the
path shards with no data-path collective).  The ensemble split
by member is checked the same way.  Needs two GPUs (skipped otherwise; the
host-side arithmetic is covered on CPU by tests/test_sharding.py under gloo).
Run: gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nd, ret):
    import pathlib as pl
    import sys

    root = pl.Path(__file__).resolve().parent.parent
    for q in (root, root / "tests"):
        if str(q) not in sys.path:
            sys.path.insert(0, str(q))
    import torch
    import torch.distributed as dist

    import bench
    import scenarios as sc
    from pywatershed_b200 import Engine, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        p, f, start = sc.load_drb()
        hv, sv = ["pkwater_equiv", "soil_moist", "sroff_vol", "hru_actet"], ["seg_outflow", "seg_stor_change"]
        ok = True
        # ---- (1) one domain, split by outlet tree
        pt, ft = sc.tile(p, {k: v[:nd] for k, v in f.items()}, 5)
        nhru, nseg = pt["hru_area"].shape[0], pt["tosegment"].shape[0]
        seg_rank, hru_rank = sharding.partition_basins(pt["tosegment"], pt["hru_segment"], world)
        q, hi, si = sharding.shard_parameters(pt, seg_rank, hru_rank, rank)
        e = Engine(q, start, device=rank)
        e.select_outputs(hv, sv)
        e.set_block_days(8)
        res, viol = e.run_host(np.ascontiguousarray(ft["prcp"][:, hi]), np.ascontiguousarray(ft["tmax"][:, hi]),
                               np.ascontiguousarray(ft["tmin"][:, hi]))
        e.close()
        full = {k: sharding.gather_outputs(np.asarray(res[k]), hi, nhru) for k in hv}
        full.update({k: sharding.gather_outputs(np.asarray(res[k]), si, nseg) for k in sv})
        tot_viol = sharding.reduce_budget_violations(viol)
        # ---- (2) an ensemble, split by member, shared forcings
        nm = 6
        mine = sharding.partition_members(nm, world, rank)
        pe = bench.ensemble_parameters(p, mine)
        fb = {k: np.ascontiguousarray(v[:nd]) for k, v in f.items()}
        e = Engine(pe, start, device=rank, forcing_period=765)
        e.select_outputs(hv, sv)
        e.set_block_days(8)
        eres, eviol = e.run_host(fb["prcp"], fb["tmax"], fb["tmin"])
        e.close()
        hidx = np.concatenate([np.arange(765) + m * 765 for m in mine])
        sidx = np.concatenate([np.arange(456) + m * 456 for m in mine])
        efull = {k: sharding.gather_outputs(np.asarray(eres[k]), hidx, nm * 765) for k in hv}
        efull.update({k: sharding.gather_outputs(np.asarray(eres[k]), sidx, nm * 456) for k in sv})
        if rank == 0:
            one = Engine(pt, start, device=0)
            one.select_outputs(hv, sv)
            ref, rviol = one.run_host(ft["prcp"], ft["tmax"], ft["tmin"])
            one.close()
            ok = ok and all(np.array_equal(full[k], np.asarray(ref[k])) for k in hv + sv)
            ok = ok and np.array_equal(tot_viol, rviol)
            one = Engine(bench.ensemble_parameters(p, np.arange(nm)), start, device=0, forcing_period=765)
            one.select_outputs(hv, sv)
            eref, _ = one.run_host(fb["prcp"], fb["tmax"], fb["tmin"])
            one.close()
            ok = ok and all(np.array_equal(efull[k], np.asarray(eref[k])) for k in hv + sv)
            ret["ok"] = bool(ok)
            ret["nlocal"] = int(hi.size)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_sharded_run_equals_unsharded():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 30, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
    assert 0 < ret["nlocal"] < 5 * 765


# ---------------------------------------------------------------- HRUs split, one router
def _split_case(nd, ntile=4):
    import bench
    import scenarios as sc

    p, f, start = sc.load_drb()
    pt, ft = sc.tile(p, {k: v[:nd] for k, v in f.items()}, ntile)
    pt, _ = bench.chain_tiles(p, pt, ntile)               # ONE outlet tree: basin sharding cannot split it
    return pt, ft, start


def _run_split(pt, ft, start, rank, world, device, block_days, ring=None):
    import torch

    from pywatershed_b200 import sharding

    hv = ["pkwater_equiv", "soil_moist", "sroff_vol"]
    run = sharding.SplitColumnsOneRouter(pt, start, rank, world, device, hru_vars=hv, block_days=block_days)
    if ring is not None:
        run.LAT_RING_MAX = ring                           # exchange buffers come round again
    dev = torch.device("cuda", device)
    hi = run.hru_idx
    forc = [torch.from_numpy(np.ascontiguousarray(ft[k][:, hi])).to(dev) for k in ("prcp", "tmax", "tmin")]
    nd = forc[0].shape[0]
    seg = torch.zeros((nd, 5, run.nseg), dtype=torch.float64, device=dev) if run.route is not None else None
    hru = torch.zeros((nd, len(hv), hi.size), dtype=torch.float64, device=dev)
    run.run_device(*forc, seg_out=seg, hru_out=hru)
    sv = list(run.seg_vars) if run.route is not None else []
    run.close()
    return hv, sv, hi, hru.cpu().numpy(), (seg.cpu().numpy() if seg is not None else None)


def _reference(pt, ft, start, hv, sv):
    from pywatershed_b200 import Engine

    one = Engine(pt, start, device=0)
    one.select_outputs(hv, sv)
    ref, _ = one.run_host(ft["prcp"], ft["tmax"], ft["tmin"])
    one.close()
    return ref


def test_split_columns_one_router_on_one_gpu_equals_the_chain():
    """world = 1: columns through an Engine without a channel, lateral inflow restated in
    torch (sharding.lateral_inflow), PRMSChannel alone on a second handle, blocks of 7
    days pipelined -- every segment variable equals the fused chain bit for bit."""
    pt, ft, start = _split_case(30)
    ref = None
    for ring in (None, 2):
        hv, sv, hi, hru, seg = _run_split(pt, ft, start, 0, 1, 0, 7, ring)
        ref = ref or _reference(pt, ft, start, hv, sv)
        assert float(seg[-1, sv.index("seg_outflow")].sum()) > 0.0
        for k, v in enumerate(hv):
            np.testing.assert_array_equal(hru[:, k], np.asarray(ref[v]), err_msg=v)
        for k, v in enumerate(sv):
            np.testing.assert_array_equal(seg[:, k], np.asarray(ref[v]), err_msg=v)


def _worker_split(rank, world, port, nd, ret):
    import pathlib as pl
    import sys

    root = pl.Path(__file__).resolve().parent.parent
    for q in (root, root / "tests"):
        if str(q) not in sys.path:
            sys.path.insert(0, str(q))
    import torch
    import torch.distributed as dist

    from pywatershed_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pt, ft, start = _split_case(nd)
        hv, sv, hi, hru, seg = _run_split(pt, ft, start, rank, world, rank, 8)
        nhru = pt["hru_area"].shape[0]
        full = {v: sharding.gather_outputs(hru[:, k], hi, nhru) for k, v in enumerate(hv)}
        if rank == 0:
            ref = _reference(pt, ft, start, hv, sv)
            ok = all(np.array_equal(full[v], np.asarray(ref[v])) for v in hv)
            ok = ok and all(np.array_equal(seg[:, k], np.asarray(ref[v])) for k, v in enumerate(sv))
            ret["ok"] = bool(ok)
            ret["nlocal"] = int(hi.size)
            ret["flow"] = float(seg[-1, sv.index("seg_outflow")].sum())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_split_columns_one_router_equals_unsharded():
    """The exchange-step variant on two GPUs (NCCL reduce of seg_lateral_inflow to the
    router): one chained tree of four tiles, equal to the unsharded chain bit for bit."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_split, args=(2, port, 30, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
    assert 0 < ret["nlocal"] < 4 * 765 and ret["flow"] > 0.0

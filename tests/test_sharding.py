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

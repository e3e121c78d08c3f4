"""Golden vectors for FlowGraph (base/flow_graph.py) from the UNMODIFIED
reference: the drb_2yr PRMSChannel network as PRMSChannelFlowNodes (numpy
calc_method) with a PassThroughFlowNode and an ObsInFlowNode inserted in series
(hydrology/prms_channel_flow_graph.py:_build_flow_graph_inputs, the helper
behind prms_channel_flow_graph_postprocess / _to_model_dict), forced by lateral
inflows [ndays, nnodes] and, for the ObsIn node, a series of observed flows with
stretches of negative values (= pass the inflow through).  Build container
only.  Writes tests/golden/drb_flowgraph_golden.npz.

    python oracle/gen_golden_flowgraph.py
"""
import pathlib as pl
import sys
import warnings

import numpy as np

HERE = pl.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent / "tests"))
import gen_golden as gg  # noqa: E402
import prms_oracle as po  # noqa: E402
import ref_harness as rh  # noqa: E402
import scenarios  # noqa: E402

# where the new nodes go: above these nhm_seg ids (a mid-basin reach and the reach below it)
PASS_ABOVE_NHM_SEG = 1829   # the reference's own docstring example (flow_graph.py)
OBSIN_ABOVE_NHM_SEG = 1766
ND = 731


def graph_inputs():
    """Lateral inflows [ND, nseg + 2] (float32-representable, so the fixture can
    store them exactly in half the space) and the ObsIn node's observations."""
    p, f, start = scenarios.load_drb()
    o = po.Oracle(p, start=start)
    res, _ = o.run(f["prcp"][:ND], f["tmax"][:ND], f["tmin"][:ND], (), ("seg_lateral_inflow", "seg_outflow"))
    lat = res["seg_lateral_inflow"].astype(np.float32).astype(np.float64)
    nseg = lat.shape[1]
    inflows = np.zeros((ND, nseg + 2))
    inflows[:, :nseg] = lat
    inflows[:, nseg] = 1.5      # a lateral inflow at the pass-through node
    inflows[:, nseg + 1] = 0.25  # ... and at the ObsIn node
    return p, inflows, res["seg_outflow"]


def main():
    import pandas as pd

    pws = rh.import_reference()
    control, params, _, _ = rh.load_drb(pws)
    control.options["budget_type"] = "error"
    from pywatershed.hydrology import prms_channel_flow_graph as pcfg

    p, inflows, seg_outflow = graph_inputs()
    nseg = inflows.shape[1] - 2
    nhm_seg = np.asarray(params.parameters["nhm_seg"])
    i_obs = int(np.where(nhm_seg == OBSIN_ABOVE_NHM_SEG)[0][0])
    # observed flows: 0.8 x the routed flow of the reach, rounded to 3 decimals; negative
    # (missing) in two stretches and on scattered days
    rng = np.random.default_rng(20260101)
    obs = np.round(0.8 * seg_outflow[:, i_obs], 3)
    obs[100:160] = -999.0
    obs[400:403] = -1.0
    obs[rng.integers(0, ND, 40)] = -999.0
    times = control.start_time + np.arange(ND) * control.time_step
    obs_df = pd.DataFrame({"gage_a": obs}, index=pd.DatetimeIndex(times))
    obs_params = pws.Parameters(dims={"npoi": 1}, coords={}, data_vars={"poi_gage_id": np.array(["gage_a"])},
                                metadata={"poi_gage_id": {"dims": ["npoi"]}}, validate=False)
    new_makers = {"pass": pws.PassThroughFlowNodeMaker(), "obsin": pws.ObsInFlowNodeMaker(obs_params, obs_df)}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        params_fg, makers = pcfg._build_flow_graph_inputs(
            params, params, new_makers, ["pass", "obsin"], [0, 0], [9001, 9002],
            [PASS_ABOVE_NHM_SEG, OBSIN_ABOVE_NHM_SEG])
        # the numpy node calculation is the parity target (the default is numba with fastmath)
        makers["prms_channel"] = pws.PRMSChannelFlowNodeMaker(params, params, calc_method="numpy")
        adapter = rh.make_mem_adapter(pws, "inflows", inflows, times, control)
        fg = pws.FlowGraph(control=control, discretization=None, parameters=params_fg, inflows=adapter,
                           node_maker_dict=makers, budget_type="error")
    names = list(fg.get_variables())
    assert names == po.FLOW_GRAPH_VARS, names
    nn = inflows.shape[1]
    out = {k: np.empty((ND, nn)) for k in names}
    zero_sum = np.empty(ND, dtype=bool)
    for t in range(ND):
        control.advance()
        fg.advance()
        fg.calculate(1.0)
        for k in names:
            out[k][t] = fg[k]
        zero_sum[t] = fg.budget._zero_sum
    assert zero_sum.all()
    to_graph_index = np.asarray(params_fg.parameters["to_graph_index"])
    kind = np.zeros(nn, dtype=np.int32)
    kind[nseg], kind[nseg + 1] = 1, 2
    obs_full = np.zeros((ND, nn))
    obs_full[:, nseg + 1] = obs
    # sample: the new nodes, the reaches around them, the outlets, and a spread of others
    around = set([nseg, nseg + 1])
    for j in (nseg, nseg + 1):
        around.add(int(to_graph_index[j]))
        around.update(np.where(to_graph_index == j)[0].tolist())
        k = int(to_graph_index[j])
        for _ in range(6):  # a few reaches downstream
            if k < 0:
                break
            around.add(k)
            k = int(to_graph_index[k])
    around.update(np.where(to_graph_index < 0)[0].tolist())
    around.update(gg.seg_sample(np.asarray(params.parameters["tosegment"])).tolist())
    sample = np.array(sorted(a for a in around if a >= 0), dtype=np.int64)
    z = {"node_sample": sample, "to_graph_index": to_graph_index.astype(np.int64), "kind": kind,
         "inflows_f32": inflows.astype(np.float32), "obs": obs, "obs_node": np.int64(nseg + 1),
         "node_order": np.asarray(fg._node_order, dtype=np.int64)}
    assert (z["inflows_f32"].astype(np.float64) == inflows).all()
    for k in names:
        z["ts_" + k] = out[k][:, sample]
        z["sum_" + k] = np.nansum(out[k], axis=1)
    z["last_" + "node_outflows"] = out["node_outflows"][-1]
    # the reference's own Muskingum coefficients (numpy's vectorised power is not libm's
    # pow in the last bit for some reaches): with them the restatements match bit for bit
    mk = makers["prms_channel"]
    pad = lambda a, fill: np.concatenate([np.asarray(a), np.full(nn - nseg, fill, dtype=np.asarray(a).dtype)])  # noqa: E731
    coef = dict(tsi=pad(mk._tsi, 1).astype(np.int32), ts=pad(mk._ts, 1.0), c0=pad(mk._c0, 0.0),
                c1=pad(mk._c1, 0.0), c2=pad(mk._c2, 0.0))
    for k, v in coef.items():
        z["coef_" + k] = v
    np.savez_compressed(gg.GOLDEN / "drb_flowgraph_golden.npz", **z)
    print("flow graph:", nn, "nodes,", len(sample), "sampled; last-day sum(node_outflows) =",
          repr(float(out["node_outflows"][-1].sum())), "; sink/source days:",
          int((out["node_sink_source"][:, nseg + 1] != 0).sum()))
    # the oracle's restatement against the reference, here and now
    o = po.Oracle(scenarios.flow_graph_params(p, to_graph_index), start=scenarios.load_drb()[2])
    got = o.run_flow_graph(inflows, kind, obs_full)
    o2 = po.Oracle(scenarios.flow_graph_params(p, to_graph_index), start=scenarios.load_drb()[2])
    o2.set_channel_coefficients(**coef)
    got2 = o2.run_flow_graph(inflows, kind, obs_full)
    for k in names:
        a, b = got2[k], out[k]
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        assert same.all(), (k, int((~same).sum()))
    print("  oracle with the reference's Muskingum coefficients: every value of every variable bit-identical")
    for k in names:
        a, b = got[k], out[k]
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        print(f"  oracle vs reference {k}: {int((~same).sum())} of {a.size} values differ in bits; "
              f"max abs {np.nanmax(np.abs(a - b)):.3g}")


if __name__ == "__main__":
    main()

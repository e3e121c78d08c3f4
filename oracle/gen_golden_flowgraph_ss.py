"""Golden vectors for FlowGraph with SourceSinkFlowNodes
(hydrology/source_sink_flow_node.py) from the UNMODIFIED reference: the drb_2yr
PRMSChannel network as PRMSChannelFlowNodes (numpy calc_method) with two
SourceSinkFlowNodes inserted -- a SINK above one reach (requests that exceed
what the minimum flow allows on many days, so every branch of
calculate_subtimestep is taken: full sink, sink cut back to the minimum flow,
no sink below the minimum flow, and positive requests = sources) and a SOURCE
above another -- plus a PassThroughFlowNode, over 240 days.  Build container
only.  Writes tests/golden/drb_flowgraph_ss_golden.npz.

    python oracle/gen_golden_flowgraph_ss.py
"""
import pathlib as pl
import sys
import warnings

import numpy as np

HERE = pl.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent / "tests"))
import gen_golden as gg  # noqa: E402
import gen_golden_flowgraph as gf  # noqa: E402
import prms_oracle as po  # noqa: E402
import ref_harness as rh  # noqa: E402
import scenarios  # noqa: E402

ND = 240
SINK_ABOVE_NHM_SEG = 1766
SOURCE_ABOVE_NHM_SEG = 1829
PASS_ABOVE = -1  # filled in main(): a headwater reach


def main():
    import pandas as pd

    pws = rh.import_reference()
    control, params, _, _ = rh.load_drb(pws)
    control.options["budget_type"] = "error"
    from pywatershed.hydrology import prms_channel_flow_graph as pcfg

    p, inflows_full, seg_outflow = gf.graph_inputs()
    nseg = inflows_full.shape[1] - 2
    nhm_seg = np.asarray(params.parameters["nhm_seg"])
    i_sink = int(np.where(nhm_seg == SINK_ABOVE_NHM_SEG)[0][0])
    pass_above = int(nhm_seg[np.where(np.bincount(np.asarray(params.parameters["tosegment"])[
        np.asarray(params.parameters["tosegment"]) > 0] - 1, minlength=nseg) == 0)[0][7]])
    nn = nseg + 3
    inflows = np.zeros((ND, nn))
    inflows[:, :nseg] = inflows_full[:ND, :nseg]
    inflows[:, nseg] = 0.125      # lateral inflow at the sink node
    inflows[:, nseg + 1] = 0.0    # ... the source node
    inflows[:, nseg + 2] = 2.0    # ... the pass-through node
    # requests: the sink asks for 60 % of the reach's routed flow (rounded), the
    # minimum flow is that reach's median flow, so on low-flow days nothing or only a
    # part is taken; a few days ask for a source instead
    rng = np.random.default_rng(20260102)
    q = seg_outflow[:ND, i_sink]
    sink_req = -np.round(0.6 * q, 3)
    sink_req[rng.integers(0, ND, 25)] = 3.5
    sink_req[50:60] = 0.0
    source_req = np.round(5.0 + 2.0 * np.sin(np.arange(ND) / 9.0), 3)
    source_req[120:130] = -1.0e3  # a sink far beyond the flow of the reach: cut back to flow_min = 0
    flow_min = np.array([float(np.round(np.median(q), 3)), 0.0])
    times = control.start_time + np.arange(ND) * control.time_step
    ss_df = pd.DataFrame({"sink_a": sink_req, "source_b": source_req}, index=pd.DatetimeIndex(times))
    ss_params = pws.Parameters(dims={"nss": 2}, coords={}, data_vars={"flow_min": flow_min},
                               metadata={"flow_min": {"dims": ["nss"]}}, validate=False)
    from pywatershed.hydrology.source_sink_flow_node import SourceSinkFlowNodeMaker

    new_makers = {"ss": SourceSinkFlowNodeMaker(ss_params, ss_df), "pass": pws.PassThroughFlowNodeMaker()}
    control.edit_end_time(times[-1]) if hasattr(control, "edit_end_time") else None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        params_fg, makers = pcfg._build_flow_graph_inputs(
            params, params, new_makers, ["ss", "ss", "pass"], [0, 1, 0], [9001, 9002, 9003],
            [SINK_ABOVE_NHM_SEG, SOURCE_ABOVE_NHM_SEG, pass_above])
        makers["prms_channel"] = pws.PRMSChannelFlowNodeMaker(params, params, calc_method="numpy")
        full_times = control.start_time + np.arange(control.n_times) * control.time_step
        pad = np.zeros((control.n_times, nn))
        pad[:ND] = inflows
        adapter = rh.make_mem_adapter(pws, "inflows", pad, full_times, control)
        fg = pws.FlowGraph(control=control, discretization=None, parameters=params_fg, inflows=adapter,
                           node_maker_dict=makers, budget_type="error")
    names = list(fg.get_variables())
    out = {k: np.empty((ND, nn)) for k in names}
    for t in range(ND):
        control.advance()
        fg.advance()
        fg.calculate(1.0)
        for k in names:
            out[k][t] = fg[k]
        assert fg.budget._zero_sum
    to_graph_index = np.asarray(params_fg.parameters["to_graph_index"])
    kind = np.zeros(nn, dtype=np.int32)
    kind[nseg], kind[nseg + 1], kind[nseg + 2] = 3, 3, 1
    req = np.zeros((ND, nn))
    req[:, nseg], req[:, nseg + 1] = sink_req, source_req
    fmin_full = np.zeros(nn)
    fmin_full[nseg], fmin_full[nseg + 1] = flow_min
    around = set(range(nseg, nn))
    for j in range(nseg, nn):
        around.update(np.where(to_graph_index == j)[0].tolist())
        k = int(to_graph_index[j])
        for _ in range(8):
            if k < 0:
                break
            around.add(k)
            k = int(to_graph_index[k])
    around.update(np.where(to_graph_index < 0)[0].tolist())
    around.update(gg.seg_sample(np.asarray(params.parameters["tosegment"])).tolist())
    sample = np.array(sorted(a for a in around if a >= 0), dtype=np.int64)
    z = {"node_sample": sample, "to_graph_index": to_graph_index.astype(np.int64), "kind": kind,
         "inflows_f32": inflows.astype(np.float32), "requests": req[:, nseg:nseg + 2],
         "flow_min": flow_min, "pass_above_nhm_seg": np.int64(pass_above)}
    assert (z["inflows_f32"].astype(np.float64) == inflows).all()
    for k in names:
        z["ts_" + k] = out[k][:, sample]
        z["sum_" + k] = np.nansum(out[k], axis=1)
    mk = makers["prms_channel"]
    padc = lambda a, fill: np.concatenate([np.asarray(a), np.full(nn - nseg, fill, dtype=np.asarray(a).dtype)])  # noqa: E731
    coef = dict(tsi=padc(mk._tsi, 1).astype(np.int32), ts=padc(mk._ts, 1.0), c0=padc(mk._c0, 0.0),
                c1=padc(mk._c1, 0.0), c2=padc(mk._c2, 0.0))
    for k, v in coef.items():
        z["coef_" + k] = v
    np.savez_compressed(gg.GOLDEN / "drb_flowgraph_ss_golden.npz", **z)
    ss = out["node_sink_source"][:, nseg]
    print("flow graph with source/sink nodes:", nn, "nodes,", len(sample), "sampled;",
          "sink node: full sink on", int((np.abs(ss - sink_req) < 1e-12).sum()), "days, cut back on",
          int(((ss < 0) & (np.abs(ss - sink_req) > 1e-9)).sum()), ", none on", int((ss == 0).sum()),
          ", source on", int((ss > 0).sum()))
    o = po.Oracle(scenarios.flow_graph_params(p, to_graph_index), start=scenarios.load_drb()[2])
    o.set_channel_coefficients(**coef)
    got = o.run_flow_graph(inflows, kind, req, fmin_full)
    for k in names:
        a, b = got[k], out[k]
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        assert same.all(), (k, int((~same).sum()), np.nanmax(np.abs(a - b)))
    print("  oracle with the reference's Muskingum coefficients: every value of every variable bit-identical")


if __name__ == "__main__":
    main()

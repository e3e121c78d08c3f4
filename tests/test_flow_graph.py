"""FlowGraph (base/flow_graph.py) on CPU: the oracle's restatement against the
fixture generated from the unmodified reference (oracle/gen_golden_flowgraph.py:
the drb_2yr network as PRMSChannelFlowNodes + a PassThroughFlowNode + an
ObsInFlowNode), and the host-side graph construction / HRU-to-segment mapping
of pywatershed_b200.flow_graph against the same fixture and the oracle."""
import numpy as np
import pytest

import prms_oracle as po
import scenarios

ND = 731


@pytest.fixture(scope="module")
def golden():
    return np.load(scenarios.GOLDEN / "drb_flowgraph_golden.npz")


def _graph_inputs(z):
    inflows = z["inflows_f32"].astype(np.float64)
    nn = inflows.shape[1]
    obs = np.zeros((ND, nn))
    obs[:, int(z["obs_node"])] = z["obs"]
    return inflows, z["kind"], obs


def test_oracle_flow_graph_bit_exact_with_the_reference_coefficients(drb, golden):
    """With the reference's own Muskingum coefficients (numpy's vectorised power is
    not the libm's pow in the last bit on 28 of 456 reaches) every sampled value of
    every FlowGraph variable is the reference's, bit for bit, over 731 days."""
    p, _, start = drb
    z = golden
    inflows, kind, obs = _graph_inputs(z)
    o = po.Oracle(scenarios.flow_graph_params(p, z["to_graph_index"]), start=start)
    o.set_channel_coefficients(z["coef_tsi"], z["coef_ts"], z["coef_c0"], z["coef_c1"], z["coef_c2"])
    got = o.run_flow_graph(inflows, kind, obs)
    s = z["node_sample"]
    for k in po.FLOW_GRAPH_VARS:
        np.testing.assert_array_equal(got[k][:, s], z["ts_" + k], err_msg=k)
        np.testing.assert_allclose(np.nansum(got[k], axis=1), z["sum_" + k], rtol=1e-12, atol=1e-9, err_msg=k)
    np.testing.assert_array_equal(got["node_outflows"][-1], z["last_node_outflows"])
    # the ObsIn node: out goes the observation, or the first hour's inflow when it is missing
    j = int(z["obs_node"])
    have = z["obs"] >= 0
    np.testing.assert_array_equal(got["node_outflows"][have, j], z["obs"][have])
    assert (got["node_sink_source"][:, j] != 0).any() and (got["node_storage_changes"][:, j] == 0).all()


def test_oracle_flow_graph_with_its_own_coefficients(drb, golden):
    p, _, start = drb
    z = golden
    inflows, kind, obs = _graph_inputs(z)
    got = po.Oracle(scenarios.flow_graph_params(p, z["to_graph_index"]), start=start).run_flow_graph(
        inflows, kind, obs)
    s = z["node_sample"]
    for k in po.FLOW_GRAPH_VARS:
        np.testing.assert_allclose(got[k][:, s], z["ts_" + k], rtol=1e-9, atol=1e-10, err_msg=k)


def test_oracle_source_sink_nodes_bit_exact_with_the_reference_coefficients(drb):
    """SourceSinkFlowNode (hydrology/source_sink_flow_node.py:60-92), every branch
    of calculate_subtimestep taken, against the fixture of
    oracle/gen_golden_flowgraph_ss.py (unmodified reference, 240 days)."""
    p, _, start = drb
    z = np.load(scenarios.GOLDEN / "drb_flowgraph_ss_golden.npz")
    inflows = z["inflows_f32"].astype(np.float64)
    nd, nn = inflows.shape
    nseg = nn - 3
    req = np.zeros((nd, nn))
    req[:, nseg:nseg + 2] = z["requests"]
    fmin = np.zeros(nn)
    fmin[nseg:nseg + 2] = z["flow_min"]
    o = po.Oracle(scenarios.flow_graph_params(p, z["to_graph_index"]), start=start)
    o.set_channel_coefficients(z["coef_tsi"], z["coef_ts"], z["coef_c0"], z["coef_c1"], z["coef_c2"])
    got = o.run_flow_graph(inflows, z["kind"], req, fmin)
    s = z["node_sample"]
    for k in po.FLOW_GRAPH_VARS:
        np.testing.assert_array_equal(got[k][:, s], z["ts_" + k], err_msg=k)
        np.testing.assert_allclose(np.nansum(got[k], axis=1), z["sum_" + k], rtol=1e-12, atol=1e-9, err_msg=k)
    ss = got["node_sink_source"][:, nseg]
    assert (ss < 0).any() and (ss > 0).any() and (ss == 0).any()
    assert (got["node_storage_changes"][:, nseg:nseg + 2] == 0).all()


def test_source_sink_maker_series_and_flow_min():
    import pywatershed_b200 as pwsb

    t = np.datetime64("1979-01-01") + np.arange(5)
    mk = pwsb.SourceSinkFlowNodeMaker({"flow_min": np.array([2.0, 0.5])},
                                      [(t, np.arange(5.0)), (t[::-1], -np.arange(5.0))])
    assert mk.kind == 3 and mk.flow_min(1) == 0.5
    np.testing.assert_array_equal(mk.series(1, t[1:3].astype("datetime64[s]")), [-3.0, -2.0])
    with pytest.raises(KeyError, match="do not cover"):
        mk.series(0, t + 3)
    import pandas as pd

    df = pd.DataFrame({"a": np.arange(5.0), "b": np.ones(5)}, index=pd.DatetimeIndex(t))
    mk = pwsb.SourceSinkFlowNodeMaker({"flow_min": np.zeros(2)}, df)
    np.testing.assert_array_equal(mk.series(0, t), np.arange(5.0))


def test_graph_construction_follows_the_reference_helper(drb, golden):
    """_build_flow_graph_inputs (prms_channel_flow_graph.py:985-1100): the same
    to_graph_index as the reference's helper produced for the fixture."""
    import pywatershed_b200 as pwsb
    from pywatershed_b200.flow_graph import _build_flow_graph_inputs

    p, _, _ = drb
    obs_par = {"poi_gage_id": np.array(["gage_a"])}
    makers = {"pass": pwsb.PassThroughFlowNodeMaker(), "obsin": pwsb.ObsInFlowNodeMaker(obs_par, {})}
    pfg, node_makers = _build_flow_graph_inputs(p, p, makers, ["pass", "obsin"], [0, 0], [9001, 9002],
                                                [1829, 1766])
    np.testing.assert_array_equal(pfg.parameters["to_graph_index"], golden["to_graph_index"])
    assert list(pfg.parameters["node_maker_name"][-2:]) == ["pass", "obsin"]
    assert set(node_makers) == {"prms_channel", "pass", "obsin"}
    # new nodes in series: the second one above the first (a non-positive entry = index of a new node)
    pfg2, _ = _build_flow_graph_inputs(p, p, makers, ["pass", "obsin"], [0, 0], [9001, 9002], [1829, 0])
    t = np.asarray(pfg2.parameters["to_graph_index"])
    nseg = t.shape[0] - 2
    assert t[nseg + 1] == nseg and t[nseg] == golden["to_graph_index"][nseg]
    with pytest.raises(AssertionError):
        _build_flow_graph_inputs(p, p, makers, ["pass", "obsin"], [0, 0], [1, 2], [1829, 1829])


def test_hru_segment_flow_adapter_is_the_channels_lateral_inflow(drb):
    """HruSegmentFlowAdapter (prms_channel_flow_graph.py:389-472) = PRMSChannel's
    seg_lateral_inflow (prms_channel.py:396-421), bit for bit."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    nd = 40
    res, _ = po.Oracle(p, start=start).run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd],
                                           ("sroff_vol", "ssres_flow_vol", "gwres_flow_vol"),
                                           ("seg_lateral_inflow",))
    ad = pwsb.HruSegmentFlowAdapter(p, res["sroff_vol"], res["ssres_flow_vol"], res["gwres_flow_vol"])
    np.testing.assert_array_equal(ad.data, res["seg_lateral_inflow"])


def test_flow_graph_descriptors_and_checks(drb):
    import pywatershed_b200 as pwsb

    fg = pwsb.FlowGraph
    assert fg.get_dimensions() == ("nnodes",)
    assert fg.get_inputs() == ("inflows",)
    assert fg.get_parameters() == ("node_maker_name", "node_maker_index", "node_maker_id", "to_graph_index")
    assert fg.get_variables() == po.FLOW_GRAPH_VARS
    assert fg.get_mass_budget_terms() == {
        "inputs": ["inflows"], "outputs": ["outflows"],
        "storage_changes": ["node_storage_changes", "node_negative_sink_source"]}
    t0 = np.datetime64("1979-01-01T00:00:00")
    control = pwsb.Control(t0, t0 + np.timedelta64(9, "D"), np.timedelta64(24, "h"))
    par = pwsb.Parameters(dims={"nnodes": 3}, data_vars={
        "node_maker_name": np.array(["pass"] * 3), "node_maker_index": np.arange(3),
        "to_graph_index": np.array([1, -1, -1])})
    with pytest.raises(ValueError, match="Disconnected nodes"):
        pwsb.FlowGraph(control, None, par, np.zeros(3), {"pass": pwsb.PassThroughFlowNodeMaker()})
    with pytest.warns(UserWarning, match="Disconnected nodes"):
        pwsb.FlowGraph(control, None, par, np.zeros(3), {"pass": pwsb.PassThroughFlowNodeMaker()},
                       allow_disconnected_nodes=True)

    class Mine(pwsb.FlowNodeMaker):
        pass

    with pytest.raises(NotImplementedError, match="no CPU path"):
        pwsb.FlowGraph(control, None, par, np.zeros(3), {"pass": Mine()}, allow_disconnected_nodes=True)


def test_hru_node_flow_exchange_and_model_dict_helper(drb):
    """HruNodeFlowExchange (prms_channel_flow_graph.py:501-620) for a block of days =
    the reference's per-day loop: inflows per node in ascending HRU order (the
    PRMS segments' part is HruSegmentFlowAdapter's / PRMSChannel's lateral inflow,
    bit for bit), HRUs without a segment go to `sinks`, new nodes get nothing; and
    prms_channel_flow_graph_to_model_dict (:746-895) appends the two entries."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    nd = 12
    res, _ = po.Oracle(p, start=start).run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd],
                                           ("sroff_vol", "ssres_flow_vol", "gwres_flow_vol"),
                                           ("seg_lateral_inflow",))
    t0 = np.datetime64(start, "s")
    control = pwsb.Control(t0, t0 + np.timedelta64(nd - 1, "D"), np.timedelta64(24, "h"),
                           options={"budget_type": "error"})
    md = {"control": control, "dis_hru": p, "model_order": ["prms_runoff"],
          "prms_runoff": {"class": pwsb.PRMSRunoff, "parameters": p, "dis": "dis_hru"}}
    out = pwsb.prms_channel_flow_graph_to_model_dict(
        md, p, "dis_hru", p, {"pass": pwsb.PassThroughFlowNodeMaker()}, ["pass"], [0], [9001], [1829])
    assert out["model_order"] == ["prms_runoff", "inflow_exchange", "prms_channel_flow_graph"]
    assert md["model_order"] == ["prms_runoff"]  # the caller's dict is not edited
    assert out["inflow_exchange"]["class"] is pwsb.HruNodeFlowExchange
    assert out["prms_channel_flow_graph"]["class"] is pwsb.FlowGraph
    nseg = np.asarray(p["tosegment"]).shape[0]
    ex = pwsb.HruNodeFlowExchange(control, p, out["inflow_exchange"]["parameters"])
    assert ex.nnodes == nseg + 1 and ex.budget.basis == "global"
    blk = ex.block(res["sroff_vol"], res["ssres_flow_vol"], res["gwres_flow_vol"])
    np.testing.assert_array_equal(blk["inflows"][:, :nseg], res["seg_lateral_inflow"])
    assert (blk["inflows"][:, nseg] == 0).all()
    no_seg = np.asarray(p["hru_segment"]) == 0
    tot = (0 + res["sroff_vol"] + res["ssres_flow_vol"] + res["gwres_flow_vol"]) / 86400.0
    np.testing.assert_array_equal(blk["sinks"][:, no_seg], tot[:, no_seg])
    assert (blk["sinks"][:, ~no_seg] == 0).all() and no_seg.sum() == 4
    np.testing.assert_array_equal(blk["inflows_vol"], blk["inflows"] * 86400.0)
    # the global budget of the exchange closes: what comes in goes to nodes or sinks
    lhs = (res["sroff_vol"] + res["ssres_flow_vol"] + res["gwres_flow_vol"]).sum(axis=1)
    np.testing.assert_allclose(blk["inflows_vol"].sum(axis=1) + blk["sinks_vol"].sum(axis=1), lhs, rtol=1e-12)

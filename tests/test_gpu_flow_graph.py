"""FlowGraph on the routing kernel (k_route_chunks<G, true>, through the C ABI:
pws_set_flow_graph / pws_run_flow_graph_host) against the oracle's restatement
(bit for bit: routing has no transcendental function) and against the fixture
generated from the unmodified reference (rtol 1e-9 / atol 1e-10).  GPU only."""
import numpy as np
import pytest

import prms_oracle as po
import scenarios

pytestmark = pytest.mark.gpu
ND = 731


@pytest.fixture(scope="module")
def golden():
    return np.load(scenarios.GOLDEN / "drb_flowgraph_golden.npz")


def _control(start, nd, **opts):
    import pywatershed_b200 as pwsb

    t0 = np.datetime64(start, "s")
    return pwsb.Control(t0, t0 + np.timedelta64(nd - 1, "D"), np.timedelta64(24, "h"),
                        options={"budget_type": "error", **opts})


def _flow_graph(drb, z, inflows, nd=ND, **kw):
    import pywatershed_b200 as pwsb
    from pywatershed_b200.flow_graph import _build_flow_graph_inputs

    p, _, start = drb
    control = _control(start, nd)
    times = control.start_time + np.arange(ND) * control.time_step
    makers = {"pass": pwsb.PassThroughFlowNodeMaker(),
              "obsin": pwsb.ObsInFlowNodeMaker({"poi_gage_id": np.array(["gage_a"])},
                                               {"gage_a": (times, z["obs"])})}
    pfg, node_makers = _build_flow_graph_inputs(p, p, makers, ["pass", "obsin"], [0, 0], [9001, 9002],
                                                [1829, 1766])
    return pwsb.FlowGraph(control, None, pfg, inflows, node_makers, budget_type="error", **kw), control


def test_flow_graph_equals_the_oracle_bit_for_bit_and_the_reference(drb, golden):
    p, _, start = drb
    z = golden
    inflows = z["inflows_f32"].astype(np.float64)
    nn = inflows.shape[1]
    obs = np.zeros((ND, nn))
    obs[:, int(z["obs_node"])] = z["obs"]
    ref = po.Oracle(scenarios.flow_graph_params(p, z["to_graph_index"]), start=start).run_flow_graph(
        inflows, z["kind"], obs)
    fg, control = _flow_graph(drb, z, inflows)
    got = fg.run_all()
    assert fg.budget._zero_sum
    s = z["node_sample"]
    for k in po.FLOW_GRAPH_VARS:
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)  # NaN == NaN for node_storages
        np.testing.assert_allclose(got[k][:, s], z["ts_" + k], rtol=1e-9, atol=1e-10, err_msg=k)
    # -1 * 0.0 is -0.0 in the reference too (flow_graph.py:643)
    assert np.signbit(got["node_negative_sink_source"][:, 0]).all()
    fg.finalize()


def test_flow_graph_day_by_day_adapter_equals_blocks(drb, golden):
    """Inflows from an adapter that only knows advance() / current (the reference's
    GraphInflowAdapter pattern, flow_graph.py docstring) against the blocked run;
    and blocks of 7 days against blocks of a year."""
    z = golden
    nd = 60
    inflows = z["inflows_f32"].astype(np.float64)
    fg, _ = _flow_graph(drb, z, inflows, nd=nd)
    want = fg.run_all()
    fg.finalize()

    class Adapter:
        def __init__(self, control):
            self.control = control
            self._cur = np.full(inflows.shape[1], np.nan)

        def advance(self):
            self._cur[:] = inflows[self.control.itime_step]

        @property
        def current(self):
            return self._cur

    fg2, control2 = _flow_graph(drb, z, None, nd=nd)
    fg2._inflows_src = Adapter(control2)
    got = fg2.run_all()
    fg2.finalize()
    fg3, _ = _flow_graph(drb, z, inflows, nd=nd, block_days=7)
    got3 = fg3.run_all()
    fg3.finalize()
    for k in po.FLOW_GRAPH_VARS:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
        np.testing.assert_array_equal(got3[k], want[k], err_msg=k)


def test_pass_through_node_leaves_prms_channel_unchanged(drb):
    """The reference's own FlowGraph example (flow_graph.py docstring): a pass-through
    node inserted above nhm_seg 1829 -- the flows on the PRMSChannel nodes are those of
    PRMSChannel alone (there: within 1e-10; here: the same bits), through the
    postprocess helper, which forces the graph with the three HRU outflow volumes."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    nd = 120
    res, _ = po.Oracle(p, start=start).run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd],
                                           ("sroff_vol", "ssres_flow_vol", "gwres_flow_vol"), po.SEG_VARS)
    control = _control(start, nd)
    fg = pwsb.prms_channel_flow_graph_postprocess(
        control, None, p, p, {"pass_throughs": pwsb.PassThroughNodeMaker()}, ["pass_throughs"], [0], [9001],
        [1829], budget_type="error",
        inputs={k: res[k] for k in ("sroff_vol", "ssres_flow_vol", "gwres_flow_vol")})
    got = fg.run_all()
    fg.finalize()
    nseg = res["seg_outflow"].shape[1]
    np.testing.assert_array_equal(got["node_outflows"][:, :nseg], res["seg_outflow"])
    np.testing.assert_array_equal(got["node_upstream_inflows"][:, :nseg], res["seg_upstream_inflow"])
    # the pass-through node hands on what the reach above it receives
    below = int(np.where(np.asarray(p["nhm_seg"]) == 1829)[0][0])
    np.testing.assert_array_equal(got["node_outflows"][:, nseg], got["node_upstream_inflows"][:, below])
    assert (got["node_storage_changes"][:, nseg] == 0).all() and (got["node_sink_source"] == 0).all()


def test_flow_graph_budget_raises_on_an_imbalance(drb, golden):
    """A positive control of the global budget: an ObsIn node's sink / source term
    left out of the balance makes it fail on the first day with an observation."""
    z = golden
    inflows = z["inflows_f32"].astype(np.float64)
    fg, control = _flow_graph(drb, z, inflows, nd=30)
    fg.budget.terms["storage_changes"] = ["node_storage_changes"]
    with pytest.raises(ValueError, match="global flux balance"):
        fg.run_all()
    fg.finalize()


def test_flow_graph_netcdf_output(drb, golden, tmp_path):
    z = golden
    nd = 12
    inflows = z["inflows_f32"].astype(np.float64)
    fg, control = _flow_graph(drb, z, inflows, nd=nd)
    fg.initialize_netcdf(tmp_path, output_vars=["node_outflows", "node_sink_source"])
    rows = []
    for _ in range(nd):
        control.advance()
        fg.advance()
        fg.calculate(1.0)
        fg.output()
        rows.append(fg["node_outflows"].copy())
    fg.finalize()
    from pywatershed_b200.netcdf_in import read_time_variable

    _, a = read_time_variable(tmp_path / "node_outflows.nc", "node_outflows")
    np.testing.assert_array_equal(a, np.array(rows))
    assert (tmp_path / "node_sink_source.nc").exists() and not (tmp_path / "outflows.nc").exists()


def test_source_sink_nodes_equal_the_oracle_bit_for_bit_and_the_reference(drb):
    """SourceSinkFlowNode (hydrology/source_sink_flow_node.py): a sink whose requests
    exceed what the minimum flow leaves on 97 of 240 days (cut back), are refused on
    low-flow days and are sources on others, a source node, and a pass-through node,
    against the oracle (same bits) and the fixture from the unmodified reference
    (oracle/gen_golden_flowgraph_ss.py)."""
    import pywatershed_b200 as pwsb
    from pywatershed_b200.flow_graph import _build_flow_graph_inputs

    z = np.load(scenarios.GOLDEN / "drb_flowgraph_ss_golden.npz")
    p, _, start = drb
    inflows = z["inflows_f32"].astype(np.float64)
    nd, nn = inflows.shape
    nseg = nn - 3
    req = np.zeros((nd, nn))
    req[:, nseg:nseg + 2] = z["requests"]
    fmin = np.zeros(nn)
    fmin[nseg:nseg + 2] = z["flow_min"]
    ref = po.Oracle(scenarios.flow_graph_params(p, z["to_graph_index"]), start=start).run_flow_graph(
        inflows, z["kind"], req, fmin)
    control = _control(start, nd)
    times = control.start_time + np.arange(nd) * control.time_step
    makers = {"ss": pwsb.SourceSinkFlowNodeMaker({"flow_min": z["flow_min"]},
                                                 [(times, z["requests"][:, 0]), (times, z["requests"][:, 1])]),
              "pass": pwsb.PassThroughFlowNodeMaker()}
    pfg, node_makers = _build_flow_graph_inputs(p, p, makers, ["ss", "ss", "pass"], [0, 1, 0],
                                                [9001, 9002, 9003], [1766, 1829, int(z["pass_above_nhm_seg"])])
    np.testing.assert_array_equal(pfg.parameters["to_graph_index"], z["to_graph_index"])
    fg = pwsb.FlowGraph(control, None, pfg, inflows, node_makers, budget_type="error")
    got = fg.run_all()
    assert fg.budget._zero_sum
    fg.finalize()
    s = z["node_sample"]
    for k in po.FLOW_GRAPH_VARS:
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
        np.testing.assert_allclose(got[k][:, s], z["ts_" + k], rtol=1e-9, atol=1e-10, err_msg=k)
    ss = got["node_sink_source"][:, nseg]
    want = z["requests"][:, 0]
    assert ((ss < 0) & (ss != want)).sum() > 50 and (ss == want).sum() > 20 and (ss > 0).any()
    # the node stores nothing and its sink / source is the daily mean of what it added or took
    assert (got["node_storage_changes"][:, nseg:nseg + 2] == 0).all()
    assert (ss >= want - 1e-9 * np.abs(want)).all()  # a sink is only ever cut back, never enlarged


def test_flow_graph_behind_the_prms_column_in_a_model(drb, tmp_path):
    """prms_channel_flow_graph_to_model_dict (prms_channel_flow_graph.py:746-895): the
    PRMS column without PRMSChannel + `inflow_exchange` + a FlowGraph with a
    pass-through node above nhm_seg 1829, as ONE Model.  The flows on the PRMSChannel
    nodes are those of the NHM model with PRMSChannel (the reference's notebook
    checks 1e-10; here: the same bits), through run() (blocks) and day by day, with
    the exchange's and the graph's global budgets closed and netCDF files written."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    nd = 75
    chain = ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoff",
             "PRMSSoilzone", "PRMSGroundwater"]
    nhru, nseg = p["hru_area"].shape[0], p["tosegment"].shape[0]
    par = pwsb.Parameters(dims={"nhru": nhru, "nsegment": nseg, "nmonth": 12, "ndoy": 366},
                          coords={"nhm_id": p["nhm_id"], "nhm_seg": p["nhm_seg"]},
                          data_vars={k: v for k, v in p.items() if k not in ("nhm_id", "nhm_seg")})

    def control():
        return _control(start, nd, calc_method="cuda")

    def build(ctl):
        md = {"control": ctl, "dis_both": par, "model_order": [n.lower() for n in chain]}
        for n in chain:
            md[n.lower()] = {"class": getattr(pwsb, n), "parameters": par, "dis": "dis_both"}
        md = pwsb.prms_channel_flow_graph_to_model_dict(
            md, par, "dis_both", par, {"pass_throughs": pwsb.PassThroughNodeMaker()}, ["pass_throughs"],
            [0], [9001], [1829], graph_budget_type="error")
        m = pwsb.Model(md, find_input_files=False)
        for k in ("prcp", "tmax", "tmin"):
            m.processes["PRMSAtmosphere"].set_input_to_adapter(k, f[k][:nd])
        return m

    # the NHM model with PRMSChannel: the flows to reproduce
    ctl0 = control()
    m0 = pwsb.Model([getattr(pwsb, n) for n in chain + ["PRMSChannel"]], control=ctl0, parameters=par,
                    find_input_files=False)
    for k in ("prcp", "tmax", "tmin"):
        m0.processes["PRMSAtmosphere"].set_input_to_adapter(k, f[k][:nd])
    want = []
    for _ in range(nd):
        m0.advance()
        m0.calculate()
        want.append(m0.processes["PRMSChannel"]["seg_outflow"].copy())
    want = np.array(want)
    m0.finalize()

    # day by day
    m = build(control())
    assert list(m.processes)[-2:] == ["inflow_exchange", "prms_channel_flow_graph"]
    fg, ex = m.processes["prms_channel_flow_graph"], m.processes["inflow_exchange"]
    got = []
    for _ in range(nd):
        m.advance()
        m.calculate()
        got.append(fg["node_outflows"].copy())
        assert fg.budget._zero_sum and ex.budget._zero_sum
    got = np.array(got)
    m.finalize()
    np.testing.assert_array_equal(got[:, :nseg], want)
    below = int(np.where(np.asarray(p["nhm_seg"]) == 1829)[0][0])
    assert (got[:, nseg] > 0).any()

    # run(): blocks, fast path, netCDF
    m = build(control())
    m.run(netcdf_dir=tmp_path, finalize=False, output_vars=["node_outflows", "inflows", "sroff_vol"])
    fg = m.processes["prms_channel_flow_graph"]
    np.testing.assert_array_equal(fg["node_outflows"][:nseg], want[-1])
    np.testing.assert_array_equal(fg["node_upstream_inflows"][below], got[-1, nseg])
    acc = fg.budget.accumulations
    np.testing.assert_allclose(acc["outputs"]["outflows"].sum(), np.where(
        np.asarray(fg.outflow_mask), got, 0.0).sum(), rtol=1e-12)
    m.finalize()
    from pywatershed_b200.netcdf_in import read_time_variable

    _, a = read_time_variable(tmp_path / "node_outflows.nc", "node_outflows")
    np.testing.assert_array_equal(a, got)
    assert (tmp_path / "sroff_vol.nc").exists() and (tmp_path / "inflows.nc").exists()

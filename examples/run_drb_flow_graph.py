"""The Delaware River Basin NHM configuration with a FlowGraph behind the PRMS
column instead of PRMSChannel (the reference's examples/06_flow_graph_starfit
pattern, with the node kinds that run on the device): a source / sink node that
withdraws water above a minimum flow above nhm_seg 1766 and a pass-through node
above nhm_seg 1829, in ONE Model.

    python examples/run_drb_flow_graph.py [output_dir]

The same model_dict works with the reference's own Control / Parameters objects.
"""
import pathlib as pl
import sys
import time

import numpy as np

ROOT = pl.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import pywatershed_b200 as pws  # noqa: E402

CHAIN = ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoff", "PRMSSoilzone",
         "PRMSGroundwater"]


def main():
    out_dir = pl.Path(sys.argv[1]) if len(sys.argv) > 1 else None
    z = np.load(ROOT / "tests" / "golden" / "drb_inputs.npz")
    p = {k[2:]: z[k] for k in z.files if k.startswith("p_")}
    nhru, nseg = p["hru_area"].shape[0], p["tosegment"].shape[0]
    params = pws.Parameters(
        dims={"nhru": nhru, "nsegment": nseg, "nmonth": 12, "ndoy": 366},
        coords={"nhm_id": p.pop("nhm_id"), "nhm_seg": p.pop("nhm_seg")}, data_vars=p)
    start = np.datetime64(str(z["start"]), "s")
    nd = z["prcp"].shape[0]
    control = pws.Control(start, start + np.timedelta64(nd - 1, "D"), np.timedelta64(24, "h"),
                          options={"budget_type": "error", "calc_method": "cuda"})
    times = control.start_time + np.arange(nd) * control.time_step

    # the PRMS column, then the graph
    model_dict = {"control": control, "dis_both": params, "model_order": [n.lower() for n in CHAIN]}
    for n in CHAIN:
        model_dict[n.lower()] = {"class": getattr(pws, n), "parameters": params, "dis": "dis_both"}
    withdrawal = np.full(nd, -150.0)          # cfs, every day
    withdrawal[200:230] = 40.0                # a month of releases instead
    makers = {
        "withdrawals": pws.SourceSinkFlowNodeMaker({"flow_min": np.array([300.0])}, [(times, withdrawal)]),
        "pass_throughs": pws.PassThroughNodeMaker(),
    }
    model_dict = pws.prms_channel_flow_graph_to_model_dict(
        model_dict, params, "dis_both", params, makers, ["withdrawals", "pass_throughs"], [0, 0],
        [9001, 9002], [1766, 1829], graph_budget_type="error")
    model = pws.Model(model_dict, find_input_files=False)
    for k in ("prcp", "tmax", "tmin"):
        model.processes["PRMSAtmosphere"].set_input_to_adapter(k, z[k])

    t0 = time.perf_counter()
    model.run(netcdf_dir=out_dir, finalize=False,
              output_vars=["node_outflows", "node_sink_source", "inflows"] if out_dir else None)
    dt = time.perf_counter() - t0
    fg = model.processes["prms_channel_flow_graph"]
    taken = fg.budget.accumulations["storage_changes"]["node_negative_sink_source"]
    print(f"{nd} days x {nhru} HRUs + {fg.nnodes} nodes in {dt:.2f} s")
    print(f"last day: outflow of the basin {fg['outflows'].sum():.1f} cfs; the withdrawal node took "
          f"{taken[nseg]:.1f} cfs-days over the run (asked for {-withdrawal[withdrawal < 0].sum():.0f}), "
          f"its last sink / source {fg['node_sink_source'][nseg]:.2f} cfs")
    model.finalize()
    if out_dir is not None:
        print("wrote", sorted(f.name for f in pl.Path(out_dir).glob("*.nc")))


if __name__ == "__main__":
    main()

"""FlowGraph: routing over a graph of FlowNodes, on the routing kernel.

The reference (pywatershed/base/flow_graph.py:131-657) steps one Python object
per node and hour.  Here the graph is a PRMSChannel-shaped problem -- nodes are
the segments of an engine handle, `to_graph_index` is `tosegment` -- and the
node KINDS the reference ships as plain flow solutions are cases of the routing
kernel (k_route_chunks<G, true>, include/pws_b200.h: pws_set_flow_graph):

  * PRMSChannelFlowNode (hydrology/prms_channel_flow_graph.py:30-160): the
    Muskingum-Mann recurrence, coefficients as PRMSChannelFlowNodeMaker._init_data
    (:230-330) derives them;
  * PassThroughFlowNode (hydrology/pass_through_flow_node.py:6-53);
  * ObsInFlowNode (hydrology/obsin_flow_node.py:8-90);
  * SourceSinkFlowNode (hydrology/source_sink_flow_node.py:10-125).

Same constructor signatures, variable names, budget terms and helper functions
(prms_channel_flow_graph_postprocess, prms_channel_flow_graph_to_model_dict,
HruSegmentFlowAdapter, HruNodeFlowExchange; inside a Model the exchange and the
graph are evaluated per block of days behind the PRMS column, model.py:_tail_block).  A node maker of any other class
(StarfitFlowNodeMaker, user-defined nodes computed in Python) is rejected: there
is no CPU path in this package.

Execution is time-blocked like Model: `advance()` / `calculate()` on day d
compute a block of days at once when the inflows are known ahead (an array
[ntime, nnodes], a netCDF file, an adapter with `.data`), and one day at a time
when they come from an adapter that only knows `advance()` / `.current`.
"""
from __future__ import annotations

import pathlib as pl
from warnings import warn

import numpy as np

from .budget import Budget
from .parameters import Parameters, as_param_dict

KIND_MUSKINGUM, KIND_PASS, KIND_OBSIN, KIND_SOURCESINK = 0, 1, 2, 3


class FlowNode:
    """base/flow_graph.py:16-99: the per-node interface.  Nodes are not stepped
    in Python here; the class exists so that `isinstance` checks keep working."""


class FlowNodeMaker:
    """base/flow_graph.py:102-127"""

    kind = None
    name = "FlowNodeMaker"

    def __init__(self, discretization: Parameters = None, parameters: Parameters = None):
        self.name = "FlowNodeMaker"


class PRMSChannelFlowNodeMaker(FlowNodeMaker):
    """hydrology/prms_channel_flow_graph.py:163-330"""

    kind = KIND_MUSKINGUM

    def __init__(self, discretization: Parameters, parameters: Parameters, calc_method=None,
                 verbose: bool = None):
        self.name = "PRMSChannelFlowNodeMaker"
        if calc_method not in (None, "cuda", "numba", "numpy"):
            raise ValueError(f"Invalid choice of calc_method: {calc_method}")
        self._calc_method = calc_method
        merged = {}
        for p in (discretization, parameters):
            if p is not None:
                merged.update(as_param_dict(p))
        missing = [k for k in self.get_parameters() if k not in merged and k != "tosegment_nhm"]
        if missing:
            raise ValueError(f"PRMSChannelFlowNodeMaker: parameters missing: {sorted(missing)}")
        self._params = merged
        self.nsegment = int(np.asarray(merged["seg_length"]).shape[0])

    @staticmethod
    def get_dimensions() -> tuple:
        return ("nsegment",)

    @staticmethod
    def get_parameters() -> tuple:
        return ("mann_n", "seg_depth", "seg_length", "seg_slope", "segment_type", "tosegment",
                "tosegment_nhm", "x_coef", "segment_flow_init")


class PassThroughFlowNodeMaker(FlowNodeMaker):
    """hydrology/pass_through_flow_node.py:56-77"""

    kind = KIND_PASS

    def __init__(self) -> None:
        self.name = "PassThroughFlowNodeMaker"


class ObsInFlowNodeMaker(FlowNodeMaker):
    """hydrology/obsin_flow_node.py:93-118: `parameters` holds `poi_gage_id`,
    `obs_data` is a pandas DataFrame (index = dates, one column per gage id) as
    pyPRMS.Streamflow gives it, or a mapping gage id -> (times, values)."""

    kind = KIND_OBSIN

    def __init__(self, parameters: Parameters, obs_data) -> None:
        self.name = "ObsInNodeMaker"
        self._parameters = parameters
        self._obs_data = obs_data

    def series(self, index: int, times: np.ndarray) -> np.ndarray:
        """The observations of node `index` on the days `times` (datetime64)."""
        gid = np.asarray(as_param_dict(self._parameters)["poi_gage_id"])[index]
        if isinstance(gid, bytes):
            gid = gid.decode()
        col = self._obs_data[gid]
        if hasattr(col, "values") and not isinstance(col, (tuple, list)):  # pandas Series
            src_t = np.asarray(col.index.values).astype("datetime64[D]")
            src_v = np.asarray(col.values, dtype=np.float64)
        else:
            src_t = np.asarray(col[0]).astype("datetime64[D]")
            src_v = np.asarray(col[1], dtype=np.float64)
        return _values_on_days(src_t, src_v, times, f"observations of {gid!r}")


class SourceSinkFlowNodeMaker(FlowNodeMaker):
    """hydrology/source_sink_flow_node.py:128-166: `parameters` holds `flow_min`
    [n], `source_sink_df` a pandas DataFrame (index = dates) whose columns are
    collated with `flow_min` by POSITION (the names are not used), or a
    sequence of (times, values) pairs.  Sources are positive, sinks negative; a
    sink only takes what leaves `flow_min` in the channel."""

    kind = KIND_SOURCESINK

    def __init__(self, parameters: Parameters, source_sink_df) -> None:
        self.name = "SourceSinkNodeMaker"
        self._parameters = parameters
        self._source_sink_df = source_sink_df

    def flow_min(self, index: int) -> float:
        return float(np.asarray(as_param_dict(self._parameters)["flow_min"])[index])

    def series(self, index: int, times: np.ndarray) -> np.ndarray:
        """The requests of node `index` on the days `times` (datetime64)."""
        df = self._source_sink_df
        if hasattr(df, "columns"):
            col = df[df.columns[index]]
            src_t = np.asarray(col.index.values).astype("datetime64[D]")
            src_v = np.asarray(col.values, dtype=np.float64)
        else:
            src_t = np.asarray(df[index][0]).astype("datetime64[D]")
            src_v = np.asarray(df[index][1], dtype=np.float64)
        return _values_on_days(src_t, src_v, times, f"requests of source/sink column {index}")


def _values_on_days(src_t, src_v, times, what):
    days = times.astype("datetime64[D]")
    order = np.argsort(src_t)
    src_t, src_v = src_t[order], src_v[order]
    j = np.searchsorted(src_t, days)
    if (j >= src_t.size).any() or (src_t[np.minimum(j, src_t.size - 1)] != days).any():
        raise KeyError(f"{what} do not cover the simulation days")
    return src_v[j]


# the reference exports both spellings
PassThroughNodeMaker = PassThroughFlowNodeMaker
ObsInNodeMaker = ObsInFlowNodeMaker


class FlowGraph:
    """base/flow_graph.py:131-657"""

    _all_time = False
    VARIABLES = ("outflows", "node_upstream_inflows", "node_outflows", "node_storage_changes",
                 "node_storages", "node_sink_source", "node_negative_sink_source")

    def __init__(self, control, discretization, parameters, inflows, node_maker_dict: dict,
                 addtl_output_vars: list = None, params_not_to_netcdf: list = None,
                 budget_type="defer", allow_disconnected_nodes: bool = False,
                 type_check_nodes: bool = False, verbose: bool = None, device: int = None,
                 block_days: int = None):
        self.name = "FlowGraph"
        self.control = control
        self.discretization = discretization
        self.params = parameters
        self._params = parameters
        self._param_dict = as_param_dict(parameters)
        for k in ("node_maker_name", "node_maker_index", "to_graph_index"):
            if k not in self._param_dict:
                raise ValueError(f"FlowGraph: parameters missing: ['{k}']")  # base/process.py:233-237
        self._node_maker_dict = node_maker_dict
        for fnm in node_maker_dict.values():
            assert isinstance(fnm, FlowNodeMaker)
        if addtl_output_vars:
            raise NotImplementedError("addtl_output_vars: the node kinds of this package have no extra variables")
        self._addtl_output_vars = []
        self._params_not_to_netcdf = params_not_to_netcdf
        self._allow_disconnected_nodes = allow_disconnected_nodes
        self._verbose = verbose
        self._inflows_src = inflows
        self._input_variables_dict = {"inflows": inflows}
        if budget_type == "defer":  # conservative_process.py:66-77
            budget_type = control.options.get("budget_type", "warn") if hasattr(control, "options") else "warn"
        self._budget_type = budget_type
        self._device = device if device is not None else int(getattr(control, "options", {}).get("device", 0) or 0)
        self._block_days_user = block_days
        self._init_graph()
        self._vars = {k: np.full(self.nnodes, np.nan) for k in self.VARIABLES}
        self.inflows = np.full(self.nnodes, np.nan)
        self.budget = None
        if budget_type is not None:
            self.budget = Budget(control, self, self.get_mass_budget_terms(), budget_type, basis="global",
                                 description="FlowGraph")
        self._engine = None
        self._block = None
        self._block_t0, self._block_n = 0, 0
        self._itime = -1
        self._netcdf = None
        self._netcdf_initialized = False
        self._model = None

    # ---- descriptors
    @staticmethod
    def get_dimensions() -> tuple:
        return ("nnodes",)

    @staticmethod
    def get_parameters() -> tuple:
        return ("node_maker_name", "node_maker_index", "node_maker_id", "to_graph_index")

    @staticmethod
    def get_inputs() -> tuple:
        return ("inflows",)

    @staticmethod
    def get_init_values() -> dict:
        return {k: np.nan for k in FlowGraph.VARIABLES}

    @classmethod
    def get_variables(cls) -> tuple:
        return list(cls.get_init_values().keys())

    @staticmethod
    def get_mass_budget_terms():
        return {"inputs": ["inflows"], "outputs": ["outflows"],
                "storage_changes": ["node_storage_changes", "node_negative_sink_source"]}

    @staticmethod
    def get_restart_variables() -> tuple:
        return ()

    @classmethod
    def description(cls) -> dict:
        return {"class_name": cls.__name__, "mass_budget_terms": cls.get_mass_budget_terms(),
                "dimensions": cls.get_dimensions(), "parameters": cls.get_parameters(),
                "inputs": cls.get_inputs(), "variables": tuple(cls.get_variables())}

    dimensions = property(lambda self: self.get_dimensions())
    parameters = property(lambda self: self.get_parameters())
    inputs = property(lambda self: self.get_inputs())
    variables = property(lambda self: tuple(self.get_variables()))
    init_values = property(lambda self: self.get_init_values())

    def get_outflow_mask(self):
        return self._outflow_mask

    @property
    def outflow_mask(self):
        return self._outflow_mask

    # ---- graph
    def _init_graph(self):
        """base/flow_graph.py:456-540: the DAG and its node kinds."""
        p = self._param_dict
        tgi = np.asarray(p["to_graph_index"], dtype=np.int64)
        self.nnodes = int(tgi.shape[0])
        names = [n.decode() if isinstance(n, bytes) else str(n) for n in np.asarray(p["node_maker_name"]).tolist()]
        index = np.asarray(p["node_maker_index"], dtype=np.int64)
        if len(names) != self.nnodes or index.shape[0] != self.nnodes:
            raise ValueError("node_maker_name, node_maker_index and to_graph_index must be collated")
        if ((tgi < -1) | (tgi >= self.nnodes)).any():
            raise ValueError("to_graph_index out of range")
        self._outflow_mask = tgi == -1
        connected = np.zeros(self.nnodes, dtype=bool)
        connected[tgi[tgi >= 0]] = True
        connected[tgi >= 0] = True
        if self.nnodes > 1 and not connected.all():
            if not self._allow_disconnected_nodes:
                raise ValueError("Disconnected nodes present in FlowGraph.")
            warn("Disconnected nodes present in FlowGraph.")
        self._to_graph_index = tgi
        kind = np.zeros(self.nnodes, dtype=np.int32)
        obs_col = np.full(self.nnodes, -1, dtype=np.int32)
        flow_min = np.zeros(self.nnodes)
        # stand-in channel parameters for the nodes that are not Muskingum nodes
        seg = {"mann_n": np.full(self.nnodes, 0.04), "seg_depth": np.ones(self.nnodes),
               "seg_length": np.full(self.nnodes, 1000.0), "seg_slope": np.full(self.nnodes, 0.001),
               "x_coef": np.full(self.nnodes, 0.2), "segment_flow_init": np.zeros(self.nnodes),
               "segment_type": np.zeros(self.nnodes, dtype=np.int64)}
        self._obs_nodes = []  # (node, maker, index within the maker)
        for j, (nm, ix) in enumerate(zip(names, index)):
            if nm not in self._node_maker_dict:
                raise KeyError(f"node_maker_name {nm!r} is not in node_maker_dict")
            mk = self._node_maker_dict[nm]
            if mk.kind is None:
                raise NotImplementedError(
                    f"FlowNodeMaker {type(mk).__name__}: only PRMSChannelFlowNodeMaker, PassThroughFlowNodeMaker, "
                    "ObsInFlowNodeMaker and SourceSinkFlowNodeMaker nodes run on the device; there is no CPU path")
            kind[j] = mk.kind
            if mk.kind == KIND_MUSKINGUM:
                for k in seg:
                    seg[k][j] = np.asarray(mk._params[k])[ix]
            elif mk.kind in (KIND_OBSIN, KIND_SOURCESINK):
                obs_col[j] = len(self._obs_nodes)
                self._obs_nodes.append((j, mk, int(ix)))
                if mk.kind == KIND_SOURCESINK:
                    flow_min[j] = mk.flow_min(int(ix))
        self._kind, self._obs_col, self._flow_min = kind, obs_col, flow_min
        seg["tosegment"] = (tgi + 1).astype(np.int64)
        seg["hru_segment"] = np.zeros(1, dtype=np.int64)
        seg["hru_area"] = np.ones(1)
        self._engine_params = seg

    def _ensure_engine(self):
        if self._engine is not None:
            return
        from .engine import Engine

        p = Engine._filled(self._engine_params)
        self._engine = Engine(p, self.control.start_time, device=self._device, channel=True,
                              fill_missing=True,
                              flow_graph=(self._kind, self._obs_col, len(self._obs_nodes), self._flow_min))
        nt = self.control.n_times
        times = self.control.start_time + np.arange(nt) * self.control.time_step
        self._obs = None
        if self._obs_nodes:
            self._obs = np.empty((nt, len(self._obs_nodes)))
            for c, (_, mk, ix) in enumerate(self._obs_nodes):
                self._obs[:, c] = mk.series(ix, times)
        self._inflow_data = self._all_time_inflows()

    def _all_time_inflows(self):
        """[ntime, nnodes] when the inflows are known ahead, else None (an
        adapter that is advanced day by day)."""
        src, nt = self._inflows_src, self.control.n_times
        arr = times = None
        if isinstance(src, np.ndarray):
            arr = np.broadcast_to(src, (nt, self.nnodes)) if src.ndim == 1 else src
        elif isinstance(src, (str, pl.Path)):
            from .netcdf_in import read_time_variable

            times, arr = read_time_variable(pl.Path(src), "inflows")
        elif hasattr(src, "data") and getattr(src, "data", None) is not None and not callable(src.data):
            arr = np.asarray(src.data)
            times = getattr(getattr(src, "_nc_read", None), "times", None)
        if arr is None:
            if not (hasattr(src, "advance") and hasattr(src, "current")):
                raise TypeError("inflows must be an array, a netCDF path or an adapter (advance / current)")
            return None
        if times is not None:
            times = np.asarray(times).astype("datetime64[s]")
            i0 = np.where(times == self.control.start_time.astype("datetime64[s]"))[0]
            if not len(i0):
                raise ValueError("Control start_time is not in the input data time")
            arr = arr[i0[0]:]
        if arr.ndim != 2 or arr.shape[1] != self.nnodes or arr.shape[0] < nt:
            raise ValueError(f"inflows must be [>= {nt}, {self.nnodes}], got {arr.shape}")
        return np.ascontiguousarray(arr[:nt], dtype=np.float64)

    # ---- access
    def __getitem__(self, key):
        if key in self._vars:
            return self._vars[key]
        if key == "inflows":
            return self.inflows
        return getattr(self, key)

    def __getattr__(self, key):
        v = self.__dict__.get("_vars")
        if v is not None and key in v:
            return v[key]
        raise AttributeError(key)

    # ---- stepping
    def _compute_block(self, it):
        self._ensure_engine()
        eng = self._engine
        if eng.day != it:
            raise RuntimeError(f"FlowGraph engine is at day {eng.day}, block starts at {it}")
        if self._inflow_data is not None:
            n = min(self._block_days_user or 366, self.control.n_times - it)
            lat = self._inflow_data[it:it + n]
        else:
            n = 1
            self._inflows_src.advance()
            lat = np.asarray(self._inflows_src.current, dtype=np.float64)[None, :].copy()
        obs = self._obs[it:it + n] if self._obs is not None else None
        blk = eng.run_flow_graph(lat, obs)
        blk["inflows"] = lat
        blk["node_negative_sink_source"] = -1 * blk["node_sink_source"]  # flow_graph.py:643
        self._block, self._block_t0, self._block_n = blk, it, n

    def advance(self):
        """base/process.py:379-395 (the control is advanced by the caller or the Model)"""
        it = self.control.itime_step
        if it == self._itime:
            return
        self._itime = it
        if self._block is None or not (self._block_t0 <= it < self._block_t0 + self._block_n):
            self._compute_block(it)
        self.inflows[:] = self._block["inflows"][it - self._block_t0]

    def calculate(self, time_length: float = 1.0, n_substeps: int = 24) -> None:
        if n_substeps != 24:
            raise NotImplementedError("FlowGraph on the device routes 24 substeps per day")
        it = self.control.itime_step
        if it < 0 or it != self._itime:
            raise RuntimeError("advance() before calculate()")
        j = it - self._block_t0
        for k in self.VARIABLES:
            if k == "node_storages":
                continue  # nan for these node kinds (the FlowNode.storage properties)
            self._vars[k][:] = self._block[k][j]
        if self.budget is not None:
            self.budget.advance()
            self.budget.calculate()

    def run_all(self):
        """Every remaining day (control included); returns var -> [ndays, nnodes]."""
        out = {k: [] for k in self.VARIABLES}
        while self.control.itime_step + 1 < self.control.n_times:
            self.control.advance()
            self.advance()
            self.calculate(1.0)
            for k in self.VARIABLES:
                out[k].append(self._vars[k].copy())
        return {k: np.array(v) for k, v in out.items()}

    def output(self):
        """base/process.py:593-613: the current day to the netCDF files."""
        if not self._netcdf:
            return
        it = self.control.itime_step
        t_days = (self.control.current_time - np.datetime64("1970-01-01T00:00:00")) / np.timedelta64(1, "D")
        for f in self._netcdf:
            f.write(it, it + 1, [t_days], {k: self._vars[k][None, :] for k in f.vars})

    def initialize_netcdf(self, output_dir=None, separate_files: bool = None, budget_args: dict = None,
                          output_vars: list = None, extra_coords: dict = None) -> None:
        """base/flow_graph.py:542-567: <output_dir>/<variable>.nc over (time, node_coord), or
        one FlowGraph.nc with separate_files=False."""
        if self._netcdf_initialized:
            warn(f"{self.name} class previously initialized netcdf output")
            return
        from .netcdf_out import _File

        if output_dir is None:
            output_dir = self.control.options.get("netcdf_output_dir")
        if output_dir is None:
            raise ValueError("no netcdf output_dir")
        if separate_files is None:
            separate_files = self.control.options.get("netcdf_output_separate_files", True)
        names = [v for v in self.VARIABLES if output_vars is None or v in output_vars]
        coord = np.arange(self.nnodes)
        out = pl.Path(output_dir)
        if separate_files:
            self._netcdf = [_File(out / f"{v}.nc", [v], "node_coord", coord, self.name) for v in names]
        else:
            self._netcdf = [_File(out / f"{self.name}.nc", names, "node_coord", coord, self.name)]
        self._netcdf_initialized = True

    def finalize(self):
        if self._netcdf:
            for f in self._netcdf:
                f.close()
            self._netcdf = None
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        self._block = None


# ---------------------------------------------------------------- HRU -> node
class HruSegmentFlowAdapter:
    """hydrology/prms_channel_flow_graph.py:389-472: lateral inflows of the PRMS
    segments (cfs) from the three HRU outflow volumes.  The inputs are arrays
    [ntime, nhru], netCDF paths or adapters with `.data`; `.data` is the whole
    [ntime, nsegment] series, `.current` the control's day."""

    def __init__(self, parameters: Parameters, sroff_vol, ssres_flow_vol, gwres_flow_vol, control=None):
        self._variable = "inflows"
        self._parameters = parameters
        p = as_param_dict(parameters)
        self._hru_segment = np.asarray(p["hru_segment"], dtype=np.int64) - 1
        self._nhru = int(self._hru_segment.shape[0])
        self._nsegment = int(np.asarray(p["tosegment"]).shape[0])
        self._inputs = {"sroff_vol": sroff_vol, "ssres_flow_vol": ssres_flow_vol,
                        "gwres_flow_vol": gwres_flow_vol}
        self.control = control if control is not None else getattr(sroff_vol, "control", None)
        self._data = None
        self._current_value = np.full(self._nsegment, np.nan)
        self._it = -1

    @staticmethod
    def lateral(hru_segment0, nsegment, vols, s_per_time=86400.0):
        """[ndays, nsegment]: (sroff_vol + ssres_flow_vol + gwres_flow_vol) / s, added to
        each segment in ascending HRU order (:448-472)."""
        total = (0 + vols[0] + vols[1] + vols[2]) / s_per_time
        out = np.zeros((total.shape[0], nsegment))
        for ihru in np.where(hru_segment0 >= 0)[0]:
            out[:, hru_segment0[ihru]] += total[:, ihru]
        return out

    @property
    def data(self):
        if self._data is None:
            vols = []
            for k, v in self._inputs.items():
                if isinstance(v, (str, pl.Path)):
                    from .netcdf_in import read_time_variable

                    _, a = read_time_variable(pl.Path(v), k)
                else:
                    a = np.asarray(v.data if hasattr(v, "data") and not isinstance(v, np.ndarray) else v)
                vols.append(np.asarray(a, dtype=np.float64))
            self._data = self.lateral(self._hru_segment, self._nsegment, vols)
        return self._data

    def advance(self):
        self._it = self.control.itime_step if self.control is not None else self._it + 1
        self._current_value[:] = self.data[self._it]

    @property
    def current(self):
        return self._current_value


class HruNodeFlowExchange:
    """hydrology/prms_channel_flow_graph.py:501-620: maps the three HRU outflow
    volumes of the PRMS column onto the lateral inflows of a FlowGraph's nodes
    inside a Model (HRUs without a segment go to `sinks`; nodes that are not PRMS
    segments get nothing).  Global-basis budget.  Under `Model` it is evaluated
    for a whole block of days at once (`block()`); it has no state."""

    _all_time = False
    VARIABLES = ("inflows", "inflows_vol", "sinks", "sinks_vol")

    def __init__(self, control, discretization, parameters, sroff_vol=None, ssres_flow_vol=None,
                 gwres_flow_vol=None, budget_type="defer", verbose: bool = None, name: str = None):
        self.name = name or "HruNodeFlowExchange"
        self.control = control
        self.params = parameters
        p = {}
        for src in (discretization, parameters):
            if src is not None:
                p.update(as_param_dict(src))
        for k in self.get_parameters():
            if k not in p:
                raise ValueError(f"{self.name}: parameters missing: ['{k}']")  # base/process.py:233-237
        self._param_dict = p
        self._hru_segment = np.asarray(p["hru_segment"], dtype=np.int64) - 1
        self.nhru = int(self._hru_segment.shape[0])
        self.nnodes = int(np.asarray(p["node_coord"]).shape[0])
        self._input_variables_dict = {"sroff_vol": sroff_vol, "ssres_flow_vol": ssres_flow_vol,
                                      "gwres_flow_vol": gwres_flow_vol}
        self._vars = {"inflows": np.full(self.nnodes, np.nan), "inflows_vol": np.full(self.nnodes, np.nan),
                      "sinks": np.full(self.nhru, np.nan), "sinks_vol": np.full(self.nhru, np.nan)}
        if budget_type == "defer":
            budget_type = control.options.get("budget_type", "warn") if hasattr(control, "options") else "warn"
        self._budget_type = budget_type
        self._model = None
        self._netcdf_initialized = False
        self.budget = None
        if budget_type is not None:
            self.budget = Budget(control, self, self.get_mass_budget_terms(), budget_type, basis="global",
                                 description=self.name)

    @staticmethod
    def get_dimensions() -> tuple:
        return ("nhru", "nnodes")

    @staticmethod
    def get_parameters() -> tuple:
        return ("hru_segment", "node_coord")

    @staticmethod
    def get_inputs() -> tuple:
        return ("sroff_vol", "ssres_flow_vol", "gwres_flow_vol")

    @staticmethod
    def get_init_values() -> dict:
        return {k: np.nan for k in HruNodeFlowExchange.VARIABLES}

    @classmethod
    def get_variables(cls) -> tuple:
        return tuple(cls.VARIABLES)

    @staticmethod
    def get_mass_budget_terms():
        return {"inputs": ["sroff_vol", "ssres_flow_vol", "gwres_flow_vol"],
                "outputs": ["inflows_vol", "sinks_vol"], "storage_changes": []}

    @staticmethod
    def get_restart_variables() -> tuple:
        return ()

    def block(self, sroff_vol, ssres_flow_vol, gwres_flow_vol) -> dict:
        """The exchange for [ndays, nhru] input volumes (:593-618), day by day the
        arithmetic of _calculate: sum of the three / seconds per step, added per
        node in ascending HRU order."""
        s_per_time = float(self.control.time_step / np.timedelta64(1, "s"))
        total = (0 + np.asarray(sroff_vol, dtype=np.float64) + np.asarray(ssres_flow_vol, dtype=np.float64)
                 + np.asarray(gwres_flow_vol, dtype=np.float64)) / s_per_time
        nd = total.shape[0]
        inflows = np.zeros((nd, self.nnodes))
        sinks = np.zeros((nd, self.nhru))
        for ihru in range(self.nhru):
            iseg = self._hru_segment[ihru]
            if iseg < 0:
                sinks[:, ihru] += total[:, ihru]
            else:
                inflows[:, iseg] += total[:, ihru]
        return {"inflows": inflows, "inflows_vol": inflows * s_per_time, "sinks": sinks,
                "sinks_vol": sinks * s_per_time}

    def __getitem__(self, key):
        if key in self._vars:
            return self._vars[key]
        if key in self._input_variables_dict and self._model is not None:
            for proc in self._model.processes.values():
                if proc is not self and key in getattr(proc, "_vars", {}):
                    return proc._vars[key]
        if key in self._param_dict:
            return self._param_dict[key]
        return getattr(self, key)

    def __getattr__(self, key):
        v = self.__dict__.get("_vars")
        if v is not None and key in v:
            return v[key]
        raise AttributeError(key)

    def advance(self):
        pass

    def calculate(self, time_length: float = 1.0, **kwargs):
        if self._model is None:
            raise RuntimeError("HruNodeFlowExchange computes inside a Model")

    def output(self):
        pass

    def finalize(self):
        pass


def prms_channel_flow_graph_to_model_dict(model_dict: dict, prms_channel_dis, prms_channel_dis_name: str,
                                          prms_channel_params, new_nodes_maker_dict: dict,
                                          new_nodes_maker_names: list, new_nodes_maker_indices: list,
                                          new_nodes_maker_ids: list, new_nodes_flow_to_nhm_seg: list,
                                          addtl_output_vars: list = None, graph_budget_type="defer",
                                          allow_disconnected_nodes: bool = False,
                                          type_check_nodes: bool = False) -> dict:
    """hydrology/prms_channel_flow_graph.py:746-895: add an exchange
    (`inflow_exchange`) and a FlowGraph (`prms_channel_flow_graph`) of
    PRMSChannel's segments + new nodes to a model_dict whose processes produce
    sroff_vol / ssres_flow_vol / gwres_flow_vol; the new nodes get no lateral
    inflow."""
    params_fg, makers = _build_flow_graph_inputs(
        prms_channel_dis, prms_channel_params, new_nodes_maker_dict, new_nodes_maker_names,
        new_nodes_maker_indices, new_nodes_maker_ids, new_nodes_flow_to_nhm_seg)
    nnodes = params_fg.dims["nnodes"]
    pc = as_param_dict(prms_channel_params)
    ex_data = dict(pc)
    ex_data["node_coord"] = np.arange(nnodes)
    params_exchange = Parameters(dims={"nnodes": nnodes, "nhru": int(np.asarray(ex_data["hru_segment"]).shape[0])},
                                 coords={"node_coord": ex_data.pop("node_coord")}, data_vars=ex_data)
    graph_dict = {
        "inflow_exchange": {"class": HruNodeFlowExchange, "parameters": params_exchange,
                            "dis": prms_channel_dis_name},
        "prms_channel_flow_graph": {"class": FlowGraph, "node_maker_dict": makers, "parameters": params_fg,
                                    "dis": None, "budget_type": graph_budget_type,
                                    "addtl_output_vars": addtl_output_vars,
                                    "allow_disconnected_nodes": allow_disconnected_nodes},
    }
    out = dict(model_dict)
    out["model_order"] = list(model_dict["model_order"]) + list(graph_dict.keys())
    out.update(graph_dict)
    return out


def _build_flow_graph_inputs(prms_channel_dis, prms_channel_params, new_nodes_maker_dict: dict,
                             new_nodes_maker_names: list, new_nodes_maker_indices: list,
                             new_nodes_maker_ids: list, new_nodes_flow_to_nhm_seg: list):
    """hydrology/prms_channel_flow_graph.py:985-1100: the collated FlowGraph
    parameters for PRMSChannel's segments + new nodes placed above given nhm_seg
    (or, with a non-positive entry -k, above new node k, in series)."""
    assert not any(isinstance(v, PRMSChannelFlowNodeMaker) for v in new_nodes_maker_dict.values())
    msg = "Inconsistency in collated inputs"
    assert len(new_nodes_maker_names) == len(new_nodes_maker_indices), msg
    assert len(new_nodes_maker_names) == len(new_nodes_flow_to_nhm_seg), msg
    assert len(new_nodes_flow_to_nhm_seg) == len(np.unique(new_nodes_flow_to_nhm_seg)), (
        "Cant have more than one new node flowing to an existing or new node.")
    dis, par = as_param_dict(prms_channel_dis), as_param_dict(prms_channel_params)
    tosegment = np.asarray(dis["tosegment"] if "tosegment" in dis else par["tosegment"], dtype=np.int64) - 1
    nhm_seg = np.asarray(dis["nhm_seg"] if "nhm_seg" in dis else par["nhm_seg"])
    nseg = int(tosegment.shape[0])
    nnew = len(new_nodes_maker_names)
    nnodes = nseg + nnew
    node_maker_name = np.array(["prms_channel"] * nseg + list(new_nodes_maker_names))
    node_maker_index = np.array(list(range(nseg)) + list(new_nodes_maker_indices), dtype=np.int64)
    node_maker_id = np.append(nhm_seg, np.array(new_nodes_maker_ids))
    to_graph_index = np.zeros(nnodes, dtype=np.int64)
    to_graph_index[:nseg] = tosegment
    pending, placed = {}, {}
    for ii, target in enumerate(new_nodes_flow_to_nhm_seg):
        if target < 1:
            pending[ii] = -1 * target
            continue
        above = np.where(nhm_seg == target)[0]
        if not len(above):
            raise ValueError(f"nhm_seg {target} is not in the discretization")
        below = np.where(tosegment == above[0])[0]
        to_graph_index[nseg + ii] = above[0]
        to_graph_index[below] = nseg + ii
        placed[ii] = nseg + ii
    for _ in range(len(pending)):
        for ii in list(pending):
            tgt = pending[ii]
            if tgt not in placed:
                continue
            flows_to = placed[tgt]
            flows_from = np.where(to_graph_index == flows_to)[0]
            to_graph_index[nseg + ii] = flows_to
            to_graph_index[flows_from] = nseg + ii
            placed[ii] = nseg + ii
            del pending[ii]
    if pending:
        raise ValueError("Unable to connect some new nodes in-series.")
    params_flow_graph = Parameters(
        dims={"nnodes": nnodes}, coords={"node_coord": np.arange(nnodes)},
        data_vars={"node_maker_name": node_maker_name, "node_maker_index": node_maker_index,
                   "node_maker_id": node_maker_id, "to_graph_index": to_graph_index},
        metadata={k: {"dims": ["nnodes"]} for k in ("node_coord", "node_maker_name", "node_maker_index",
                                                      "node_maker_id", "to_graph_index")})
    node_maker_dict = {"prms_channel": PRMSChannelFlowNodeMaker(prms_channel_dis, prms_channel_params)}
    node_maker_dict.update(new_nodes_maker_dict)
    return params_flow_graph, node_maker_dict


def prms_channel_flow_graph_postprocess(control, input_dir, prms_channel_dis, prms_channel_params,
                                        new_nodes_maker_dict: dict, new_nodes_maker_names: list,
                                        new_nodes_maker_indices: list, new_nodes_maker_ids: list,
                                        new_nodes_flow_to_nhm_seg: list, addtl_output_vars: list = None,
                                        allow_disconnected_nodes: bool = False, budget_type="defer",
                                        type_check_nodes: bool = False, inputs: dict = None) -> FlowGraph:
    """hydrology/prms_channel_flow_graph.py:575-720: a FlowGraph of PRMSChannel's
    segments + new nodes, forced by the sroff_vol / ssres_flow_vol / gwres_flow_vol
    files of `input_dir` (or the arrays of `inputs`); the new nodes get no lateral
    inflow."""
    params_fg, makers = _build_flow_graph_inputs(
        prms_channel_dis, prms_channel_params, new_nodes_maker_dict, new_nodes_maker_names,
        new_nodes_maker_indices, new_nodes_maker_ids, new_nodes_flow_to_nhm_seg)
    src = {}
    for key in ("sroff_vol", "ssres_flow_vol", "gwres_flow_vol"):
        src[key] = inputs[key] if inputs is not None else pl.Path(input_dir) / f"{key}.nc"
    seg_inflows = HruSegmentFlowAdapter(prms_channel_params, control=control, **src).data
    nt = control.n_times
    lat = np.zeros((nt, params_fg.dims["nnodes"]))
    lat[:, :seg_inflows.shape[1]] = seg_inflows[:nt]
    return FlowGraph(control=control, discretization=prms_channel_dis, parameters=params_fg, inflows=lat,
                     node_maker_dict=makers, addtl_output_vars=addtl_output_vars, budget_type=budget_type,
                     type_check_nodes=type_check_nodes, allow_disconnected_nodes=allow_disconnected_nodes)

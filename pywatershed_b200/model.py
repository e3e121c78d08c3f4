"""Model: the time loop over the NHM process chain, behind the reference's
Model API (pywatershed/base/model.py:31-755): same constructor styles, same
`run / advance / calculate / output / finalize / initialize_netcdf`, processes
reachable as `model.processes[ClassName][var]`.

Differences that stay invisible to a caller:
  * The chain is computed in *time blocks*: on the first `calculate()` of a
    block the whole block runs on the GPU (pws_run_host: forcings staged
    H2D in sub-blocks, outputs drained D2H asynchronously); the following days
    are served from the drained block.  Public variables keep stable array
    identities (the reference's tests check `id(proc.var)`), refreshed in
    place for the current day.
  * `Model.run()` uses the fast path: whole blocks go to netCDF with
    `var[t0:t1, :] = block` instead of one write per variable per day.
"""
from __future__ import annotations

import pathlib as pl
from warnings import warn

import numpy as np

from . import processes as P
from ._process_meta import PROCESS_META, VARIABLE_META
from .control import Control
from .engine import Engine, EtEngine, Registry
from .parameters import Parameters, PrmsParameters, as_param_dict
from .prms_files import read_cbh

COLUMN_CHAIN = ("PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoff",
                "PRMSSoilzone", "PRMSGroundwater")
FORCINGS = ("prcp", "tmax", "tmin")
BUDGET_INDEX = {"PRMSCanopy": 0, "PRMSSnow": 1, "PRMSRunoff": 2, "PRMSSoilzone": 3,
                "PRMSGroundwater": 4, "PRMSChannel": 5}


class Model:
    def __init__(self, process_list_or_model_dict, control: Control = None,
                 parameters=None, find_input_files: bool = True, write_control=False):
        self.control = control
        self.parameters = parameters
        self._find_input_files = find_input_files
        self._netcdf_initialized = False
        self._netcdf = None
        proc_classes, proc_kwargs = self._parse_spec(process_list_or_model_dict)
        names = [c.__name__ for c in proc_classes]
        order = [n for n in P.PROCESS_ORDER_NHM if n in names]  # base/model.py:294-317
        unknown = [n for n in names if n not in P.PROCESS_ORDER_NHM]
        if unknown:
            raise NotImplementedError(f"processes outside the NHM chain: {unknown}")
        if "input_dir" not in self.control.options and find_input_files:
            # base/model.py:509-518
            raise ValueError("Control object must have an input_dir specified in its options")
        self._mode = self._classify(order)
        self.processes = {}
        for n in order:
            cls = P.PROCESS_CLASSES[n]
            kw = dict(proc_kwargs.get(n, {}))
            dis = kw.pop("dis", None)
            par = kw.pop("parameters", self.parameters)
            proc = cls(self.control, dis, par, **kw)
            proc._model = self
            self.processes[n] = proc
        self.process_order = order
        self._param_dict = {}
        for n in order:
            self._param_dict.update(self.processes[n]._param_dict)
        self._build_tail(proc_kwargs)
        self._engine = None
        self._forcings = None
        self._block = None
        self._block_t0 = 0
        self._block_n = 0
        self._block_viol = None
        self._calculated_step = -1
        self._init_runtime()
        # True when a single process is stepped by its caller, who advances
        # the Control (autotest/test_prms_snow.py:122-132)
        self._control_is_external = False
        # days per host-side block (what one engine call brings back); None = sized
        # from the domain when the engine exists (about 2 GB of results, <= 366 days)
        self.block_days = self.control.options.get("block_days") or None
        if write_control:
            self._write_control(write_control)

    def _build_tail(self, proc_kwargs):
        """A FlowGraph behind the PRMS column (model_dict entries made by
        flow_graph.prms_channel_flow_graph_to_model_dict, the reference's
        nhm + obsin / starfit configurations): `inflow_exchange` maps the column's
        three outflow volumes to node inflows, the FlowGraph routes them.  Both are
        evaluated per block of days right after the column's block."""
        self._tail = []
        spec = getattr(self, "_tail_spec", [])
        if not spec:
            return
        from .flow_graph import FlowGraph, HruNodeFlowExchange

        if self._mode != "chain":
            raise NotImplementedError("a FlowGraph in a Model follows the full PRMS column without PRMSChannel")
        exchange = graph = None
        for key, cls, kw in spec:
            kw = dict(kw)
            dis = kw.pop("dis", None)
            par = kw.pop("parameters", self.parameters)
            if issubclass(cls, HruNodeFlowExchange):
                exchange = cls(self.control, dis, par, name=key, **kw)
                obj = exchange
            else:
                if exchange is None:
                    raise ValueError("model_dict: the FlowGraph needs an inflow exchange before it")
                kw.pop("addtl_output_vars", None)
                nn = int(np.asarray(as_param_dict(par)["to_graph_index"]).shape[0])
                graph = cls(self.control, dis, par, np.zeros((self.control.n_times, nn)), device=None, **kw)
                graph.name = key
                graph._inflow_valid = 0  # days of inflows the column has produced so far
                obj = graph
            obj._model = self
            self.processes[key] = obj
            self.process_order.append(key)
            self._tail.append(obj)
        if graph is None:
            raise ValueError("model_dict: an inflow exchange without a FlowGraph")
        self._tail_exchange, self._tail_graph = exchange, graph

    def _tail_block(self, blk, t0, n):
        """The exchange and the FlowGraph for the block the column just computed."""
        ex, fg = self._tail_exchange, self._tail_graph
        blk.update(ex.block(blk["sroff_vol"], blk["ssres_flow_vol"], blk["gwres_flow_vol"]))
        fg._ensure_engine()
        fg._inflow_data[t0:t0 + n] = blk["inflows"]
        if fg._engine.day != t0:
            raise RuntimeError(f"FlowGraph engine is at day {fg._engine.day}, block starts at {t0}")
        obs = fg._obs[t0:t0 + n] if fg._obs is not None else None
        res = fg._engine.run_flow_graph(np.ascontiguousarray(blk["inflows"]), obs)
        res["node_negative_sink_source"] = -1 * res["node_sink_source"]
        for k in fg.VARIABLES:
            if k in res:
                blk[k] = res[k]
        # global-basis budgets of the two, for every day of the block at once
        # (base/budget.py:388-416); a failing day raises / warns like Budget.calculate
        for obj in (ex, fg):
            b = obj.budget
            if b is None:
                continue
            tot = {c: sum(np.asarray(blk[v], dtype=np.float64).sum(axis=1) for v in b.terms[c])
                   if b.terms[c] else 0.0 for c in b.components}
            ok = np.isclose(tot["inputs"], tot["outputs"] + tot["storage_changes"], rtol=b.rtol, atol=b.atol)
            if not np.all(ok):
                day = t0 + int(np.argmin(ok))
                msg = (f"The global flux balance not equal to the change in global storage: {b.description} "
                       f"(day {day})")
                if b.imbalance_fatal:
                    raise ValueError(msg)
                warn(msg, UserWarning)

    def _init_runtime(self):
        self._engine_kwargs = None
        self._pending_writes = []      # futures of the netCDF writer thread (<= 2 blocks in flight)
        self._writer = None
        self._buf_slot = 0             # result buffers alternate so a block can be written while the next computes
        self._device_acc_on = None     # what the engine's "accumulate" option was last set to
        self._device_acc_cache = (None, {})
        self._days_on_device = 0       # days computed so far (rows an all-time array may be missing)

    # ---------------------------------------------------------------- spec
    def _parse_spec(self, spec):
        if isinstance(spec, dict):
            # model_dict style (base/model.py:56-120): {'control':..., 'dis_x':...,
            # 'proc': {'class': C, 'parameters': p, 'dis': 'dis_x'}, 'model_order': [...]}
            md = spec
            ctl = [v for v in md.values() if isinstance(v, Control) or type(v).__name__ == "Control"]
            if ctl:
                self.control = ctl[0]
            classes, kwargs = [], {}
            keys = md.get("model_order", [k for k, v in md.items() if isinstance(v, dict) and "class" in v])
            from .flow_graph import FlowGraph, HruNodeFlowExchange

            self._tail_spec = []
            for k in keys:
                d = md[k]
                cls = d["class"]
                if isinstance(cls, type) and issubclass(cls, (FlowGraph, HruNodeFlowExchange)):
                    kw = {kk: vv for kk, vv in d.items() if kk != "class"}
                    if isinstance(kw.get("dis"), str):
                        kw["dis"] = md[kw["dis"]]
                    self._tail_spec.append((k, cls, kw))
                    continue
                cls = P.PROCESS_CLASSES[cls.__name__ if not isinstance(cls, str) else cls]
                classes.append(cls)
                kw = {kk: vv for kk, vv in d.items() if kk not in ("class",)}
                if isinstance(kw.get("dis"), str):
                    kw["dis"] = md[kw["dis"]]
                kwargs[cls.__name__] = kw
            if self.control is None:
                raise ValueError("model_dict has no Control")
            return classes, kwargs
        classes = []
        for c in spec:
            name = c if isinstance(c, str) else c.__name__
            if name not in P.PROCESS_CLASSES:
                raise NotImplementedError(f"process {name} is outside the NHM chain")
            classes.append(P.PROCESS_CLASSES[name])  # a reference class maps to ours by name
        if self.control is None or self.parameters is None:
            raise ValueError("a process list needs control= and parameters=")
        return classes, {}

    @staticmethod
    def _classify(order):
        """"chain" / "chain+channel": the whole NHM column on the fused kernel;
        "channel": routing alone; "sub": any other subset (a single process or
        a partial chain), the missing processes replaced by input arrays."""
        if not order:
            raise ValueError("no processes")
        slots = [P.base_name(n) for n in order]
        if len(set(slots)) != len(slots):
            raise ValueError(f"two variants of the same process in one model: {order}")
        if "PRMSEt" in slots:
            if slots != ["PRMSEt"]:
                # its variables hru_actet / unused_potet are PRMSSoilzone's too: the two do
                # not live in one model (no NHM configuration of the reference uses PRMSEt)
                raise NotImplementedError("PRMSEt runs on its own, fed from arrays or files")
            return "et"
        has_col = [n in slots for n in COLUMN_CHAIN]
        if all(has_col):
            return "chain+channel" if "PRMSChannel" in order else "chain"
        if order == ["PRMSChannel"]:
            return "channel"
        return "sub"

    def _external_inputs(self):
        """Inputs of the model's processes that no process of the model
        produces (base/model.py:_solve_inputs).  The soltab tables are always
        computed from the parameters."""
        produced = set()
        for proc in self.processes.values():
            produced.update(proc.get_variables())
        ext = []
        for proc in self.processes.values():
            for k in proc.get_inputs():
                if k not in produced and not k.startswith("soltab") and k not in ext:
                    ext.append(k)
        return ext

    @classmethod
    def _around(cls, proc):
        """The private one-process model behind a process that is stepped on
        its own (control.advance(); proc.advance(); proc.calculate(dt))."""
        self = cls.__new__(cls)
        self.control = proc.control
        self.parameters = proc.params
        self._find_input_files = False
        self._netcdf_initialized = False
        self._netcdf = None
        self._mode = cls._classify([proc.name])
        self.processes = {proc.name: proc}
        self.process_order = [proc.name]
        self._param_dict = dict(proc._param_dict)
        self._engine = None
        self._forcings = None
        self._block = None
        self._block_t0 = self._block_n = 0
        self._block_viol = None
        self._calculated_step = -1
        self._tail = []
        self._init_runtime()
        self._control_is_external = True
        # days per host-side block (what one engine call brings back); None = sized
        # from the domain when the engine exists (about 2 GB of results, <= 366 days)
        self.block_days = self.control.options.get("block_days") or None
        return self

    @staticmethod
    def model_dict_from_yaml(yaml_file) -> dict:
        """A model dictionary from a yaml file (base/model.py:541-593): string
        values ending in .yaml/.yml are Control files, .nc are discretization
        Parameters; dict values with a "class" key are processes whose
        "parameters" is a parameter file.  Paths are relative to the yaml file.
        Beyond the reference, a parameter / discretization file may also be a
        PRMS ASCII parameter file (.param)."""
        import yaml

        yaml_file = pl.Path(yaml_file)

        def rel(path):
            path = pl.Path(path)
            return path if path.is_absolute() else (yaml_file.parent / path).resolve()

        def params(path):
            path = rel(path)
            if path.suffix == ".param":
                return PrmsParameters.load(path)
            if path.suffix == ".nc":
                return Parameters.from_netcdf(path)
            raise ValueError("Unsupported file extension for control (.yml/.yaml) and parameter "
                             "(.nc) file paths in model yaml file")

        with yaml_file.open("r") as fh:
            md = yaml.load(fh, Loader=yaml.Loader)
        for key, val in md.items():
            if isinstance(val, str):
                if val.endswith((".yml", ".yaml")):
                    md[key] = Control.from_yaml(rel(val))
                else:
                    md[key] = params(val)
            elif isinstance(val, dict) and "class" in val:
                if val["class"] not in P.PROCESS_CLASSES:
                    raise NotImplementedError(f"process {val['class']} is outside the NHM chain")
                val["class"] = P.PROCESS_CLASSES[val["class"]]
                val["parameters"] = params(val["parameters"])
        return md

    @classmethod
    def from_yaml(cls, yaml_file):
        """Model from a yaml file (base/model.py:595-664)."""
        return cls(cls.model_dict_from_yaml(yaml_file))

    # ---------------------------------------------------------------- inputs
    def _load_input(self, name, nunits, keep_f32=False):
        """An 'adaptable' (base/adapter.py:143-190): ndarray [ntime, n], path to a CBH
        ASCII / netCDF file, or None -> input_dir/<name>.{nc,cbh}.  float64, or with
        `keep_f32` the float32 of the source (the reference's forcing files): the
        engine uploads those as they are and widens them on the device."""
        nt = self.control.n_times
        src = None
        for proc in self.processes.values():
            v = proc._input_variables_dict.get(name)
            if v is not None:
                src = v
        if src is None:
            if not self._find_input_files:
                raise ValueError(f"input '{name}' was not provided and find_input_files=False")
            d = pl.Path(self.control.options["input_dir"])
            for cand in (d / f"{name}.nc", d / f"{name}.cbh"):
                if cand.exists():
                    src = cand
                    break
            if src is None:
                raise FileNotFoundError(f"no {name}.nc / {name}.cbh in {d}")
        if hasattr(src, "data") and not isinstance(src, np.ndarray):  # adapter-like
            arr, times = np.asarray(src.data), getattr(getattr(src, "_nc_read", None), "times", None)
        elif isinstance(src, (str, pl.Path)):
            src = pl.Path(src)
            if src.suffix == ".cbh":
                times, arr = read_cbh(src, nunits)
            else:
                times, arr = _read_netcdf_var(src, name, keep_f32=keep_f32)
        else:
            arr, times = np.asarray(src, dtype=np.float64), None
        if times is not None:
            times = np.asarray(times).astype("datetime64[s]")
            i0 = np.where(times == self.control.start_time.astype("datetime64[s]"))[0]
            if not len(i0):
                raise ValueError("Control start_time is not in the input data time")
            arr = arr[i0[0]:]
        if arr.ndim != 2 or arr.shape[1] != nunits or arr.shape[0] < nt:
            raise ValueError(f"input '{name}' must be [>= {nt}, {nunits}], got {arr.shape}")
        dt = np.float32 if keep_f32 and arr.dtype == np.float32 else np.float64
        return np.ascontiguousarray(arr[:nt], dtype=dt)

    def _ensure_engine(self):
        if self._engine is not None:
            return
        pd = self._param_dict
        # dprst_flag lives on Runoff / Soilzone / Groundwater only (base/process.py:339-358:
        # an option exists where the constructor names it); their NoDprst variants say False
        flags = {bool(self.processes[n]._dprst_flag) for n in self.process_order
                 if getattr(self.processes[n], "_dprst_flag", None) is not None}
        if len(flags) > 1:
            raise ValueError("dprst_flag must be the same for every process of the model "
                             "(mixing PRMSRunoff with PRMSSoilzoneNoDprst, ...)")
        dprst = flags.pop() if flags else True
        # adjust_parameters: Soilzone's decides whether the device edits the soil
        # parameters / initial storages (prms_soilzone.py:361-463); Channel's only warns
        # or raises (its slope edit comes after the velocity, prms_channel.py:234-260).
        # The checks themselves run here, on the parameters, with the reference's messages.
        sz_mode, ch_mode = "no", "no"
        for n in self.process_order:
            if P.base_name(n) == "PRMSSoilzone":
                sz_mode = self.processes[n]._adjust_parameters or "warn"
            if n == "PRMSChannel":
                ch_mode = self.processes[n]._adjust_parameters or "warn"
        P.check_adjusted_parameters(pd, sz_mode if any(
            P.base_name(n) == "PRMSSoilzone" for n in self.process_order) else None,
            ch_mode if "PRMSChannel" in self.process_order else None, dprst)
        adj = sz_mode != "no"
        start = self.control.start_time
        dev = int(self.control.options.get("device", 0) or 0)
        if self._mode == "et":
            self._engine = EtEngine(pd, start, device=dev)
            self._forcings = {k: self._load_input(k, self._engine.nhru) for k in EtEngine.INPUTS}
            self._engine_kwargs = {}
            self._full_selection = (list(EtEngine.VARS), [])
            self._selection_is_full = True
            self._block_days_user = self.block_days
            if self.block_days is None:
                self.block_days = self._days_for(EtEngine.INPUTS + EtEngine.VARS, [], 366)
            self.block_days = int(self.block_days)
            return
        if self._mode == "channel":
            # the HRU column is not run; its tables get stand-in values
            kw = dict(device=dev, channel=True, fill_missing=True)
            self._engine = Engine(pd, start, **kw)
            nh = self._engine.nhru
            self._forcings = {k: self._load_input(k, nh) for k in
                              ("sroff_vol", "ssres_flow_vol", "gwres_flow_vol")}
        elif self._mode == "sub":
            kw = dict(device=dev, channel="PRMSChannel" in self.processes, adjust_parameters=adj,
                      fill_missing=True, dprst_flag=bool(dprst))
            self._engine = Engine(pd, start, **kw)
            nh = self._engine.nhru
            self._forcings = {k: self._load_input(k, nh) for k in self._external_inputs()}
        else:
            kw = dict(device=dev, channel=self._mode == "chain+channel", adjust_parameters=adj,
                      dprst_flag=bool(dprst), fill_missing=not dprst)
            self._engine = Engine(pd, start, **kw)
            nh = self._engine.nhru
            self._forcings = {k: self._page_locked(self._load_input(k, nh, keep_f32=True))
                              for k in FORCINGS}
        self._engine_kwargs = kw
        reg = self._engine.reg
        hv = list(reg.hru_vars) if self._mode != "channel" else [
            "channel_sroff_vol", "channel_ssres_flow_vol", "channel_gwres_flow_vol"]
        sv = list(reg.seg_vars) if self._engine.nseg else []
        if self._mode == "sub":
            mine = set()
            for proc in self.processes.values():
                mine.update(proc.get_variables())
            hv = [v for v in reg.hru_vars if v in mine]
            sv = sv if "PRMSChannel" in self.processes else []
        self._full_selection = (list(hv), list(sv))
        self._selection_is_full = True
        self._block_days_user = self.block_days
        if self.block_days is None:
            self.block_days = self._days_for(hv, sv, 366)
        self.block_days = int(self.block_days)
        if self._mode != "channel":
            self._engine.select_outputs(hv, sv)
        else:
            self._engine.select_outputs([], sv)
        # all-time variables are materialised when somebody looks (processes.TimeseriesArray)
        if "PRMSSolarGeometry" in self.processes and self._mode != "channel":
            for ts in self.processes["PRMSSolarGeometry"]._vars.values():
                ts._loader = self._load_soltab
                ts.mark_stale()
        if "PRMSAtmosphere" in self.processes:
            for ts in self.processes["PRMSAtmosphere"]._vars.values():
                if self._mode in ("chain", "chain+channel"):
                    ts._loader = self._replay_all_time
                else:  # a partial chain: every block is complete, the rows are kept as they come
                    ts.data
        # budget terms are summed on the device during run() (pws_select_accumulations)
        if self._mode in ("chain", "chain+channel"):
            th, tsg = self._budget_terms()
            self._engine.select_accumulations(th, tsg)
            self._set_device_acc(False)

    def _page_locked(self, arr):
        """The forcings of a run in page-locked memory (from the pool of
        engine._PINNED, handed back when the model goes away): pws_run_host then
        uploads them in place instead of staging every block through pinned slots."""
        from .engine import _PINNED

        buf = _PINNED.take(arr.nbytes)
        self.__dict__.setdefault("_pinned", []).append(buf)
        out = buf.numpy()[:arr.nbytes].view(arr.dtype).reshape(arr.shape)
        np.copyto(out, arr)
        return out

    def __del__(self):
        try:
            from .engine import _PINNED

            for buf in self.__dict__.pop("_pinned", []):
                _PINNED.give(buf)
        except Exception:
            pass

    def _days_for(self, hv, sv, cap):
        """Days per host-side block so that one block of these variables is ~2 GB."""
        per_day = 8.0 * (len(hv) * self._engine.nhru + len(sv) * self._engine.nseg) + 1.0
        return int(max(8, min(cap, 2.0e9 / per_day)))

    def _budget_terms(self):
        """(HRU variables, segment variables) that are a term of some budget of the model."""
        reg = self._engine.reg
        th, tsg = [], []
        for proc in self.processes.values():
            b = getattr(proc, "budget", None)
            if b is None:
                continue
            for comp in b.components:
                for v in b.terms[comp]:
                    if v in reg.hru_index and v not in th:
                        th.append(v)
                    elif v in reg.seg_index and v not in tsg:
                        tsg.append(v)
        return th, tsg

    def _set_device_acc(self, on: bool):
        if self._device_acc_on is not on and self._mode in ("chain", "chain+channel"):
            self._engine.set_accumulate(on)
            self._device_acc_on = on

    def _device_accumulations(self) -> dict:
        """Budget-term sums held on the device (days computed by run())."""
        eng = self._engine
        if eng is None or not getattr(eng, "acc_hru", None):
            return {}
        if self._device_acc_cache[0] != eng.day:
            self._device_acc_cache = (eng.day, eng.accumulations())
        return self._device_acc_cache[1]

    def _load_soltab(self, ts):
        sg = self.processes["PRMSSolarGeometry"]
        for k, v in self._engine.soltab().items():
            t = sg._vars[k]
            if t._array is None:
                t._array = v
            else:
                t._array[:] = v
            t._stale = False

    def _replay_all_time(self, ts):
        """Fill `ts` (an all-time Atmosphere variable) for the days computed so
        far: a second handle replays them from the initial conditions with only
        that variable selected.  The reference holds these arrays from the first
        advance on (prms_atmosphere.py:264-285); here they cost nothing until read."""
        nd = self._days_on_device
        if nd <= 0:
            return
        eng = Engine(self._param_dict, self.control.start_time, **self._engine_kwargs)
        try:
            eng.select_outputs([ts.variable_name], [])
            f = self._forcings
            blk, _ = eng.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd], budgets=False)
            ts._array[:nd] = blk[ts.variable_name]
        finally:
            eng.close()

    # ---------------------------------------------------------------- blocks
    def _reduced_selection(self):
        """What run() has to bring back for a block nobody looks into: only what
        goes to netCDF (the variable files, and every budget term when budget
        files are written).  The budget accumulations are summed on the device,
        the all-time Atmosphere variables are recomputed if somebody reads them,
        and everything else is only observable for the last day of a run, i.e.
        in its last block."""
        hv_full, sv = self._full_selection
        want = set()
        if self._netcdf is not None:
            for _, _, f in self._netcdf.files:
                want.update(f.vars)
            if self._netcdf.budget_files:
                for proc in self.processes.values():
                    b = getattr(proc, "budget", None)
                    if b is not None:
                        for comp in b.components:
                            want.update(b.terms[comp])
        if self._tail:
            want.update(("sroff_vol", "ssres_flow_vol", "gwres_flow_vol"))
        return [v for v in hv_full if v in want], [v for v in sv if v in want]

    def _select(self, full: bool):
        if self._mode in ("channel", "et") or full == self._selection_is_full:
            return
        hv, sv = self._full_selection if full else self._reduced_selection()
        self._engine.select_outputs(hv, sv)
        self._selection_is_full = full

    def _compute_block(self, t0, n, reduced: bool = False, device_acc: bool = False):
        self._ensure_engine()
        eng = self._engine
        if eng.day != t0:
            raise RuntimeError(f"engine is at day {eng.day}, block starts at {t0}")
        self._select(full=not reduced)
        self._set_device_acc(device_acc)
        if self._mode == "et":
            blk, viol = eng.run({k: v[t0:t0 + n] for k, v in self._forcings.items()}), None
        elif self._mode == "channel":
            blk, viol = self._channel_block(t0, n)
        elif self._mode == "sub":
            eng.set_block_days(min(n, 16))
            blk, viol = eng.run_host_inputs({k: v[t0:t0 + n] for k, v in self._forcings.items()},
                                            [P.base_name(x) for x in self.process_order], ndays=n)
        else:
            eng.set_block_days(0)  # the library sizes its device blocks (~2 GB of results each)
            f = self._forcings
            # result buffers alternate between two sets: block b may still be on its
            # way to the netCDF files while block b+1 is computed
            self._buf_slot ^= 1
            blk, viol = eng.run_host(f["prcp"][t0:t0 + n], f["tmax"][t0:t0 + n], f["tmin"][t0:t0 + n],
                                     pinned_outputs=True, reuse_buffers=1 + self._buf_slot)
        if self._tail:
            self._tail_block(blk, t0, n)
        self._block, self._block_t0, self._block_n, self._block_viol = blk, t0, n, viol
        self._days_on_device = max(self._days_on_device, t0 + n)
        # all-time variables of Atmosphere: rows are kept when the array exists, else
        # recomputed on demand; budget terms: added to the host-side accumulations
        # unless the device sums them.  Plain copies and sums over [n, nhru] arrays --
        # numpy releases the GIL for them, so a few threads share the memory traffic.
        jobs = []
        if "PRMSAtmosphere" in self.processes:
            atm = self.processes["PRMSAtmosphere"]
            for var in atm.get_variables():
                ts = atm._vars[var]
                if var in blk and ts.allocated:
                    jobs.append((ts.set_rows, (t0, blk[var])))
                else:
                    ts.mark_stale()
        if not device_acc:
            dt = float(self.control.time_step / np.timedelta64(1, "D"))
            for name, proc in self.processes.items():
                if getattr(proc, "budget", None) is not None:
                    jobs.extend((proc.budget.accumulate_term, a) for a in proc.budget.block_terms(blk, dt))
        if device_acc and self._tail:
            on_device = set(getattr(eng, "acc_hru", ()) or ()) | set(getattr(eng, "acc_seg", ()) or ())
            dt = float(self.control.time_step / np.timedelta64(1, "D"))
            for obj in self._tail:
                if obj.budget is not None:
                    jobs.extend((obj.budget.accumulate_term, a) for a in obj.budget.block_terms(blk, dt)
                                if a[1] not in on_device)
        big = n * max(self._engine.nhru, 1) >= 1 << 18
        if big and len(jobs) > 1:
            list(_pool().map(lambda j: j[0](*j[1]), jobs))
        else:
            for fn, a in jobs:
                fn(*a)

    def _channel_block(self, t0, n):
        import torch

        eng = self._engine
        f = self._forcings
        hs = np.asarray(self._param_dict["hru_segment"]) - 1
        vols = [f[k][t0:t0 + n] for k in ("sroff_vol", "ssres_flow_vol", "gwres_flow_vol")]
        lat = np.zeros((n, eng.nseg))
        ch = [np.where(hs[None, :] >= 0, v, 0.0) for v in vols]
        contrib = (ch[0] + ch[1] + ch[2]) / 86400.0
        for h in np.where(hs >= 0)[0]:  # HRU order, as prms_channel.py:396-421
            lat[:, hs[h]] += contrib[:, h]
        dev = torch.device("cuda", eng.device)
        dl = torch.from_numpy(lat).to(dev)
        out = torch.empty((n, len(eng.sel_seg), eng.nseg), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        eng.run_channel_device(n, dl.data_ptr(), out.data_ptr())
        eng.day += n
        o = out.cpu().numpy()
        blk = {v: o[:, k, :] for k, v in enumerate(eng.sel_seg)}
        blk.update({"channel_sroff_vol": ch[0], "channel_ssres_flow_vol": ch[1],
                    "channel_gwres_flow_vol": ch[2]})
        return blk, None

    # ---------------------------------------------------------------- stepping
    def advance(self):
        """base/model.py:723-737"""
        if self._engine is None:
            self._ensure_engine()
        if not self._control_is_external:
            self.control.advance()
        if "PRMSSolarGeometry" in self.processes:
            for ts in self.processes["PRMSSolarGeometry"]._vars.values():
                ts.advance()
        if "PRMSAtmosphere" in self.processes:
            self._ensure_block(self.control.itime_step)
            self._serve_all_time_current()

    def _serve_all_time_current(self, lazy=False):
        """`.current` of the all-time Atmosphere variables for the control's day:
        from the block when it holds them; else from `.data` (which recomputes the
        array), unless `lazy` -- run() passes over blocks nobody looks into."""
        it = self.control.itime_step
        j = it - self._block_t0
        for var, ts in self.processes["PRMSAtmosphere"]._vars.items():
            if self._block is not None and var in self._block and 0 <= j < self._block_n:
                ts.set_current(self._block[var][j])
            elif not lazy or (ts.allocated and not ts._stale):
                ts.advance()

    def _ensure_block(self, it, limit=None, fast=False):
        """Make `_block` cover day `it`.  `fast` (run()): every block but the
        run's last few days is computed with the reduced selection and large
        blocks, its budget terms summed on the device; the last days form a block
        with every variable, because that is what a caller can look at afterwards."""
        if self._block is not None and self._block_t0 <= it < self._block_t0 + self._block_n:
            return
        self._ensure_engine()
        self._drain_writes(keep=1)  # the buffers of the block before the current one are reused now
        remaining = self.control.n_times - it
        if limit is not None:
            remaining = min(remaining, limit)
        n_full = min(self.block_days, remaining)
        if not fast or self._mode in ("channel", "sub", "et"):
            self._compute_block(it, n_full)
            return
        # after run() a caller sees ONE day, the last (the process arrays hold a day):
        # only that day is drained with every variable
        n_full = 1
        if remaining <= n_full:
            self._compute_block(it, remaining, reduced=False, device_acc=True)
            return
        hv, sv = self._reduced_selection()
        cap = self._block_days_user or self._days_for(hv, sv, 1024)
        n = min(cap, remaining - n_full)
        self._compute_block(it, n, reduced=True, device_acc=True)

    def calculate(self):
        """base/model.py:739-743: after this call every process shows day itime_step."""
        it = self.control.itime_step
        if it < 0:
            raise RuntimeError("advance() before calculate()")
        self._ensure_block(it)
        j = it - self._block_t0
        blk = self._block
        for name, proc in self.processes.items():
            if proc._all_time:
                continue
            for var in proc.get_variables():
                if var in blk:
                    proc._vars[var][:] = blk[var][j]
        if self._tail:
            self._tail_graph.inflows[:] = blk["inflows"][j]
        self._calculated_step = it
        for name, proc in self.processes.items():
            b = getattr(proc, "budget", None)
            if b is None:
                continue
            nv = None
            if self._block_viol is not None and P.base_name(name) in BUDGET_INDEX:
                nv = int(self._block_viol[j, BUDGET_INDEX[P.base_name(name)]])
            b.calculate(nv)

    def output(self):
        """base/model.py:745-749"""
        if self._netcdf is not None:
            it = self.control.itime_step
            self._netcdf.write_days(self, it, it + 1)

    def _write_async(self, t0, t1):
        """Days [t0, t1) of the current block go to the netCDF files on the
        writer thread while the next block is computed (its results land in the
        other set of buffers)."""
        if self._writer is None:
            from concurrent.futures import ThreadPoolExecutor

            self._writer = ThreadPoolExecutor(max_workers=1)
        blk, b0 = self._block, self._block_t0
        self._pending_writes.append(self._writer.submit(self._netcdf.write_days, self, t0, t1, blk, b0))

    def _drain_writes(self, keep=0):
        while len(self._pending_writes) > keep:
            self._pending_writes.pop(0).result()

    def finalize(self):
        self._drain_writes()
        if self._writer is not None:
            self._writer.shutdown()
            self._writer = None
        if self._netcdf is not None:
            self._netcdf.close()
            self._netcdf = None
        self._netcdf_initialized = False
        # the process arrays hold copies of the last day: the drained block and its
        # page-locked buffers can go back to the pool (a later step computes a new block)
        self._block, self._block_n = None, 0
        if self._engine is not None and hasattr(self._engine, "release_buffers"):
            self._engine.release_buffers()
        if self._tail:
            self._tail_graph.finalize()

    # ---------------------------------------------------------------- netcdf
    def initialize_netcdf(self, output_dir=None, separate_files: bool = None,
                          budget_args: dict = None, output_vars: list = None):
        """base/model.py:644-676"""
        from .netcdf_out import ModelNetcdfOutput

        if self._netcdf_initialized:
            warn("Model class previously initialized netcdf output", RuntimeWarning)
            return
        opts = self.control.options
        output_dir = output_dir if output_dir is not None else opts.get("netcdf_output_dir")
        if output_dir is None:
            raise ValueError("no netcdf output_dir")
        if separate_files is None:
            separate_files = opts.get("netcdf_output_separate_files", True)
        if output_vars is None:
            output_vars = opts.get("netcdf_output_var_names")
        self._ensure_engine()
        self._netcdf = ModelNetcdfOutput(self, pl.Path(output_dir), separate_files, output_vars,
                                         budget_args)
        self._netcdf_initialized = True
        for proc in self.processes.values():
            proc._netcdf_initialized = True

    def _initialize_netcdf_for(self, proc, output_dir=None, separate_files=None,
                               budget_args=None, output_vars=None, **kw):
        if not self._netcdf_initialized:
            self.initialize_netcdf(output_dir, separate_files, budget_args, output_vars)

    # ---------------------------------------------------------------- run
    def run(self, netcdf_dir=None, finalize: bool = True, n_time_steps: int = None,
            output_vars: list = None):
        """base/model.py:678-721, time-blocked."""
        if netcdf_dir:
            self.initialize_netcdf(netcdf_dir, output_vars=output_vars)
        if n_time_steps is None:
            n_time_steps = self.control.n_times - (self.control.itime_step + 1)
        self._ensure_engine()
        done = 0
        while done < n_time_steps:
            it = self.control.itime_step + 1
            self._ensure_block(it, limit=n_time_steps - done, fast=True)
            t1 = min(self._block_t0 + self._block_n, it + (n_time_steps - done))
            nd = t1 - it
            # budgets for the whole block: any violating day raises / warns as the
            # reference would on that day
            if self._block_viol is not None:
                v = self._block_viol[it - self._block_t0:t1 - self._block_t0]
                if v.any():
                    bad_day = int(np.argwhere(v.any(axis=1))[0, 0])
                    self._drain_writes()
                    if self._netcdf is not None and bad_day > 0:
                        self._netcdf.write_days(self, it, it + bad_day)
                    for _ in range(bad_day + 1):
                        self.advance()
                    self.calculate()  # raises ValueError / warns through Budget
                    if self._netcdf is not None:
                        self._netcdf.write_days(self, it + bad_day, it + bad_day + 1)
                    done += bad_day + 1
                    continue
            if self._netcdf is not None:
                self._write_async(it, t1)
            self.control.advance(nd)
            if "PRMSSolarGeometry" in self.processes:
                for ts in self.processes["PRMSSolarGeometry"]._vars.values():
                    if ts.allocated:
                        ts.advance()
            if "PRMSAtmosphere" in self.processes:
                self._serve_all_time_current(lazy=True)
            self.calculate_last_only()
            done += nd
        if finalize:
            self.finalize()

    def calculate_last_only(self):
        """Expose the last computed day in the process arrays (what a caller
        inspecting the model after run() sees)."""
        it = self.control.itime_step
        j = it - self._block_t0
        for name, proc in self.processes.items():
            if proc._all_time:
                continue
            for var in proc.get_variables():
                if var in self._block:
                    proc._vars[var][:] = self._block[var][j]
            b = getattr(proc, "budget", None)
            if b is not None and self._block_viol is not None and P.base_name(name) in BUDGET_INDEX:
                b._n_violations = int(self._block_viol[j, BUDGET_INDEX[P.base_name(name)]])
                b._zero_sum = b._n_violations == 0
                b._itime_accumulated = it
        if self._tail:
            # (their global budgets were checked for every day of the block in _tail_block)
            self._tail_graph.inflows[:] = self._block["inflows"][j]
            for obj in self._tail:
                if obj.budget is not None:
                    obj.budget._n_violations, obj.budget._zero_sum = 0, True
                    obj.budget._itime_accumulated = it
        self._calculated_step = it

    # ---------------------------------------------------------------- restart
    def save_state(self, path=None) -> dict:
        """A restart point at the model's current day: every prognostic value of
        the chain (the union of the processes' get_restart_variables(),
        prms_snow.py:269-293, PRMSChannel's outflow_ts / seg_inflow included) plus
        the budget accumulations.  The reference declares restart but raises
        "restart capability not implemented" (prms_soilzone.py:331); here
        `Model.load_state()` on a freshly built model of the same control and
        parameters continues the run bit for bit.  `path`: also write an .npz."""
        self._ensure_engine()
        if self._mode not in ("chain", "chain+channel") or self._tail:
            raise NotImplementedError("save_state needs the full NHM chain (without a FlowGraph behind it)")
        it = self.control.itime_step
        if self._engine.day != it + 1:
            raise RuntimeError(
                f"save_state at a block boundary only: the control is at day {it}, the device at "
                f"{self._engine.day - 1} (run(n_time_steps=...) ends on one; a model stepped day by day "
                "computes ahead of the control)")
        st = self._engine.save_state()
        for name, proc in self.processes.items():
            b = getattr(proc, "budget", None)
            if b is None:
                continue
            for comp, terms in b.accumulations.items():
                for v, a in terms.items():
                    if a is not None:
                        st[f"_acc/{name}/{comp}/{v}"] = np.asarray(a)
        if path is not None:
            np.savez(path, **st)
        return st

    def load_state(self, state):
        """Continue from `save_state()`'s dict (or its .npz file)."""
        if not isinstance(state, dict):
            with np.load(state) as z:
                state = {k: z[k] for k in z.files}
        self._ensure_engine()
        day = int(state["_day"])
        self._engine.reset()  # zeroes the device-side accumulations
        self._engine.load_state({k: v for k, v in state.items() if not k.startswith("_acc/")})
        while self.control.itime_step < day - 1:
            self.control.advance()
        self._days_on_device = day
        self._block, self._block_n = None, 0
        self._device_acc_cache = (None, {})
        for name, proc in self.processes.items():
            b = getattr(proc, "budget", None)
            if b is None:
                continue
            for comp in b.components:
                for v in b.terms[comp]:
                    key = f"_acc/{name}/{comp}/{v}"
                    b._accumulations[comp][v] = np.array(state[key]) if key in state else None
        if "PRMSAtmosphere" in self.processes:
            for ts in self.processes["PRMSAtmosphere"]._vars.values():
                ts.mark_stale()

    def _write_control(self, where):
        path = pl.Path(where) if not isinstance(where, bool) else pl.Path("control.yaml")
        import yaml

        opts = {k: (str(v) if isinstance(v, pl.Path) else v) for k, v in self.control.options.items()}
        path.write_text(yaml.safe_dump({
            "start_time": str(self.control.start_time), "end_time": str(self.control.end_time),
            "time_step": str(self.control.time_step), "options": opts}))


def _copy_rows(dst, t0, n, src):
    dst[t0:t0 + n] = src


_POOL = None


def _pool():
    """A small thread pool for the per-block host copies / sums of Model."""
    global _POOL
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor

        _POOL = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) // 2)))
    return _POOL


def _read_netcdf_var(path, name, keep_f32=False):
    """[time, unit] variable + decoded times from a netCDF file."""
    from .netcdf_in import read_time_variable

    return read_time_variable(path, name, keep_f32=keep_f32)

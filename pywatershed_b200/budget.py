"""Budget: per-process mass balance, same observable behaviour as the
reference's Budget (pywatershed/base/budget.py:25-783) for the terms a
ConservativeProcess declares in get_mass_budget_terms().

The closure verdict itself comes from the device (k_hru_column counts, per
day and process, the units failing np.allclose(inputs, outputs +
storage_changes, rtol=atol=1e-5); k_channel_budget does the global-basis
check), so a fast Model.run() never touches per-unit sums on the host; the
`_inputs_sum` / `_outputs_sum` / `_storage_changes_sum` / `balance` views are
computed lazily from the process's current-day arrays when somebody looks.
"""
from __future__ import annotations

from warnings import warn

import numpy as np


class Budget:
    def __init__(self, control, process, terms: dict, budget_type, basis: str = "unit",
                 rtol: float = 1e-5, atol: float = 1e-5, description: str = None,
                 units: str = None):
        self.name = "Budget"
        self.control = control
        self._proc = process
        self.terms = {k: list(v) for k, v in terms.items()}
        self.components = ("inputs", "outputs", "storage_changes")
        self.basis = basis
        self.rtol, self.atol = rtol, atol
        self.description = description or type(process).__name__
        self.imbalance_fatal = budget_type == "error"
        self.budget_type = budget_type
        self._zero_sum = True
        self._n_violations = 0
        self._itime_accumulated = -1
        self._accumulations = {
            c: {v: None for v in self.terms[c]} for c in self.components}
        self._acc_baseline = {}
        self._accum_start_time = getattr(control, "init_time", None)
        self.time_unit = "D"
        self._unit_desc = "volumes"
        # units: the common unit of the terms' metadata (base/budget.py:79-100)
        from ._process_meta import VARIABLE_META

        all_vars = [v for c in self.components for v in self.terms[c]]
        self.meta = {v: dict(VARIABLE_META[v]) for v in all_vars if v in VARIABLE_META}
        if units is None and len(self.meta) == len(all_vars) and all_vars:
            us = [m.get("units") for m in self.meta.values()]
            if any(u != us[0] for u in us):
                raise ValueError("Units not consistent over all terms")
            units = us[0]
        self.units = units
        # the derived output variables of <Process>_budget.nc (base/budget.py:102-127)
        self.output_vars_desc = {
            c + "_sum": f"Sum of {d} ({basis})"
            for c, d in (("inputs", "input fluxes"), ("outputs", "output fluxes"),
                         ("storage_changes", "storage changes")) if self.terms[c]}
        for k, desc in self.output_vars_desc.items():
            t0 = self.terms[k[:-4]][0]
            if t0 in self.meta:
                self.meta[k] = dict(self.meta[t0], desc=desc, units=self.units)

    # ---- component views (dict var -> current-day array), like budget[component]
    def __getitem__(self, key):
        if key in self.components:
            return {v: self._proc[v] for v in self.terms[key]}
        return getattr(self, key)

    def _sum(self, component):
        vals = [self._proc[v] for v in self.terms[component]]
        if self.basis == "unit":
            return sum(vals)
        return sum(np.sum(v) for v in vals)

    @property
    def _inputs_sum(self):
        return self._sum("inputs")

    @property
    def _outputs_sum(self):
        return self._sum("outputs")

    @property
    def _storage_changes_sum(self):
        return self._sum("storage_changes")

    @property
    def balance(self):
        return self._inputs_sum - self._outputs_sum

    @property
    def _balance(self):
        return self.balance

    # the reference's public names for the three sums (base/budget.py:188-198)
    @property
    def inputs_sum(self):
        return self._inputs_sum

    @property
    def outputs_sum(self):
        return self._outputs_sum

    @property
    def storage_changes_sum(self):
        return self._storage_changes_sum

    @property
    def inputs(self):
        return self["inputs"]

    @property
    def outputs(self):
        return self["outputs"]

    @property
    def storage_changes(self):
        return self["storage_changes"]

    @staticmethod
    def get_components():
        return ("inputs", "outputs", "storage_changes")

    @property
    def accumulations(self):
        """component -> term -> per-unit sum over the steps so far
        (base/budget.py:262-269).  Blocks computed by Model.run() are summed on
        the device (pws_select_accumulations) and fetched when somebody looks;
        blocks served day by day are summed on the host."""
        model = getattr(self._proc, "_model", None)
        dev = model._device_accumulations() if model is not None else {}
        out = {}
        for c in self.components:
            out[c] = {}
            for v in self.terms[c]:
                host, d = self._accumulations[c][v], dev.get(v)
                base = self._acc_baseline.get((c, v))
                if d is not None and base is not None:
                    d = d - base
                out[c][v] = host if d is None else (d if host is None else host + d)
        return out

    def _n_units(self):
        for c in self.components:
            for v in self.terms[c]:
                try:
                    return int(np.asarray(self._proc[v]).shape[-1])
                except Exception:
                    continue
        return 1

    def set_initial_accumulations(self, init_accumulations, accum_start_time=None):
        """base/budget.py:204-228: start the accumulations from zero, or from
        `init_accumulations[component][term]` (a restart), counted since
        `accum_start_time`."""
        self.reset_accumulations()
        if init_accumulations is None:
            self._accum_start_time = getattr(self.control, "init_time", None)
            return
        self._accum_start_time = accum_start_time
        for c in self.components:
            for v in self.terms[c]:
                if v in init_accumulations.get(c, {}):
                    self._accumulations[c][v] = np.array(init_accumulations[c][v], dtype=np.float64)

    def reset_accumulations(self):
        """base/budget.py:230-237: every accumulation back to zero, counted from
        the last accumulated time.  Sums living on the device are not touched
        (other budgets of the model may share the term): their current value
        becomes the baseline subtracted from now on."""
        self._accum_start_time = self._time_accumulated()
        model = getattr(self._proc, "_model", None)
        dev = model._device_accumulations() if model is not None else {}
        n = self._n_units()
        for c in self.components:
            for v in self.terms[c]:
                self._accumulations[c][v] = np.zeros(n) if dev.get(v) is None else None
                if dev.get(v) is not None:
                    self._acc_baseline[(c, v)] = np.array(dev[v])

    def _time_accumulated(self):
        return self.control.current_time if self._itime_accumulated >= 0 else None

    @property
    def _accumulations_sum(self):
        return self._sum_component_accumulations()

    def _sum_component_accumulations(self):
        """base/budget.py:283-307: per component, the sum of its terms'
        accumulations -- per unit, or one number for a global-basis budget."""
        acc = self.accumulations
        out = {}
        for c in self.components:
            tot = None
            for v in self.terms[c]:
                a = acc[c][v]
                if a is None:
                    continue
                a = np.asarray(a, dtype=np.float64)
                a = a.copy() if self.basis == "unit" else a.sum()
                tot = a if tot is None else tot + a
            out[c] = tot
        return out

    # ---- stepping
    def advance(self):
        pass

    def block_terms(self, block: dict, dt_days: float):
        """(component, term, [ndays, n] array, dt) for every budget term present in `block`."""
        return [(c, v, block[v], dt_days) for c in self.components for v in self.terms[c] if v in block]

    def accumulate_term(self, c, v, arr, dt_days: float):
        """Add a block [ndays, n] of one term (times the step in days,
        base/budget.py:262-269) to its accumulation."""
        add = np.asarray(arr, dtype=np.float64).sum(axis=0) * dt_days
        cur = self._accumulations[c][v]
        self._accumulations[c][v] = add if cur is None else cur + add

    def accumulate_block(self, block: dict, dt_days: float):
        for a in self.block_terms(block, dt_days):
            self.accumulate_term(*a)

    def calculate(self, n_violations: int | None = None):
        """Verdict for the current day.  `n_violations` = the device's count of
        units out of balance; when None the check is done here with numpy
        exactly as base/budget.py:336-416 does."""
        it = self.control.itime_step
        if self._itime_accumulated >= it and it >= 0:
            raise ValueError("Can not accumulate twice per timestep")
        self._itime_accumulated = it
        if n_violations is None:
            lhs = self._inputs_sum
            rhs = self._outputs_sum + self._storage_changes_sum
            ok = np.allclose(lhs, rhs, rtol=self.rtol, atol=self.atol)
            n_violations = 0 if ok else int(np.sum(~np.isclose(lhs, rhs, rtol=self.rtol, atol=self.atol)))
        self._n_violations = int(n_violations)
        self._zero_sum = self._n_violations == 0
        if not self._zero_sum and self.budget_type is not None:
            if self.basis == "unit":
                msg = ("The flux unit balance not equal to the change in unit storage at time "
                       f"{self.control.current_time} for {self.description}: "
                       f"{self._n_violations} unit(s)")
            else:
                msg = ("The global flux balance not equal to the change in global storage: "
                       f"{self.description}")
            if self.imbalance_fatal:
                raise ValueError(msg)
            warn(msg, UserWarning)

    def __repr__(self):
        """The reference's report (base/budget.py:422-648): the step's rates and
        the accumulated volumes, spatial sums per term in three columns, with
        the balance line under each table."""
        it = self.control.itime_step
        if it < 0:
            return f"Budget of {self.description} only initialized. Check back later."
        cols = []
        acc = self.accumulations
        for c in self.components:
            keys = self.terms[c]
            cols.append((c, keys, max([len(k) for k in keys] + [len(c) + 14]) + 12))
        indent, sep = " " * 9, " " * 4
        width = sum(w for _, _, w in cols) + 2 * len(sep)

        def fmt(x):
            return "--" if x is None else np.format_float_scientific(float(np.sum(x)), precision=4)

        def table(values):
            lines = []
            for r in range(max(len(k) for _, k, _ in cols)):
                line = indent
                for (c, keys, w) in cols:
                    cell = f"{keys[r]}: {fmt(values[c][keys[r]])}" if r < len(keys) else ""
                    line += cell.ljust(w) + sep
                lines.append(line.rstrip())
            return lines

        def balance(sums):
            op = "=" if self._zero_sum else "!=!"
            line = "Balance: "
            for (c, _, w), o in zip(cols, ("", "-", op)):
                line += (o + " " + fmt(sums[c])).ljust(w) + sep
            return line.rstrip()

        head = ("Individual spatial unit budget" if self.basis == "unit" else "Global budget")
        out = ["*-" * ((len(indent) + width) // 2),
               f"{head} for {self.description} (units: {self.units}).",
               ("Budget is checked on each spatial unit. This summary shows spatial sums for the "
                "entire model domain." if self.basis == "unit" else
                "Budget is checked on full domain: spatially summed fluxes and storage changes are "
                "checked for balance."),
               f"@ time: {self.control.current_time} (itime_step: {it})", "",
               "This timestep, rates:",
               indent + sep.join(f"{n} rates".ljust(w) for n, (_, _, w) in zip(
                   ("input", "output", "storage change"), cols)),
               indent + sep.join("-" * w for _, _, w in cols)]
        out += table({c: {v: self._proc[v] for v in self.terms[c]} for c in self.components})
        out += [indent + "-" * width,
                balance({"inputs": self._inputs_sum, "outputs": self._outputs_sum,
                         "storage_changes": self._storage_changes_sum}), "",
                f"Accumulated volumes (since {self._accum_start_time}):",
                indent + sep.join(f"{n} volumes".ljust(w) for n, (_, _, w) in zip(
                    ("input", "output", "storage change"), cols)),
                indent + sep.join("-" * w for _, _, w in cols)]
        out += table(acc)
        out += [indent + "-" * width, balance(self._sum_component_accumulations()), ""]
        return "\n".join(out)

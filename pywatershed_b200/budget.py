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
                 rtol: float = 1e-5, atol: float = 1e-5, description: str = None):
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
                out[c][v] = host if d is None else (d if host is None else host + d)
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
        return (f"Budget({self.description}, basis={self.basis}, zero_sum={self._zero_sum}, "
                f"terms={self.terms})")

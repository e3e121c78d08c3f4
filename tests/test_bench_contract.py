"""The JSON line of `bench.py --impl reference` (the CPU arm the driver runs
beside the GPU arm): keys, units and the bounded sample, on this machine's host
cores.  The GPU arm's line is checked by tests/test_gpu_fullsize.py."""
import json
import pathlib as pl
import subprocess
import sys

import numpy as np

ROOT = pl.Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--ref", "port",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=900,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("simulated HRU-days/sec")
    assert d["unit"] == "HRU-days/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "CONUS" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "DRB tile" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "HRU-days/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    # the other CPU paths ride along: the reference package's own numba / numpy when it was
    # packed into oracle/_ref (this container), known answers included
    if (ROOT / "oracle" / "_ref" / "pywatershed" / "__init__.py").exists():
        for k in ("numba", "numpy"):
            o = d["other_cpu_paths"][k]
            assert o["value"] > 0 and "pywatershed" in o["sample"] and f"calc_method={k}" in o["sample"]
        assert d["other_cpu_paths"]["numba"]["value"] > d["other_cpu_paths"]["numpy"]["value"]


def test_gpu_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm must not print a number."""
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1",
                        "--workload", "ucb", "--no-cpu"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT)
    assert r.returncode != 0
    assert not any(ln.lstrip().startswith("{") for ln in r.stdout.splitlines())


def test_ensemble_parameters_follow_the_recipe(drb):
    """bench.ensemble_parameters: member m scales the
    listed soil / snow parameters by U(0.8, 1.2) from seed m, inside the
    reference's parameter bounds; everything else is the base domain's."""
    import bench

    p, f, start = drb
    pt = bench.ensemble_parameters(p, [3, 900])
    again = bench.ensemble_parameters(p, [900])
    for k, (lo, hi) in bench.ENSEMBLE_PARAMS.items():
        v = np.asarray(pt[k], dtype=np.float64)
        assert v.min() >= lo and v.max() <= hi, k
        base = np.asarray(p[k], dtype=np.float64)
        ratio = v[..., :765] / np.where(base == 0, 1.0, base)
        inside = (base != 0) & (v[..., :765] > lo) & (v[..., :765] < hi)
        assert ((ratio[inside] >= 0.8 - 1e-12) & (ratio[inside] <= 1.2 + 1e-12)).all(), k
        np.testing.assert_array_equal(v[..., 765:], np.asarray(again[k]))  # a member does not depend on its rank
    np.testing.assert_array_equal(pt["hru_area"][:765], pt["hru_area"][765:])
    assert pt["tosegment"][456:].max() <= 2 * 456 and pt["hru_segment"][765:].max() > 456


def test_every_bench_leg_selects_the_outputs_it_reports():
    """A handle's default output selection is EMPTY (nothing is written): a bench leg
    that builds an Engine must select outputs before it times anything, and its line
    must carry a checksum of what was written.  (The channel-only workload once ran
    without a selection; its figures were no-output figures.)"""
    import inspect

    import bench

    for fn in (bench.run_channel_only, bench.run_split_columns):
        src = inspect.getsource(fn)
        assert "checksum_seg_outflow_last_day" in src, fn.__name__
    src = inspect.getsource(bench.run_channel_only)
    assert src.index("select_outputs") < src.index("def one_pass")
    main = inspect.getsource(bench.main)
    assert main.count("Engine(") == main.count("def engine_for")  # engines come from one factory ...
    leg = main[main.index("def device_leg"):main.index("main_leg =")]
    assert "eng.select_outputs(hru_vars, seg_vars)" in leg and "seg_last" in leg  # ... and select + checksum

"""Stress of overlapped routing launches (option route_overlap) at full size: the
chained configs[4] network (56,088 segments in one tree, 123 chunks, 8,856 levels)
routed over 1,100 days in launches of 256 days issued back to back, against ONE
launch: every value of every segment variable of the last 76 days must be the same
bits, and again with the launches serialised."""
import pathlib as pl
import sys

ROOT = pl.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import numpy as np
import torch

import bench
import scenarios
from pywatershed_b200 import Engine

p0, f0, start = scenarios.load_drb()
ntile, nd = 123, 1100
pt, _ = scenarios.tile(p0, {k: v[:2] for k, v in f0.items()}, ntile)
pt, depth = bench.chain_tiles(p0, pt, ntile)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(11)
nseg = pt["tosegment"].shape[0]
lat = torch.exp(torch.randn((nd, nseg), dtype=torch.float64, device=dev, generator=gen))


def run(T, overlap):
    e = Engine(pt, start, device=0)
    e.select_outputs((), e.reg.seg_vars)
    e._opt("route_overlap", overlap)
    out = torch.zeros((nd, 5, nseg), dtype=torch.float64, device=dev)
    d = 0
    while d < nd:
        n = min(T, nd - d)
        e.run_channel_device(n, lat.data_ptr() + d * nseg * 8, out.data_ptr() + d * 5 * nseg * 8, sync=False)
        d += n
    e.sync()
    e.close()
    return out


one = run(nd, 0)
sums = [float(one[:, k].sum()) for k in range(5)]
print("sums over all days per segment variable:", sums)
assert all(np.isfinite(sums)) and sums[3] > 0.0, "the single launch wrote nothing"
import prms_oracle as po
ref = po.Oracle(pt, start=start).run_channel(lat.cpu().numpy())
oc = one.cpu().numpy()
for k, v in enumerate(po.SEG_VARS):
    np.testing.assert_array_equal(oc[:, k], ref[v], err_msg=v)
print("one launch of", nd, "days equals the oracle bit for bit")
for T, ov in ((256, 1), (256, 0), (37, 1)):
    got = run(T, ov)
    same = bool(torch.equal(got, one))
    print(f"launches of {T} days, route_overlap {ov}: identical to one launch: {same}")
    assert same
print("levels", depth * ntile, "segments", nseg, "checksum seg_outflow last day", float(one[-1, 3].sum()))

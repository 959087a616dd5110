"""GPU: the K2 scatter-aggregate variants (gib_scatter_variant) at the C4 single-GPU shape, bond-type-grouped and
dst-sorted message layouts; bytes = SURVEY.md 8(d)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from graphinvent_b200._lib import lib  # noqa: E402

S, E, width = 155648, 352256, 100
ld = 112
nbytes = E * width * 4 + S * width * 4 + (S + 1) * 4
g = torch.Generator(device="cpu").manual_seed(1)
t = torch.multinomial(torch.tensor([0.84, 0.14, 0.02]), E, replacement=True, generator=g)
order = torch.argsort(t, stable=True)
ent_grouped = torch.empty(E, dtype=torch.int32)
ent_grouped[order] = torch.arange(E, dtype=torch.int32)
pk = bench.peaks()
res = {}
for v in (0, 1, 2, 3):
    lib.gib_scatter_variant(v)
    ms_g = bench._time_scatter(S, E, ld, ent_grouped)
    ms_s = bench._time_scatter(S, E, ld, torch.arange(E, dtype=torch.int32))
    res[v] = {"grouped_ms": ms_g, "grouped_gbs": nbytes / ms_g / 1e6, "grouped_frac": nbytes / ms_g / 1e6 / pk["hbm"],
              "sorted_ms": ms_s, "sorted_gbs": nbytes / ms_s / 1e6, "sorted_frac": nbytes / ms_s / 1e6 / pk["hbm"]}
    print(v, json.dumps(res[v]), flush=True)
lib.gib_scatter_variant(2)
print(json.dumps({"k2_variants": res, "bytes": nbytes, "peak_gbs": pk["hbm"]}))

"""Canonical neighbour tie order (ascending (d^2, dy, dx): the oracle default and the CUDA path) against the restated rstar 0.7.1
order (ORC_KNN=rstar, oracle/rstar_port.hpp) on the nine diff.rs configurations and BASELINE C1-C3: fraction of output pixels
whose colour / source coordinate differ, mean neighbourhood cost, and the total-variation distance between the two 64-bin
histograms of per-pixel cost.  Inputs: the reference's images with its JPEGs decoded as jpeg-decoder 0.1.22 does.  CPU only; writes tests/golden/tie_order_compare.json (quoted in DESIGN.md section 2).
Usage: python tests/golden/compare_tie_orders.py [case ...]"""
import sys, os, json, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import numpy as np
from tests import fullsize_cases as F
# inputs as the reference decodes them (jpeg-decoder restated, oracle/jpeg_port.py): with the rstar order this side of the
# comparison reproduces the reference's nine hashes exactly, the canonical side is bit-identical to the CUDA path
F.set_decoder("jpegport")
names = sys.argv[1:] or list(F.DIFF_HASHES) + ["c1_single_example_500", "c2_multi_example_500", "c3_guided_500", "c3_style_transfer_500"]
def hist_distance(a, b):
    a = a[np.isfinite(a)]; b = b[np.isfinite(b)]
    hi = max(a.max(), b.max(), 1e-9)
    ha, _ = np.histogram(a, bins=64, range=(0, hi)); hb, _ = np.histogram(b, bins=64, range=(0, hi))
    return 0.5 * np.abs(ha / ha.sum() - hb / hb.sum()).sum()
out = {}
for name in names:
    spec = F.SPECS[name]()
    os.environ.pop("ORC_KNN", None)
    t0=time.time(); a = F.to_oracle(spec).run(); ta=time.time()-t0
    os.environ["ORC_KNN"] = "rstar"
    t0=time.time(); b = F.to_oracle(spec).run(); tb=time.time()-t0
    os.environ.pop("ORC_KNN", None)
    sa, sb = a.resolved()[1], b.resolved()[1]
    r = dict(color_mismatch=float(np.mean((a.color() != b.color()).any(axis=2))), coord_mismatch=float(np.mean((a.coord() != b.coord()).any(axis=2))),
             mean_score_canonical=float(np.nanmean(sa[np.isfinite(sa)])), mean_score_rstar=float(np.nanmean(sb[np.isfinite(sb)])),
             cost_histogram_tv=float(hist_distance(sa, sb)), secs=(round(ta,1), round(tb,1)))
    out[name] = r
    print(name, r, flush=True)
json.dump(out, open('tests/golden/tie_order_compare.json','w'), indent=1)

"""Full-size runs of the BASELINE.json configs C4/C5 on one GPU (not a pytest file): python tests/gpu_configs.py [c4] [c5] [c5small]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.helpers import Case, O
from texture_synthesis_b200 import capi
from texture_synthesis_b200.synth import synth_texture, border_inpaint_mask

def report(name, g, dt, out):
    st = g.stats()
    flat, score = g.resolved()
    print(f"{name}: {dt*1e3:.1f} ms -> {out*out/dt/1e6:.2f} Mpx/s; items {st['work_items']} phases {st['phases']} resolve_ms {st['gpu_ms_resolve']:.1f} "
          f"analysis_ms {st['gpu_ms_analysis']:.1f} sched_wait_ms {st['host_ms_schedule']:.1f} gpu_total {st['gpu_ms_total']:.1f} "
          f"texels {st['texels_fetched']/1e9:.2f}G/{st['texels_nominal']/1e9:.2f}G resolved {len(flat)} unique {len(np.unique(flat))}", flush=True)

which = sys.argv[1:] or ["c4", "c5small"]
if "c4" in which:
    # C4: 1024^2 synthetic example, CatmullRom-upscaled to 2048^2 (inpaint forces example = output size), border mask, tiling
    ex = capi.resize(synth_texture(1024, 1024, 2), 2048, 2048, capi.FILTER_CATMULLROM)
    mask = border_inpaint_mask(2048, 2048, 0.17)
    pyr = capi.pyramid_build(ex, 5)
    g = capi.Generator(2048, 2048, mask, ex, 0)
    g.upload_inputs([pyr])
    for it in range(2):
        g.reset()
        t0 = time.perf_counter(); g.resolve_resident(capi.make_params(seed=0, tiling=True)); dt = time.perf_counter() - t0
        report("C4 inpaint+tiling 2048^2", g, dt, 2048)
    col = g.color(); keep = mask[..., 0] == 255
    print("  locked pixels untouched:", bool((col[keep] == ex[keep]).all()))
    del g
for name, out in (("c5small", 4096), ("c5", 8192)):
    if name in which:
        ex = synth_texture(1024, 1024, 2)
        pyr = capi.pyramid_build(ex, 5)
        g = capi.Generator(out, out)
        g.upload_inputs([pyr])
        for it in range(2 if out == 4096 else 1):
            g.reset()
            t0 = time.perf_counter(); g.resolve_resident(capi.make_params(seed=0)); dt = time.perf_counter() - t0
            report(f"C5 {out}^2 from 1024^2", g, dt, out)
        co = g.coord(); col = g.color()
        print("  colour == example[coord]:", bool((col == pyr[-1][co[..., 1], co[..., 0]]).all()))
        del g

"""Times the pieces of the end-to-end (host buffer) path: python tests/gpu_e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.helpers import Case
from texture_synthesis_b200 import capi
case = Case("e2e", 2048, 2048, [(512, 512)], seed=0).build()
out = np.empty((2048, 2048, 4), np.uint8)
for it in range(3):
    t0 = time.perf_counter()
    g = capi.Generator(2048, 2048)
    t1 = time.perf_counter()
    g.resolve(case.gpu_params(), case.pyramids)
    t2 = time.perf_counter()
    capi._check(g.L.tsb_generator_read_color(g.h, out.ctypes.data))
    t3 = time.perf_counter()
    st = g.stats()
    print(f"create {1e3*(t1-t0):.1f} ms, resolve {1e3*(t2-t1):.1f} ms (stats wall {st['wall_ms_total']:.1f}, gpu_total {st['gpu_ms_total']:.1f}), read_color {1e3*(t3-t2):.1f} ms", flush=True)
    g.close()
g = capi.Generator(2048, 2048)
for it in range(3):
    g.reset()
    t1 = time.perf_counter()
    g.resolve(case.gpu_params(), case.pyramids)
    t2 = time.perf_counter()
    capi._check(g.L.tsb_generator_read_color(g.h, out.ctypes.data))
    t3 = time.perf_counter()
    st = g.stats()
    print(f"reuse: resolve {1e3*(t2-t1):.1f} ms (stats wall {st['wall_ms_total']:.1f}, gpu_total {st['gpu_ms_total']:.1f}), read_color {1e3*(t3-t2):.1f} ms", flush=True)

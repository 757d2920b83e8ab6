"""Step time (fused residual norm + one V-cycle, graph replay, CUDA events) with the tail program (csrc/tail.cu) off and
on for several thresholds / CTAs per SM, on BASELINE configs[1] (irregular 1025^2 mesh, NN-built 4-level hierarchy,
V(3,3)) and on a structured 2049^2 6-level V(1,1) hierarchy (the shape of the headline run's coarse tail).
    python tools/tail_sweep.py [--skip-c2] > gpurun_out/tail_sweep.jsonl
One JSON line per measurement."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(h, params, rhs, steps, torch, _lib):
    lib = h.lib
    lev0 = h.levels[0]
    st = _lib.stream_handle(torch)

    def step():
        _lib.check(lib.mg_sell_residual_norm2(ctypes.byref(lev0.A.struct), lev0.x.data_ptr(), lev0.b.data_ptr(),
                                              h._norm_ws.data_ptr(), h._norm_out.data_ptr(), st))
        h.vcycle(params)
    h.set_rhs(rhs)
    h.zero_x()
    for _ in range(5):
        step()
    h.zero_x()
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        best = ms if best is None else min(best, ms)
    return best, 2 + h.last_launches, float(h.residual_norm())


def sweep(name, h, params, rhs, thresholds, steps, torch, _lib):
    lib = h.lib
    rows = [lv.n for lv in h.levels]
    base = None
    for thr in thresholds:
        for ctas in ((2,) if thr == 0 else (1, 2, 3)):
            lib.mg_set_tail_max_rows(int(thr))
            lib.mg_set_tail_ctas_per_sm(ctas)
            ms, launches, res = measure(h, params, rhs, steps, torch, _lib)
            o, b, l = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
            lib.mg_vcycle(h._level_structs, h.nlevels, ctypes.byref(params), _lib.stream_handle(torch))
            torch.cuda.synchronize()
            lib.mg_tail_last_stats(ctypes.byref(o), ctypes.byref(b), ctypes.byref(l))
            if thr == 0:
                base = ms
            print(json.dumps({"case": name, "levels_rows": rows, "tail_max_rows": int(thr), "ctas_per_sm": ctas,
                              "ms_per_step": round(ms, 5), "speedup_vs_off": round(base / ms, 4),
                              "launches_per_step": launches, "tail_ops": o.value, "tail_barriers": b.value,
                              "tail_launches": l.value, "residual_after": res}), flush=True)
    lib.mg_set_tail_max_rows(0)
    lib.mg_set_tail_ctas_per_sm(2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-c2", action="store_true")
    ap.add_argument("--skip-structured", action="store_true")
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    import torch
    from learnmultigrid_b200 import _lib
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.solvers.Multigrid import NeuralMG_2D, SemiGeometricMG
    from learnmultigrid_b200.neural2d import MassSurrogate
    if not a.skip_structured:
        t0 = time.perf_counter()
        n = 2048
        A, rhs, Qs = P.structured_laplacian_2d(n), P.structured_rhs_2d(n), P.structured_hierarchy_2d(n, 6, "linear")
        mg = SemiGeometricMG(A, rhs, Qs)
        h = mg._hierarchy(6, "GaussSeidel", "multicolor", None, True)
        params = h.make_params(nu_pre=1, nu_post=1)
        print(json.dumps({"case": "structured 2049^2 6-level V(1,1)", "build_s": round(time.perf_counter() - t0, 2)}),
              flush=True)
        sweep("structured 2049^2 6-level V(1,1)", h, params, rhs, [0, 20000, 70000, 300000, 1100000], a.steps,
              torch, _lib)
        del h, mg, A, Qs
    if not a.skip_c2:
        t0 = time.perf_counter()
        pb = P.irregular_p1_2d(1024, seed=42)
        nmg = NeuralMG_2D(pb["A"], pb["rhs"], MassSurrogate(), pb["M"], np.ones(43), np.zeros(43))
        nmg.define_hierarchy(4)
        mg = SemiGeometricMG(pb["A"], pb["rhs"], nmg.l_hierarchy)
        h = mg._hierarchy(4, "GaussSeidel", "multicolor", None, True)
        params = h.make_params(nu_pre=3, nu_post=3)
        print(json.dumps({"case": "C2 irregular 1025^2 NN 4-level V(3,3)", "build_s": round(time.perf_counter() - t0, 2)}),
              flush=True)
        sweep("C2 irregular 1025^2 NN 4-level V(3,3)", h, params, pb["rhs"], [0, 20000, 70000, 300000, 2000000],
              a.steps, torch, _lib)


if __name__ == "__main__":
    main()

"""Fused persistent pipeline vs the three stand-alone kernels: bit-identity and time (device-resident)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regularizepsf_b200 as rp
from regularizepsf_b200 import _native
from regularizepsf_b200.device import DeviceCube

def run(B, P, H, W, steps=20):
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((H, W), P)]
    g = torch.Generator(device="cuda").manual_seed(1)
    kernel = torch.randn((len(coords), P, P), dtype=torch.complex64, device="cuda", generator=g)
    t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
    frames = torch.rand((B, H, W), device="cuda", generator=g) * 1000
    nt = t._native_transform("float32")
    plan = nt.plan(H, W, 0, 0, H, B)
    lib = _native.load()
    has_fused = lib.rpsf_plan_set_fused(plan, 2) == 0
    info = nt.plan_info(plan)
    res = {}
    for mode, name in ((1, "3-kernel"), (2, "fused")):
        if mode == 2 and not has_fused:
            res[name] = res["3-kernel"]
            continue
        _native.check(lib.rpsf_plan_set_fused(plan, mode))
        out = torch.empty_like(frames)
        for _ in range(3):
            t._apply_device(frames, "float32", 0, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            t._apply_device(frames, "float32", 0, out=out)
        e1.record(); torch.cuda.synchronize()
        res[name] = (out.clone(), 1e3 * e0.elapsed_time(e1) / steps / B)
    stats = ""
    if info["fused_pipeline"] and os.environ.get("FUSED_STATS"):
        import ctypes
        buf = (ctypes.c_uint64 * 8)()
        _native.check(lib.rpsf_plan_fused_stats(plan, 1, None))
        _native.check(lib.rpsf_plan_set_fused(plan, 2))
        t._apply_device(frames, "float32", 0, out=out)
        _native.check(lib.rpsf_plan_fused_stats(plan, 0, buf))
        v = [int(x) for x in buf]
        stats = (f"\n    wait share of role time: K1 {v[0] / max(v[3], 1):.2f} K2(group leader only, x{1}) {v[1] * 16 / max(v[4], 1):.2f} "
                 f"K3(lane 0 team) {v[2] / max(v[5], 1):.2f}; mean role kcycles/warp: "
                 f"K1 {v[3] / 1e3 / max(1, 16 * int(os.environ.get('N1', 46))):.0f} K2 {v[4] / 1e3 / max(1, 16 * int(os.environ.get('N2', 62))):.0f} "
                 f"K3 {v[5] / 1e3 / max(1, 16 * int(os.environ.get('N3', 40))):.0f}; units not prefetched {v[6]}/{v[7]}")
    if info["fused_pipeline"] and os.environ.get("FUSED_TRACE"):
        import ctypes
        import numpy as np
        nb = len({c[0] for c in coords})
        n = B * nb
        buf = (ctypes.c_uint64 * (3 * n))()
        _native.check(lib.rpsf_plan_fused_trace(plan, 1, None, 0))
        t._apply_device(frames, "float32", 0, out=out)
        _native.check(lib.rpsf_plan_fused_trace(plan, 0, buf, 3 * n))
        tr = np.array(buf, dtype=np.float64).reshape(3, n)
        t0 = tr[tr > 0].min()
        tr = (tr - t0) / 1e3
        stats += "\n    band timeline (us since first event): seq  K1-done  K2-done  K3-freed"
        for q in list(range(0, min(n, 2 * nb))) + list(range(max(2 * nb, n - nb), n)):
            stats += f"\n      {q:4d} {tr[0, q]:8.1f} {tr[1, q]:8.1f} {tr[2, q]:8.1f}"
    same = torch.equal(res["3-kernel"][0], res["fused"][0])
    diff = float((res["3-kernel"][0] - res["fused"][0]).abs().max())
    print(f"B={B} P={P} {H}x{W} fused={info['fused_pipeline']}: 3-kernel {res['3-kernel'][1]:.1f} us/frame, fused {res['fused'][1]:.1f} us/frame, "
          f"bit-identical {same} (max diff {diff:.3g})" + stats, flush=True)

if __name__ == "__main__":
    cases = [(1, 64, 256, 192), (2, 128, 512, 384), (1, 256, 1024, 1024), (1, 256, 2048, 2048), (8, 256, 2048, 2048), (8, 128, 1024, 1024), (1, 128, 1024, 1024)]
    if len(sys.argv) > 1:
        cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for c in cases:
        run(*c)

#!/usr/bin/env python
"""Device-timed throughput of the Z-stack kernel (dcb_unwarp_stack_backward_f32)
on BASELINE-shaped stacks.  Usage: bench_stack.py [--reps N]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import discorpy_b200 as dcb                                    # noqa: E402
from discorpy_b200 import _cabi                                # noqa: E402
from bench import ClockSampler                                 # noqa: E402

COEF_DOT_05 = [1.00227490554, -2.99523692178e-05, 8.99519088e-08,
               -1.57066461911e-10, 8.08880211618e-14]
CASES = {
    # name: (D, H, W, xc, yc, fact, coord_round)
    "cfg2x16": (16, 4096, 4096, 2050.37, 2040.81, [COEF_DOT_05[i] / 3.0 ** i for i in range(5)], 1),
    "cfg2x64": (64, 4096, 4096, 2050.37, 2040.81, [COEF_DOT_05[i] / 3.0 ** i for i in range(5)], 1),
    "cfg4shard": (64, 2560, 2560, 1283.4, 1275.9, [1.0, -2e-5, 6e-8, -1e-10, 5e-14], 0),
    "cfg4deep": (512, 2560, 2560, 1283.4, 1275.9, [1.0, -2e-5, 6e-8, -1e-10, 5e-14], 0),
    "cfg4chunk": (64, 2560, 2560, 1283.4, 1275.9, [1.0, -2e-5, 6e-8, -1e-10, 5e-14], 1),
    "cfg5shard": (8, 8192, 8192, 4100.3, 4090.8,
                  [1.0, -1e-5, 3e-8, -2e-11, 5e-15, -8e-19, 6e-23, -2e-27, 3e-32], 1),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cases", default=",".join(c for c in CASES if c != "cfg4deep"))
    ap.add_argument("--blends", default="exact,lerp64,lerp32")
    ap.add_argument("--orders", default="1")
    args = ap.parse_args()
    dcb.set_device(0)
    peak = 6451.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    fn = _cabi.load().dcb_unwarp_stack_backward_f32
    stream = dcb.current_stream()
    sh = ctypes.c_void_p(stream.handle)
    blends = {"exact": dcb.BLEND_EXACT, "lerp64": dcb.BLEND_LERP64, "lerp32": dcb.BLEND_LERP32}
    for name in args.cases.split(","):
        D, H, W, xc, yc, fact, cr = CASES[name]
        src = dcb.DeviceArray((D, H, W)).fill_synthetic(seed=4)
        dst = dcb.DeviceArray((D, H, W))
        model = _cabi.make_radial(xc, yc, fact)
        for order in [int(o) for o in args.orders.split(",")]:
            for bname in args.blends.split(","):
                if order == 0 and bname != "exact":
                    continue
                if order == 0 and cr == 0:
                    continue
                opt = _cabi.make_options(order, blends[bname], dcb.PATH_AUTO)

                def run():
                    _cabi.check(fn(ctypes.c_void_p(src.ptr), ctypes.c_void_p(dst.ptr), D, H, W, 0, H,
                                   src.pitch, src.slice_stride, dst.pitch, dst.slice_stride, 0, H, cr,
                                   ctypes.byref(model), ctypes.byref(opt), sh))
                for _ in range(2):
                    run()
                sampler = ClockSampler(0)
                sampler.start()
                e0, e1 = dcb.Event(), dcb.Event()
                e0.record(stream)
                for _ in range(args.reps):
                    run()
                e1.record(stream)
                e1.sync()
                ms = e0.elapsed_ms(e1) / args.reps
                clocks = sampler.stop()
                px = D * H * W
                gbs = 8.0 * px / (ms * 1e-3) / 1e9
                print(json.dumps({"case": name, "order": order, "blend": bname, "coord_round": cr,
                                  "D": D, "H": H, "W": W, "ms": ms, "Mpix_s": px / 1e6 / (ms * 1e-3),
                                  "GBs": gbs, "frac": gbs / peak, "us_per_4096sq": ms * 1e3 * 4096 * 4096 / px,
                                  "plan": dcb.last_plan(), "clocks": clocks,
                                  "lib": os.path.basename(_cabi.LIB_PATH)}), flush=True)
        del src, dst
        dcb.device.device_pool.clear()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric : Mpixels/s of backward unwarp, 4096 x 4096 float32, 5-term polynomial,
         order 1 (BASELINE configs[1]), whole job over N GPUs (weak scaling:
         every rank unwarps its own batch of independent images; the only
         collective is the broadcast of the <=160-byte coefficient block).
Step   : one pass of the hot path over one batch of IMAGES_PER_STEP distinct
         synthetic images per GPU (each its own kernel launch through the C
         ABI).  The batch is 2 x IMAGES_PER_STEP x 64 MiB >> the 126 MB L2, so
         every launch reads its source from HBM.
value  : device-timed, inputs already resident in HBM.
e2e    : the same metric through the public Python API
         (discorpy_b200.post.postprocessing.unwarp_image_backward -> C ABI
         dcb_unwarp_image_backward_host_f32) with host buffers: per image a
         64 MiB host->device copy from pinned memory, the kernel, and a 64 MiB
         device->host copy of the result, all inside the timed region (the
         library pipelines the three in row bands).
extras : secondary device-timed figures that explain the headline: the other
         blend modes of the single-image kernel and the Z-stack kernel on the
         same 4096^2 geometry (geometry evaluated once per tile and reused for
         every slice -- the path unwarp_chunk_slices_backward takes).
--impl reference : the reference's own CPU code path (NumPy float64
         coordinate temporaries + scipy.ndimage.map_coordinates, restated in
         oracle/oracle_np.py because /root/reference does not travel to the GPU
         box), split over row blocks on all host cores; one image per step.

PyTorch is used only for torch.distributed plumbing when N > 1.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 4096
XC, YC = 2050.37, 2040.81
COEF_DOT_05 = [1.00227490554, -2.99523692178e-05, 8.99519088e-08,
               -1.57066461911e-10, 8.08880211618e-14]
FACT = [COEF_DOT_05[i] / 3.0 ** i for i in range(5)]       # SURVEY.md 8d, cfg 2
IMAGES_PER_STEP = 16
E2E_IMAGES_PER_STEP = 4
METRIC = "Mpixels/s backward unwarp 4096x4096 fp32"
UNIT = "Mpixels/s"
WORKLOAD = ("configs[1]: single 4096x4096 fp32 synthetic image, 5-term backward "
            "polynomial, order 1, 1xB200 per rank")
ALGO_BYTES_PER_PX = 8.0          # read each source pixel once + write each output pixel once


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ---------------------------------------------------------------------------
# clocks during the timed region (NVML in a thread; nvidia-smi CSV fallback)
# ---------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.mask = 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.handle).gpu
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle) \
                    if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((mhz, util))
                self.mask |= int(mask)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        mhz = sorted(m for m, _ in self.samples)
        reasons = [name for bit, name in self.REASONS.items()
                   if self.mask & bit and name != "gpu_idle"]
        return {"sm_mhz": float(mhz[len(mhz) // 2]), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(mhz)}


# ---------------------------------------------------------------------------
# CPU leg: the reference's code path on all host cores
# ---------------------------------------------------------------------------
_CPU_IMG = None


def _cpu_rows(args):
    from oracle import oracle_np
    row0, nrows = args
    return oracle_np.unwarp_rows_scipy(_CPU_IMG, XC, YC, FACT, row0, nrows)


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuReference:
    """One 4096^2 image per call, row blocks spread over a fork pool."""

    def __init__(self, cores=None, size=H):
        import multiprocessing as mp
        global _CPU_IMG
        self.cores = cores or _host_cores()
        self.size = size
        rng = np.random.default_rng(2)
        _CPU_IMG = rng.random((size, W), dtype=np.float32)
        nblk = max(self.cores * 4, 1)
        edges = np.linspace(0, size, nblk + 1).astype(int)
        self.blocks = [(int(a), int(b - a)) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        self.pool = mp.get_context("fork").Pool(self.cores) if self.cores > 1 else None

    def run_once(self):
        t0 = time.perf_counter()
        if self.pool is None:
            parts = [_cpu_rows(b) for b in self.blocks]
        else:
            parts = self.pool.map(_cpu_rows, self.blocks)
        out = np.concatenate(parts)
        dt = time.perf_counter() - t0
        assert out.shape == (self.size, W)
        return dt

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    ref = CpuReference()
    for _ in range(args.warmup):
        ref.run_once()
    times = [ref.run_once() for _ in range(args.steps)]
    ref.close()
    total = sum(times)
    mpix = H * W / 1e6
    value = mpix * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": 1,
                   "note": "host CPU only; N GPUs are not used by the reference"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": "port",
                         "sample": "%d x one 4096x4096 image, NumPy coordinates + "
                                   "scipy.ndimage.map_coordinates over %d row blocks in a "
                                   "%d-process pool; %s" % (args.steps, len(ref.blocks),
                                                            ref.cores, cpu_model_name())},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_gpu_arm(args, rank, local_rank, world):
    import ctypes
    import discorpy_b200 as dcb
    from discorpy_b200 import _cabi, multigpu
    import discorpy_b200.post.postprocessing as post

    dist = None
    if world > 1:
        import torch
        import torch.distributed as tdist
        torch.cuda.set_device(local_rank)
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = tdist
    dcb.set_device(local_rank)
    numa_cores = dcb.bind_host_to_device(local_rank) if world > 1 else None   # before any pinned allocation
    post.config["blend"] = {"exact": dcb.BLEND_EXACT, "lerp64": dcb.BLEND_LERP64,
                            "lerp32": dcb.BLEND_LERP32}[args.blend]
    post.config["path"] = {"auto": dcb.PATH_AUTO, "direct": dcb.PATH_DIRECT,
                           "tma": dcb.PATH_TMA}[args.path]

    # the one collective of the path: coefficient block from rank 0 (NCCL / NVLink)
    params = dict(xcenter=XC, ycenter=YC, list_fact=FACT) if rank == 0 else None
    if dist is not None:
        params = multigpu.broadcast_params(params, src=0)
    else:
        params = multigpu.unpack_params(multigpu.pack_params(**params))
    xc, yc, fact = params["xcenter"], params["ycenter"], params["list_fact"]

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
        dcb.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = dcb.current_stream()
    nimg = args.images_per_step
    srcs = [dcb.DeviceArray((H, W)).fill_synthetic(seed=2 + rank, offset=i * H * W)
            for i in range(nimg)]
    dsts = [dcb.DeviceArray((H, W)) for _ in range(nimg)]
    model = _cabi.make_radial(xc, yc, fact)
    opt = _cabi.make_options(1, post.config["blend"], post.config["path"])
    fn = _cabi.load().dcb_unwarp_image_backward_f32
    sh = ctypes.c_void_p(stream.handle)

    def step():
        for s, d in zip(srcs, dsts):
            rc = fn(ctypes.c_void_p(s.ptr), ctypes.c_void_p(d.ptr), H, W, s.pitch, d.pitch,
                    ctypes.byref(model), ctypes.byref(opt), sh)
            if rc != 0:
                _cabi.check(rc)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    if sampler:
        sampler.start()
    launches0 = dcb.launch_count()
    e0, e1 = dcb.Event(), dcb.Event()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    e1.sync()
    barrier()
    launches = dcb.launch_count() - launches0
    ms_local = e0.elapsed_ms(e1)
    ms = max_over_ranks(ms_local)
    plan = dcb.last_plan()
    clocks = sampler.stop() if sampler else None     # NVML polling must not sit in the e2e loop

    # ---- end to end through the public API, host buffers ----------------------
    e2e_n = args.e2e_images_per_step
    host_in = []
    for i in range(min(e2e_n, 4)):
        a = dcb.pinned_empty((H, W), np.float32)
        srcs[i % nimg].to_host(out=a)
        host_in.append(a)
    # warm-up in the timed loop's own pattern: one result is still referenced while the next
    # call runs, so BOTH pinned output blocks the loop alternates between exist before the
    # clock starts (a cold cudaHostAlloc of 64 MiB takes ~28 ms, tools/e2e_probe.py)
    out = None
    for i in range(max(3, args.warmup)):
        out = post.unwarp_image_backward(host_in[i % len(host_in)], xc, yc, fact)
    barrier()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    checksum = 0.0
    for s in range(e2e_steps):
        for i in range(e2e_n):
            out = post.unwarp_image_backward(host_in[(s * e2e_n + i) % len(host_in)], xc, yc, fact)
            checksum += float(out[17, 33])
    dcb.synchronize()
    e2e_s_local = time.perf_counter() - t0
    e2e_s = max_over_ranks(e2e_s_local)
    barrier()

    # ---- secondary device-timed figures (explain the headline, do not replace it) ----
    extras = {}
    if rank == 0 and not args.no_extras:
        def time_images(blend, order=1, reps=2):
            o = _cabi.make_options(order, blend, post.config["path"])
            def once():
                for s_, d_ in zip(srcs, dsts):
                    _cabi.check(fn(ctypes.c_void_p(s_.ptr), ctypes.c_void_p(d_.ptr), H, W, s_.pitch,
                                   d_.pitch, ctypes.byref(model), ctypes.byref(o), sh))
            once()
            a, b = dcb.Event(), dcb.Event()
            a.record(stream)
            for _ in range(reps):
                once()
            b.record(stream)
            b.sync()
            return a.elapsed_ms(b) * 1e3 / (reps * nimg)        # us per image
        extras["single_image_kernel_us"] = {
            "exact": time_images(dcb.BLEND_EXACT), "lerp64": time_images(dcb.BLEND_LERP64),
            "lerp32": time_images(dcb.BLEND_LERP32), "order0": time_images(dcb.BLEND_EXACT, 0)}
        # the same 4096^2 geometry as a Z-stack of slices sharing the model (a3, the path
        # unwarp_chunk_slices_backward takes): geometry evaluated once per tile
        depth = args.stack_depth
        stack = dcb.DeviceArray((depth, H, W)).fill_synthetic(seed=2)
        sout = dcb.DeviceArray((depth, H, W))
        sfn = _cabi.load().dcb_unwarp_stack_backward_f32
        stack_us = {}
        for name, blend, order in (("exact", dcb.BLEND_EXACT, 1), ("lerp32", dcb.BLEND_LERP32, 1),
                                   ("order0", dcb.BLEND_EXACT, 0)):
            o = _cabi.make_options(order, blend, post.config["path"])
            def once():
                _cabi.check(sfn(ctypes.c_void_p(stack.ptr), ctypes.c_void_p(sout.ptr), depth, H, W, 0,
                                H, stack.pitch, stack.slice_stride, sout.pitch, sout.slice_stride, 0,
                                H, 1, ctypes.byref(model), ctypes.byref(o), sh))
            once()
            a, b = dcb.Event(), dcb.Event()
            a.record(stream)
            for _ in range(3):
                once()
            b.record(stream)
            b.sync()
            stack_us[name] = a.elapsed_ms(b) * 1e3 / (3 * depth)  # us per 4096^2 slice
        extras["stack_kernel_us_per_slice"] = stack_us
        extras["stack_depth"] = depth
        peak_, _ = measured_peak()
        extras["roofline_frac"] = {
            "single_image": {k: ALGO_BYTES_PER_PX * H * W / (v * 1e-6) / 1e9 / peak_
                             for k, v in extras["single_image_kernel_us"].items()},
            "stack": {k: ALGO_BYTES_PER_PX * H * W / (v * 1e-6) / 1e9 / peak_
                      for k, v in stack_us.items()}}
        del stack, sout

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    mpix_img = H * W / 1e6
    value = world * nimg * args.steps * mpix_img / (ms * 1e-3)
    kernel_s = (ms_local * 1e-3) / (nimg * args.steps)
    peak, peak_kind = measured_peak()
    achieved = ALGO_BYTES_PER_PX * H * W / kernel_s / 1e9
    e2e_value = world * e2e_n * e2e_steps * mpix_img / e2e_s
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step_per_gpu": nimg,
                   "l2": "inputs larger than L2: %d distinct 64 MiB source/destination pairs "
                         "per step (%.1f GiB) vs 126 MB L2" % (nimg, 2 * nimg * 64 / 1024.0),
                   "blend": args.blend, "path": args.path,
                   "plan": plan, "collective": "one broadcast of the coefficient block"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": e2e_n * H * W * 4, "d2h_bytes_per_step": e2e_n * H * W * 4,
                "steps": e2e_steps, "images_per_step_per_gpu": e2e_n,
                "host_cores_bound_per_rank": len(numa_cores) if numa_cores else None,
                "api": "discorpy_b200.post.postprocessing.unwarp_image_backward(pinned ndarray)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": TRAFFIC_BYTES_PER_LAUNCH,
                     "traffic_source": TRAFFIC_SOURCE,
                     "peak_kind": peak_kind,
                     "kernel": "remap_image_kernel<RADIAL, order 1, blend %s, 5 terms>" % args.blend,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX * H * W,
                     "kernel_us": kernel_s * 1e6},
    }
    if extras:
        line["extras"] = extras
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference()
        ref.run_once()
        times = []
        t_begin = time.perf_counter()
        while len(times) < 3 or (time.perf_counter() - t_begin < 10 and len(times) < 20):
            times.append(ref.run_once())
        ref.close()
        best = min(times)
        line["cpu_baseline"] = {
            "value": mpix_img / best, "unit": UNIT, "cores": ref.cores, "kind": "port",
            "sample": "%d x one 4096x4096 image (best taken), NumPy coordinates + "
                      "scipy.ndimage.map_coordinates over %d row blocks in a %d-process pool; %s"
                      % (len(times), len(ref.blocks), ref.cores, cpu_model_name())}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant
# kernel, from the committed `ncu --set full` capture (profiles/); None until
# a capture exists for the current kernel.
TRAFFIC_BYTES_PER_LAUNCH = 69821696 + 17368832
TRAFFIC_SOURCE = ("profiles/r1/ncu_image_v14.txt: dram__bytes_read.sum 69.8 MB + "
                  "dram__bytes_write.sum 17.4 MB of one launch (most of the 64 MiB output is "
                  "still dirty in the 126 MB L2 when the profiled launch ends)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blend", default="exact", choices=["exact", "lerp64", "lerp32"])
    ap.add_argument("--path", default="auto", choices=["auto", "direct", "tma"])
    ap.add_argument("--images-per-step", type=int, default=IMAGES_PER_STEP)
    ap.add_argument("--e2e-images-per-step", type=int, default=E2E_IMAGES_PER_STEP)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary figures (blend variants, Z-stack kernel)")
    ap.add_argument("--stack-depth", type=int, default=32)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: relaunch under torchrun
        import subprocess
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(port)] + sys.argv
        sys.exit(subprocess.call(cmd))
    run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric : Mpixels/s of backward unwarp, 4096 x 4096 float32, 5-term polynomial,
         order 1 (BASELINE configs[1]), whole job over N GPUs (weak scaling:
         every rank unwarps its own batch of independent images; the only
         collective is the broadcast of the <=256-byte coefficient block).
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
         library pipelines the three in row bands).  `e2e_pageable` is the same
         call with ordinary (pageable) NumPy arrays in and out -- what a drop-in
         caller passes; `pcie` holds the copy rates measured on the spot.
extras : secondary device-timed figures: the other blend modes of the
         single-image kernel, the Z-stack kernel on the same 4096^2 geometry,
         config 3 (2048^2 radial -> perspective), and on every N the two
         sharded workloads north_star names -- config 4 (2048 x 2560^2 stack,
         slice semantics, Z-sharded: strong scaling) and config 5 (64 x 8192^2,
         9 terms, 64/N images per rank) -- each with an on-device parity spot
         check against the oracle.  For N > 1 `exchange` runs the one optional
         exchange step (one sinogram assembled from all ranks) in its fused
         (peer stores) and collective (NCCL all-gather) forms against the oracle.
--impl reference : the reference's own CPU code path (NumPy float64
         coordinate temporaries + scipy.ndimage.map_coordinates, restated in
         oracle/oracle_np.py because /root/reference does not travel to the GPU
         box), split over row blocks on all host cores; one image per step.

No PyTorch: for N > 1 the launcher's environment (RANK, LOCAL_RANK, WORLD_SIZE,
MASTER_ADDR, MASTER_PORT) is read by discorpy_b200.multigpu.NcclComm, NCCL
through the C ABI.
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


def config_dict(blend="exact"):
    """The `config` object -- identical in both arms so that the driver can compare them."""
    return {"workload": WORKLOAD, "order": 1, "blend": blend,
            "l2": "GPU arm: inputs larger than L2 (%d distinct 64 MiB source/destination pairs "
                  "per step per GPU = %.1f GiB, vs 126 MB L2); reference arm: host CPU only"
                  % (IMAGES_PER_STEP, 2 * IMAGES_PER_STEP * 64 / 1024.0),
            "collective": "one broadcast of the coefficient block (NCCL through the C ABI)"}


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
        "config": config_dict(),
        "detail": {"images_per_step": 1,
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
FACT3 = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]                          # SURVEY.md 8d, cfg 3 / 4
PERS3 = [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]
FACT9 = [1.0, -1e-5, 3e-8, -2e-11, 5e-15, -8e-19, 6e-23, -2e-27, 3e-32]   # cfg 5


def _event_ms(dcb, stream, fn, reps):
    fn()
    a, b = dcb.Event(), dcb.Event()
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    b.sync()
    return a.elapsed_ms(b) / reps


def pcie_probe(dcb, _cabi, ctypes, nbytes=64 << 20, reps=6):
    """Pinned host <-> device copy rates on this box: each direction alone and both at once."""
    h_in = dcb.pinned_empty((nbytes // 4,), np.float32)
    h_out = dcb.pinned_empty((nbytes // 4,), np.float32)
    h_in[:] = 1.0
    d_a, d_b = dcb.DeviceArray((1, nbytes // 4)), dcb.DeviceArray((1, nbytes // 4))
    s1, s2 = dcb.Stream(), dcb.Stream()

    def h2d(st):
        _cabi.call("dcb_h2d", ctypes.c_void_p(d_a.ptr), ctypes.c_void_p(h_in.ctypes.data), nbytes,
                   ctypes.c_void_p(st.handle))

    def d2h(st):
        _cabi.call("dcb_d2h", ctypes.c_void_p(h_out.ctypes.data), ctypes.c_void_p(d_b.ptr), nbytes,
                   ctypes.c_void_p(st.handle))

    def wall(fn):
        fn()
        dcb.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dcb.synchronize()
        return (time.perf_counter() - t0) / reps

    t_up = wall(lambda: h2d(s1))
    t_dn = wall(lambda: d2h(s1))
    t_both = wall(lambda: (h2d(s1), d2h(s2)))
    return {"h2d_gbs": nbytes / t_up / 1e9, "d2h_gbs": nbytes / t_dn / 1e9,
            "duplex_gbs_each_way": nbytes / t_both / 1e9, "bytes_per_copy": nbytes}


def run_gpu_arm(args, rank, local_rank, world):
    import ctypes
    import discorpy_b200 as dcb
    from discorpy_b200 import _cabi, multigpu
    import discorpy_b200.post.postprocessing as post

    comm = None
    if world > 1:
        comm = multigpu.NcclComm.from_env()       # binds LOCAL_RANK's device, ncclCommInitRank
    else:
        dcb.set_device(local_rank)
    numa_cores = dcb.bind_host_to_device(local_rank) if world > 1 else None   # before any pinned allocation
    post.config["blend"] = {"exact": dcb.BLEND_EXACT, "lerp64": dcb.BLEND_LERP64,
                            "lerp32": dcb.BLEND_LERP32}[args.blend]
    post.config["path"] = {"auto": dcb.PATH_AUTO, "direct": dcb.PATH_DIRECT,
                           "tma": dcb.PATH_TMA}[args.path]

    # the one collective of the path: coefficient block from rank 0 (NCCL / NVLink)
    params = dict(xcenter=XC, ycenter=YC, list_fact=FACT) if rank == 0 else None
    params = multigpu.broadcast_params(params, src=0, comm=comm)
    xc, yc, fact = params["xcenter"], params["ycenter"], params["list_fact"]

    def barrier():
        if comm is not None:
            comm.barrier()
        dcb.synchronize()

    def max_over_ranks(x):
        return x if comm is None else comm.allreduce_max([x])[0]

    peak, peak_kind = measured_peak()
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

    # the very first launch of a (model, geometry) pair also builds its plan (remap_image.cuh):
    # time that cold call apart, it is not part of the steady-state metric
    # (a 256 x 256 call first: the process's first launch also loads the kernels' modules and
    # allocates the scheduler counters, which is not what a new calibration costs)
    tiny_s, tiny_d = dcb.DeviceArray((256, 256)).fill(0.0), dcb.DeviceArray((256, 256))
    t0 = time.perf_counter()
    _cabi.check(fn(ctypes.c_void_p(tiny_s.ptr), ctypes.c_void_p(tiny_d.ptr), 256, 256, tiny_s.pitch,
                   tiny_d.pitch, ctypes.byref(model), ctypes.byref(opt), sh))
    dcb.synchronize()
    first_call_us = (time.perf_counter() - t0) * 1e6
    dcb.plan_cache_clear()
    dcb.synchronize()
    t0 = time.perf_counter()
    _cabi.check(fn(ctypes.c_void_p(srcs[0].ptr), ctypes.c_void_p(dsts[0].ptr), H, W, srcs[0].pitch,
                   dsts[0].pitch, ctypes.byref(model), ctypes.byref(opt), sh))
    dcb.synchronize()
    cold_call_us = (time.perf_counter() - t0) * 1e6

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
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e = {}
    if args.e2e_steps > 0:
        pinned_in, page_in = [], []
        for i in range(min(e2e_n, 4)):
            a = dcb.pinned_empty((H, W), np.float32)
            srcs[i % nimg].to_host(out=a)
            pinned_in.append(a)
            page_in.append(np.array(a))                   # an ordinary (pageable) copy

        def e2e_run(inputs):
            # warm-up in the timed loop's own pattern: one result is still referenced while the
            # next call runs, so the output blocks the loop alternates between exist before the
            # clock starts (a cold cudaHostAlloc of 64 MiB takes ~28 ms, tools/e2e_probe.py)
            out = None
            for i in range(max(3, args.warmup)):
                out = post.unwarp_image_backward(inputs[i % len(inputs)], xc, yc, fact)
            barrier()
            t0 = time.perf_counter()
            checksum = 0.0
            for s in range(e2e_steps):
                for i in range(e2e_n):
                    out = post.unwarp_image_backward(inputs[(s * e2e_n + i) % len(inputs)], xc, yc, fact)
                    checksum += float(out[17, 33])
            dcb.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            barrier()
            return world * e2e_n * e2e_steps * (H * W / 1e6) / dt

        e2e["pinned"] = e2e_run(pinned_in)
        os.environ["DCB_REGISTER"] = "0"          # every array new to the library: staged copies
        e2e["pageable"] = e2e_run(page_in)
        del os.environ["DCB_REGISTER"]
        for a in page_in * 2:                     # frame buffers the caller reuses: the second
            post.unwarp_image_backward(a, xc, yc, fact)   # sighting page-locks them in place
        e2e["pageable_reused"] = e2e_run(page_in)
        del pinned_in, page_in
    pcie = pcie_probe(dcb, _cabi, ctypes) if (rank == 0 and args.e2e_steps > 0) else None

    # ---- secondary device-timed figures (explain the headline, do not replace it) ----
    extras = {}
    if rank == 0 and not args.no_extras:
        def time_images(blend, order=1, reps=2):
            o = _cabi.make_options(order, blend, post.config["path"])
            def once():
                for s_, d_ in zip(srcs, dsts):
                    _cabi.check(fn(ctypes.c_void_p(s_.ptr), ctypes.c_void_p(d_.ptr), H, W, s_.pitch,
                                   d_.pitch, ctypes.byref(model), ctypes.byref(o), sh))
            return _event_ms(dcb, stream, once, reps) * 1e3 / nimg        # us per image
        extras["single_image_kernel_us"] = {
            "exact": time_images(dcb.BLEND_EXACT), "lerp64": time_images(dcb.BLEND_LERP64),
            "lerp32": time_images(dcb.BLEND_LERP32), "order0": time_images(dcb.BLEND_EXACT, 0)}
        extras["single_image_cold_call_us"] = cold_call_us
        extras["first_call_of_process_us"] = first_call_us
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
            stack_us[name] = _event_ms(dcb, stream, once, 3) * 1e3 / depth   # us per 4096^2 slice
        extras["stack_kernel_us_per_slice"] = stack_us
        extras["stack_depth"] = depth
        extras["roofline_frac"] = {
            "single_image": {k: ALGO_BYTES_PER_PX * H * W / (v * 1e-6) / 1e9 / peak
                             for k, v in extras["single_image_kernel_us"].items()},
            "stack": {k: ALGO_BYTES_PER_PX * H * W / (v * 1e-6) / 1e9 / peak
                      for k, v in stack_us.items()}}
        del stack, sout
        extras["cfg3"] = bench_cfg3(dcb, _cabi, ctypes, stream, peak)
    del srcs, dsts
    dcb.device_pool_clear()
    if not args.no_extras:
        extras["cfg4"] = bench_cfg4(dcb, _cabi, ctypes, stream, peak, rank, world, comm, args)
        dcb.device_pool_clear()
        extras["cfg5"] = bench_cfg5(dcb, _cabi, ctypes, stream, peak, rank, world, comm, args)
        dcb.device_pool_clear()
    exchange = bench_exchange(dcb, multigpu, post, comm, rank, world) if world > 1 else None

    if rank != 0:
        if comm is not None:
            comm.close()
        return

    mpix_img = H * W / 1e6
    value = world * nimg * args.steps * mpix_img / (ms * 1e-3)
    kernel_s = (ms_local * 1e-3) / (nimg * args.steps)
    achieved = ALGO_BYTES_PER_PX * H * W / kernel_s / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.blend),
        "detail": {"images_per_step_per_gpu": nimg, "path": args.path, "plan": plan,
                   "plan_cache": "per-tile boxes and verified row patches are built once per "
                                 "(model, geometry) and reused (remap_image.cuh); the cold call "
                                 "is in extras.single_image_cold_call_us"},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": TRAFFIC_BYTES_PER_LAUNCH,
                     "traffic_source": TRAFFIC_SOURCE,
                     "peak_kind": peak_kind,
                     "kernel": "remap_image_kernel<RADIAL, order 1, blend %s, 5 terms>" % args.blend,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX * H * W,
                     "kernel_us": kernel_s * 1e6,
                     "colimit": COLIMIT},
    }
    if e2e:
        bytes_each_way = e2e_n * H * W * 4
        line["e2e"] = {"value": e2e["pinned"], "unit": UNIT,
                       "h2d_bytes_per_step": bytes_each_way, "d2h_bytes_per_step": bytes_each_way,
                       "steps": e2e_steps, "images_per_step_per_gpu": e2e_n,
                       "host_cores_bound_per_rank": len(numa_cores) if numa_cores else None,
                       "api": "discorpy_b200.post.postprocessing.unwarp_image_backward(pinned ndarray)"}
        line["e2e_pageable"] = {"value": e2e["pageable"], "unit": UNIT,
                                "api": "the same call with ordinary (pageable) numpy arrays, each "
                                       "new to the library (staged through pinned memory by a "
                                       "pool of host threads)",
                                "ratio_to_pinned": e2e["pageable"] / e2e["pinned"],
                                "reused_buffers_value": e2e["pageable_reused"],
                                "reused_buffers_ratio_to_pinned": e2e["pageable_reused"] / e2e["pinned"],
                                "reused_buffers": "the same four ordinary arrays passed again and "
                                                  "again: page-locked in place on their second "
                                                  "sighting (device.maybe_register)"}
        if pcie:
            # one image = 64 MiB each way; the duplex rate bounds the pipelined call
            pcie["e2e_gbs_each_way"] = e2e["pinned"] / world * 1e6 * 4 / 1e9
            pcie["frac"] = pcie["e2e_gbs_each_way"] / pcie["duplex_gbs_each_way"]
            pcie["bytes_per_step"] = 2 * bytes_each_way
            line["pcie"] = pcie
    if extras:
        line["extras"] = extras
    if exchange:
        line["exchange"] = exchange
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference()
        ref.run_once()
        times = []
        t_begin = time.perf_counter()
        while len(times) < 3 or (time.perf_counter() - t_begin < 10 and len(times) < 20):
            times.append(ref.run_once())
        ref.close()
        best = min(times)
        one = CpuReference(cores=1)                      # the reference as shipped: one thread
        t_single = min(one.run_once() for _ in range(2))
        one.close()
        line["cpu_baseline"] = {
            "value": mpix_img / best, "unit": UNIT, "cores": ref.cores, "kind": "port",
            "single_process_value": mpix_img / t_single,
            "sample": "%d x one 4096x4096 image (best taken), NumPy coordinates + "
                      "scipy.ndimage.map_coordinates over %d row blocks in a %d-process pool; "
                      "single_process_value: the same image in one process (the reference is "
                      "single-threaded), best of 2; %s"
                      % (len(times), len(ref.blocks), ref.cores, cpu_model_name())}
    print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()


# ---------------------------------------------------------------------------
# the other BASELINE configs (extras)
# ---------------------------------------------------------------------------
def bench_cfg3(dcb, _cabi, ctypes, stream, peak):
    """configs[2]: radial -> perspective on 2048^2 (demo_05.py:127,147), device-timed over 12
    rotating buffer sets (576 MB >> L2), against the 8 B/px a fused kernel would move."""
    from oracle import oracle_c
    n, S = 12, 2048
    srcs = [dcb.DeviceArray((S, S)).fill_synthetic(seed=3, offset=i * S * S) for i in range(n)]
    tmps = [dcb.DeviceArray((S, S)) for _ in range(n)]
    dsts = [dcb.DeviceArray((S, S)) for _ in range(n)]
    rad = _cabi.make_radial(1030.2, 1019.6, FACT3)
    per = _cabi.make_persp(PERS3)
    opt = _cabi.make_options(1, dcb.BLEND_EXACT, dcb.PATH_AUTO)
    fn = _cabi.load().dcb_unwarp_image_backward_perspective_f32
    sh = ctypes.c_void_p(stream.handle)

    def once():
        for s, t, d in zip(srcs, tmps, dsts):
            _cabi.check(fn(ctypes.c_void_p(s.ptr), ctypes.c_void_p(d.ptr), ctypes.c_void_p(t.ptr), S, S,
                           s.pitch, d.pitch, t.pitch, ctypes.byref(rad), ctypes.byref(per),
                           ctypes.byref(opt), sh))
    us = _event_ms(dcb, stream, once, 3) * 1e3 / n
    src0 = srcs[0].to_host()
    want = oracle_c.correct_perspective_image(
        oracle_c.unwarp_image_backward(src0, 1030.2, 1019.6, FACT3, 1), PERS3, 1)
    bad = int(np.count_nonzero(dsts[0].to_host() != want))
    # end to end: the public call on pinned host arrays (banded two-stage host pipeline)
    import discorpy_b200.post.postprocessing as post
    pins = []
    for i in range(4):
        a = dcb.pinned_empty((S, S), np.float32)
        a[:] = src0 if i == 0 else srcs[i].to_host()
        pins.append(a)
    got = post.unwarp_image_backward_perspective(pins[0], 1030.2, 1019.6, FACT3, PERS3)
    bad_e2e = int(np.count_nonzero(got != want))
    for i in range(3):
        post.unwarp_image_backward_perspective(pins[i], 1030.2, 1019.6, FACT3, PERS3)
    t0 = time.perf_counter()
    for i in range(16):
        post.unwarp_image_backward_perspective(pins[i % 4], 1030.2, 1019.6, FACT3, PERS3)
    e2e_ms = (time.perf_counter() - t0) / 16 * 1e3
    return {"workload": "configs[2]: perspective+radial combined unwarp, 2048x2048 fp32",
            "e2e_ms_per_image": e2e_ms, "e2e_Mpixels_per_s": S * S / (e2e_ms * 1e3),
            "e2e_api": "post.unwarp_image_backward_perspective(pinned ndarray): upload, both passes and "
                       "download in row bands (dcb_unwarp_image_backward_perspective_host_f32)",
            "e2e_parity_ok": bad_e2e == 0,
            "kernel_us_per_image": us, "Mpixels_per_s": S * S / us,
            "roofline_frac_vs_8B_per_px": ALGO_BYTES_PER_PX * S * S / (us * 1e-6) / 1e9 / peak,
            "launches_per_image": 2, "hbm_bytes_per_px_moved": 16,
            "parity_ok": bad == 0, "pixels_differing_from_oracle": bad}


def bench_cfg4(dcb, _cabi, ctypes, stream, peak, rank, world, comm, args):
    """configs[3]: 2048 slices x 2560^2, unwarp_slice_backward semantics (float64 coordinates),
    every output row of every slice, the stack Z-sharded over the ranks (strong scaling); the
    shard is generated on the device (stateless hash), nothing crosses PCIe or NVLink."""
    from discorpy_b200 import multigpu
    from oracle import oracle_np
    D, S = args.cfg4_depth, 2560
    xc, yc = 1283.4, 1275.9
    lo, hi = multigpu.shard_range(D, rank, world)
    nz = hi - lo
    stack = dcb.DeviceArray((nz, S, S)).fill_synthetic(seed=4, offset=lo * S * S)
    out = dcb.DeviceArray((nz, S, S))
    model = _cabi.make_radial(xc, yc, FACT3)
    opt = _cabi.make_options(1, dcb.BLEND_EXACT, dcb.PATH_AUTO)
    sfn = _cabi.load().dcb_unwarp_stack_backward_f32
    sh = ctypes.c_void_p(stream.handle)

    def once():
        _cabi.check(sfn(ctypes.c_void_p(stack.ptr), ctypes.c_void_p(out.ptr), nz, S, S, 0, S,
                        stack.pitch, stack.slice_stride, out.pitch, out.slice_stride, 0, S, 0,
                        ctypes.byref(model), ctypes.byref(opt), sh))
    if comm is not None:
        comm.barrier()
    ms_local = _event_ms(dcb, stream, once, 2)
    ms = ms_local if comm is None else comm.allreduce_max([ms_local])[0]
    # parity: D' = 8 slices of this rank's shard on the host, three sinograms against the oracle
    dsub = min(8, nz)
    sub = dcb.DeviceArray((dsub, S, S))
    _cabi.call("dcb_d2d", ctypes.c_void_p(sub.ptr), ctypes.c_void_p(stack.ptr), dsub * stack.slice_stride, sh)
    osub = dcb.DeviceArray((dsub, S, S))
    _cabi.call("dcb_d2d", ctypes.c_void_p(osub.ptr), ctypes.c_void_p(out.ptr), dsub * out.slice_stride, sh)
    host, got = sub.to_host(), osub.to_host()
    worst, cpu_s = 0.0, 0.0
    for index in (3, S // 2 + 7, S - 2):
        t0 = time.perf_counter()
        want = oracle_np.unwarp_slice_backward(host, xc, yc, FACT3, index)
        cpu_s += time.perf_counter() - t0
        worst = max(worst, float(np.max(np.abs(got[:, index, :] - want))))
    worst = worst if comm is None else comm.allreduce_max([worst])[0]
    res = {"workload": "configs[3]: 3D stack %d slices x 2560x2560 fp32, unwarp_slice_backward "
                       "semantics, every row, Z-sharded over %d GPU(s)" % (D, world),
           "scaling": "strong", "slices_per_gpu": nz, "ms": ms,
           "Mpixels_per_s": D * S * S / 1e6 / (ms * 1e-3),
           "roofline_frac_per_gpu": ALGO_BYTES_PER_PX * nz * S * S / (ms_local * 1e-3) / 1e9 / peak,
           "parity_ok": worst <= 1e-5, "max_abs_diff_vs_oracle": worst,
           "parity_sample": "3 sinograms x %d slices per rank, tolerance 1e-5" % dsub}
    if world == 1:
        # the reference loops `depth` times per index (postprocessing.py:226-228): linear in both
        res["cpu_scaled_s"] = cpu_s / (3 * dsub) * D * S
        res["cpu_scaling"] = ("oracle unwarp_slice_backward on %d slices x 3 indices, one process, "
                              "scaled by (D / %d) x (2560 / 3)" % (dsub, dsub))
    return res


def bench_cfg5(dcb, _cabi, ctypes, stream, peak, rank, world, comm, args):
    """configs[4]: 64 images of 8192^2, 9-term fisheye-strength model, 64/N images per rank
    through the Z-stack kernel (the images of a batch share the model)."""
    from discorpy_b200 import multigpu
    from oracle import oracle_c
    B, S = args.cfg5_batch, 8192
    xc, yc = 4100.3, 4090.8
    lo, hi = multigpu.shard_range(B, rank, world)
    nz = hi - lo
    batch = dcb.DeviceArray((nz, S, S)).fill_synthetic(seed=5, offset=lo * S * S)
    out = dcb.DeviceArray((nz, S, S))
    model = _cabi.make_radial(xc, yc, FACT9)
    opt = _cabi.make_options(1, dcb.BLEND_EXACT, dcb.PATH_AUTO)
    sfn = _cabi.load().dcb_unwarp_stack_backward_f32
    sh = ctypes.c_void_p(stream.handle)

    def once():
        _cabi.check(sfn(ctypes.c_void_p(batch.ptr), ctypes.c_void_p(out.ptr), nz, S, S, 0, S,
                        batch.pitch, batch.slice_stride, out.pitch, out.slice_stride, 0, S, 1,
                        ctypes.byref(model), ctypes.byref(opt), sh))
    if comm is not None:
        comm.barrier()
    ms_local = _event_ms(dcb, stream, once, 2)
    ms = ms_local if comm is None else comm.allreduce_max([ms_local])[0]
    # parity: this rank's first image against the C oracle, bit for bit
    one_in, one_out = dcb.DeviceArray((S, S)), dcb.DeviceArray((S, S))
    _cabi.call("dcb_d2d", ctypes.c_void_p(one_in.ptr), ctypes.c_void_p(batch.ptr), batch.slice_stride, sh)
    _cabi.call("dcb_d2d", ctypes.c_void_p(one_out.ptr), ctypes.c_void_p(out.ptr), out.slice_stride, sh)
    src = one_in.to_host()
    t0 = time.perf_counter()
    want = oracle_c.unwarp_image_backward(src, xc, yc, FACT9, 1, nthreads=max(1, _host_cores() // world))
    cpu_s = time.perf_counter() - t0
    bad = float(np.count_nonzero(one_out.to_host() != want))
    bad = bad if comm is None else comm.allreduce_max([bad])[0]
    res = {"workload": "configs[4]: fisheye-strength 9-term polynomial, 8192x8192 fp32 batch=%d, "
                       "%d image(s) per GPU over %d GPU(s)" % (B, nz, world),
           "scaling": "strong", "images_per_gpu": nz, "ms": ms,
           "Mpixels_per_s": B * S * S / 1e6 / (ms * 1e-3),
           "roofline_frac_per_gpu": ALGO_BYTES_PER_PX * nz * S * S / (ms_local * 1e-3) / 1e9 / peak,
           "parity_ok": bad <= 16, "pixels_differing_from_oracle": int(bad),
           "parity_sample": "first image of every rank, bit for bit (<= 16 float32-coordinate "
                            "flips per 67 Mpixel allowed, tests/test_gpu_parity.py)"}
    if world == 1:
        res["cpu_scaled_s"] = cpu_s * B
        res["cpu_scaling"] = ("C restatement of the reference arithmetic (oracle_c, %d threads) on "
                              "one image x %d" % (max(1, _host_cores() // world), B))
    return res


def bench_exchange(dcb, multigpu, post, comm, rank, world):
    """The one optional exchange step: ONE sinogram of a Z-sharded stack on rank 0,
    fused (every rank's remap kernel stores into rank 0's buffer over NVLink) and collective
    (local rows + NCCL all-gather through the C ABI), both against the oracle on rank 0."""
    from discorpy_b200.device import DeviceArray, synthetic_host
    from oracle import oracle_np
    D, S = 8 * world + 5, 640                  # uneven shards on purpose
    scale = 2560.0 / S
    params = dict(xcenter=S / 2 + 3.4, ycenter=S / 2 - 4.1,
                  list_fact=[FACT3[i] * scale ** i for i in range(5)], list_coef=[])
    lo, hi = multigpu.shard_range(D, rank, world)
    shard = DeviceArray((hi - lo, S, S)).fill_synthetic(seed=4, offset=lo * S * S)
    window = multigpu.SinogramWindow(D, S, owner=0, comm=comm)
    full = synthetic_host(D * S * S, seed=4).reshape(D, S, S) if rank == 0 else None
    fused_bad = coll_bad = 0
    rows = DeviceArray((hi - lo, S))
    for index in (0, S // 3, S - 1):
        multigpu.unwarp_slice_backward_sharded(shard, params, index, window)
        window.fence()
        post._unwarp_slice_into(shard, params["xcenter"], params["ycenter"], params["list_fact"],
                                index, multigpu._Rows(rows.ptr, rows.pitch, rows.shape))
        gathered = multigpu.gather_rows(rows, D, comm=comm)
        if rank == 0:
            want = oracle_np.unwarp_slice_backward(full, params["xcenter"], params["ycenter"],
                                                   params["list_fact"], index)
            fused_bad += int(np.count_nonzero(window.array.to_host() != want))
            coll_bad += int(np.count_nonzero(gathered.to_host() != want))
        comm.barrier()

    def timed(fn, reps=20):
        fn()
        comm.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        comm.barrier()
        return comm.allreduce_max([(time.perf_counter() - t0) / reps])[0] * 1e3

    def fused():
        multigpu.unwarp_slice_backward_sharded(shard, params, S // 2, window)
        window.fence()

    def collective():
        post._unwarp_slice_into(shard, params["xcenter"], params["ycenter"], params["list_fact"],
                                S // 2, multigpu._Rows(rows.ptr, rows.pitch, rows.shape))
        multigpu.gather_rows(rows, D, comm=comm)
        dcb.current_stream().sync()

    res = {"fused_ms": timed(fused), "collective_ms": timed(collective),
           "fused_ok": fused_bad == 0, "collective_ok": coll_bad == 0,
           "stack": "%d slices of %dx%d over %d ranks, 3 indices checked bit for bit against the "
                    "oracle on rank 0" % (D, S, S, world)}
    window.close()
    return res


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant
# kernel, from the committed `ncu --set full` capture (profiles/); None until
# a capture exists for the current kernel.
TRAFFIC_BYTES_PER_LAUNCH = 78657280 + 18953216
TRAFFIC_SOURCE = ("profiles/r2/ncu_image_r2zp_exact.txt: dram__bytes_read.sum 78.7 MB + "
                  "dram__bytes_write.sum 19.0 MB of one launch (most of the 64 MiB output is "
                  "still dirty in the 126 MB L2 when the profiled launch ends)")
# What the SM side of one launch costs at 100 % of each pipe (us), from the instruction mix of
# the committed capture (profiles/r2/ncu_image_r2zp_exact.txt: thread instructions per pixel by
# pipe) and the pipe rates measured in round 1 (profiles/r1/microbench_*.txt: fp64 60.1 and XU
# 15.6 thread-ops per clock per SM, issue 128): 16.78 Mpx / 148 SMs / 1.965 GHz x ops / rate.
# issue_fp64_dispatch_us is the co-limit the ablation runs of round 2 point to
# (profiles/r2/ablation_raw2_r2a1.txt): an fp64 instruction holds a sub-partition's issue port for
# its whole dispatch -- 2.13 cycles, 3.07 with three distinct source registers -- so the 17.6 fp64
# operations per pixel (8 of them three-register forms) cost 45 issue cycles, the other 38.8
# instructions one each.
def _colimit(ops_per_px, rate):
    return H * W / 148.0 / 1965.0 * ops_per_px / rate          # us


COLIMIT = {"fp64_us": _colimit(17.6, 60.1), "xu_us": _colimit(0.2, 15.6),
           "issue_us": _colimit(56.4, 128.0),
           "issue_fp64_dispatch_us": _colimit(38.8 + 8 * 3.07 + 9.6 * 2.13, 128.0),
           "source": "profiles/r2/ncu_image_r2zp_exact.txt instruction mix; pipe rates "
                     "profiles/r1/microbench_v1.txt, microbench_fp64_operand_forms.txt; "
                     "profiles/r2/ablation_raw2_r2a1.txt"}


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
    ap.add_argument("--cfg4-depth", type=int, default=2048)
    ap.add_argument("--cfg5-batch", type=int, default=64)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: one process per GPU, the launcher's environment by hand
        import subprocess
        port = 29500 + (os.getpid() % 2000)
        procs = []
        for r in range(args.gpus):
            env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(args.gpus),
                       MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
            procs.append(subprocess.Popen([sys.executable] + sys.argv, env=env))
        sys.exit(max(p.wait() for p in procs))
    run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

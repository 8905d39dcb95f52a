"""GPU: the streaming Z-stack entry (discorpy_b200/post/streaming.py) returns exactly
what unwarp_chunk_slices_backward returns (same kernel, same window) and what the
oracle computes, for in-memory, memory-mapped and pinned sources."""
import os

import numpy as np
import pytest

from oracle import oracle_np as orc

import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
from discorpy_b200.post import streaming

pytestmark = pytest.mark.gpu
FACT = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]


@pytest.mark.parametrize("dtype", ["float32", "uint16"])
@pytest.mark.parametrize("spb", [1, 3, 7, 64])
def test_stream_equals_chunk_call_and_oracle(dtype, spb):
    rng = np.random.default_rng(7)
    d, h, w = 11, 90, 133                      # W % 4 != 0: pitched device rows
    if dtype == "float32":
        stack = rng.random((d, h, w), dtype=np.float32)
    else:
        stack = rng.integers(0, 65535, (d, h, w), dtype=np.uint16)
    xc, yc, start, stop = 66.3, 44.1, 10, 71
    want = post.unwarp_chunk_slices_backward(stack, xc, yc, FACT, start, stop)
    got = streaming.unwarp_chunk_slices_backward_stream(stack, xc, yc, FACT, start, stop,
                                                        slices_per_block=spb)
    assert got.dtype == want.dtype and np.array_equal(got, want)
    ref = orc.unwarp_chunk_slices_backward(stack, xc, yc, FACT, start, stop)
    assert np.array_equal(got, ref)


def test_stream_from_memmap_into_memmap(tmp_path):
    rng = np.random.default_rng(8)
    d, h, w = 9, 64, 128
    src = np.lib.format.open_memmap(tmp_path / "in.npy", mode="w+", dtype=np.float32,
                                    shape=(d, h, w))
    src[:] = rng.random((d, h, w), dtype=np.float32)
    src.flush()
    src = np.load(tmp_path / "in.npy", mmap_mode="r")
    dst = np.lib.format.open_memmap(tmp_path / "out.npy", mode="w+", dtype=np.float32,
                                    shape=(d, h, w))
    out = streaming.unwarp_chunk_slices_backward_stream(src, 63.7, 30.2, FACT, out=dst,
                                                        slices_per_block=4)
    assert out is dst
    want = orc.unwarp_chunk_slices_backward(np.asarray(src), 63.7, 30.2, FACT, 0, h - 1)
    assert np.array_equal(np.asarray(dst), want)


def test_stream_pinned_direct_dma():
    rng = np.random.default_rng(9)
    d, h, w = 10, 72, 256
    src = dcb.pinned_empty((d, h, w), np.float32)
    src[:] = rng.random((d, h, w), dtype=np.float32)
    dst = dcb.pinned_empty((d, 41, w), np.float32)
    streaming.unwarp_chunk_slices_backward_stream(src, 120.5, 35.5, FACT, 20, 60, out=dst,
                                                  slices_per_block=3)
    want = orc.unwarp_chunk_slices_backward(np.asarray(src), 120.5, 35.5, FACT, 20, 60)
    assert np.array_equal(np.asarray(dst), want)


def test_stream_argument_errors():
    with pytest.raises(ValueError, match="3D"):
        streaming.unwarp_chunk_slices_backward_stream(np.zeros((4, 4), np.float32), 1, 1, [1.0])
    st = np.zeros((2, 8, 8), np.float32)
    with pytest.raises(ValueError, match="out of the range"):
        streaming.unwarp_chunk_slices_backward_stream(st, 4, 4, [1.0], 0, 8)
    with pytest.raises(NotImplementedError):
        streaming.unwarp_chunk_slices_backward_stream(st.astype(np.float64), 4, 4, [1.0])


@pytest.mark.gpu
def test_sinogram_exchange_two_ranks():
    """SURVEY.md 8e, the optional exchange step: two ranks (one per GPU) assemble one sinogram on
    rank 0 by peer stores of the remap kernel and by an NCCL all-gather; both bit-identical to the
    oracle (tests/two_rank_sinogram_check.py).  Needs two GPUs on the box."""
    import subprocess
    import sys
    import discorpy_b200
    if discorpy_b200.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29617",
           os.path.join(root, "tests", "two_rank_sinogram_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, (res.stdout[-1500:], res.stderr[-1500:])
    assert '"ok": true' in res.stdout


@pytest.mark.gpu
def test_sinogram_window_single_rank():
    """The fused sinogram path with one rank (no process group): the window is a plain device buffer
    and unwarp_slice_backward_sharded must equal unwarp_slice_backward and the oracle."""
    from discorpy_b200 import multigpu
    rng = np.random.default_rng(11)
    stack = rng.random((9, 120, 200), dtype=np.float32)
    params = dict(xcenter=101.3, ycenter=58.6, list_fact=FACT)
    window = multigpu.SinogramWindow(9, 200, owner=0)
    assert window.rows == (0, 9) and window.world == 1
    shard = dcb.DeviceArray.from_host(stack)
    for index in (0, 57, 119):
        multigpu.unwarp_slice_backward_sharded(shard, params, index, window)
        window.fence()
        got = window.array.to_host()
        want = orc.unwarp_slice_backward(stack, 101.3, 58.6, FACT, index)
        assert np.array_equal(got, want)
        assert np.array_equal(got, post.unwarp_slice_backward(stack, 101.3, 58.6, FACT, index))
    with pytest.raises(ValueError):
        multigpu.unwarp_slice_backward_sharded(dcb.DeviceArray((4, 120, 200)), params, 3, window)
    window.close()


@pytest.mark.gpu
def test_chunk_function_delegates_large_host_stacks_to_the_pipeline():
    """unwarp_chunk_slices_backward hands large float32 host stacks to the streaming pipeline
    (config["stream_bytes"]); same numbers either way, and integer stacks keep the device-side
    widening path."""
    rng = np.random.default_rng(5)
    stack = rng.random((9, 96, 160), dtype=np.float32)
    want = orc.unwarp_chunk_slices_backward(stack, 83.7, 44.2, FACT, 10, 80)
    direct = post.unwarp_chunk_slices_backward(stack, 83.7, 44.2, FACT, 10, 80)
    old = post.config["stream_bytes"]
    post.config["stream_bytes"] = 1
    try:
        piped = post.unwarp_chunk_slices_backward(stack, 83.7, 44.2, FACT, 10, 80)
        ints = (stack * 60000).astype(np.uint16)
        got_int = post.unwarp_chunk_slices_backward(ints, 83.7, 44.2, FACT, 10, 80)
    finally:
        post.config["stream_bytes"] = old
    assert np.array_equal(direct, want) and np.array_equal(piped, want)
    assert got_int.dtype == np.uint16
    assert np.array_equal(got_int, orc.unwarp_chunk_slices_backward(ints, 83.7, 44.2, FACT, 10, 80))


@pytest.mark.gpu
def test_chunk_rows_outside_the_reference_window_gpu():
    """Chunk rows that sample outside the reference's row window (golden outputs of the real
    reference, tests/golden/chunk_window.npz): one-shot and streaming entries both reproduce SciPy's
    reflection into the cropped slices."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "chunk_window.npz"))
    for i in range(int(z["n"])):
        stack, par, ref = z["stack%d" % i], z["par%d" % i], z["ref%d" % i]
        xc, yc, a, b, fact = float(par[0]), float(par[1]), int(par[2]), int(par[3]), [float(v) for v in par[4:]]
        got = post.unwarp_chunk_slices_backward(stack, xc, yc, fact, a, b)
        assert got.dtype == ref.dtype and np.array_equal(got, ref), i
        piped = streaming.unwarp_chunk_slices_backward_stream(stack, xc, yc, fact, a, b)
        assert piped.dtype == ref.dtype and np.array_equal(piped, ref), i

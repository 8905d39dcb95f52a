"""
Streaming a Z-stack that does not fit HBM (or is not in host memory at all)
through the GPU -- SURVEY.md 8f rank 3, the data-format side of
``unwarp_chunk_slices_backward`` (reference ``postprocessing.py:255-313``; the
tomography examples ``examples/example_04.py:85-102`` hold 600 projections in
host memory and unwarp a chunk of rows of all of them).

``unwarp_chunk_slices_backward_stream`` takes any object that can be sliced
along its first axis into NumPy arrays -- an ``ndarray``, a ``numpy.memmap`` over
a raw / ``.npy`` file, an ``h5py`` dataset (``losa.load_hdf_file`` returns one,
``loadersaver.py:248-329``) -- and runs blocks of slices through a three-stage
pipeline on three CUDA streams::

    upload   block k+1   pinned staging -> HBM          (stream `up`)
    compute  block k     dcb_unwarp_stack_backward_f32  (stream `run`)
    download block k-1   HBM -> pinned staging -> out   (stream `down`)

with two staging buffers per direction, so PCIe runs in both directions while
the kernel works and the device holds only two blocks at a time.  Numerics are
those of ``unwarp_chunk_slices_backward`` (the same kernel, the same row window,
coordinates rounded to float32).  There is no CPU fallback.
"""
import ctypes
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .. import _cabi
from .. import device as _dev
from . import postprocessing as _post


def _vp(ptr):
    return ctypes.c_void_p(ptr)


_pool = None


def _copy_slices(dst, src, n, rows=None):
    """dst[:n] = src[:n] (with dtype conversion), one slice per task on a small thread
    pool: NumPy releases the GIL while it copies, and one core moves only ~4 GB/s --
    far less than the PCIe link the staging buffers feed."""
    if (n > 0 and dst.dtype == np.float32 and src.dtype == np.float32
            and isinstance(dst, np.ndarray) and isinstance(src, np.ndarray)):
        d, s_ = dst[:n], (src[:n] if rows is None else src[:n, rows[0]:rows[1], :])
        # each slice (window) one contiguous run of bytes on both sides: the library's copy
        # pool moves them with non-temporal stores (csrc/api.cu: CopyPool)
        run = int(np.prod(d.shape[1:])) * 4
        if (d.shape == s_.shape and run > 0 and d[0].flags.c_contiguous and s_[0].flags.c_contiguous
                and (n == 1 or (d.strides[0] >= run and s_.strides[0] >= run))):
            _cabi.call("dcb_host_copy_2d", _vp(d.ctypes.data), d.strides[0] if n > 1 else run,
                       _vp(s_.ctypes.data), s_.strides[0] if n > 1 else run, run, n)
            return
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) - 1)))

    def one(i):
        if rows is None:
            dst[i] = src[i]
        else:
            dst[i] = src[i, rows[0]:rows[1], :]
    list(_pool.map(one, range(n)))


def unwarp_chunk_slices_backward_stream(mat3D, xcenter, ycenter, list_fact,
                                        start_index=0, stop_index=None, out=None,
                                        slices_per_block=None, block_bytes=256 << 20):
    """
    Rows ``start_index .. stop_index`` (inclusive, like the reference) of every
    slice of ``mat3D``, unwarped, streamed block by block.

    Parameters
    ----------
    mat3D : array_like, (depth, height, width)
        Sliceable along axis 0 (ndarray, memmap, h5py dataset); float32 or
        uint8 / int8 / uint16 / int16.
    out : array_like, optional
        Destination ``(depth, stop-start+1, width)`` of ``mat3D``'s dtype,
        sliceable the same way (e.g. a writable memmap).  Allocated if None.
    slices_per_block : int, optional
        Slices per pipeline block; by default as many as fit ``block_bytes`` of
        source window.

    Returns
    -------
    array_like
        ``out``.
    """
    shape = tuple(mat3D.shape)
    if len(shape) < 3:
        raise ValueError("Input must be a 3D data")
    (depth, height, width) = shape
    if stop_index is None:
        stop_index = height - 1
    index_list = np.arange(height, dtype=np.int16)
    if stop_index == -1:
        stop_index = height
    if (start_index not in index_list) or (stop_index not in index_list):
        raise ValueError("Selected index is out of the range")
    start_index, stop_index = int(start_index), int(stop_index)
    if stop_index < start_index:
        raise ValueError("Selected index is out of the range")
    dtype = np.dtype(mat3D.dtype)
    if dtype == np.float32:
        flags, out_dtype = 0, np.dtype(np.float32)
    elif dtype in _post._INT_IMAGE_DTYPES:
        flags, out_dtype = _cabi.FLAG_ROUND_INT, dtype
    else:
        raise NotImplementedError(
            "dtype %s is not implemented on the CUDA path (float32, uint8, int8, "
            "uint16 and int16 are); there is no CPU fallback" % dtype)
    nrows = stop_index - start_index + 1
    yd1 = _post._row_yd(height, width, xcenter, ycenter, list_fact, start_index)
    yd2 = _post._row_yd(height, width, xcenter, ycenter, list_fact, stop_index)
    y0 = int(np.int16(np.floor(np.amin(yd1))))
    y1 = int(np.int16(np.ceil(np.amax(yd2)))) + 1          # reference :289-301
    wrows = y1 - y0
    if wrows <= 0:
        raise ValueError("empty row window [%d, %d): the model maps the last row of the chunk "
                         "above the first one (the reference's result is undefined here)" % (y0, y1))
    if out is None:
        out = np.empty((depth, nrows, width), dtype=out_dtype)
    if tuple(out.shape) != (depth, nrows, width):
        raise ValueError("out must have shape %s" % ((depth, nrows, width),))
    if depth == 0:
        return out
    if _post._rows_leave_window(height, width, xcenter, ycenter, list_fact, start_index,
                                stop_index, y0, y1):
        # rows sampling outside the reference's row window: SciPy reflects them into the cropped
        # slice; the (untuned) explicit-coordinate path of postprocessing.py restates that
        for z in range(depth):
            out[z:z + 1] = _post._chunk_outside_window(mat3D[z:z + 1], xcenter, ycenter, list_fact,
                                                       start_index, stop_index, y0, y1)
        return out
    if slices_per_block is None:
        slices_per_block = max(1, int(block_bytes // max(1, wrows * width * 4)))
    nb = int(min(max(1, slices_per_block), depth))
    model = _cabi.make_radial(xcenter, ycenter, list_fact)
    opt = _cabi.make_options(1, _post.config["blend"], _post.config["path"], flags)
    _dev.ensure_init()
    up, run, down = _dev.Stream(), _dev.Stream(), _dev.Stream()
    pitch_in = _dev._pitch_for(width)
    dense = pitch_in == width * 4
    in_pin = out_pin = None
    # the pooled buffers below are used on three private streams: anything the calling thread
    # queued earlier on its own stream against a buffer the pool hands back must have finished
    _dev.current_stream().sync()
    d_in = [_dev.DeviceArray((nb, wrows, width)) for _ in range(2)]
    d_out = [_dev.DeviceArray((nb, nrows, width)) for _ in range(2)]
    ev_up = [_dev.Event() for _ in range(2)]       # block uploaded
    ev_run = [_dev.Event() for _ in range(2)]      # block computed (d_in free again, d_out ready)
    ev_down = [_dev.Event() for _ in range(2)]     # block downloaded (d_out free again)
    blocks = [(z, min(z + nb, depth)) for z in range(0, depth, nb)]

    # a pinned float32 ndarray needs no staging: the DMA engine reads the row window of
    # every slice straight from it (and writes straight into a pinned float32 `out`)
    direct_in = (isinstance(mat3D, np.ndarray) and dtype == np.float32
                 and mat3D.flags.c_contiguous and _dev.is_pinned(mat3D))
    direct_out = (isinstance(out, np.ndarray) and out.dtype == np.float32
                  and out.flags.c_contiguous and _dev.is_pinned(out))

    if not direct_in:
        in_pin = [_dev.pinned_empty((nb, wrows, width), np.float32) for _ in range(2)]
    if not direct_out:
        out_pin = [_dev.pinned_empty((nb, nrows, width), np.float32) for _ in range(2)]

    def upload(k):
        z0, z1 = blocks[k]
        b = k & 1
        n = z1 - z0
        if direct_in:
            if k >= 2:
                up.wait(ev_run[b])
            for i in range(n):
                src_ptr = mat3D.ctypes.data + ((z0 + i) * height + y0) * width * 4
                _cabi.call("dcb_h2d_2d", _vp(d_in[b].ptr + i * d_in[b].slice_stride),
                           d_in[b].pitch, _vp(src_ptr), width * 4, width * 4, wrows,
                           _vp(up.handle))
            ev_up[b].record(up)
            return
        if k >= 2:
            ev_up[b].sync()        # the DMA that last read this staging buffer is done
        if isinstance(mat3D, np.ndarray):             # (memmaps are ndarrays too)
            _copy_slices(in_pin[b], mat3D[z0:z1], n, rows=(y0, y1))   # read + widen to float32
        else:                                         # h5py & co: one read call per block
            in_pin[b][:n] = mat3D[z0:z1, y0:y1, :]
        if k >= 2:
            up.wait(ev_run[b])     # the kernel that last read d_in[b] is done
        if dense:
            _cabi.call("dcb_h2d", _vp(d_in[b].ptr), _vp(in_pin[b].ctypes.data),
                       n * wrows * width * 4, _vp(up.handle))
        else:
            _cabi.call("dcb_h2d_2d", _vp(d_in[b].ptr), d_in[b].pitch,
                       _vp(in_pin[b].ctypes.data), width * 4, width * 4, n * wrows,
                       _vp(up.handle))
        ev_up[b].record(up)

    def compute(k):
        z0, z1 = blocks[k]
        b = k & 1
        run.wait(ev_up[b])
        if k >= 2:
            run.wait(ev_down[b])   # the download that last read d_out[b] is done
        _cabi.call("dcb_unwarp_stack_backward_f32", _vp(d_in[b].ptr), _vp(d_out[b].ptr),
                   z1 - z0, height, width, y0, wrows, d_in[b].pitch, d_in[b].slice_stride,
                   d_out[b].pitch, d_out[b].slice_stride, start_index, nrows, 1,
                   ctypes.byref(model), ctypes.byref(opt), _vp(run.handle))
        ev_run[b].record(run)

    def download(k):
        z0, z1 = blocks[k]
        b = k & 1
        n = z1 - z0
        down.wait(ev_run[b])
        if direct_out:
            _cabi.call("dcb_d2h_2d", _vp(out.ctypes.data + z0 * nrows * width * 4), width * 4,
                       _vp(d_out[b].ptr), d_out[b].pitch, width * 4, n * nrows,
                       _vp(down.handle))
        elif d_out[b].pitch == width * 4:
            _cabi.call("dcb_d2h", _vp(out_pin[b].ctypes.data), _vp(d_out[b].ptr),
                       n * nrows * width * 4, _vp(down.handle))
        else:
            _cabi.call("dcb_d2h_2d", _vp(out_pin[b].ctypes.data), width * 4,
                       _vp(d_out[b].ptr), d_out[b].pitch, width * 4, n * nrows,
                       _vp(down.handle))
        ev_down[b].record(down)

    def drain(k):
        z0, z1 = blocks[k]
        b = k & 1
        ev_down[b].sync()
        if direct_out:
            return
        if isinstance(out, np.ndarray):
            _copy_slices(out[z0:z1], out_pin[b], z1 - z0)             # narrows integer stacks
        else:
            res = out_pin[b][:z1 - z0]
            out[z0:z1] = res if out_dtype == np.float32 else res.astype(out_dtype)

    nblk = len(blocks)
    upload(0)
    for k in range(nblk):
        compute(k)
        if k + 1 < nblk:
            upload(k + 1)          # host staging of k+1 overlaps the kernel of k
        download(k)
        if k >= 1:
            drain(k - 1)           # host copy-out of k-1 overlaps the DMA of k
    drain(nblk - 1)
    return out

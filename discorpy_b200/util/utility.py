"""
B200 counterpart of the one image-unwarping function that lives outside
``discorpy.post.postprocessing``: ``discorpy.util.utility.
unwarp_color_image_backward`` (reference ``discorpy/util/utility.py:278-342``),
the call real pipelines make on camera frames
(``examples/apply_correction_to_images.py:37``, ``docs/source/usage/tips.rst``).

The reference evaluates one coordinate map and then calls
``scipy.ndimage.map_coordinates`` once per colour channel in a Python loop.
Here the channels become the planes of a (C, H, W) stack and go through the
Z-stack kernel in ONE launch: geometry, floor and the fp64 weights are
evaluated once per tile and reused for every channel.  uint8 / uint16 frames
are supported with SciPy's integer rounding (see ``post.postprocessing.
_as_f32_image``).  No CPU fallback.
"""
import numpy as np

from ..post import postprocessing as _post


def find_point_to_point(points, xcenter, ycenter, list_fact, output_order="xy"):
    """Corresponding point in the other space for one (row, column) point and a
    forward / backward model -- reference ``utility.py:192-226``; host side (one point)."""
    xi, yi = points[1] - xcenter, points[0] - ycenter
    ri = np.sqrt(xi * xi + yi * yi)
    factor = np.float64(np.sum(list_fact * np.power(ri, np.arange(len(list_fact)))))
    xo = xcenter + factor * xi
    yo = ycenter + factor * yi
    return (xo, yo) if output_order == "xy" else (yo, xo)


def _calc_pad(pad, height, width, xcenter, ycenter, list_fact):
    """Pad widths (top, bottom, left, right) -- reference ``utility.py:229-275``.

    ``pad=True`` asks for the automatic width, which the reference derives from
    a forward model fitted by ``discorpy.proc`` (CPU code that is outside this
    package): it is delegated to the installed reference when there is one."""
    if isinstance(pad, bool):
        if not pad:
            return 0, 0, 0, 0
        try:
            from discorpy.util import utility as _ref_util
        except ImportError as exc:
            raise NotImplementedError(
                "pad=True needs discorpy.proc (transform_coef_backward_and_"
                "forward) from the reference install; pass an int or a "
                "4-tuple instead") from exc
        return _ref_util._calc_pad(pad, height, width, xcenter, ycenter,
                                   list_fact)
    if isinstance(pad, int):
        return pad, pad, pad, pad
    if isinstance(pad, (tuple, list)):
        if len(pad) != 4:
            raise ValueError("Incorrect format!!! Please use a tuple/list of "
                             "(top_pad, bottom_pad, left_pad, right_pad)")
        return tuple(pad)
    raise ValueError("Invalid format of the 'pad' parameter!!!")


def unwarp_color_image_backward(mat, xcenter, ycenter, list_fact, order=1,
                                mode="reflect", pad=False,
                                pad_mode='constant'):
    """
    Unwarp a color image using a backward model (reference
    ``utility.py:278-342``).

    Parameters
    ----------
    mat : array_like
        2D (H, W) or 3D (H, W, C) array; float32, uint8, int8, uint16, int16.
    xcenter, ycenter : float
        Center of distortion.
    list_fact : list of float
        Polynomial coefficients of the backward model.
    order : int, optional
        Spline order 0..5; 0 and 1 take the tuned float32 kernels (all channels in one launch),
        2..5 SciPy's float64 prefilter + sampler on the GPU, one channel at a time.
    mode : str, optional
        SciPy boundary mode; the coordinates are clipped before sampling, so it only matters
        for orders >= 2 (the prefilter's boundary condition).
    pad : bool, int, or tuple of int
        Keeps the original view; see the reference.
    pad_mode : str
        ``numpy.pad`` mode.

    Returns
    -------
    array_like
        2D/3D array of the (padded) shape and the input dtype (C-contiguous;
        the reference returns the same values as a moved-axis view).
    """
    mat = np.asarray(mat)
    (height, width) = mat.shape[:2]
    order = _post._check_order_mode(order, mode)
    t_pad, b_pad, l_pad, r_pad = _calc_pad(pad, height, width, xcenter,
                                           ycenter, list_fact)
    num_dim = len(mat.shape)
    if num_dim == 2:
        pad_width = [(t_pad, b_pad), (l_pad, r_pad)]
    else:
        pad_width = [(t_pad, b_pad), (l_pad, r_pad), (0, 0)]
    if t_pad or b_pad or l_pad or r_pad:
        mat = np.pad(mat, pad_width, mode=pad_mode)
    xcenter = xcenter + l_pad
    ycenter = ycenter + t_pad
    if num_dim == 2:
        return _post.unwarp_image_backward(mat, xcenter, ycenter, list_fact,
                                           order=order, mode=mode)
    if num_dim != 3:
        raise ValueError("Input must be a 2D or 3D (H, W, C) array")
    if _post._wants_spline(mat, order):
        # spline orders >= 2 / float64 frames: one prefilter + remap per channel
        # (the reference loops over the channels too, utility.py:337-341)
        planes = [_post.unwarp_image_backward(mat[:, :, i], xcenter, ycenter,
                                              list_fact, order=order, mode=mode)
                  for i in range(mat.shape[-1])]
        return np.ascontiguousarray(np.moveaxis(np.asarray(planes), 0, 2))
    return _post._unwarp_frame_hwc(mat, xcenter, ycenter, list_fact, order)

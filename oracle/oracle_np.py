"""
oracle_np -- CPU restatement (NumPy) of Discorpy's image-unwarping hot path.

*** TEST INFRASTRUCTURE, NOT PRODUCT ***
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs
(``cpu_baseline`` / ``--impl reference``) may import this module. The product
(``discorpy_b200``) never imports anything under ``oracle/``.

Parity status: PINNED.  ``oracle/make_golden.py`` imported the real reference
from ``/root/reference`` (discorpy 1.7.0 + the installed SciPy
``scipy.ndimage.map_coordinates``) in the build container and compared every
function below bit-for-bit with it; the resulting fixtures are committed under
``tests/golden/`` and re-checked by ``tests/test_oracle.py``.

Each function cites the reference lines it restates (paths relative to
``/root/reference/``).  Nothing here is copied: the coordinate formulas are
re-derived, and the order-0 / order-1 sampler is an independent vectorised
restatement of what SciPy's ``_nd_image.geometric_transform`` does for
pre-clipped coordinates (the sampler lives in a third-party dependency whose
source is not in the reference tree; SciPy is unpinned by the reference,
``requirements.txt:3``; the build container has scipy 1.18.1 / numpy 2.3.5).
"""
import numpy as np

__all__ = [
    "radial_factor", "radial_coords", "radial_coords_row", "persp_coords",
    "sample", "unwarp_image_backward", "unwarp_slice_backward",
    "unwarp_chunk_slices_backward", "correct_perspective_image",
    "unwarp_image_backward_perspective", "mapping", "chunk_row_window",
    "unwarp_color_image_backward",
]


# --------------------------------------------------------------------------
# coordinates
# --------------------------------------------------------------------------
def radial_factor(ru, list_fact):
    """F = sum_i a_i * ru**i, accumulated term by term in list order.

    Restates ``discorpy/post/postprocessing.py:142-143`` (and ``:217-218``,
    ``:292-293``, ``:305-306``): the reference stacks the N terms and reduces
    over axis 0, which NumPy evaluates as ((t0 + t1) + t2) + ... ; ``ru**i`` is
    NumPy's own power ufunc, so the oracle shares its libm/SVML rounding.
    """
    ru = np.asarray(ru, dtype=np.float64)
    acc = None
    for i, a in enumerate(list_fact):
        term = a * ru ** i
        acc = term if acc is None else acc + term
    if acc is None:                     # empty coefficient list -> sum of nothing
        acc = np.zeros_like(ru)
    return acc


def radial_coords(height, width, xcenter, ycenter, list_fact, row0=0,
                  nrows=None, round32=True):
    """(yd, xd) source coordinates of output rows ``row0 .. row0+nrows-1``.

    ``postprocessing.py:138-145`` (image) and ``:302-309`` (chunk).  float64
    math, clip to the image, then (``round32``) one rounding to float32.
    """
    if nrows is None:
        nrows = height - row0
    xu = np.arange(width) - xcenter
    yu = np.arange(row0, row0 + nrows) - ycenter
    xu_mat, yu_mat = np.meshgrid(xu, yu)
    ru = np.sqrt(xu_mat ** 2 + yu_mat ** 2)
    fact = radial_factor(ru, list_fact)
    xd = np.clip(xcenter + fact * xu_mat, 0, width - 1)
    yd = np.clip(ycenter + fact * yu_mat, 0, height - 1)
    if round32:
        xd = np.float32(xd)
        yd = np.float32(yd)
    return yd, xd


def radial_coords_row(height, width, xcenter, ycenter, list_fact, index):
    """Unrounded float64 (yd, xd) of one output row -- ``postprocessing.py:214-220``."""
    xu = np.arange(0, width) - xcenter
    yu = index - ycenter
    ru = np.sqrt(xu ** 2 + yu ** 2)
    fact = radial_factor(ru, list_fact)
    xd = np.clip(xcenter + fact * xu, 0, width - 1)
    yd = np.clip(ycenter + fact * yu, 0, height - 1)
    return yd, xd


def persp_coords(height, width, list_coef):
    """float32 (yd, xd) of the projective backward map -- ``postprocessing.py:444-459``."""
    c1, c2, c3, c4, c5, c6, c7, c8 = list_coef
    xu_mat, yu_mat = np.meshgrid(np.arange(width), np.arange(height))
    den = c7 * xu_mat + c8 * yu_mat + 1.0
    xd = (c1 * xu_mat + c2 * yu_mat + c3) / den
    yd = (c4 * xu_mat + c5 * yu_mat + c6) / den
    xd = np.float32(np.clip(xd, 0, width - 1))
    yd = np.float32(np.clip(yd, 0, height - 1))
    return yd, xd


# --------------------------------------------------------------------------
# sampler: scipy.ndimage.map_coordinates(order in {0, 1}) on clipped coords
# --------------------------------------------------------------------------
def _cast_like_scipy(val, dtype):
    """double -> output dtype the way ``_nd_image`` does it.

    Floating outputs: one IEEE round-to-nearest.  Integer outputs: round half
    away from zero (``val + 0.5`` / ``val - 0.5`` then truncate), as SURVEY.md
    section 8(a1) records for u8/u16/i16, saturating at the type's range
    (measured against SciPy with overshooting cubic splines).
    """
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        return val.astype(dtype)
    if dtype.kind in "iu":
        # (spline orders >= 2 overshoot: SciPy saturates at the type's range)
        info = np.iinfo(dtype)
        shifted = np.where(val >= 0, val + 0.5, val - 0.5)
        return np.clip(np.trunc(shifted), info.min, info.max).astype(dtype)
    raise TypeError("unsupported dtype %s" % dtype)


def reflect_coordinate(cc, n):
    """SciPy's ``map_coordinate`` for mode 'reflect' (``ni_interpolation.c``),
    applied to a float64 coordinate array before the spline start index is
    taken: (d c b a | a b c d | d c b a).  Coordinates inside ``[0, n-1]`` are
    returned unchanged.  Only ``unwarp_chunk_slices_backward`` can produce
    coordinates outside the array: its row window comes from the first and the
    last row of the chunk (``postprocessing.py:289-301``), which a strongly
    off-centre model need not respect."""
    cc = np.array(cc, dtype=np.float64, copy=True)
    if n <= 1:
        cc[(cc < 0) | (cc > n - 1)] = 0.0
        return cc
    sz2 = 2.0 * n
    neg = cc < 0
    v = cc[neg]
    far = v < -sz2
    v[far] = sz2 * np.trunc(-v[far] / sz2) + v[far]
    v = np.where(v < -n, v + sz2, -v - 1.0)
    cc[neg] = v
    pos = cc > n - 1
    pos &= ~neg
    v = cc[pos]
    v = v - sz2 * np.trunc(v / sz2)
    v = np.where(v >= n, sz2 - v - 1.0, v)
    cc[pos] = v
    return cc


def reflect_index(idx, n):
    """Integer tap index -> index inside ``[0, n)`` by SciPy's 'reflect' rule."""
    idx = np.array(idx, dtype=np.intp, copy=True)
    if n <= 1:
        return np.zeros_like(idx)
    s2 = 2 * n
    neg = idx < 0
    v = idx[neg]
    far = v < -s2
    v[far] = s2 * ((-v[far]) // s2) + v[far]
    idx[neg] = np.where(v < -n, v + s2, -v - 1)
    pos = (idx >= n) & ~neg
    v = idx[pos]
    v = v - s2 * (v // s2)
    idx[pos] = np.where(v >= n, s2 - v - 1, v)
    return idx


def sample(mat, yd, xd, order=1, out_dtype=None):
    """Sample 2-D ``mat`` at coordinates (yd, xd) that already lie inside
    ``[0, H-1] x [0, W-1]``.

    Restates the boundary the reference crosses at ``postprocessing.py:147,
    227-228, 251, 491`` (``scipy.ndimage.map_coordinates`` -> C
    ``geometric_transform``): the coordinate is widened to double; order 1
    takes ``floor`` and the four taps weighted ``(1-ty)(1-tx), (1-ty)tx,
    ty(1-tx), ty*tx`` -- each tap multiplied first by its y-weight then by its
    x-weight, the four products added in that order, all in double, one final
    cast.  Order 0 takes ``floor(c + 0.5)``.  The +1 neighbour of the last
    row/column folds back onto it (weight 0), which is what every SciPy
    boundary mode yields for in-range coordinates.
    """
    mat = np.asarray(mat)
    if mat.ndim != 2:
        raise ValueError("oracle sampler is 2-D only")
    h, w = mat.shape
    out_dtype = mat.dtype if out_dtype is None else np.dtype(out_dtype)
    shape = np.shape(yd)
    y = np.asarray(yd, dtype=np.float64).ravel()
    x = np.asarray(xd, dtype=np.float64).ravel()
    src = mat.astype(np.float64, copy=False)
    if order == 0:
        yi = np.floor(y + 0.5).astype(np.intp)
        xi = np.floor(x + 0.5).astype(np.intp)
        val = 0.0 + src[yi, xi]       # SciPy accumulates from 0.0: a -0.0 pixel comes out +0.0
    elif order == 1:
        y0f = np.floor(y)
        x0f = np.floor(x)
        ty = y - y0f
        tx = x - x0f
        y0 = y0f.astype(np.intp)
        x0 = x0f.astype(np.intp)
        # taps that leave the array are reflected (SciPy's default boundary, the
        # reference never passes another one on this path): for in-range
        # coordinates only the +1 tap of the last row / column can, and it folds
        # back onto that row / column with weight 0
        y0, y1 = reflect_index(y0, h), reflect_index(y0 + 1, h)
        x0, x1 = reflect_index(x0, w), reflect_index(x0 + 1, w)
        wy0 = 1.0 - ty
        wx0 = 1.0 - tx
        val = 0.0 + (src[y0, x0] * wy0) * wx0
        val = val + (src[y0, x1] * wy0) * tx
        val = val + (src[y1, x0] * ty) * wx0
        val = val + (src[y1, x1] * ty) * tx
    else:
        raise NotImplementedError("oracle sampler covers order 0 and 1")
    return _cast_like_scipy(val, out_dtype).reshape(shape)


# --------------------------------------------------------------------------
# the public functions of the path
# --------------------------------------------------------------------------
def unwarp_image_backward(mat, xcenter, ycenter, list_fact, order=1,
                          mode="reflect"):
    """``postprocessing.py:111-148``.  ``mode`` is accepted for signature
    parity; with pre-clipped coordinates it does not influence order 0/1."""
    (height, width) = np.shape(mat)
    yd, xd = radial_coords(height, width, xcenter, ycenter, list_fact)
    return sample(mat, yd, xd, order)


def unwarp_slice_backward(mat3D, xcenter, ycenter, list_fact, index):
    """``postprocessing.py:188-229``: one unwarped sinogram, float32 output,
    float64 (unrounded) coordinates.  The reference crops a row window and
    shifts yd by its origin (``:221-223``); the subtraction is exact, so
    sampling the whole slice at the unshifted coordinate is the same number."""
    mat3D = np.asarray(mat3D)
    if mat3D.ndim < 3:
        raise ValueError("Input must be a 3D data")
    (depth, height, width) = mat3D.shape
    yd, xd = radial_coords_row(height, width, xcenter, ycenter, list_fact,
                               index)
    sino = np.zeros((depth, width), dtype=np.float32)
    for i in range(depth):
        # map_coordinates returns the slice's own dtype (:227-228), i.e. integer
        # stacks are rounded to integers BEFORE the store into the float32 sinogram
        sino[i] = sample(mat3D[i], yd, xd, 1)
    return sino


def chunk_row_window(height, width, xcenter, ycenter, list_fact, start_index,
                     stop_index):
    """Row window ``[yd_min, yd_max)`` the reference crops to in
    ``postprocessing.py:289-301`` (from the first and last chunk row only)."""
    yd1, _ = radial_coords_row(height, width, xcenter, ycenter, list_fact,
                               start_index)
    yd2, _ = radial_coords_row(height, width, xcenter, ycenter, list_fact,
                               stop_index)
    yd_min = int(np.int16(np.floor(np.amin(yd1))))
    yd_max = int(np.int16(np.ceil(np.amax(yd2)))) + 1
    return yd_min, yd_max


def unwarp_chunk_slices_backward(mat3D, xcenter, ycenter, list_fact,
                                 start_index, stop_index):
    """``postprocessing.py:255-313``: rows start..stop inclusive of every
    slice, float32-rounded coordinates, output dtype = input dtype.  The
    sampling really happens inside the cropped window (``:310-312``): taps that
    fall outside it fold back onto its last row, which the restatement keeps by
    sampling the cropped view with window-relative coordinates."""
    mat3D = np.asarray(mat3D)
    if mat3D.ndim < 3:
        raise ValueError("Input must be a 3D data")
    (depth, height, width) = mat3D.shape
    index_list = np.arange(height, dtype=np.int16)
    if stop_index == -1:
        stop_index = height
    if (start_index not in index_list) or (stop_index not in index_list):
        raise ValueError("Selected index is out of the range")
    yd_min, yd_max = chunk_row_window(height, width, xcenter, ycenter,
                                      list_fact, start_index, stop_index)
    if yd_max <= yd_min:
        # a model that maps the last chunk row above the first one: the reference hands SciPy an
        # EMPTY slice and SciPy reads past it (the values returned differ from run to run)
        raise ValueError("empty row window [%d, %d): the reference's result is undefined here"
                         % (yd_min, yd_max))
    nrows = stop_index - start_index + 1
    yd, xd = radial_coords(height, width, xcenter, ycenter, list_fact,
                           row0=start_index, nrows=nrows)
    yd = yd - yd_min          # float32 - int16 scalar -> float32 (:308-309)
    # rows the window does not hold are reflected into it by SciPy (mode 'reflect')
    yd = reflect_coordinate(yd, yd_max - yd_min)
    return np.asarray([sample(mat3D[i, yd_min:yd_max, :], yd, xd, 1)
                       for i in range(depth)])


def mapping(mat, xmat, ymat):
    """``postprocessing.py:232-252`` (``_mapping``)."""
    return sample(mat, ymat, xmat, 1)


def correct_perspective_image(mat, list_coef, order=1, mode="reflect",
                              map_index=None):
    """``postprocessing.py:462-492``."""
    if len(list_coef) != 8:
        raise ValueError("!!! Eight coefficients are required !!!")
    (height, width) = np.shape(mat)
    if map_index is None:
        yd, xd = persp_coords(height, width, list_coef)
    else:
        yd, xd = map_index
    out = sample(mat, np.reshape(yd, -1), np.reshape(xd, -1), order)
    return out.reshape((height, width))


def unwarp_image_backward_perspective(mat, xcenter, ycenter, list_fact,
                                      list_coef, order=1, mode="reflect"):
    """The combined entry named by BASELINE.json's north_star.  It does not
    exist in the reference; its definition is the two-pass composition of
    ``examples/readthedocs_demo/demo_05.py:127`` then ``:147`` (intermediate
    rounded to ``mat.dtype``)."""
    tmp = unwarp_image_backward(mat, xcenter, ycenter, list_fact, order, mode)
    return correct_perspective_image(tmp, list_coef, order, mode)


# --------------------------------------------------------------------------
# CPU-baseline helpers (bench.py's cpu_baseline / --impl reference legs only)
# --------------------------------------------------------------------------
def unwarp_rows_scipy(mat, xcenter, ycenter, list_fact, row0, nrows, order=1,
                      mode="reflect"):
    """Rows ``row0 .. row0+nrows-1`` of ``unwarp_image_backward`` computed the
    way the reference computes them -- NumPy float64 temporaries for the
    coordinates (``postprocessing.py:138-145``) then SciPy's own
    ``map_coordinates`` (``:147``).  This, not the vectorised ``sample``
    above, is what the CPU baseline times: it is the reference's code path
    split over row blocks so that several cores can be used."""
    from scipy.ndimage import map_coordinates
    (height, width) = np.shape(mat)
    yd, xd = radial_coords(height, width, xcenter, ycenter, list_fact,
                           row0=row0, nrows=nrows)
    indices = np.reshape(yd, (-1, 1)), np.reshape(xd, (-1, 1))
    out = map_coordinates(mat, indices, order=order, mode=mode)
    return out.reshape((nrows, width))


def unwarp_color_image_backward(mat, xcenter, ycenter, list_fact, order=1,
                                mode="reflect", pad=0, pad_mode="constant"):
    """``discorpy/util/utility.py:278-342`` for ``pad`` given as an int or a
    (top, bottom, left, right) tuple (``:267-275``; ``pad=True`` needs
    ``discorpy.proc`` and is outside the oracle): ``np.pad`` (``:315-319``),
    centre shifted by the pad (``:321-322``), the image coordinate map
    (``:323-330``, same expressions as ``postprocessing.py:138-145``) and one
    ``map_coordinates`` per channel (``:332-341``).  The result keeps the
    reference's layout: channels moved back to the last axis (a view)."""
    mat = np.asarray(mat)
    if isinstance(pad, bool):
        if pad:
            raise NotImplementedError("pad=True is outside the oracle")
        t_pad = b_pad = l_pad = r_pad = 0
    elif isinstance(pad, int):
        t_pad = b_pad = l_pad = r_pad = pad
    else:
        t_pad, b_pad, l_pad, r_pad = pad
    if mat.ndim == 2:
        pad_width = [(t_pad, b_pad), (l_pad, r_pad)]
    else:
        pad_width = [(t_pad, b_pad), (l_pad, r_pad), (0, 0)]
    mat_pad = np.pad(mat, pad_width, mode=pad_mode)
    (height, width) = mat_pad.shape[:2]
    yd, xd = radial_coords(height, width, xcenter + l_pad, ycenter + t_pad,
                           list_fact)
    if mat.ndim == 2:
        return sample(mat_pad, yd, xd, order)
    planes = [sample(mat_pad[:, :, i], yd, xd, order)
              for i in range(mat_pad.shape[-1])]
    return np.moveaxis(np.asarray(planes), 0, 2)


def unwarp_image_forward(mat, xcenter, ycenter, list_fact):
    """Forward-model scatter (reference ``postprocessing.py:151-185``): every source pixel is
    written to its rounded, clipped target; vacant targets stay 0 and, where several sources
    land on one target, the last one in C order wins (what NumPy's ``out[yu, xu] = mat`` does).
    Restated with an explicit winner table instead of the fancy assignment."""
    mat = np.asarray(mat)
    (height, width) = mat.shape
    xd = np.arange(width) - xcenter
    yd = np.arange(height) - ycenter
    xd_mat, yd_mat = np.meshgrid(xd, yd)
    rd = np.sqrt(xd_mat ** 2 + yd_mat ** 2)
    fact = None
    for i, a in enumerate(list_fact):                      # :176-177, summed first to last
        term = a * rd ** i
        fact = term if fact is None else fact + term
    xu = np.intp(np.round(np.clip(xcenter + fact * xd_mat, 0, width - 1)))
    yu = np.intp(np.round(np.clip(ycenter + fact * yd_mat, 0, height - 1)))
    target = (yu * width + xu).ravel()
    winner = np.full(height * width, -1, dtype=np.int64)
    np.maximum.at(winner, target, np.arange(height * width, dtype=np.int64))
    out = np.zeros(height * width, dtype=mat.dtype)
    hit = winner >= 0
    out[hit] = mat.ravel()[winner[hit]]
    return out.reshape(height, width)

"""The qualitative properties the reference's own unit tests assert for the
hot functions (tests/test_postprocessing.py:77-123, 205-238 of the reference),
restated against the drop-in module.  GPU needed."""
import numpy as np
import pytest
import scipy.ndimage as ndi

import discorpy_b200.post.postprocessing as post

pytestmark = pytest.mark.gpu

HEI, WID = 64, 64


def test_unwarp_image_backward_profile():
    x0, y0 = WID // 2, HEI // 2
    mat = np.zeros((HEI, WID), dtype=np.float32)
    mat[4:-3, 4:-3] = 1.0
    out = post.unwarp_image_backward(mat, x0, y0, [1.0, 3.0e-3])
    vals = np.mean(out, axis=0)[11:-10]
    pos = len(vals) // 2
    assert vals[0] < vals[pos] and vals[-1] < vals[pos]


def _stripes():
    mat = np.zeros((HEI, WID), dtype=np.float32)
    mat[:, 6:-8:8] = 1.0
    mat = np.float32(ndi.binary_dilation(np.int16(mat), iterations=1))
    mat3d = np.zeros((10, HEI, WID), dtype=np.float32)
    mat3d[:] = mat
    return mat3d


def test_unwarp_slice_backward_changes_the_sinogram():
    x0, y0 = WID // 2, HEI // 2
    mat3d = _stripes()
    out = post.unwarp_slice_backward(mat3d, x0, y0, [1.0, 3.0e-3], y0)
    assert out.shape == (10, WID) and out.dtype == np.float32
    assert np.max(mat3d[:, y0, :] - out) > 0.1


def test_unwarp_chunk_slices_backward_first_and_last_rows():
    x0, y0 = WID // 2, HEI // 2
    mat3d = _stripes()
    out = post.unwarp_chunk_slices_backward(mat3d, x0, y0, [1.0, 3.0e-3],
                                            y0 - 5, y0 + 5)
    assert out.shape == (10, 11, WID)
    assert np.max(mat3d[:, y0 - 5, :] - out[:, 0, :]) > 0.1
    assert np.max(mat3d[:, y0 + 5, :] - out[:, -1, :]) > 0.1


def test_correct_perspective_image_roundtrip():
    # forward / backward homographies of a mild keystone, built by hand
    # (the reference test derives them with proc.calc_perspective_coefficients)
    fwd = np.array([[1.0, 0.08, -2.0], [0.0, 1.05, -1.0], [0.0, 1.5e-3, 1.0]])
    bwd = np.linalg.inv(fwd)
    bwd /= bwd[2, 2]
    fcoef = list(fwd.ravel()[:8])
    bcoef = list(bwd.ravel()[:8])
    mat = np.zeros((HEI, WID), dtype=np.float32)
    mat[HEI // 2 - 3:HEI // 2 + 3, 10:-10] = 1.0
    warped = post.correct_perspective_image(mat, bcoef)
    back = post.correct_perspective_image(warped, fcoef)
    assert warped.shape == mat.shape and warped.dtype == np.float32
    assert np.argmax(np.sum(back, axis=1)) in range(HEI // 2 - 4, HEI // 2 + 4)
    assert abs(float(back.sum()) - float(mat.sum())) / float(mat.sum()) < 0.2

"""Small cases for compute-sanitizer (tools/gpu_sanitize.sh): the single-image kernel with several
launches back to back (programmatic dependent launch, pooled tile scheduling, plan build), its
blend variants, the Z-stack kernel in both tile shapes and both coordinate semantics, the
perspective kernel, the banded host-buffer pipelines and the colour path."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
import discorpy_b200.util.utility as util

dcb.set_device(0)
rng = np.random.default_rng(0)
mat = rng.random((700, 1100), dtype=np.float32)
fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
dev = dcb.DeviceArray.from_host(mat)
for order in (1, 0):
    outs = [post.unwarp_image_backward(dev, 551.3, 349.2, fact, order=order) for _ in range(3)]
    [o.to_host() for o in outs]
for blend in (dcb.BLEND_LERP32, dcb.BLEND_LERP64):
    post.config["blend"] = blend
    post.unwarp_image_backward(dev, 551.3, 349.2, fact).to_host()
post.config["blend"] = dcb.BLEND_EXACT
neg = mat - 0.5                      # negative pixels: the exact rows of the patch path
post.unwarp_image_backward(dcb.DeviceArray.from_host(neg), 551.3, 349.2, fact).to_host()
stack = rng.random((5, 96, 400), dtype=np.float32)
post.unwarp_chunk_slices_backward(stack, 201.3, 47.2, fact, 0, 95)
post.unwarp_slice_backward(stack, 201.3, 47.2, fact, 40)
shear = [1.0, -4e-4, 3e-6, -2e-9]    # strong shear: the 64 x 32 tile shape
post.unwarp_chunk_slices_backward(stack, 201.3, 47.2, shear, 0, 95)
post.unwarp_slice_backward(stack, 201.3, 47.2, shear, 40)
post.correct_perspective_image(mat, [1.02, 0.01, -5.0, 0.005, 1.01, -3.0, 8e-6, -5e-6])
# the banded host-buffer pipelines: radial, projective and both stages, several bands, pinned and
# pageable sources, and the band kernels storing straight into the page-locked result
coef = [1.02, 0.01, -5.0, 0.005, 1.01, -3.0, 8e-6, -5e-6]
pinned = dcb.pinned_empty(mat.shape, np.float32)
pinned[:] = mat
for bands, src in ((3, pinned), (5, mat)):
    post.config["bands"] = bands
    post.unwarp_image_backward(src, 551.3, 349.2, fact)
    post.correct_perspective_image(src, coef)
    post.unwarp_image_backward_perspective(src, 551.3, 349.2, fact, coef)
os.environ["DCB_PIPE_DIRECT"] = "1"
post.unwarp_image_backward_perspective(pinned, 551.3, 349.2, fact, coef)
del os.environ["DCB_PIPE_DIRECT"]
post.config["bands"] = 0
u16 = (rng.random((2200, 2200)) * 60000).astype(np.uint16)      # 9.7 MB: chunked staging of the upload
post.unwarp_image_backward(u16, 1101.3, 1099.2, fact)
rgb = (rng.random((300, 400, 3)) * 255).astype(np.uint8)
util.unwarp_color_image_backward(rgb, 201.3, 149.2, fact)
print("sanitizer case done")

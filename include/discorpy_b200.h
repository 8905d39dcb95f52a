/*
 * discorpy_b200.h -- C ABI of libdiscorpy_b200.so
 *
 * A from-scratch sm_100a (NVIDIA B200) implementation of the image-unwarping
 * hot path of Discorpy's `discorpy/post/postprocessing.py`.  The reference is
 * pure Python and has no FFI of its own: the functions
 * below are what a `discorpy.post.postprocessing` replacement binds with
 * ctypes/cffi -- see INTEGRATION.md for the stub -- and each one names the
 * reference lines it replaces.
 *
 * Conventions
 *   - plain C, no C++ / torch types; every function returns an int status
 *     (DCB_OK == 0, negative on error) and leaves a thread-local message for
 *     dcb_last_error().
 *   - pointers are DEVICE pointers unless the name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *     compute calls are asynchronous on it, the caller owns every buffer.
 *   - pitches / strides are in BYTES.  The TMA-staged path needs the source
 *     base 16-byte aligned and the source pitch and slice stride multiples of
 *     16 bytes; otherwise DCB_PATH_AUTO silently takes the direct-gather path
 *     (same numerics) and DCB_PATH_TMA returns DCB_ERR_ARG.
 *   - the library never falls back to a CPU implementation.
 */
#ifndef DISCORPY_B200_H
#define DISCORPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCB_VERSION_MAJOR 0
#define DCB_VERSION_MINOR 1
#define DCB_MAX_TERMS 16 /* polynomial coefficients a_0 .. a_15 */

enum dcb_status {
    DCB_OK = 0,
    DCB_ERR_ARG = -1,         /* bad argument (message says which) */
    DCB_ERR_CUDA = -2,        /* CUDA runtime / driver error */
    DCB_ERR_UNSUPPORTED = -3, /* valid request the library does not implement */
    DCB_ERR_NO_DEVICE = -4
};

/* How the four bilinear taps are blended (order 1).  All variants read and
 * multiply all four taps (a NaN tap contaminates like it does in SciPy). */
enum dcb_blend {
    DCB_BLEND_EXACT = 0,  /* fp64, SciPy's operation order: bit-faithful (default) */
    DCB_BLEND_LERP64 = 1, /* fp64 three-lerp form: fewer fp64 ops, <= 1 fp32 ulp from EXACT in rare cases */
    DCB_BLEND_LERP32 = 2  /* fp32 three-lerp: +-1 fp32 ulp; opt-in fast mode */
};

enum dcb_path {
    DCB_PATH_AUTO = 0,   /* TMA-staged tiles where the source footprint fits, direct gathers elsewhere */
    DCB_PATH_DIRECT = 1, /* ld.global.nc gathers only */
    DCB_PATH_TMA = 2     /* like AUTO but refuses (DCB_ERR_ARG) when the layout cannot be described to TMA */
};

/* Radial backward model: F(r) = sum_{i<n} a[i] r^i, r = |(x,y) - (xc,yc)|
 * (postprocessing.py:138-145). */
typedef struct dcb_radial {
    double xc;
    double yc;
    int32_t n;
    int32_t reserved;
    double a[DCB_MAX_TERMS];
} dcb_radial;

/* Projective backward map, c[0..7] = c1..c8 of postprocessing.py:448-455. */
typedef struct dcb_persp {
    double c[8];
} dcb_persp;

/* dcb_options.flags.  DCB_FLAG_ROUND_INT: the image holds integer pixel values
 * (uint8/uint16/int8/int16 widened exactly to float32 by the host); order-1
 * results are then rounded half away from zero to an integer while still in
 * fp64 -- what scipy.ndimage.map_coordinates does for integer outputs
 * (output dtype = input dtype, scipy/ndimage/_ni_support.py:83) -- and stored as
 * that integer in float32, which the host narrows back without loss.  Bits 0-7
 * select experimental kernel variants in -DDCB_AB builds and are ignored otherwise. */
#define DCB_FLAG_ROUND_INT 0x100

typedef struct dcb_options {
    int32_t order; /* 0 nearest, 1 bilinear */
    int32_t blend; /* enum dcb_blend */
    int32_t path;  /* enum dcb_path */
    int32_t flags; /* DCB_FLAG_* */
} dcb_options;

/* ---- library / device ---------------------------------------------------- */
int dcb_version(void);                 /* major*1000 + minor */
const char *dcb_last_error(void);      /* thread-local, never NULL */
int dcb_device_count(int *count);
int dcb_init(int device);              /* cudaSetDevice + resolve the TMA encoder; idempotent */
int dcb_device_info(int device, int *sm_count, int *cc_major, int *cc_minor,
                    size_t *total_mem, size_t *free_mem, char *name, int name_len);

/* ---- memory, streams, events (thin cudart wrappers so that the Python host
 *      needs neither torch nor cuda-python) --------------------------------- */
int dcb_malloc(void **dptr, size_t nbytes);
int dcb_free(void *dptr);
int dcb_memset(void *dptr, int value, size_t nbytes, void *stream);
int dcb_host_alloc(void **hptr_host, size_t nbytes);   /* pinned */
int dcb_host_free(void *hptr_host);
int dcb_host_register(void *hptr_host, size_t nbytes); /* pin caller memory in place */
int dcb_host_unregister(void *hptr_host);
int dcb_is_pinned(const void *hptr_host, int *pinned);
int dcb_h2d(void *dst, const void *src_host, size_t nbytes, void *stream);
int dcb_d2h(void *dst_host, const void *src, size_t nbytes, void *stream);
int dcb_d2d(void *dst, const void *src, size_t nbytes, void *stream);
int dcb_h2d_2d(void *dst, size_t dst_pitch, const void *src_host, size_t src_pitch,
               size_t width_bytes, size_t rows, void *stream);
int dcb_d2h_2d(void *dst_host, size_t dst_pitch, const void *src, size_t src_pitch,
               size_t width_bytes, size_t rows, void *stream);
int dcb_stream_create(void **stream);
int dcb_stream_destroy(void *stream);
int dcb_stream_sync(void *stream);
int dcb_device_sync(void);
int dcb_event_create(void **event);
int dcb_event_destroy(void *event);
int dcb_event_record(void *event, void *stream);
int dcb_event_sync(void *event);
int dcb_stream_wait_event(void *stream, void *event); /* later work on `stream` waits for `event` */
int dcb_event_elapsed_ms(void *start, void *stop, float *ms);

/* ---- peer windows (one process per GPU): a device buffer of one rank mapped into the others.
 * when the caller wants ONE sinogram (D x W) assembled on one GPU, every rank's
 * unwarp_slice_backward kernel (postprocessing.py:188-229, one output row per slice) writes its D/N
 * rows straight into the owner's buffer -- the stores of the remap epilogue travel over
 * NVLink / NVSwitch, there is no separate gather collective.  `dptr` must be the base address of a
 * dcb_malloc allocation; the handle is DCB_IPC_HANDLE_BYTES opaque bytes that the caller moves to
 * the other processes (torch.distributed broadcast in discorpy_b200.multigpu). */
#define DCB_IPC_HANDLE_BYTES 64
int dcb_ipc_export(const void *dptr, void *handle_host);
int dcb_ipc_open(const void *handle_host, void **peer_dptr); /* maps it, enabling peer access */
int dcb_ipc_close(void *peer_dptr);

/* ---- the hot path -------------------------------------------------------- */

/* Replaces discorpy/post/postprocessing.py:111-148 `unwarp_image_backward`
 * for a float32 image: coordinates in fp64, clipped, rounded once to fp32
 * (:144-145), then the order-0/1 sampling that scipy.ndimage.map_coordinates
 * performs (:147).  src and dst are (H, W) float32, dst must not alias src. */
int dcb_unwarp_image_backward_f32(const float *src, float *dst, int H, int W,
                                  size_t src_pitch, size_t dst_pitch,
                                  const dcb_radial *model_host,
                                  const dcb_options *opt_host, void *stream);

/* Same function for HOST buffers (what a NumPy caller holds): the image is
 * uploaded in `nbands` row bands (0 = library default), each band of output rows
 * is unwarped as soon as the last source row it can sample has arrived, and is
 * downloaded while the next one computes -- three streams, so for a pinned
 * 4096^2 image the call costs about one PCIe transfer instead of two plus the
 * kernel.  Synchronous: returns when dst_host is complete.  Pitches in bytes.
 * Device scratch and streams are cached per host thread. */
int dcb_unwarp_image_backward_host_f32(const float *src_host, float *dst_host, int H, int W,
                                       size_t src_pitch_host, size_t dst_pitch_host,
                                       const dcb_radial *model_host,
                                       const dcb_options *opt_host, int nbands);

/* The row-band schedule the host-buffer entries use for an H x W float32 image
 * (`nbands` as in those calls: 0 = library default): edges[0] = 0 < edges[1] < ...
 * < edges[*count] = H, *count <= 32; `edges` holds 33 ints.  Diagnostics / tests;
 * needs no GPU. */
int dcb_host_band_edges(int H, int W, int nbands, int *edges, int *count);

/* Host-to-host copy of `rows` rows of `width_bytes` bytes (pitches in bytes) by the
 * library's pool of host threads with non-temporal stores -- what stages pageable
 * (ordinary NumPy) data into page-locked buffers and results back out of them
 * (post/streaming.py; one core copies ~4 GB/s, far less than the PCIe link).
 * Synchronous; calls are serialised.  No reference counterpart (NumPy slicing). */
int dcb_host_copy_2d(void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                     size_t width_bytes, int rows);

/* The projective remap (`correct_perspective_image`, postprocessing.py:462-492) and
 * the radial remap followed by the projective one (demo_05.py:127,147) for HOST
 * buffers, through the same banded pipeline: a band of projective output rows is
 * launched as soon as the last row it can sample has been uploaded (or, in the
 * two-stage form, has been produced by the radial stage; the intermediate image
 * stays in HBM), and is downloaded behind its kernel.  Results equal those of
 * dcb_correct_perspective_image_f32 / dcb_unwarp_image_backward_perspective_f32. */
int dcb_correct_perspective_image_host_f32(const float *src_host, float *dst_host, int H, int W,
                                           size_t src_pitch_host, size_t dst_pitch_host,
                                           const dcb_persp *model_host,
                                           const dcb_options *opt_host, int nbands);
int dcb_unwarp_image_backward_perspective_host_f32(const float *src_host, float *dst_host, int H,
                                                   int W, size_t src_pitch_host,
                                                   size_t dst_pitch_host,
                                                   const dcb_radial *radial_host,
                                                   const dcb_persp *persp_host,
                                                   const dcb_options *opt_host, int nbands);

/* Replaces the per-slice Python loops of
 *   postprocessing.py:188-229 `unwarp_slice_backward`        (coord_round = 0,
 *       row0 = index, nrows = 1: float64 coordinates, never rounded), and
 *   postprocessing.py:255-313 `unwarp_chunk_slices_backward` (coord_round = 1,
 *       row0 = start_index, nrows = stop-start+1: fp32-rounded coordinates),
 * and, with row0 = 0, nrows = H, a whole stack / batch of independent images
 * (BASELINE configs 4 and 5).  The images are (D, H, W) float32; dst is
 * (D, nrows, W) float32.  The radial map is evaluated once per output tile and
 * reused for every slice.
 * Like the reference (:221-223, :300-301) the caller may hold only a window of
 * source rows on the device: `src` points at image row `src_row0` of slice 0
 * and every slice holds `src_rows` rows (src_row0 = 0, src_rows = H for whole
 * slices).  Tap rows are clamped into the window, and the "+1" bilinear tap
 * never goes past its last row -- exactly what sampling the reference's cropped
 * view does. */
int dcb_unwarp_stack_backward_f32(const float *src, float *dst, int D, int H, int W,
                                  int src_row0, int src_rows,
                                  size_t src_pitch, size_t src_slice_stride,
                                  size_t dst_pitch, size_t dst_slice_stride,
                                  int row0, int nrows, int coord_round,
                                  const dcb_radial *model_host,
                                  const dcb_options *opt_host, void *stream);

/* Replaces postprocessing.py:444-459 `_generate_perspective_map` +
 * :462-492 `correct_perspective_image` (map_index=None). */
int dcb_correct_perspective_image_f32(const float *src, float *dst, int H, int W,
                                      size_t src_pitch, size_t dst_pitch,
                                      const dcb_persp *model_host,
                                      const dcb_options *opt_host, void *stream);

/* Replaces the scipy.ndimage.map_coordinates call itself for caller-supplied
 * coordinates: postprocessing.py:232-252 `_mapping` and the `map_index=`
 * argument of `correct_perspective_image` (:489-491).  yd/xd are device arrays
 * of n_out coordinates, float32 (coord_is_f64 = 0) or float64 (= 1); they are
 * clamped to the image like every coordinate the reference produces.  dst
 * receives n_out float32 values.  If oob_count (device uint32, caller-zeroed,
 * may be NULL) is given it is incremented by the number of coordinates that
 * lay outside [0,H-1]x[0,W-1] (or were NaN) before clamping, so the host can
 * refuse boundary modes for which clamping is not what SciPy does. */
int dcb_map_coordinates_f32(const float *src, float *dst, int H, int W, size_t src_pitch,
                            const void *yd, const void *xd, int coord_is_f64,
                            size_t n_out, uint32_t *oob_count,
                            const dcb_options *opt_host, void *stream);

/* The combined entry BASELINE.json's north_star names
 * (`unwarp_image_backward_perspective`; not in the reference).  Defined as the
 * two-pass composition of examples/readthedocs_demo/demo_05.py:127 then :147
 * with the intermediate image rounded to float32.  `scratch` is a caller-owned
 * (H, W) float32 device buffer with pitch `scratch_pitch` for the intermediate. */
int dcb_unwarp_image_backward_perspective_f32(const float *src, float *dst, float *scratch,
                                              int H, int W, size_t src_pitch,
                                              size_t dst_pitch, size_t scratch_pitch,
                                              const dcb_radial *radial_host,
                                              const dcb_persp *persp_host,
                                              const dcb_options *opt_host, void *stream);

/* Pixel types of host frames; the remap kernels themselves work on float32. */
enum dcb_dtype { DCB_DTYPE_F32 = 0, DCB_DTYPE_U8 = 1, DCB_DTYPE_I8 = 2, DCB_DTYPE_U16 = 3, DCB_DTYPE_I16 = 4 };

/* Camera frames: dense interleaved (H, W, C) device array of `dtype` <-> C
 * float32 planes (row pitch / plane stride in bytes).  Replaces the host-side
 * per-channel slicing of discorpy/util/utility.py:337-341 (mat_pad[:, :, i])
 * and the np.moveaxis at :341; integer samples widen exactly, and narrow back
 * exactly after a DCB_FLAG_ROUND_INT remap.  C = 1 converts a 2-D image. */
int dcb_unpack_hwc_to_planes_f32(const void *src_hwc, int dtype, float *dst_planes, int H, int W,
                                 int C, size_t dst_pitch, size_t dst_plane_stride, void *stream);
int dcb_pack_planes_f32_to_hwc(const float *src_planes, void *dst_hwc, int dtype, int H, int W,
                               int C, size_t src_pitch, size_t src_plane_stride, void *stream);

/* Benchmark input generator (no reference counterpart): dst[i] =
 * top24(splitmix64(seed ^ (offset + i))) / 2^24, float32 in [0,1).  Lets
 * multi-GB stacks be created in HBM without crossing PCIe. */
int dcb_fill_synthetic_f32(float *dst, size_t n, uint64_t seed, uint64_t offset,
                           void *stream);

/* ---- diagnostics --------------------------------------------------------- */

/* ---- spline orders 2..5 and float64 images ------------
 * `order=` / `mode=` of postprocessing.py:147, :491 and util/utility.py:333,
 * :338 beyond bilinear: scipy.ndimage.map_coordinates' float64 B-spline
 * prefilter (scipy/ndimage/_interpolation.py:467-469, pre-padding :212-227)
 * and its (order+1)^2-tap interpolation, restated operation by operation
 * (oracle/oracle_spline.py) so that results are bit-identical to SciPy for the
 * same coordinates.  Two steps, so that one prefiltered image can serve several
 * maps:
 *   1. dcb_spline_prefilter: (H, W) float32 or float64 image -> float64
 *      coefficients at the start of `workspace` (dense (H+2p) x (W+2p) doubles,
 *      p = 12 for nearest / grid-constant, else 0; order <= 1: a widened copy);
 *   2. dcb_spline_remap: samples them through the radial map, the projective
 *      map or caller-supplied coordinates (clamped into the image; *oob_count
 *      counts those that were outside) into a float32 or float64 destination.
 * Integer images travel as float32: DCB_FLAG_ROUND_INT rounds half away from
 * zero and saturates at [sat_lo, sat_hi] like SciPy's integer outputs. */
enum dcb_mode {
    DCB_MODE_REFLECT = 0, DCB_MODE_GRID_MIRROR = 1, DCB_MODE_CONSTANT = 2,
    DCB_MODE_GRID_CONSTANT = 3, DCB_MODE_NEAREST = 4, DCB_MODE_MIRROR = 5,
    DCB_MODE_GRID_WRAP = 6, DCB_MODE_WRAP = 7
};
enum dcb_map_kind { DCB_MAP_RADIAL = 0, DCB_MAP_PERSP = 1, DCB_MAP_COORDS = 2 };

int dcb_spline_workspace_bytes(int H, int W, int order, int mode, size_t *bytes);
int dcb_spline_prefilter(const void *src, int src_is_f64, int H, int W, size_t src_pitch,
                         int order, int mode, void *workspace, size_t workspace_bytes,
                         void *stream);
int dcb_spline_remap(const void *workspace, int H, int W, int order, int mode,
                     void *dst, int dst_is_f64, size_t dst_pitch,
                     int map_kind, const dcb_radial *radial_host, const dcb_persp *persp_host,
                     const void *yd, const void *xd, int coord_is_f64, size_t n_out,
                     uint32_t *oob_count, int flags, double sat_lo, double sat_hi,
                     void *stream);

/* postprocessing.py:151-185 `unwarp_image_forward` for a device-resident float32
 * image: every source pixel moves to
 * round-half-even(clip(centre + F(rd) (p - centre))); unreached outputs are 0;
 * collisions keep the source with the largest linear index (NumPy's
 * last-assignment-wins), made deterministic with an atomicMax pass.
 * `workspace`: H*W uint32 on the device. */
int dcb_unwarp_image_forward_f32(const float *src, float *dst, int H, int W, size_t src_pitch,
                                 size_t dst_pitch, const dcb_radial *model_host,
                                 uint32_t *workspace, void *stream);

/* Number of compute-kernel launches issued by this library in the calling process
 * (all threads) since load / since the last reset.  Plan builds of the single-image kernel (one
 * small kernel the first time a (model, geometry) pair is seen) are counted separately, see
 * dcb_plan_cache_clear. */
int dcb_launch_count(uint64_t *count);
int dcb_launch_count_reset(void);

/* What the last compute call on this thread decided: path actually used
 * (enum dcb_path, never AUTO), staged box width/height, grid size, dynamic
 * shared memory bytes. */
int dcb_last_plan(int *path, int *box_w, int *box_h, int *grid, int *smem_bytes);

/* ---- multi-GPU plumbing: one process per GPU, NCCL (csrc/mg.cu) --------------------------
 * The path shards without a data-plane collective (every slice / image is independent); the
 * ranks exchange the parameter block (one broadcast; reference: the scalars every call of
 * postprocessing.py:111, :188, :255 takes), optionally the rows of one assembled sinogram
 * (postprocessing.py:224-229 on a sharded stack), and benchmark scalars.  NCCL is loaded at run
 * time (libnccl.so.2, or $DCB_NCCL_LIB); DCB_ERR_UNSUPPORTED when it cannot be found.
 *   dcb_mg_unique_id  rank 0: a fresh NCCL unique id (128 bytes) to hand to the other ranks by
 *                     whatever the launcher offers (discorpy_b200/multigpu.py: a TCP exchange
 *                     on MASTER_ADDR); *nccl_version, when not NULL, the library version
 *   dcb_mg_init       every rank, after dcb_init(local device): ncclCommInitRank
 *   dcb_mg_bcast      device buffer, asynchronous on `stream`
 *   dcb_mg_bcast_host <= 4096 host bytes, synchronous (the parameter block)
 *   dcb_mg_allgather  nbytes_per_rank from every rank, rank order, asynchronous on `stream`
 *   dcb_mg_allreduce_max_f64  up to 8 host doubles, in place, synchronous (max over ranks of a
 *                     device-timed duration)
 *   dcb_mg_barrier    device idle on every rank (cudaDeviceSynchronize + a 1-element all-reduce)
 *   dcb_mg_finalize   destroys the communicator */
#define DCB_MG_UNIQUE_ID_BYTES 128
int dcb_mg_unique_id(void *id_out, int *nccl_version);
int dcb_mg_init(const void *id, int world, int rank);
int dcb_mg_info(int *world, int *rank);
int dcb_mg_bcast(void *dev_buf, size_t nbytes, int root, void *stream);
int dcb_mg_bcast_host(void *host_buf, size_t nbytes, int root);
int dcb_mg_allgather(const void *send_dev, void *recv_dev, size_t nbytes_per_rank, void *stream);
int dcb_mg_allreduce_max_f64(double *host_values, int count);
int dcb_mg_barrier(void);
int dcb_mg_finalize(void);

/* Drops every cached plan of the single-image kernel (per-tile staged boxes and verified row
 * patches, built once per (model, geometry) and reused for later frames; csrc/remap_image.cuh,
 * csrc/api.cu "Plan cache") on all devices; *plans_built, when not NULL, receives the number of
 * plans built by this process so far.  Environment: DCB_PLAN_CACHE=0 builds a plan per launch,
 * DCB_PLAN_CACHE_MB bounds the cache (default 512). */
int dcb_plan_cache_clear(uint64_t *plans_built);

/* Diagnostics of the single-image kernel's patch path (csrc/remap_image.cuh, RowPatch): counting is
 * switched on with enable != 0 (a small device buffer is allocated on the current device) and off
 * with 0; out[0..7], when not NULL, receives since the last reset: [3] patch rows redone exactly
 * because of the blend certificate, [4] tiles holding values the certified blend does not cover,
 * and from the plans BUILT while counting was on: [5] tile rows verified for the patch path in
 * full, [6] rows verified in part (such rows take the exact path), [7] rows in total; why rows
 * were not verified: [0] their tile is not eligible (partial width, box not staged, centre outside
 * the float32 binades), [1] the row's y binade, [2] no 32-pixel segment passed.  reset != 0 clears the counters after reading. */
int dcb_image_stats(int enable, uint64_t *out, int reset);

/* Timeline of the LAST single-image launch made while dcb_image_stats counting was on: for CTA b
 * (b < nctas <= 1024) out[8 b + 0..5] = %globaltimer (ns) at CTA start, when its first tile was
 * ready to be sampled, when its first sampling warp finished its last tile,
 * (SM id | tiles of this CTA << 32), the time that warp waited for later tiles in total and the
 * longest such wait (ns); [6], [7] reserved.  nctas = -1: the per-warp event log of the first four
 * CTAs instead (4 x 10 x 64 time stamps, -DDCB_IMG_TIMELINE builds).  Diagnostics behind
 * profiles/r2/timeline_*.txt; an ordinary build reports zeros. */
int dcb_image_timeline(uint64_t *out, int nctas);

/* Device self-test of the custom fp64 square root used by the radial kernels
 * (probe points, Z-stack geometry) against IEEE sqrt on n pseudo-random inputs; *mismatch receives the number
 * of inputs whose result differs from the correctly rounded one. */
int dcb_selftest_sqrt(size_t n, uint64_t seed, uint64_t *mismatch);

/* Same inputs through the single-image kernel's 5-operation square root
 * (one ulp, not correctly rounded): *differ = results that are not the IEEE
 * value, *beyond_one_ulp = results more than one ulp away (expected 0). */
int dcb_selftest_sqrt_fast(size_t n, uint64_t seed, uint64_t *differ, uint64_t *beyond_one_ulp);

/* TMA probe (diagnostics): loads the (box_w x box_h) box whose first element is
 * (x0, y0, z0) of the (D, H, W) float32 tensor `src` (row pitch / slice stride
 * in bytes, multiples of 16) through cp.async.bulk.tensor into shared memory
 * and copies it to `out` (box_w*box_h floats, device).  *status: 0 = ok,
 * 1 = the copy never completed.  Out-of-range elements arrive as 0. */
int dcb_selftest_tma(const float *src, int D, int H, int W, size_t pitch, size_t slice_stride,
                     int box_w, int box_h, int x0, int y0, int z0, float *out, int *status);

/* Micro-benchmarks used to size the kernels (DESIGN.md "fp64 budget"):
 * which = 0 DFMA chain, 1 f32<->f64 conversions, 2 MUFU.RSQ64H, 3 the radial
 * coordinate evaluation alone (5 terms), 4 FFMA, 5 DFMA + 2 conversions:
 * giga-ops/s.  which = 6 DFMA, 7 f32->f64->f32 round trip, 8 MUFU.RSQ64H,
 * 9 LDS.64: dependent-issue latency in clock cycles (one warp, one chain). */
int dcb_microbench(int which, double *gops);

#ifdef __cplusplus
}
#endif
#endif /* DISCORPY_B200_H */

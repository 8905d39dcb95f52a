// api.cu -- C ABI of libdiscorpy_b200.so (declared in include/discorpy_b200.h):
// argument validation, launch planning (tile grid, staged-box size, TMA
// descriptor) and kernel dispatch.  Host side only; kernels are in remap.cuh.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include <pthread.h>

#include "api_common.hpp"
#include "remap.cuh"
#include "remap_image.cuh"
#include "remap_stack.cuh"
#include "convert.cuh"
#include "spline.cuh"
#include "forward.cuh"

using namespace dcb;

// ---------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

struct LastPlan {
    int path, bw, bh, grid, smem;
};
static thread_local LastPlan g_last_plan = {DCB_PATH_DIRECT, 0, 0, 0, 0};

void dcb::count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int dcb::fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ---------------------------------------------------------------------------
// driver entry point for the TMA descriptor encoder (no link-time libcuda)
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static EncodeTiledFn tma_encoder();
void *dcb::tma_encode_fn() { return reinterpret_cast<void *>(tma_encoder()); }

static EncodeTiledFn tma_encoder() {
    std::call_once(g_encode_once, [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    });
    return g_encode;
}

struct DevProps {
    int sm_count = 0;
    int smem_optin = 0;
    bool ok = false;
};
static DevProps g_props[64];
static std::mutex g_props_mu;

static int device_props(DevProps *out) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(DCB_ERR_ARG, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_props_mu);
    if (!g_props[dev].ok) {
        CUDA_TRY(cudaDeviceGetAttribute(&g_props[dev].sm_count, cudaDevAttrMultiProcessorCount, dev));
        CUDA_TRY(cudaDeviceGetAttribute(&g_props[dev].smem_optin,
                                        cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        g_props[dev].ok = true;
    }
    *out = g_props[dev];
    return DCB_OK;
}

// ---------------------------------------------------------------------------
// launch planning
// ---------------------------------------------------------------------------
// Upper bounds on how far the source coordinate moves per output pixel:
//   gmain  >= |d xd / d x| , |d yd / d y|      gcross >= |d xd / d y| , |d yd / d x|
// For the radial map xd = xc + F(r) xu: d xd/d x = F + F' xu^2 / r and
// d xd/d y = F' xu yu / r, so |F| + |F' r| and |F' r| / 2 bound them; both are
// sampled on a 1-D grid of r over the radii the requested rows can reach.
static void radial_slopes(const dcb_radial &m, int W, int row0, int nrows, double *gmain,
                          double *gcross) {
    const double xs[2] = {0.0 - m.xc, (double)(W - 1) - m.xc};
    const double ys[2] = {(double)row0 - m.yc, (double)(row0 + nrows - 1) - m.yc};
    double rmax = 0.0;
    for (double x : xs)
        for (double y : ys) rmax = std::max(rmax, std::sqrt(x * x + y * y));
    double gm = 0.0, gc = 0.0;
    const int S = 512;
    for (int i = 0; i <= S; ++i) {
        const double r = rmax * i / S;
        double f = 0.0, fp = 0.0;  // F and F' by Horner
        for (int k = m.n - 1; k >= 0; --k) {
            fp = fp * r + f;
            f = f * r + m.a[k];
        }
        gm = std::max(gm, std::fabs(f) + std::fabs(fp * r));
        gc = std::max(gc, 0.5 * std::fabs(fp * r));
    }
    *gmain = gm;
    *gcross = gc;
}

static void persp_slopes(const dcb_persp &m, int H, int W, double *gmain, double *gcross) {
    // The Jacobian of a projective map is monotone along lines where the
    // denominator keeps its sign; sample a 9x9 grid and keep the maxima.
    double gm = 0.0, gc = 0.0;
    const double *c = m.c;
    for (int iy = 0; iy <= 8; ++iy)
        for (int ix = 0; ix <= 8; ++ix) {
            const double x = (W - 1) * ix / 8.0, y = (H - 1) * iy / 8.0;
            const double den = c[6] * x + c[7] * y + 1.0;
            const double nx = c[0] * x + c[1] * y + c[2], ny = c[3] * x + c[4] * y + c[5];
            if (den == 0.0) continue;
            const double xx = (c[0] * den - nx * c[6]) / (den * den);
            const double xy = (c[1] * den - nx * c[7]) / (den * den);
            const double yx = (c[3] * den - ny * c[6]) / (den * den);
            const double yy = (c[4] * den - ny * c[7]) / (den * den);
            gm = std::max(gm, std::max(std::fabs(xx), std::fabs(yy)));
            gc = std::max(gc, std::max(std::fabs(xy), std::fabs(yx)));
        }
    *gmain = gm;
    *gcross = gc;
}

typedef void (*StackKernel)(const RemapParams, const CUtensorMap);
static int image_sched_slot(unsigned **out);   // two self-resetting counters per launch, below

// Launch planning for the Z-stack kernel (remap_stack.cuh).  kerns[0] is the 128 x 16 tile
// instantiation, kerns[1] the 64 x 32 one.
static int plan_and_launch_stack(const StackKernel (&kerns)[2], bool widen, RemapParams &p, double gmain,
                                 double gcross,
                                 int path_req, size_t src_pitch_bytes, size_t src_slice_bytes,
                                 cudaStream_t stream) {
    DevProps props;
    int rc = device_props(&props);
    if (rc != DCB_OK) return rc;
    if (!(gmain < 64.0) || !(gcross < 64.0)) gmain = gcross = 64.0;  // NaN / absurd
    const int max_stage = 24 * 1024;
    // footprint bound of a tw x th tile: +3 footprint slack, +3 because the box start is aligned
    // down to 4 floats
    auto bound = [&](int tw_, int th_, long long *w, long long *h) {
        const double tw = std::min(tw_, p.W) - 1, th = std::min(th_, p.nrows) - 1;
        *w = (long long)std::ceil(gmain * tw + gcross * th) + 3 + 3;
        *h = (long long)std::ceil(gmain * th + gcross * tw) + 3;
    };
    // Tile shape: 128 x 16 whenever its footprint bound can be staged; otherwise the shape whose
    // bound overshoots its largest stageable box the least (a sheared map stretches the source
    // rows of a wide tile: config 5 needs up to 165 rows for 128 x 16 output pixels).
    // DCB_STK_SHAPE=0/1 forces a shape (A/B runs).
    int shape = 0;
    {
        long long wa, ha, wb, hb;
        bound(kTileW, kStkTileH, &wa, &ha);
        bound(64, 32, &wb, &hb);
        const bool a_fits = wa <= 256 && ha <= 256 && wa * ha * 4 <= max_stage;
        if (!a_fits) {
            const double over_a = std::max(1.0, wa / 144.0) * std::max(1.0, ha / 42.0);
            const double over_b = std::max(1.0, wb / 96.0) * std::max(1.0, hb / 64.0);
            shape = over_b < over_a ? 1 : 0;
        }
        if (const char *env = getenv("DCB_STK_SHAPE")) shape = atoi(env) ? 1 : 0;
    }
    const StackKernel kern = kerns[shape];
    const int TW = shape ? 64 : kTileW, TH = shape ? 32 : kStkTileH;
    p.tiles_x = (p.W + TW - 1) / TW;
    p.tiles_y = (p.nrows + TH - 1) / TH;
    const long long tiles_xy = (long long)p.tiles_x * p.tiles_y;

    // --- can this layout be described to TMA? -------------------------------
    const bool layout_ok = ((uintptr_t)p.src % 16 == 0) && (src_pitch_bytes % 16 == 0) &&
                           (src_slice_bytes % 16 == 0) && tma_encoder() != nullptr;
    if (path_req == DCB_PATH_TMA && !layout_ok)
        return fail(DCB_ERR_ARG,
                    "DCB_PATH_TMA needs a 16-byte aligned source with pitch and slice stride "
                    "multiples of 16 bytes (and a driver exporting cuTensorMapEncodeTiled)");
    bool staged = layout_ok && path_req != DCB_PATH_DIRECT;

    // --- staged box: bound of the tile's source footprint --------------------
    const int src_rows = p.ylast - p.yorg + 1;
    int bw = 0, bh = 0, nstage = 0;
    if (staged) {
        long long need_w, need_h;
        bound(TW, TH, &need_w, &need_h);
        need_w = std::min<long long>(need_w, (long long)p.W + 3);
        need_h = std::min<long long>(need_h, src_rows);
        bw = (int)((need_w + 3) / 4 * 4);
        bh = (int)need_h;
        if (bw > 256 || bh > 256 || (long long)bw * bh * 4 > max_stage) {
            // footprint bound too large (strong magnification somewhere): stage the largest box
            // a stage holds (144 x 42 / 96 x 64); tiles that do not fit it fall back to direct
            // gathers.  (Round 1 staged only TH + 8 rows here: 15 % of config 4's tiles missed.)
            bw = std::min(256, (std::min(TW + (shape ? 32 : 16), (p.W + 3) / 4 * 4 + 4)));
            bh = std::min(src_rows, max_stage / (bw * 4));
        }
        if (const char *env = getenv("DCB_STK_BOX")) {   // diagnostics: "w,h" forces the staged box
            int ew = 0, eh = 0;
            if (sscanf(env, "%d,%d", &ew, &eh) == 2 && ew >= 4 && ew <= 256 && eh >= 1 && eh <= 256) {
                bw = ew / 4 * 4;
                bh = std::min(eh, src_rows);
            }
        }
        bw = std::max(bw, 4);
        bh = std::max(bh, 1);
        // ring depth: as many slices in flight as ~96 KB per CTA hold (two CTAs per SM)
        const long long stage = ((long long)bw * bh * 4 + 127) / 128 * 128;
        // (the fp64 blends keep two float64 copies of a box = 4 stages' worth next to the ring)
        nstage = (int)std::max<long long>(
            2, std::min<long long>(kStkMaxStages, (DCB_STK_SMEM_KB * 1024) / stage - (widen ? 4 : 0)));
    }
    p.bw = bw;
    p.bh = bh;
    p.nstage = nstage;
    p.box_bytes = (unsigned)(bw * bh * 4);
    p.stage_bytes = (p.box_bytes + 127u) / 128u * 128u;
    const size_t tail = 2 * kStkMaxStages * sizeof(uint64_t) + 2 * 4 * kWarps * sizeof(int);
    size_t smem = (size_t)(nstage + (widen && nstage > 0 ? 4 : 0)) * p.stage_bytes + tail;

    // --- z chunking: enough items to balance the machine, long enough chunks to
    //     amortise the per-tile geometry ----------------------------------------------
    int zchunk = 1;
    if (p.D > 1) {
        const long long want_items = (long long)props.sm_count * 2 * 8;
        long long nchunks = std::max<long long>(1, (want_items + tiles_xy - 1) / tiles_xy);
        nchunks = std::min<long long>(nchunks, p.D);
        zchunk = (int)((p.D + nchunks - 1) / nchunks);
        zchunk = std::min(zchunk, 64);
    }
    p.zchunk = zchunk;
    const long long nzc = (p.D + zchunk - 1) / zchunk;
    const long long nitems = tiles_xy * nzc;
    if (nitems > INT_MAX) return fail(DCB_ERR_UNSUPPORTED, "too many work items (%lld)", nitems);
    p.ntiles = (int)nitems;

    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (staged) {
        const cuuint64_t gdim[3] = {(cuuint64_t)p.W, (cuuint64_t)src_rows, (cuuint64_t)p.D};
        const cuuint64_t gstr[2] = {(cuuint64_t)src_pitch_bytes,
                                    (cuuint64_t)(p.D > 1 ? src_slice_bytes
                                                         : src_pitch_bytes * (size_t)src_rows)};
        const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        CUresult cr = tma_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)p.src, gdim,
                                    gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            if (path_req == DCB_PATH_TMA)
                return fail(DCB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
            staged = false;
            p.nstage = 0;
            smem = tail;
        }
    }
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute((const void *)kern,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)kern, kThreads, smem));
    if (occ < 1) return fail(DCB_ERR_CUDA, "kernel does not fit on an SM (smem %zu)", smem);
    const int grid = (int)std::min<long long>(nitems, (long long)occ * props.sm_count);

    rc = image_sched_slot(&p.sched);
    if (rc != DCB_OK) return rc;
    kern<<<grid, kThreads, smem, stream>>>(p, tmap);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    g_last_plan = {staged ? DCB_PATH_TMA : DCB_PATH_DIRECT, bw, bh, grid, (int)smem};
    return DCB_OK;
}

typedef void (*ImageKernel)(const ImageParams, const CUtensorMap);
static unsigned long long *g_image_stats = nullptr;   // dcb_image_stats: device counters or NULL
static unsigned long long *g_image_stats_fwd() { return g_image_stats; }

// Launch planning for the single-image kernel (remap_image.cuh).
struct ImageKernelSel {
    ImageKernel kern;
    bool wide;  // unused since round 2 (every kernel samples the raw float32 stages)
    int th;     // tile height the kernel was instantiated for
};

// Does the map send any pixel of the output rows' border outside the image (so that part of the
// output is clipped and takes the slow row path)?  Sampled on the four edges; the maps of this
// path are monotone enough that clipping starts at the border.
static bool map_clips_at_border(const ImageParams &p, int map_kind) {
    const int y0 = p.row0, y1 = p.row0 + p.nrows - 1;
    const int sx = std::max(1, p.W / 64), sy = std::max(1, p.nrows / 64);
    auto outside = [&](int x, int y) -> bool {
        double xd, yd;
        if (map_kind == MAP_RADIAL) {
            const double xu = x - p.rad.xc, yu = y - p.rad.yc, r = std::sqrt(xu * xu + yu * yu);
            double f = 0.0;
            for (int i = p.rad.n - 1; i >= 0; --i) f = f * r + p.rad.a[i];
            xd = p.rad.xc + f * xu;
            yd = p.rad.yc + f * yu;
        } else {
            const double den = p.per.c[6] * x + p.per.c[7] * y + 1.0;
            xd = (p.per.c[0] * x + p.per.c[1] * y + p.per.c[2]) / den;
            yd = (p.per.c[3] * x + p.per.c[4] * y + p.per.c[5]) / den;
        }
        return !(xd >= 0.0 && xd <= (double)(p.W - 1) && yd >= 0.0 && yd <= (double)(p.H - 1));
    };
    for (int x = 0; x < p.W; x += sx)
        if (outside(x, y0) || outside(x, y1)) return true;
    for (int y = y0; y <= y1; y += sy)
        if (outside(0, y) || outside(p.W - 1, y)) return true;
    return outside(p.W - 1, y0) || outside(p.W - 1, y1);
}

// ---------------------------------------------------------------------------
// Plan cache of the single-image kernel (remap_image.cuh, TilePlan): the per-tile staged boxes and
// verified row patches depend on the model and the launch geometry alone, so they are built once
// (image_plan_kernel, ~35 us for 4096^2) and kept on the device for every later frame unwarped
// with the same calibration -- the reference's usage (one calibration, many projections).
// Least recently used plans are dropped beyond DCB_PLAN_CACHE_MB (default 512); DCB_PLAN_CACHE=0
// builds a fresh plan for every launch (stream-ordered allocation).
// ---------------------------------------------------------------------------
namespace {
struct PlanKey {
    int device, map_kind, th, H, W, row0, nrows, yorg, ylast, bw, bh, fast, n, grid, nstatic;
    double xc, yc, a[DCB_MAX_TERMS], c[8];
};
struct PlanEntry {
    void *dptr = nullptr;
    size_t bytes = 0;
    cudaEvent_t ready = nullptr;
    bool ready_done = false;
    uint64_t last_use = 0;
};
std::map<std::string, PlanEntry> g_plans;
std::mutex g_plans_mu;
uint64_t g_plan_clock = 0;
size_t g_plan_bytes = 0;
std::atomic<uint64_t> g_plan_builds{0};

// Plan memory: TilePlan<TH>[ntiles] | int2 boxes[ntiles] | int starts[grid + 1] | unsigned cost[ntiles]
struct PlanLayout {
    size_t boxes, starts, cost, bytes;
};
PlanLayout plan_layout(int th, int ntiles, int grid) {
    PlanLayout l;
    l.boxes = (size_t)ntiles * image_rec_bytes(th);
    l.starts = l.boxes + (size_t)ntiles * sizeof(int2);
    l.cost = (l.starts + (size_t)(grid + 1) * sizeof(int) + 15) / 16 * 16;
    l.bytes = l.cost + (size_t)ntiles * sizeof(unsigned);
    return l;
}
template <int MAP, int TH>
void launch_plan_kernel(const ImageParams &p, void *out, const PlanLayout &l,
                        unsigned long long *stats, cudaStream_t st) {
    char *base = reinterpret_cast<char *>(out);
    image_plan_kernel<MAP, TH><<<p.ntiles, kThreads, 0, st>>>(
        p, reinterpret_cast<TilePlan<TH> *>(out), reinterpret_cast<int2 *>(base + l.boxes),
        reinterpret_cast<unsigned *>(base + l.cost), stats);
}
// builds the plan of launch p (p.ntiles tiles, `grid` CTAs, p.nstatic statically split tiles)
void launch_plan(int map_kind, int th, const ImageParams &p, int grid, void *out,
                 unsigned long long *stats, cudaStream_t st) {
    const PlanLayout l = plan_layout(th, p.ntiles, grid);
    if (map_kind == MAP_RADIAL)
        th == 32 ? launch_plan_kernel<MAP_RADIAL, 32>(p, out, l, stats, st)
                 : launch_plan_kernel<MAP_RADIAL, 16>(p, out, l, stats, st);
    else
        th == 32 ? launch_plan_kernel<MAP_PERSP, 32>(p, out, l, stats, st)
                 : launch_plan_kernel<MAP_PERSP, 16>(p, out, l, stats, st);
    char *base = reinterpret_cast<char *>(out);
    image_plan_ranges_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const unsigned *>(base + l.cost),
                                                 p.nstatic, grid,
                                                 reinterpret_cast<int *>(base + l.starts));
}
// Plan memory comes from slabs the library keeps (first fit, neighbours coalesced on release): a
// cudaMalloc of 8.6 MB took 2-6 ms on the round's boxes -- ten times the plan kernels -- and a
// caller trying calibrations pays it per model (tools/plan_probe.py).  Guarded by g_plans_mu.
struct PlanSlab {
    int device;
    char *base;
    size_t bytes;
    std::map<size_t, size_t> free_;   // offset -> length
};
std::vector<PlanSlab> g_plan_slabs;
constexpr size_t kPlanSlabBytes = (size_t)64 << 20;
void *plan_alloc(int device, size_t bytes) {
    bytes = (bytes + 255) / 256 * 256;
    for (int pass = 0; pass < 2; ++pass) {
        for (auto &sl : g_plan_slabs) {
            if (sl.device != device) continue;
            for (auto it = sl.free_.begin(); it != sl.free_.end(); ++it) {
                if (it->second < bytes) continue;
                const size_t off = it->first, len = it->second;
                sl.free_.erase(it);
                if (len > bytes) sl.free_[off + bytes] = len - bytes;
                return sl.base + off;
            }
        }
        if (pass == 1) break;
        PlanSlab sl;
        sl.device = device;
        sl.bytes = std::max(kPlanSlabBytes, bytes);
        void *d = nullptr;
        if (cudaMalloc(&d, sl.bytes) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        sl.base = reinterpret_cast<char *>(d);
        sl.free_[0] = sl.bytes;
        g_plan_slabs.push_back(std::move(sl));
    }
    return nullptr;
}
// (the caller has made sure no kernel still reads the block: cudaDeviceSynchronize)
void plan_release(void *ptr, size_t bytes) {
    bytes = (bytes + 255) / 256 * 256;
    char *q = reinterpret_cast<char *>(ptr);
    for (auto &sl : g_plan_slabs) {
        if (q < sl.base || q >= sl.base + sl.bytes) continue;
        size_t off = (size_t)(q - sl.base), len = bytes;
        auto next = sl.free_.lower_bound(off);
        if (next != sl.free_.end() && off + len == next->first) {
            len += next->second;
            next = sl.free_.erase(next);
        }
        if (next != sl.free_.begin()) {
            auto prev = std::prev(next);
            if (prev->first + prev->second == off) {
                off = prev->first;
                len += prev->second;
                sl.free_.erase(prev);
            }
        }
        sl.free_[off] = len;
        return;
    }
}

size_t plan_cache_limit() {
    static size_t lim = [] {
        const char *e = getenv("DCB_PLAN_CACHE_MB");
        long mb = e ? atol(e) : 512;
        return (size_t)std::max(16L, mb) << 20;
    }();
    return lim;
}
bool plan_cache_enabled() {
    static bool on = [] {
        const char *e = getenv("DCB_PLAN_CACHE");
        return !(e != nullptr && e[0] == '0');
    }();
    return on;
}
}  // namespace

static unsigned long long *g_image_stats_fwd();

// Counters of the single-image kernel's tile pool: each launch takes the next of 1024 slots of its
// device (two words, zero when the launch starts; the last CTA of a launch zeroes them again), so
// launches in flight on different streams never share one.
static unsigned *g_sched_slots[64];
static std::atomic<unsigned> g_sched_next{0};
static std::mutex g_sched_mu;
constexpr int kSchedSlots = 1024;
static int image_sched_slot(unsigned **out) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(DCB_ERR_ARG, "device index %d out of range", dev);
    {
        std::lock_guard<std::mutex> lk(g_sched_mu);
        if (g_sched_slots[dev] == nullptr) {
            unsigned *d = nullptr;
            CUDA_TRY(cudaMalloc((void **)&d, kSchedSlots * 2 * sizeof(unsigned)));
            CUDA_TRY(cudaMemset(d, 0, kSchedSlots * 2 * sizeof(unsigned)));
            g_sched_slots[dev] = d;
        }
    }
    *out = g_sched_slots[dev] + 2 * (g_sched_next.fetch_add(1, std::memory_order_relaxed) % kSchedSlots);
    return DCB_OK;
}

// Finds or builds the plan of this launch on `stream`; *transient receives a buffer the caller
// must release with cudaFreeAsync after the launch (cache disabled), else NULL.
static int get_image_plan(int map_kind, int th, ImageParams &p, int grid, cudaStream_t stream,
                          void **transient) {
    *transient = nullptr;
    const PlanLayout lay = plan_layout(th, p.ntiles, grid);
    const size_t bytes = lay.bytes;
    p.plan_ready = 0;
    auto bind = [&](void *d) {
        p.plan = d;
        p.plan_boxes = reinterpret_cast<const int2 *>(reinterpret_cast<char *>(d) + lay.boxes);
        p.plan_starts = reinterpret_cast<const int *>(reinterpret_cast<char *>(d) + lay.starts);
    };
    if (!plan_cache_enabled()) {
        void *d = nullptr;
        CUDA_TRY(cudaMallocAsync(&d, bytes, stream));
        launch_plan(map_kind, th, p, grid, d, g_image_stats_fwd(), stream);
        CUDA_TRY(cudaGetLastError());
        g_plan_builds.fetch_add(1, std::memory_order_relaxed);
        bind(d);
        *transient = d;
        return DCB_OK;
    }
    PlanKey key;
    memset(&key, 0, sizeof(key));
    CUDA_TRY(cudaGetDevice(&key.device));
    key.map_kind = map_kind, key.th = th, key.H = p.H, key.W = p.W, key.row0 = p.row0;
    key.nrows = p.nrows, key.yorg = p.yorg, key.ylast = p.ylast, key.bw = p.bw, key.bh = p.bh;
    key.fast = p.fast;
    key.grid = grid, key.nstatic = p.nstatic;
    if (map_kind == MAP_RADIAL) {
        key.n = p.rad.n, key.xc = p.rad.xc, key.yc = p.rad.yc;
        memcpy(key.a, p.rad.a, sizeof(key.a));
    } else {
        memcpy(key.c, p.per.c, sizeof(key.c));
    }
    const std::string k(reinterpret_cast<const char *>(&key), sizeof(key));
    std::lock_guard<std::mutex> lk(g_plans_mu);
    auto it = g_plans.find(k);
    if (it == g_plans.end()) {
        // make room: drop least recently used plans (cudaFree waits for kernels still reading them)
        while (!g_plans.empty() && g_plan_bytes + bytes > plan_cache_limit()) {
            auto old = g_plans.begin();
            for (auto jt = g_plans.begin(); jt != g_plans.end(); ++jt)
                if (jt->second.last_use < old->second.last_use) old = jt;
            cudaDeviceSynchronize();   // kernels on any stream may still read it
            plan_release(old->second.dptr, old->second.bytes);
            cudaEventDestroy(old->second.ready);
            g_plan_bytes -= old->second.bytes;
            g_plans.erase(old);
        }
        PlanEntry e;
        const bool trace = getenv("DCB_TRACE_PLAN") != nullptr;
        const auto tr0 = std::chrono::steady_clock::now();
        e.dptr = plan_alloc(key.device, bytes);
        if (e.dptr == nullptr) return fail(DCB_ERR_CUDA, "out of device memory for a plan of %zu bytes", bytes);
        if (trace)
            fprintf(stderr, "[dcb] plan memory (%zu bytes) %.0f us\n", bytes,
                    std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tr0).count());
        e.bytes = bytes;
        cudaError_t ce = cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming);
        if (ce != cudaSuccess) {
            plan_release(e.dptr, bytes);
            return fail(DCB_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(ce));
        }
        launch_plan(map_kind, th, p, grid, e.dptr, g_image_stats_fwd(), stream);
        ce = cudaGetLastError();
        if (ce == cudaSuccess) ce = cudaEventRecord(e.ready, stream);
        if (ce != cudaSuccess) {
            cudaDeviceSynchronize();
            plan_release(e.dptr, bytes);
            cudaEventDestroy(e.ready);
            return fail(DCB_ERR_CUDA, "plan kernel launch failed: %s", cudaGetErrorString(ce));
        }
        if (trace)
            fprintf(stderr, "[dcb] plan build enqueued after %.0f us\n",
                    std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tr0).count());
        g_plan_builds.fetch_add(1, std::memory_order_relaxed);
        g_plan_bytes += bytes;
        it = g_plans.emplace(k, e).first;
    } else if (!it->second.ready_done) {
        // built on some stream, maybe not finished: order this stream behind it
        if (cudaEventQuery(it->second.ready) == cudaSuccess)
            it->second.ready_done = true;
        else
            CUDA_TRY(cudaStreamWaitEvent(stream, it->second.ready, 0));
    }
    // complete before this launch is enqueued: the kernel may read it ahead of its grid dependency
    p.plan_ready = it->second.ready_done ? 1 : 0;
    it->second.last_use = ++g_plan_clock;
    bind(it->second.dptr);
    return DCB_OK;
}

static int plan_and_launch_image(const ImageKernelSel &sel, ImageParams &p, double gmain,
                                 double gcross, int path_req, size_t src_pitch_bytes,
                                 cudaStream_t stream, int map_kind) {
    DevProps props;
    int rc = device_props(&props);
    if (rc != DCB_OK) return rc;
    const int TH = sel.th;
    const long long tiles_x = (p.W + kTileW - 1) / kTileW;
    const long long tiles_y = (p.nrows + TH - 1) / TH;
    const long long ntiles = tiles_x * tiles_y;
    if (ntiles > INT_MAX) return fail(DCB_ERR_UNSUPPORTED, "too many tiles (%lld)", ntiles);
    p.tiles_y = (int)tiles_y;
    p.ntiles = (int)ntiles;
    const int src_rows = p.ylast - p.yorg + 1;

    const bool layout_ok = ((uintptr_t)p.src % 16 == 0) && (src_pitch_bytes % 16 == 0) &&
                           tma_encoder() != nullptr;
    if (path_req == DCB_PATH_TMA && !layout_ok)
        return fail(DCB_ERR_ARG,
                    "DCB_PATH_TMA needs a 16-byte aligned source with a pitch that is a multiple "
                    "of 16 bytes (and a driver exporting cuTensorMapEncodeTiled)");
    bool staged = layout_ok && path_req != DCB_PATH_DIRECT;
    int bw = 0, bh = 0;
    bool fallback_box = false;
    if (staged) {
        if (!(gmain < 64.0) || !(gcross < 64.0)) gmain = gcross = 64.0;
        const double tw = std::min(kTileW, p.W) - 1, th = std::min(TH, p.nrows) - 1;
        // footprint bound + 2 (floor and the +1 tap) + 2 (probe slack) + 3 (16-byte alignment)
        long long need_w = (long long)std::ceil(gmain * tw + gcross * th) + 8;
        long long need_h = (long long)std::ceil(gmain * th + gcross * tw) + 5;
        need_w = std::min<long long>(need_w, (long long)p.W + 3);
        need_h = std::min<long long>(need_h, src_rows);
        bw = (int)((need_w + 3) / 4 * 4);
        bh = (int)need_h;
        // 5 stages (fp64 blends: one raw + two float64 tiles) or 4 raw stages + tail, x 2 CTAs
        // (each + 1 KB reserved) within the SM's 228 KB
        const int nst = kRawStages;
        const int max_stage =
            TH >= 32 ? (int)((116736 - 1024 - image_tail_bytes(TH)) / nst / 128 * 128)
                     : 14 * 1024;
        if (bw > 256 || bh > 256 || (long long)bw * bh * 4 > max_stage) {
            // strong magnification somewhere: stage a modest box, tiles whose
            // probes do not fit are gathered straight from global memory
            bw = std::min(256, std::min(kTileW + 16, (p.W + 3) / 4 * 4 + 4));
            bh = std::min(std::min(TH + 8, src_rows), max_stage / (bw * 4));
            fallback_box = true;
        }
        bw = std::max(bw, 4);
        bh = std::max(bh, 1);
        if (kImgBoxW > 0) {   // fixed box width: the kernel samples with a compile-time pitch
            // (a box one or two rows short of the bound only sends single rows of the tallest
            // tiles down the exact path; much shorter and whole tiles miss it)
            if (bw > kImgBoxW || bh > max_stage / (kImgBoxW * 4) + 2) fallback_box = true;
            bw = kImgBoxW;
            bh = std::min(bh, max_stage / (bw * 4));
        }
    }
    // DCB_IMG_FAST=0 (diagnostics, A/B runs): every row takes the exact coordinate path
    {
        const char *env = getenv("DCB_IMG_FAST");
        // (integer images round half away from zero in fp64, which the patch path's blend
        // certificate does not cover yet: they keep the exact path)
        p.fast = ((env != nullptr && env[0] == '0') || p.rint) ? 0 : 1;
    }
    p.stats = g_image_stats;
    // uneven tiles (clipped regions, tiles that will not fit the box): deal them round-robin
    p.deal = (fallback_box || map_clips_at_border(p, map_kind)) ? 1 : 0;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (staged) {
        const cuuint64_t gdim[3] = {(cuuint64_t)p.W, (cuuint64_t)src_rows, 1};
        const cuuint64_t gstr[2] = {(cuuint64_t)src_pitch_bytes,
                                    (cuuint64_t)src_pitch_bytes * (cuuint64_t)src_rows};
        const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        CUresult cr = tma_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)p.src, gdim,
                                    gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            if (path_req == DCB_PATH_TMA)
                return fail(DCB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
            staged = false;
        }
    }
    if (!staged) bw = bh = 0;
    p.bw = bw;
    p.bh = bh;
    p.box_bytes = (unsigned)(bw * bh * 4);
    p.stage_bytes = (p.box_bytes + 127u) / 128u * 128u;
    const size_t smem = (size_t)kRawStages * p.stage_bytes + image_tail_bytes(TH);
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute((const void *)sel.kern,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)sel.kern,
                                                           kImgThreads, smem));
    if (occ < 1) return fail(DCB_ERR_CUDA, "kernel does not fit on an SM (smem %zu)", smem);
    // Tile scheduling (see remap_image_kernel): a static range per CTA for most of the tiles, a pool
    // of 2.5 tiles per CTA claimed dynamically at the end, the last tile per CTA of it in halves.
    // Uneven tiles (clipped regions, tiles that will not fit the box): everything is pooled.
    const long long slots = (long long)occ * props.sm_count;
    const int rpw = TH / kWarps;
    // DCB_IMG_POOL="pool,halves,depth" (tenths of a tile per CTA, tenths, units) for A/B runs
    static const struct PoolCfg { int pool10 = 25, halves10 = 10, depth = 3; } pc = [] {
        PoolCfg c;
        if (const char *e = getenv("DCB_IMG_POOL")) sscanf(e, "%d,%d,%d", &c.pool10, &c.halves10, &c.depth);
        c.depth = std::max(1, std::min(kRawStages, c.depth));
        return c;
    }();
    p.pool_depth = pc.depth;
    long long pool = p.deal ? ntiles : std::min<long long>(ntiles, slots * pc.pool10 / 10);
    const long long halves = std::min<long long>(pool, slots * pc.halves10 / 10);   // tiles claimed in two halves
    p.nstatic = (int)(ntiles - pool);
    p.npool_full = (int)(pool - halves);
    p.npool_units = (int)(p.npool_full + halves * (rpw >= 2 ? 2 : 1));
    const int grid = (int)std::min<long long>((long long)p.nstatic + p.npool_units, slots);
    rc = image_sched_slot(&p.sched);
    if (rc != DCB_OK) return rc;
    void *transient = nullptr;
    rc = get_image_plan(map_kind, TH, p, grid, stream, &transient);
    if (rc != DCB_OK) return rc;
    {
        // programmatic stream serialization: see the griddepcontrol pair in remap_image_kernel
        // (DCB_IMG_PDL=0 launches the ordinary way, for A/B runs)
        static const bool pdl = [] {
            const char *e = getenv("DCB_IMG_PDL");
            return !(e != nullptr && e[0] == '0');
        }();
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kImgThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 1 : 0;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, sel.kern, (const ImageParams)p, (const CUtensorMap)tmap));
    }
    CUDA_TRY(cudaGetLastError());
    if (transient != nullptr) CUDA_TRY(cudaFreeAsync(transient, stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    g_last_plan = {staged ? DCB_PATH_TMA : DCB_PATH_DIRECT, bw, bh, grid, (int)smem};
    return DCB_OK;
}

template <int MAP, int NT, int TH, int MINB>
static ImageKernelSel pick_image_kernel_nt(int order, int blend) {
    if (order == 0) return {remap_image_kernel<MAP, 0, DCB_BLEND_EXACT, NT, TH, MINB>, false, TH};
    switch (blend) {
        case DCB_BLEND_LERP64:
            return {remap_image_kernel<MAP, 1, DCB_BLEND_LERP64, NT, TH, MINB>, false, TH};
        case DCB_BLEND_LERP32:
            return {remap_image_kernel<MAP, 1, DCB_BLEND_LERP32, NT, TH, MINB>, false, TH};
        default:
            return {remap_image_kernel<MAP, 1, DCB_BLEND_EXACT, NT, TH, MINB>, false, TH};
    }
}

constexpr int kImgTileH = 32;     // 128 x 32 output tiles
#ifndef DCB_IMG_MINB
#define DCB_IMG_MINB 2
#endif
constexpr int kImgMinBlocks = DCB_IMG_MINB;  // 128 registers: the per-row loop keeps 4 fp64 chains in flight

// nterms: number of polynomial coefficients (radial map); 1..10 have kernels
// with the Horner chain unrolled at compile time, the rest use the generic one.
template <int MAP>
static ImageKernelSel pick_image_kernel(int order, int blend, int nterms, int flags) {
#ifdef DCB_AB
    // A/B builds only: flags selects an experimental variant of the 5-term radial kernel
    if (MAP == MAP_RADIAL && nterms == 5 && order == 1 && (flags & 0xf) != 0) {
        const int v = flags & 0xf;
        if (v == 1) return pick_image_kernel_nt<MAP, 5, 32, 2>(order, blend);
        if (v == 2) return pick_image_kernel_nt<MAP, 5, 16, 4>(order, blend);
        if (v == 3) return pick_image_kernel_nt<MAP, 5, 32, 3>(order, blend);
    }
#endif
    (void)flags;
    if (MAP == MAP_RADIAL) {
        switch (nterms) {
#define DCB_NT(N) \
    case N:       \
        return pick_image_kernel_nt<MAP, N, kImgTileH, kImgMinBlocks>(order, blend);
#ifdef DCB_NT_ONLY   // quick A/B builds (tools/build_ab.sh): one compile-time Horner length
            DCB_NT(DCB_NT_ONLY)
#else
            DCB_NT(1) DCB_NT(2) DCB_NT(3) DCB_NT(4) DCB_NT(5) DCB_NT(6) DCB_NT(7) DCB_NT(8)
            DCB_NT(9) DCB_NT(10)
#endif
#undef DCB_NT
            default:
                break;
        }
    }
    return pick_image_kernel_nt<MAP, 0, kImgTileH, kImgMinBlocks>(order, blend);
}

static int check_options(const dcb_options *opt, dcb_options *o) {
    if (opt == nullptr) {
        *o = {1, DCB_BLEND_EXACT, DCB_PATH_AUTO, 0};
        return DCB_OK;
    }
    *o = *opt;
    REQUIRE(o->order == 0 || o->order == 1,
            "order %d not supported by the CUDA path (0 and 1 are)", o->order);
    REQUIRE(o->blend >= 0 && o->blend <= 2, "unknown blend %d", o->blend);
    REQUIRE(o->path >= 0 && o->path <= 2, "unknown path %d", o->path);
    REQUIRE((o->flags & ~(DCB_FLAG_ROUND_INT | 0xff)) == 0, "unknown flags 0x%x", o->flags);
    return DCB_OK;
}

// the two tile shapes of the Z-stack kernel (remap_stack.cuh): [0] 128 x 16, [1] 64 x 32
struct StackKernelPair {
    StackKernel k[2];
};
template <bool ROUND32, bool RINT>
static StackKernelPair pick_stack_kernel_r(int order, int blend) {
    // order 0 copies pixels: nothing to round
    if (order == 0)
        return {{remap_stack_kernel<0, DCB_BLEND_EXACT, ROUND32, false, 4>,
                 remap_stack_kernel<0, DCB_BLEND_EXACT, ROUND32, false, 2>}};
    switch (blend) {
        case DCB_BLEND_LERP64:
            return {{remap_stack_kernel<1, DCB_BLEND_LERP64, ROUND32, RINT, 4>,
                     remap_stack_kernel<1, DCB_BLEND_LERP64, ROUND32, RINT, 2>}};
        case DCB_BLEND_LERP32:
            return {{remap_stack_kernel<1, DCB_BLEND_LERP32, ROUND32, RINT, 4>,
                     remap_stack_kernel<1, DCB_BLEND_LERP32, ROUND32, RINT, 2>}};
        default:
            return {{remap_stack_kernel<1, DCB_BLEND_EXACT, ROUND32, RINT, 4>,
                     remap_stack_kernel<1, DCB_BLEND_EXACT, ROUND32, RINT, 2>}};
    }
}
template <bool ROUND32>
static StackKernelPair pick_stack_kernel(int order, int blend, int rint) {
    return rint ? pick_stack_kernel_r<ROUND32, true>(order, blend)
                : pick_stack_kernel_r<ROUND32, false>(order, blend);
}

static int check_image_args(const void *src, const void *dst, int H, int W, size_t src_pitch,
                            size_t dst_pitch) {
    REQUIRE(src != nullptr && dst != nullptr, "null image pointer");
    REQUIRE(src != dst, "dst must not alias src");
    REQUIRE(H >= 1 && W >= 1, "image must be at least 1x1 (got %dx%d)", H, W);
    REQUIRE(H < (1 << 23) && W < (1 << 23), "image dimension exceeds 2^23");
    REQUIRE(src_pitch >= (size_t)W * 4 && src_pitch % 4 == 0, "bad source pitch %zu", src_pitch);
    REQUIRE(dst_pitch >= (size_t)W * 4 && dst_pitch % 4 == 0, "bad destination pitch %zu",
            dst_pitch);
    // (the kernels form row offsets with 32 x 32 -> 64 bit multiplies)
    REQUIRE((src_pitch >> 2) < (1ull << 31) && (dst_pitch >> 2) < (1ull << 31),
            "row pitch exceeds 2^31 elements");
    return DCB_OK;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int dcb_version(void) { return DCB_VERSION_MAJOR * 1000 + DCB_VERSION_MINOR; }
const char *dcb_last_error(void) { return g_err; }

int dcb_device_count(int *count) {
    REQUIRE(count != nullptr, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(DCB_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return DCB_OK;
}

int dcb_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(DCB_ERR_NO_DEVICE, "no CUDA device available (%s)", cudaGetErrorString(e));
    REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(0));
    int major = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10)
        return fail(DCB_ERR_UNSUPPORTED,
                    "libdiscorpy_b200 holds sm_100a code only; device %d is sm_%d*", device, major);
    tma_encoder();
    return DCB_OK;
}

int dcb_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem,
                    size_t *free_mem, char *name, int name_len) {
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem || free_mem) {
        int cur = 0;
        CUDA_TRY(cudaGetDevice(&cur));
        CUDA_TRY(cudaSetDevice(device));
        size_t f = 0, t = 0;
        CUDA_TRY(cudaMemGetInfo(&f, &t));
        CUDA_TRY(cudaSetDevice(cur));
        if (total_mem) *total_mem = t;
        if (free_mem) *free_mem = f;
    }
    if (name && name_len > 0) {
        strncpy(name, prop.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    return DCB_OK;
}

// ---- memory / streams / events ---------------------------------------------
int dcb_malloc(void **dptr, size_t nbytes) {
    REQUIRE(dptr != nullptr, "dptr is NULL");
    CUDA_TRY(cudaMalloc(dptr, nbytes ? nbytes : 1));
    return DCB_OK;
}
int dcb_free(void *dptr) {
    CUDA_TRY(cudaFree(dptr));
    return DCB_OK;
}
int dcb_memset(void *dptr, int value, size_t nbytes, void *stream) {
    CUDA_TRY(cudaMemsetAsync(dptr, value, nbytes, (cudaStream_t)stream));
    return DCB_OK;
}
int dcb_host_alloc(void **hptr, size_t nbytes) {
    REQUIRE(hptr != nullptr, "hptr is NULL");
    CUDA_TRY(cudaHostAlloc(hptr, nbytes ? nbytes : 1, cudaHostAllocPortable));
    return DCB_OK;
}
int dcb_host_free(void *hptr) {
    CUDA_TRY(cudaFreeHost(hptr));
    return DCB_OK;
}
int dcb_host_register(void *hptr, size_t nbytes) {
    CUDA_TRY(cudaHostRegister(hptr, nbytes, cudaHostRegisterPortable));
    return DCB_OK;
}
int dcb_host_unregister(void *hptr) {
    CUDA_TRY(cudaHostUnregister(hptr));
    return DCB_OK;
}
int dcb_is_pinned(const void *hptr, int *pinned) {
    REQUIRE(pinned != nullptr, "pinned is NULL");
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, hptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *pinned = 0;
        return DCB_OK;
    }
    *pinned = (attr.type == cudaMemoryTypeHost) ? 1 : 0;
    return DCB_OK;
}
int dcb_h2d(void *dst, const void *src, size_t n, void *stream) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return DCB_OK;
}
int dcb_d2h(void *dst, const void *src, size_t n, void *stream) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return DCB_OK;
}
int dcb_d2d(void *dst, const void *src, size_t n, void *stream) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DCB_OK;
}
int dcb_h2d_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t wbytes, size_t rows,
               void *stream) {
    CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, wbytes, rows, cudaMemcpyHostToDevice,
                               (cudaStream_t)stream));
    return DCB_OK;
}
int dcb_d2h_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t wbytes, size_t rows,
               void *stream) {
    CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, wbytes, rows, cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
    return DCB_OK;
}
int dcb_stream_create(void **stream) {
    REQUIRE(stream != nullptr, "stream is NULL");
    cudaStream_t s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return DCB_OK;
}
int dcb_stream_destroy(void *stream) {
    CUDA_TRY(cudaStreamDestroy((cudaStream_t)stream));
    return DCB_OK;
}
int dcb_stream_sync(void *stream) {
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return DCB_OK;
}
int dcb_device_sync(void) {
    CUDA_TRY(cudaDeviceSynchronize());
    return DCB_OK;
}
int dcb_event_create(void **event) {
    REQUIRE(event != nullptr, "event is NULL");
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    *event = (void *)e;
    return DCB_OK;
}
int dcb_event_destroy(void *event) {
    CUDA_TRY(cudaEventDestroy((cudaEvent_t)event));
    return DCB_OK;
}
int dcb_event_record(void *event, void *stream) {
    CUDA_TRY(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return DCB_OK;
}
int dcb_event_sync(void *event) {
    CUDA_TRY(cudaEventSynchronize((cudaEvent_t)event));
    return DCB_OK;
}
int dcb_stream_wait_event(void *stream, void *event) {
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
    return DCB_OK;
}
int dcb_event_elapsed_ms(void *start, void *stop, float *ms) {
    REQUIRE(ms != nullptr, "ms is NULL");
    CUDA_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return DCB_OK;
}

// ---- peer windows: a dcb_malloc buffer of one process mapped into another (8e) ----
static_assert(sizeof(cudaIpcMemHandle_t) == DCB_IPC_HANDLE_BYTES, "IPC handle size");
int dcb_ipc_export(const void *dptr, void *handle_host) {
    REQUIRE(dptr != nullptr && handle_host != nullptr, "dptr / handle is NULL");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void *>(dptr)));
    memcpy(handle_host, &h, sizeof(h));
    return DCB_OK;
}
int dcb_ipc_open(const void *handle_host, void **peer_dptr) {
    REQUIRE(handle_host != nullptr && peer_dptr != nullptr, "handle / peer_dptr is NULL");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    CUDA_TRY(cudaIpcOpenMemHandle(peer_dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DCB_OK;
}
int dcb_ipc_close(void *peer_dptr) {
    CUDA_TRY(cudaIpcCloseMemHandle(peer_dptr));
    return DCB_OK;
}

// ---- hot path -----------------------------------------------------------------
static int radial_to_dev(const dcb_radial *m, RadialDev *out) {
    REQUIRE(m != nullptr, "radial model is NULL");
    REQUIRE(m->n >= 1 && m->n <= DCB_MAX_TERMS, "number of polynomial terms %d not in 1..%d", m->n,
            DCB_MAX_TERMS);
    memset(out, 0, sizeof(*out));
    out->xc = m->xc;
    out->yc = m->yc;
    out->n = m->n;
    for (int i = 0; i < m->n; ++i) out->a[i] = m->a[i];
    return DCB_OK;
}

int dcb_unwarp_stack_backward_f32(const float *src, float *dst, int D, int H, int W, int src_row0,
                                  int src_rows, size_t src_pitch, size_t src_slice_stride,
                                  size_t dst_pitch, size_t dst_slice_stride, int row0, int nrows,
                                  int coord_round, const dcb_radial *model,
                                  const dcb_options *opt, void *stream) {
    dcb_options o;
    int rc = check_options(opt, &o);
    if (rc) return rc;
    rc = check_image_args(src, dst, H, W, src_pitch, dst_pitch);
    if (rc) return rc;
    REQUIRE(D >= 1, "depth must be >= 1");
    REQUIRE(row0 >= 0 && nrows >= 1 && row0 + nrows <= H, "rows %d..%d outside 0..%d", row0,
            row0 + nrows - 1, H - 1);
    REQUIRE(src_row0 >= 0 && src_rows >= 1 && src_row0 + src_rows <= H,
            "source window rows %d..%d outside 0..%d", src_row0, src_row0 + src_rows - 1, H - 1);
    REQUIRE(D == 1 || (src_slice_stride >= src_pitch * (size_t)src_rows && src_slice_stride % 4 == 0),
            "bad source slice stride %zu", src_slice_stride);
    REQUIRE(D == 1 || (dst_slice_stride >= dst_pitch * (size_t)nrows && dst_slice_stride % 4 == 0),
            "bad destination slice stride %zu", dst_slice_stride);
    REQUIRE(coord_round == 0 || coord_round == 1, "coord_round must be 0 or 1");
    REQUIRE((unsigned long long)(src_pitch / 4) * (unsigned long long)src_rows < (1ull << 31),
            "source window of %d rows x %zu bytes exceeds 2^31 elements per slice", src_rows,
            src_pitch);
    if (!coord_round) {
        REQUIRE(o.order == 1, "float64-coordinate (slice) sampling is order 1 only");
        if (o.blend == DCB_BLEND_LERP32) o.blend = DCB_BLEND_LERP64;
    }
    RemapParams p;
    memset(&p, 0, sizeof(p));
    rc = radial_to_dev(model, &p.rad);
    if (rc) return rc;
    p.src = src;
    p.dst = dst;
    p.src_pitch = (long long)(src_pitch / 4);
    p.src_slice = (long long)(src_slice_stride / 4);
    p.dst_pitch = (long long)(dst_pitch / 4);
    p.dst_slice = (long long)(dst_slice_stride / 4);
    p.H = H;
    p.W = W;
    p.D = D;
    p.row0 = row0;
    p.nrows = nrows;
    p.yorg = src_row0;
    p.ylast = src_row0 + src_rows - 1;
    p.rint = (o.flags & DCB_FLAG_ROUND_INT) ? 1 : 0;
    double gm, gc;
    radial_slopes(*model, W, row0, nrows, &gm, &gc);
    if (coord_round && D == 1) {
        ImageParams q;
        memset(&q, 0, sizeof(q));
        q.rad = p.rad;
        q.src = src;
        q.dst = dst;
        q.src_pitch = p.src_pitch;
        q.dst_pitch = p.dst_pitch;
        q.H = H;
        q.W = W;
        q.row0 = row0;
        q.nrows = nrows;
        q.yorg = p.yorg;
        q.ylast = p.ylast;
        q.dbg = (o.flags >> 4) & 0xf;
        q.rint = p.rint;
        const ImageKernelSel k = pick_image_kernel<MAP_RADIAL>(o.order, o.blend, model->n, o.flags);
        return plan_and_launch_image(k, q, gm, gc, o.path, src_pitch, (cudaStream_t)stream, MAP_RADIAL);
    }
    const bool widen = false;  // no float64 copy of the staged box (see StackWeights)
    if (coord_round)
        return plan_and_launch_stack(pick_stack_kernel<true>(o.order, o.blend, p.rint).k, widen, p, gm, gc,
                                     o.path, src_pitch, src_slice_stride, (cudaStream_t)stream);
    return plan_and_launch_stack(pick_stack_kernel<false>(1, o.blend, p.rint).k, widen, p, gm, gc, o.path,
                                 src_pitch, src_slice_stride, (cudaStream_t)stream);
}

int dcb_unwarp_image_backward_f32(const float *src, float *dst, int H, int W, size_t src_pitch,
                                  size_t dst_pitch, const dcb_radial *model,
                                  const dcb_options *opt, void *stream) {
    return dcb_unwarp_stack_backward_f32(src, dst, 1, H, W, 0, H, src_pitch, src_pitch * (size_t)H,
                                         dst_pitch, dst_pitch * (size_t)H, 0, H, 1, model, opt,
                                         stream);
}

}  // extern "C"

// output rows [row0, row0 + nrows) of the projective remap; dst points at row row0
int dcb::persp_rows_f32(const float *src, float *dst, int H, int W, size_t src_pitch, size_t dst_pitch,
                        int row0, int nrows, const dcb_persp *model, const dcb_options *opt, void *stream) {
    dcb_options o;
    int rc = check_options(opt, &o);
    if (rc) return rc;
    rc = check_image_args(src, dst, H, W, src_pitch, dst_pitch);
    if (rc) return rc;
    REQUIRE(model != nullptr, "perspective model is NULL");
    RemapParams p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < 8; ++i) p.per.c[i] = model->c[i];
    p.src = src;
    p.dst = dst;
    p.src_pitch = (long long)(src_pitch / 4);
    p.src_slice = p.src_pitch * H;
    p.dst_pitch = (long long)(dst_pitch / 4);
    p.dst_slice = p.dst_pitch * H;
    p.H = H;
    p.W = W;
    p.D = 1;
    p.row0 = row0;
    p.nrows = nrows;
    p.yorg = 0;
    p.ylast = H - 1;
    double gm, gc;
    persp_slopes(*model, H, W, &gm, &gc);
    ImageParams q;
    memset(&q, 0, sizeof(q));
    q.per = p.per;
    q.src = src;
    q.dst = dst;
    q.src_pitch = p.src_pitch;
    q.dst_pitch = p.dst_pitch;
    q.H = H;
    q.W = W;
    q.row0 = row0;
    q.nrows = nrows;
    q.yorg = 0;
    q.ylast = H - 1;
    q.rint = (o.flags & DCB_FLAG_ROUND_INT) ? 1 : 0;
    const ImageKernelSel k = pick_image_kernel<MAP_PERSP>(o.order, o.blend, 0, o.flags);
    return plan_and_launch_image(k, q, gm, gc, o.path, src_pitch, (cudaStream_t)stream, MAP_PERSP);
}

extern "C" {

int dcb_correct_perspective_image_f32(const float *src, float *dst, int H, int W, size_t src_pitch,
                                      size_t dst_pitch, const dcb_persp *model,
                                      const dcb_options *opt, void *stream) {
    return persp_rows_f32(src, dst, H, W, src_pitch, dst_pitch, 0, H, model, opt, stream);
}

int dcb_unwarp_image_backward_perspective_f32(const float *src, float *dst, float *scratch, int H,
                                              int W, size_t src_pitch, size_t dst_pitch,
                                              size_t scratch_pitch, const dcb_radial *radial,
                                              const dcb_persp *persp, const dcb_options *opt,
                                              void *stream) {
    REQUIRE(scratch != nullptr && scratch != src && scratch != dst,
            "scratch must be a third, distinct buffer");
    int rc = dcb_unwarp_image_backward_f32(src, scratch, H, W, src_pitch, scratch_pitch, radial, opt,
                                           stream);
    if (rc) return rc;
    return dcb_correct_perspective_image_f32(scratch, dst, H, W, scratch_pitch, dst_pitch, persp,
                                             opt, stream);
}

}  // extern "C"

template <class CT>
static int launch_map_coords(const float *src, float *dst, int H, int W, long long pitch,
                             const void *yd, const void *xd, size_t n, unsigned *oob,
                             const dcb_options &o, cudaStream_t stream) {
    DevProps props;
    int rc = device_props(&props);
    if (rc != DCB_OK) return rc;
    const size_t want = (n + 255) / 256;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)props.sm_count * 16));
    const CT *y = (const CT *)yd, *x = (const CT *)xd;
    const int rint = (o.flags & DCB_FLAG_ROUND_INT) ? 1 : 0;
    if (o.order == 0)
        map_coords_kernel<0, DCB_BLEND_EXACT, CT><<<grid, 256, 0, stream>>>(src, dst, H, W, pitch, y,
                                                                           x, n, oob, 0);
    else if (o.blend == DCB_BLEND_LERP64)
        map_coords_kernel<1, DCB_BLEND_LERP64, CT><<<grid, 256, 0, stream>>>(src, dst, H, W, pitch,
                                                                            y, x, n, oob, rint);
    else if (o.blend == DCB_BLEND_LERP32 && sizeof(CT) == 4)
        map_coords_kernel<1, DCB_BLEND_LERP32, CT><<<grid, 256, 0, stream>>>(src, dst, H, W, pitch,
                                                                            y, x, n, oob, rint);
    else
        map_coords_kernel<1, DCB_BLEND_EXACT, CT><<<grid, 256, 0, stream>>>(src, dst, H, W, pitch, y,
                                                                           x, n, oob, rint);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    g_last_plan = {DCB_PATH_DIRECT, 0, 0, grid, 0};
    return DCB_OK;
}

// ---- spline orders 2..5 / float64 images ------------------------------------------
namespace {

// correctly rounded doubles of sqrt(8)-3, sqrt(3)-2, ... (the literals SciPy holds;
// evaluating the radicals in double loses bits to cancellation)
const double kSplinePoles[6][2] = {
    {0, 0}, {0, 0},
    {-0x1.5f619980c4337p-3, 0},
    {-0x1.126145e9ecd56p-2, 0},
    {-0x1.72036f2fc0817p-2, -0x1.c1c13efa52247p-7},
    {-0x1.b8e8be69086f0p-2, -0x1.610b778d2f346p-5}};

struct SplinePlan {
    int npad, pad_const, filt_kind, tap_kind;
};

bool spline_plan(int order, int mode, SplinePlan *pl) {
    int filt, npad = 0, pad_const = 0, tap;
    switch (mode) {
        case DCB_MODE_REFLECT: case DCB_MODE_GRID_MIRROR: filt = SPL_REFLECT; tap = SPL_REFLECT; break;
        case DCB_MODE_NEAREST: filt = SPL_REFLECT; tap = SPL_MIRROR; npad = 12; break;
        case DCB_MODE_GRID_WRAP: filt = SPL_WRAP; tap = SPL_WRAP; break;
        case DCB_MODE_GRID_CONSTANT: filt = SPL_MIRROR; tap = SPL_MIRROR; npad = 12; pad_const = 1; break;
        case DCB_MODE_MIRROR: case DCB_MODE_CONSTANT: case DCB_MODE_WRAP: filt = SPL_MIRROR; tap = SPL_MIRROR; break;
        default: return false;
    }
    if (order <= 1) npad = 0, pad_const = 0;   // no prefilter, no pre-padding
    pl->npad = npad;
    pl->pad_const = pad_const;
    pl->filt_kind = filt;
    pl->tap_kind = tap;
    return true;
}

SplinePoles spline_poles(int order, int n, int kind, double *gain) {
    SplinePoles pl;
    memset(&pl, 0, sizeof(pl));
    pl.npoles = order / 2;
    double g = 1.0;
    for (int k = 0; k < pl.npoles; ++k) {
        const double z = kSplinePoles[order][k];
        pl.z[k] = z;
        pl.zp[k] = std::pow(z, (double)(kind == SPL_MIRROR ? n - 1 : n));
        volatile double a = 1.0 - 1.0 / z, b = 1.0 - z;   // SciPy's operation order, not folded
        volatile double ab = a * b;
        g = g * ab;
    }
    *gain = g;
    return pl;
}

template <class OUT>
void launch_spline(int order, const SplineParams &p, dim3 grid, cudaStream_t st) {
    switch (order) {
        case 0: spline_remap_kernel<0, OUT><<<grid, 256, 0, st>>>(p); break;
        case 1: spline_remap_kernel<1, OUT><<<grid, 256, 0, st>>>(p); break;
        case 2: spline_remap_kernel<2, OUT><<<grid, 256, 0, st>>>(p); break;
        case 3: spline_remap_kernel<3, OUT><<<grid, 256, 0, st>>>(p); break;
        case 4: spline_remap_kernel<4, OUT><<<grid, 256, 0, st>>>(p); break;
        default: spline_remap_kernel<5, OUT><<<grid, 256, 0, st>>>(p); break;
    }
}

}  // namespace

extern "C" {

int dcb_map_coordinates_f32(const float *src, float *dst, int H, int W, size_t src_pitch,
                            const void *yd, const void *xd, int coord_is_f64, size_t n_out,
                            uint32_t *oob_count, const dcb_options *opt, void *stream) {
    dcb_options o;
    int rc = check_options(opt, &o);
    if (rc) return rc;
    REQUIRE(src != nullptr && dst != nullptr && yd != nullptr && xd != nullptr, "null pointer");
    REQUIRE(H >= 1 && W >= 1 && H < (1 << 24) && W < (1 << 24), "bad image size %dx%d", H, W);
    REQUIRE(src_pitch >= (size_t)W * 4 && src_pitch % 4 == 0, "bad source pitch %zu", src_pitch);
    if (n_out == 0) return DCB_OK;
    if (coord_is_f64)
        return launch_map_coords<double>(src, dst, H, W, (long long)(src_pitch / 4), yd, xd, n_out,
                                         oob_count, o, (cudaStream_t)stream);
    return launch_map_coords<float>(src, dst, H, W, (long long)(src_pitch / 4), yd, xd, n_out,
                                    oob_count, o, (cudaStream_t)stream);
}

static size_t dtype_size(int dtype) {
    switch (dtype) {
        case DCB_DTYPE_F32: return 4;
        case DCB_DTYPE_U8:
        case DCB_DTYPE_I8: return 1;
        case DCB_DTYPE_U16:
        case DCB_DTYPE_I16: return 2;
        default: return 0;
    }
}

static int check_hwc_args(const void *a, const void *b, int dtype, int H, int W, int C, size_t pitch,
                          size_t plane) {
    REQUIRE(a != nullptr && b != nullptr, "null pointer");
    REQUIRE(dtype_size(dtype) != 0, "unknown dtype %d", dtype);
    REQUIRE(H >= 1 && W >= 1 && C >= 1 && C <= 64, "bad shape (%d, %d, %d)", H, W, C);
    REQUIRE(pitch >= (size_t)W * 4 && pitch % 4 == 0, "bad plane pitch %zu", pitch);
    REQUIRE(plane >= pitch * (size_t)H && plane % 4 == 0, "bad plane stride %zu", plane);
    return DCB_OK;
}

int dcb_unpack_hwc_to_planes_f32(const void *src_hwc, int dtype, float *dst_planes, int H, int W,
                                 int C, size_t dst_pitch, size_t dst_plane_stride, void *stream) {
    int rc = check_hwc_args(src_hwc, dst_planes, dtype, H, W, C, dst_pitch, dst_plane_stride);
    if (rc) return rc;
    DevProps props;
    rc = device_props(&props);
    if (rc != DCB_OK) return rc;
    const long long npx = (long long)H * W;
    const int grid = (int)std::min<long long>((npx + 255) / 256, (long long)props.sm_count * 16);
    const long long pitch = (long long)(dst_pitch / 4), plane = (long long)(dst_plane_stride / 4);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case DCB_DTYPE_U8: unpack_hwc_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)src_hwc, dst_planes, H, W, C, pitch, plane); break;
        case DCB_DTYPE_I8: unpack_hwc_kernel<int8_t><<<grid, 256, 0, st>>>((const int8_t *)src_hwc, dst_planes, H, W, C, pitch, plane); break;
        case DCB_DTYPE_U16: unpack_hwc_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)src_hwc, dst_planes, H, W, C, pitch, plane); break;
        case DCB_DTYPE_I16: unpack_hwc_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t *)src_hwc, dst_planes, H, W, C, pitch, plane); break;
        default: unpack_hwc_kernel<float><<<grid, 256, 0, st>>>((const float *)src_hwc, dst_planes, H, W, C, pitch, plane); break;
    }
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return DCB_OK;
}

int dcb_pack_planes_f32_to_hwc(const float *src_planes, void *dst_hwc, int dtype, int H, int W, int C,
                               size_t src_pitch, size_t src_plane_stride, void *stream) {
    int rc = check_hwc_args(src_planes, dst_hwc, dtype, H, W, C, src_pitch, src_plane_stride);
    if (rc) return rc;
    DevProps props;
    rc = device_props(&props);
    if (rc != DCB_OK) return rc;
    const long long npx = (long long)H * W;
    const int grid = (int)std::min<long long>((npx + 255) / 256, (long long)props.sm_count * 16);
    const long long pitch = (long long)(src_pitch / 4), plane = (long long)(src_plane_stride / 4);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case DCB_DTYPE_U8: pack_hwc_kernel<uint8_t><<<grid, 256, 0, st>>>(src_planes, (uint8_t *)dst_hwc, H, W, C, pitch, plane); break;
        case DCB_DTYPE_I8: pack_hwc_kernel<int8_t><<<grid, 256, 0, st>>>(src_planes, (int8_t *)dst_hwc, H, W, C, pitch, plane); break;
        case DCB_DTYPE_U16: pack_hwc_kernel<uint16_t><<<grid, 256, 0, st>>>(src_planes, (uint16_t *)dst_hwc, H, W, C, pitch, plane); break;
        case DCB_DTYPE_I16: pack_hwc_kernel<int16_t><<<grid, 256, 0, st>>>(src_planes, (int16_t *)dst_hwc, H, W, C, pitch, plane); break;
        default: pack_hwc_kernel<float><<<grid, 256, 0, st>>>(src_planes, (float *)dst_hwc, H, W, C, pitch, plane); break;
    }
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return DCB_OK;
}

int dcb_fill_synthetic_f32(float *dst, size_t n, uint64_t seed, uint64_t offset, void *stream) {
    REQUIRE(dst != nullptr, "dst is NULL");
    if (n == 0) return DCB_OK;
    DevProps props;
    int rc = device_props(&props);
    if (rc != DCB_OK) return rc;
    const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)props.sm_count * 32);
    fill_synthetic_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dst, n, seed, offset);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return DCB_OK;
}

// ---- diagnostics ---------------------------------------------------------------
int dcb_launch_count(uint64_t *count) {
    REQUIRE(count != nullptr, "count is NULL");
    *count = g_launches.load(std::memory_order_relaxed);
    return DCB_OK;
}
int dcb_launch_count_reset(void) {
    g_launches.store(0, std::memory_order_relaxed);
    return DCB_OK;
}
int dcb_plan_cache_clear(uint64_t *plans_built) {
    std::lock_guard<std::mutex> lk(g_plans_mu);
    // (kernels on any stream of any device may still read the plans)
    if (!g_plans.empty()) {
        int cur = 0, ndev = 0;
        cudaGetDevice(&cur);
        cudaGetDeviceCount(&ndev);
        std::vector<bool> seen((size_t)std::max(ndev, 1), false);
        for (auto &sl : g_plan_slabs)
            if (sl.device >= 0 && sl.device < ndev && !seen[(size_t)sl.device]) {
                seen[(size_t)sl.device] = true;
                cudaSetDevice(sl.device);
                cudaDeviceSynchronize();
            }
        cudaSetDevice(cur);
    }
    for (auto &kv : g_plans) {
        plan_release(kv.second.dptr, kv.second.bytes);
        cudaEventDestroy(kv.second.ready);
    }
    g_plans.clear();
    g_plan_bytes = 0;
    if (plans_built != nullptr) *plans_built = g_plan_builds.load(std::memory_order_relaxed);
    return DCB_OK;
}

static unsigned long long *g_image_stats_buf = nullptr;
constexpr size_t kImageStatsWords = 8 + 8 * (size_t)kTimelineCtas + (size_t)kLogCtas * 10 * kLogEvents;

int dcb_image_stats(int enable, uint64_t *out, int reset) {
    unsigned long long *&buf = g_image_stats_buf;
    if (enable && buf == nullptr) {
        CUDA_TRY(cudaMalloc((void **)&buf, kImageStatsWords * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(buf, 0, kImageStatsWords * sizeof(unsigned long long)));
    }
    if (buf != nullptr && (out != nullptr || reset)) {
        CUDA_TRY(cudaDeviceSynchronize());
        if (out != nullptr)
            CUDA_TRY(cudaMemcpy(out, buf, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (reset) CUDA_TRY(cudaMemset(buf, 0, 8 * sizeof(unsigned long long)));
    } else if (out != nullptr) {
        memset(out, 0, 8 * sizeof(uint64_t));
    }
    g_image_stats = enable ? buf : nullptr;
    return DCB_OK;
}

int dcb_image_timeline(uint64_t *out, int nctas) {
    // nctas < 0: the per-warp event log of the first CTAs instead (kLogCtas x 10 warps x kLogEvents)
    REQUIRE(out != nullptr && nctas >= -1 && nctas <= kTimelineCtas, "bad arguments");
    const size_t words = nctas < 0 ? (size_t)kLogCtas * 10 * kLogEvents : 8 * (size_t)nctas;
    const size_t first = nctas < 0 ? 8 + 8 * (size_t)kTimelineCtas : 8;
    if (g_image_stats_buf == nullptr) {
        memset(out, 0, words * sizeof(uint64_t));
        return DCB_OK;
    }
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out, g_image_stats_buf + first, words * sizeof(uint64_t),
                        cudaMemcpyDeviceToHost));
    return DCB_OK;
}

int dcb_last_plan(int *path, int *box_w, int *box_h, int *grid, int *smem_bytes) {
    if (path) *path = g_last_plan.path;
    if (box_w) *box_w = g_last_plan.bw;
    if (box_h) *box_h = g_last_plan.bh;
    if (grid) *grid = g_last_plan.grid;
    if (smem_bytes) *smem_bytes = g_last_plan.smem;
    return DCB_OK;
}

// ---- spline orders 2..5 / float64 images (helpers above the second extern "C") ----
int dcb_spline_workspace_bytes(int H, int W, int order, int mode, size_t *bytes) {
    REQUIRE(bytes != nullptr, "bytes is NULL");
    REQUIRE(H >= 1 && W >= 1, "empty image");
    REQUIRE(order >= 0 && order <= 5, "spline order not supported");
    SplinePlan pl;
    REQUIRE(spline_plan(order, mode, &pl), "boundary mode not supported");
    const size_t n = (size_t)(H + 2 * pl.npad) * (size_t)(W + 2 * pl.npad) * sizeof(double);
    *bytes = order > 1 ? 2 * n : n;   // coefficients (+ the transposed scratch copy)
    return DCB_OK;
}

int dcb_spline_prefilter(const void *src, int src_is_f64, int H, int W, size_t src_pitch,
                         int order, int mode, void *workspace, size_t workspace_bytes,
                         void *stream) {
    REQUIRE(src != nullptr && workspace != nullptr, "null pointer");
    size_t need = 0;
    int rc = dcb_spline_workspace_bytes(H, W, order, mode, &need);
    if (rc) return rc;
    REQUIRE(workspace_bytes >= need, "workspace of %zu bytes, %zu needed", workspace_bytes, need);
    const size_t esz = src_is_f64 ? 8 : 4;
    REQUIRE(src_pitch % esz == 0 && src_pitch >= (size_t)W * esz, "bad source pitch %zu", src_pitch);
    SplinePlan pl;
    spline_plan(order, mode, &pl);
    cudaStream_t st = (cudaStream_t)stream;
    const int Hc = H + 2 * pl.npad, Wc = W + 2 * pl.npad;
    double *coef = reinterpret_cast<double *>(workspace);
    double *scratch = coef + (size_t)Hc * Wc;
    // axis 0: lines of Hc samples (SciPy filters axis 0 first, _interpolation.py:185-188);
    // a line of length 1 is left alone, gain included
    double gain0 = 1.0, gain1 = 1.0;
    SplinePoles p0 = spline_poles(order > 1 ? order : 0, Hc, pl.filt_kind, &gain0);
    SplinePoles p1 = spline_poles(order > 1 ? order : 0, Wc, pl.filt_kind, &gain1);
    if (order <= 1 || Hc < 2) gain0 = 1.0;
    if (order <= 1 || Wc < 2) gain1 = 1.0;
    const dim3 pgrid((Wc + 31) / 32, std::min((Hc + 7) / 8, 4096));
    if (src_is_f64)
        spline_pad_kernel<double><<<pgrid, 256, 0, st>>>(reinterpret_cast<const double *>(src),
                                                        (long long)(src_pitch / 8), H, W, coef, Wc,
                                                        pl.npad, pl.pad_const, gain0);
    else
        spline_pad_kernel<float><<<pgrid, 256, 0, st>>>(reinterpret_cast<const float *>(src),
                                                       (long long)(src_pitch / 4), H, W, coef, Wc,
                                                       pl.npad, pl.pad_const, gain0);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (order > 1) {
        // lines of >= 256 samples go through the shared-memory staged kernel, short ones through
        // the one-thread-per-line kernel (DCB_SPLINE_STAGED=0/1 forces one of them: diagnostics)
        // (per device and cheap: set on every call, like the remap planners do)
        CUDA_TRY(cudaFuncSetAttribute((const void *)spline_filter_cols_staged_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSplSmemBytes));
        int force = -1;
        if (const char *env = getenv("DCB_SPLINE_STAGED")) force = atoi(env);
        auto filter_cols = [&](double *a, int pitch, int n, int ncols, const SplinePoles &pp) {
            const bool staged = force >= 0 ? force != 0 : n >= 256;
            if (staged)
                spline_filter_cols_staged_kernel<<<(ncols + 31) / 32, 256, kSplSmemBytes, st>>>(
                    a, pitch, n, ncols, pp, pl.filt_kind);
            else
                spline_filter_cols_kernel<<<(ncols + 31) / 32, 32, 0, st>>>(a, pitch, n, ncols, pp,
                                                                            pl.filt_kind);
            g_launches.fetch_add(1, std::memory_order_relaxed);
        };
        if (Hc >= 2) filter_cols(coef, Wc, Hc, Wc, p0);
        if (Wc >= 2) {
            // axis 1: transpose (with the gain), filter the columns of the transposed image, transpose back
            const dim3 tg((Wc + 31) / 32, (Hc + 31) / 32), tb((Hc + 31) / 32, (Wc + 31) / 32);
            spline_transpose_kernel<<<tg, 256, 0, st>>>(coef, Wc, Hc, Wc, scratch, Hc, gain1);
            filter_cols(scratch, Hc, Wc, Hc, p1);
            spline_transpose_kernel<<<tb, 256, 0, st>>>(scratch, Hc, Wc, Hc, coef, Wc, 1.0);
            g_launches.fetch_add(2, std::memory_order_relaxed);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return DCB_OK;
}

int dcb_spline_remap(const void *workspace, int H, int W, int order, int mode, void *dst,
                     int dst_is_f64, size_t dst_pitch, int map_kind, const dcb_radial *radial_host,
                     const dcb_persp *persp_host, const void *yd, const void *xd, int coord_is_f64,
                     size_t n_out, uint32_t *oob_count, int flags, double sat_lo, double sat_hi,
                     void *stream) {
    REQUIRE(workspace != nullptr && dst != nullptr, "null pointer");
    REQUIRE(H >= 1 && W >= 1, "empty image");
    REQUIRE(order >= 0 && order <= 5, "spline order not supported");
    SplinePlan pl;
    REQUIRE(spline_plan(order, mode, &pl), "boundary mode not supported");
    const size_t esz = dst_is_f64 ? 8 : 4;
    SplineParams p;
    memset(&p, 0, sizeof(p));
    p.coef = reinterpret_cast<const double *>(workspace);
    p.Hc = H + 2 * pl.npad;
    p.Wc = W + 2 * pl.npad;
    p.cpitch = p.Wc;
    p.npad = pl.npad;
    p.dst = dst;
    p.H = H;
    p.W = W;
    p.tap_kind = pl.tap_kind;
    p.rint = (flags & DCB_FLAG_ROUND_INT) ? 1 : 0;
    p.lo = sat_lo;
    p.hi = sat_hi;
    p.map = map_kind;
    dim3 grid;
    if (map_kind == DCB_MAP_COORDS) {
        REQUIRE(yd != nullptr && xd != nullptr, "null coordinate pointer");
        p.yd = yd;
        p.xd = xd;
        p.coord_f64 = coord_is_f64 ? 1 : 0;
        p.n = n_out;
        p.oob_count = oob_count;
        if (n_out == 0) return DCB_OK;
        grid = dim3((unsigned)std::min<size_t>((n_out + 255) / 256, 148 * 16));
    } else {
        REQUIRE(dst_pitch % esz == 0 && dst_pitch >= (size_t)W * esz, "bad destination pitch %zu", dst_pitch);
        p.dst_pitch = (long long)(dst_pitch / esz);
        if (map_kind == DCB_MAP_RADIAL) {
            REQUIRE(radial_host != nullptr, "radial model is NULL");
            int rc = radial_to_dev(radial_host, &p.rad);
            if (rc) return rc;
        } else {
            REQUIRE(map_kind == DCB_MAP_PERSP && persp_host != nullptr, "bad map kind / NULL model");
            for (int i = 0; i < 8; ++i) p.per.c[i] = persp_host->c[i];
        }
        grid = dim3((W + 31) / 32, (H + 7) / 8);
    }
    if (dst_is_f64)
        launch_spline<double>(order, p, grid, (cudaStream_t)stream);
    else
        launch_spline<float>(order, p, grid, (cudaStream_t)stream);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return DCB_OK;
}

int dcb_unwarp_image_forward_f32(const float *src, float *dst, int H, int W, size_t src_pitch,
                                 size_t dst_pitch, const dcb_radial *model, uint32_t *workspace,
                                 void *stream) {
    int rc = check_image_args(src, dst, H, W, src_pitch, dst_pitch);
    if (rc) return rc;
    REQUIRE(workspace != nullptr, "workspace is NULL (H*W uint32)");
    REQUIRE((long long)H * W < 0xffffffffll, "image too large for 32-bit pixel indices");
    RadialDev rad;
    rc = radial_to_dev(model, &rad);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)H * W * sizeof(uint32_t), st));
    const dim3 grid((W + 31) / 32, (H + 7) / 8);
    forward_propose_kernel<<<grid, 256, 0, st>>>(H, W, rad, workspace);
    forward_gather_kernel<<<grid, 256, 0, st>>>(src, (long long)(src_pitch / 4), dst,
                                               (long long)(dst_pitch / 4), H, W, workspace);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return DCB_OK;
}

}  // extern "C"

// hostpipe.cu -- the host-buffer side of the C ABI (no kernels here): staging copies of pageable
// memory (CopyPool, dcb_host_copy_2d) and the banded upload / compute / download pipeline behind
// dcb_unwarp_image_backward_host_f32, dcb_correct_perspective_image_host_f32 and
// dcb_unwarp_image_backward_perspective_host_f32.  The band kernels are launched through the
// device-pointer entries of api.cu.
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <pthread.h>

#include "api_common.hpp"

using namespace dcb;

namespace {
// Pageable (ordinary malloc / NumPy) source images: cudaMemcpyAsync from pageable memory is staged
// by the driver on ONE thread (measured: 6.3 ms for a 4096^2 float32 image against 1.8 ms from
// pinned memory).  Instead a few host threads copy each row band into a pinned staging buffer while
// the DMA engine is still moving the previous band.  One process-wide pool, created on first use;
// calls are serialised (the copy is memory-bound, two at once would not go faster).
class CopyPool {
public:
    struct Job {
        const char *src;
        char *dst;
        size_t src_pitch, dst_pitch, width_bytes;
        int rows;
    };
    static CopyPool &get() {
        static CopyPool *pool = nullptr;   // leaked on purpose: worker threads outlive static dtors
        static std::once_flag once;
        std::call_once(once, [] {
            pool = new CopyPool();
            pthread_atfork(nullptr, nullptr, [] { get().forked_ = true; });
        });
        return *pool;
    }
    void run(const Job &j) {
        std::lock_guard<std::mutex> serial(run_mu_);
        if (forked_ || nth_ <= 1 || (size_t)j.rows * j.width_bytes < (1u << 20)) {
            copy_rows(j, 0, j.rows);   // small band (or a forked child without the workers): inline
            return;
        }
        std::unique_lock<std::mutex> lk(mu_);
        job_ = j;
        pending_ = nth_;
        ++gen_;
        cv_go_.notify_all();
        cv_done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    CopyPool() {
        const unsigned hc = std::thread::hardware_concurrency();
        nth_ = (int)std::max(1u, std::min(8u, hc / 2));
        if (const char *env = getenv("DCB_COPY_THREADS")) nth_ = std::max(1, std::min(64, atoi(env)));
        if (nth_ > 1)
            for (int i = 0; i < nth_; ++i) std::thread(&CopyPool::worker, this, i).detach();
    }
    // The staging buffer is written once and read next by the DMA engine, never by this core:
    // non-temporal stores skip the read-for-ownership of every destination line (a third of the
    // copy's memory traffic; glibc's memcpy only does this above ~3/4 of the shared cache per call,
    // and a worker's share of a band is 1 MiB).
    static void stream_copy(char *dst, const char *src, size_t n) {
#if defined(__x86_64__) && defined(__SSE2__)
        if (n >= (64u << 10) && !nt_off()) {
            const size_t head = (size_t)(-(uintptr_t)dst) & 15u;
            memcpy(dst, src, head);
            dst += head, src += head, n -= head;
            const size_t blocks = n / 64;
            for (size_t i = 0; i < blocks; ++i, dst += 64, src += 64) {
                const __m128i a = _mm_loadu_si128((const __m128i *)src);
                const __m128i b = _mm_loadu_si128((const __m128i *)(src + 16));
                const __m128i c = _mm_loadu_si128((const __m128i *)(src + 32));
                const __m128i d = _mm_loadu_si128((const __m128i *)(src + 48));
                _mm_stream_si128((__m128i *)dst, a);
                _mm_stream_si128((__m128i *)(dst + 16), b);
                _mm_stream_si128((__m128i *)(dst + 32), c);
                _mm_stream_si128((__m128i *)(dst + 48), d);
            }
            _mm_sfence();
            n -= blocks * 64;
        }
#endif
        memcpy(dst, src, n);
    }
    static bool nt_off() {   // DCB_COPY_NT=0: plain memcpy (A/B runs)
        static const bool off = [] {
            const char *e = getenv("DCB_COPY_NT");
            return e != nullptr && e[0] == '0';
        }();
        return off;
    }
    static void copy_rows(const Job &j, int r0, int r1) {
        if (j.src_pitch == j.width_bytes && j.dst_pitch == j.width_bytes) {
            stream_copy(j.dst + (size_t)r0 * j.dst_pitch, j.src + (size_t)r0 * j.src_pitch,
                        (size_t)(r1 - r0) * j.width_bytes);
            return;
        }
        for (int r = r0; r < r1; ++r)
            stream_copy(j.dst + (size_t)r * j.dst_pitch, j.src + (size_t)r * j.src_pitch, j.width_bytes);
    }
    void worker(int id) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_go_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            const Job j = job_;
            lk.unlock();
            if (j.rows >= 2 * nth_) {   // many rows: a share of the rows each
                copy_rows(j, (int)((long long)j.rows * id / nth_), (int)((long long)j.rows * (id + 1) / nth_));
            } else {                    // a few long rows (slices of a stack): a share of every row's bytes
                const size_t c0 = j.width_bytes * (size_t)id / (size_t)nth_ / 64 * 64;
                const size_t c1 = id + 1 == nth_ ? j.width_bytes : j.width_bytes * (size_t)(id + 1) / (size_t)nth_ / 64 * 64;
                for (int r = 0; r < j.rows && c1 > c0; ++r)
                    stream_copy(j.dst + (size_t)r * j.dst_pitch + c0, j.src + (size_t)r * j.src_pitch + c0, c1 - c0);
            }
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_one();
        }
    }
    std::mutex run_mu_, mu_;
    std::condition_variable cv_go_, cv_done_;
    Job job_{};
    uint64_t gen_ = 0;
    int pending_ = 0, nth_ = 1;
    bool forked_ = false;
};

// is `p` ordinary pageable host memory (neither cudaHostAlloc'ed nor registered)?
bool is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace


extern "C" {

// ---- host-buffer entry: banded upload / compute / download pipeline -------------
namespace {

constexpr int kMaxBands = 32;

struct HostPipe {
    cudaStream_t up = nullptr, run = nullptr, down = nullptr;
    cudaEvent_t ev_up[kMaxBands], ev_run[kMaxBands], ev_free = nullptr;
    void *dsrc = nullptr, *ddst = nullptr, *dmid = nullptr;   // dmid: image between two stages
    size_t src_cap = 0, dst_cap = 0, mid_cap = 0;
    void *hstage = nullptr;   // pinned staging copy of a pageable source image
    size_t stage_cap = 0;
    int device = -1;
    bool ok = false;
    // Releases everything the pipe owns (a host thread that ends used to leave its
    // buffers, three streams and 65 events behind).  Errors are ignored: at process exit the
    // context may already be gone.
    void release() {
        if (!ok) return;
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
            return;
        }
        for (cudaStream_t q : {up, run, down})
            if (q) cudaStreamSynchronize(q);
        if (dsrc) cudaFree(dsrc);
        if (ddst) cudaFree(ddst);
        if (dmid) cudaFree(dmid);
        if (hstage) cudaFreeHost(hstage);
        for (int i = 0; i < kMaxBands; ++i) {
            if (ev_up[i]) cudaEventDestroy(ev_up[i]);
            if (ev_run[i]) cudaEventDestroy(ev_run[i]);
        }
        if (ev_free) cudaEventDestroy(ev_free);
        for (cudaStream_t q : {up, run, down})
            if (q) cudaStreamDestroy(q);
        cudaGetLastError();
        if (cur >= 0) cudaSetDevice(cur);
        ok = false;
    }
    HostPipe() {
        for (int i = 0; i < kMaxBands; ++i) ev_up[i] = ev_run[i] = nullptr;
    }
    HostPipe(const HostPipe &) = delete;
    HostPipe &operator=(const HostPipe &) = delete;
    ~HostPipe() { release(); }
};
thread_local HostPipe g_pipe;

int pipe_prepare(size_t src_bytes, size_t dst_bytes, size_t mid_bytes) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    HostPipe &hp = g_pipe;
    if (hp.ok && hp.device != dev) {  // the thread moved to another GPU: start over
        hp.release();
        hp.up = hp.run = hp.down = nullptr;
        hp.ev_free = nullptr;
        for (int i = 0; i < kMaxBands; ++i) hp.ev_up[i] = hp.ev_run[i] = nullptr;
        hp.dsrc = hp.ddst = hp.dmid = hp.hstage = nullptr;
        hp.src_cap = hp.dst_cap = hp.mid_cap = hp.stage_cap = 0;
        hp.device = -1;
        CUDA_TRY(cudaSetDevice(dev));
    }
    if (!hp.ok) {
        CUDA_TRY(cudaStreamCreateWithFlags(&hp.up, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp.run, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp.down, cudaStreamNonBlocking));
        for (int i = 0; i < kMaxBands; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&hp.ev_up[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&hp.ev_run[i], cudaEventDisableTiming));
        }
        hp.device = dev;
        hp.ok = true;
    }
    if (hp.src_cap < src_bytes) {
        if (hp.dsrc) CUDA_TRY(cudaFree(hp.dsrc));
        hp.dsrc = nullptr;
        hp.src_cap = 0;
        CUDA_TRY(cudaMalloc(&hp.dsrc, src_bytes));
        hp.src_cap = src_bytes;
    }
    if (hp.dst_cap < dst_bytes) {
        if (hp.ddst) CUDA_TRY(cudaFree(hp.ddst));
        hp.ddst = nullptr;
        hp.dst_cap = 0;
        CUDA_TRY(cudaMalloc(&hp.ddst, dst_bytes));
        hp.dst_cap = dst_bytes;
    }
    if (hp.mid_cap < mid_bytes) {
        if (hp.dmid) CUDA_TRY(cudaFree(hp.dmid));
        hp.dmid = nullptr;
        hp.mid_cap = 0;
        CUDA_TRY(cudaMalloc(&hp.dmid, mid_bytes));
        hp.mid_cap = mid_bytes;
    }
    return DCB_OK;
}

// Conservative range of source rows that output rows [r0, r1) of the radial map
// can sample: interval product of F over the radii the band reaches and yu over
// the band, padded for the sampling step of F and for the bilinear footprint.
void radial_row_range(const dcb_radial &m, int H, int W, int r0, int r1, int *lo, int *hi) {
    const double ya = (double)r0 - m.yc, yb = (double)(r1 - 1) - m.yc;
    const double xa = 0.0 - m.xc, xb = (double)(W - 1) - m.xc;
    auto nearest0 = [](double a, double b) { return (a <= 0.0 && b >= 0.0) ? 0.0 : std::min(std::fabs(a), std::fabs(b)); };
    const double ymin = nearest0(ya, yb), xmin = nearest0(xa, xb);
    const double ymax = std::max(std::fabs(ya), std::fabs(yb)), xmax = std::max(std::fabs(xa), std::fabs(xb));
    const double rmin = std::sqrt(xmin * xmin + ymin * ymin), rmax = std::sqrt(xmax * xmax + ymax * ymax);
    const int S = 2048;
    const double step = (rmax - rmin) / S;
    double fmin = 1e300, fmax = -1e300, dmax = 0.0;
    for (int i = 0; i <= S; ++i) {
        const double r = rmin + step * i;
        double f = 0.0, fp = 0.0;
        for (int k = m.n - 1; k >= 0; --k) {
            fp = fp * r + f;
            f = f * r + m.a[k];
        }
        fmin = std::min(fmin, f);
        fmax = std::max(fmax, f);
        dmax = std::max(dmax, std::fabs(fp));
    }
    const double pad = 2.0 * dmax * step;  // F between two samples (generous for a polynomial)
    fmin -= pad;
    fmax += pad;
    const double c[4] = {fmin * ya, fmin * yb, fmax * ya, fmax * yb};
    double vmin = c[0], vmax = c[0];
    for (double v : c) {
        vmin = std::min(vmin, v);
        vmax = std::max(vmax, v);
    }
    vmin += m.yc;
    vmax += m.yc;
    if (!(vmin == vmin) || !(vmax == vmax) || !(std::fabs(vmin) < 1e15) || !(std::fabs(vmax) < 1e15)) {
        *lo = 0;
        *hi = H - 1;
        return;
    }
    *lo = (int)std::max(0.0, std::min((double)(H - 1), std::floor(vmin) - 1.0));
    *hi = (int)std::max(0.0, std::min((double)(H - 1), std::ceil(vmax) + 2.0));
}

// The same for the projective map: along any line the source row (c3 x + c4 y + c5) / (c6 x + c7 y + 1)
// is a Moebius function of the line parameter, monotone while the denominator keeps its sign, so
// over a band (a rectangle) it takes its extremes at the four corners.  A denominator that
// changes sign or vanishes over the band: the whole image.
void persp_row_range(const dcb_persp &m, int H, int W, int r0, int r1, int *lo, int *hi) {
    const double *c = m.c;
    const double xs[2] = {0.0, (double)(W - 1)}, ys[2] = {(double)r0, (double)(r1 - 1)};
    double vmin = 1e300, vmax = -1e300;
    int sign = 0;
    bool ok = true;
    for (double y : ys)
        for (double x : xs) {
            const double den = c[6] * x + c[7] * y + 1.0;
            const int sg = den > 0.0 ? 1 : (den < 0.0 ? -1 : 0);
            if (sg == 0 || (sign != 0 && sg != sign)) ok = false;
            sign = sg;
            const double v = (c[3] * x + c[4] * y + c[5]) / den;
            if (!(std::fabs(v) < 1e15)) ok = false;
            vmin = std::min(vmin, v);
            vmax = std::max(vmax, v);
        }
    if (!ok) {
        *lo = 0;
        *hi = H - 1;
        return;
    }
    *lo = (int)std::max(0.0, std::min((double)(H - 1), std::floor(vmin) - 1.0));
    *hi = (int)std::max(0.0, std::min((double)(H - 1), std::ceil(vmax) + 2.0));
}

}  // namespace

// Row bands [edge[b], edge[b + 1]) of an H x W float32 image; returns their number (<= kMaxBands).
// The call ends one upload band (the rows below an output band that it samples), one kernel and one
// download band after the last upload, and the first download starts two upload bands into the
// call: images of 32 MiB and more get bands of 2, 2 and 4 MiB at both ends and ~6 MiB bands between
// them (a 4096^2 image: 128, 128, 256, 8 x 384, 256, 128, 128 rows -- 1.78 ms against 1.82 ms for
// eight equal bands; smaller bands throughout cost more in per-copy overhead than they save:
// tools/e2e_edges.py, profiles/r2/e2e_edges_*.txt); smaller images one band per 3.5 MiB, at most 8
// (2160 x 2560: 0.73 ms in 6 bands against 0.87 ms in 2, profiles/r2/e2e_small_r2z9.txt).
// nbands > 0 asks for that many equal bands; DCB_BAND_EDGES="n0,n1,..." (diagnostics, only with
// nbands <= 0) gives the band heights in 64ths of the image.
static int band_edges(int H, int W, int nbands, int *edge) {
    const bool auto_bands = nbands <= 0;
    const size_t img_bytes = (size_t)W * 4 * (size_t)H;
    const int unit = (int)std::max<size_t>(1, ((size_t)2 << 20) / ((size_t)W * 4));   // rows per 2 MiB
    if (auto_bands && img_bytes >= ((size_t)32 << 20) && H >= 16 * unit) {
        const int mid = H - 8 * unit;
        const int nmid = (int)std::max<size_t>(
            1, std::min<size_t>(std::min<size_t>(kMaxBands - 6, (size_t)mid),
                                ((size_t)mid * W * 4 + ((size_t)3 << 20)) / ((size_t)6 << 20)));
        int n = 0;
        edge[0] = 0;
        for (int h : {unit, unit, 2 * unit}) edge[n + 1] = edge[n] + h, ++n;
        for (int k = 1; k <= nmid; ++k) edge[++n] = 4 * unit + (int)((long long)mid * k / nmid);
        for (int h : {2 * unit, unit, unit}) edge[n + 1] = edge[n] + h, ++n;
        nbands = n;
    } else {
        if (auto_bands) nbands = (int)std::max<size_t>(1, std::min<size_t>(8, img_bytes / ((size_t)7 << 19)));
        nbands = std::min(std::min(nbands, kMaxBands), H);
        const int rows_per = (H + nbands - 1) / nbands;
        nbands = (H + rows_per - 1) / rows_per;
        for (int b = 0; b <= nbands; ++b) edge[b] = std::min(H, b * rows_per);
    }
    const char *env = auto_bands ? getenv("DCB_BAND_EDGES") : nullptr;
    if (env != nullptr && env[0] != 0 && H >= 64) {
        int n = 0, acc = 0;
        edge[0] = 0;
        for (const char *q = env; *q != 0 && n < kMaxBands;) {
            char *e = nullptr;
            const long v = strtol(q, &e, 10);
            if (e == q || v <= 0) break;
            acc += (int)v;
            const int row = (int)std::min<long long>(H, (long long)H * acc / 64);
            if (row > edge[n]) edge[++n] = row;
            q = (*e == ',') ? e + 1 : e;
        }
        if (n == 0 || edge[n] < H) {
            if (n < kMaxBands) edge[++n] = H; else edge[n] = H;
        }
        nbands = n;
    }
    return nbands;
}

// The schedule alone (no GPU needed): what host_pipeline would use for an H x W image.
int dcb_host_band_edges(int H, int W, int nbands, int *edges, int *count) {
    REQUIRE(H >= 1 && W >= 1 && edges != nullptr && count != nullptr, "bad arguments");
    *count = band_edges(H, W, nbands, edges);
    return DCB_OK;
}

// One stage of the host-buffer pipeline: a remap of the whole image, launched band by band.
struct PipeStage {
    // conservative range [lo, hi] of the INPUT rows that output rows [r0, r1) of this stage sample
    std::function<void(int r0, int r1, int *lo, int *hi)> rows;
    // output rows [q0, q0 + qn) of this stage: `in` is the stage's whole input image (device,
    // `in_pitch` bytes per row), `out_band` the first of the qn output rows
    std::function<int(const float *in, size_t in_pitch, float *out_band, size_t out_pitch, int q0, int qn,
                      cudaStream_t stream)> launch;
};

// Host image -> nstages remaps (1: radial or projective, 2: radial then projective) -> host image,
// as one pipeline over row bands on three streams: the image is uploaded band by band; a band of
// stage-1 output rows is launched as soon as the last source row it can sample has been enqueued,
// a band of stage-2 rows as soon as the last stage-1 row it can sample has been launched (both on
// the one compute stream, so stream order is the dependency), and every finished band of the last
// stage is downloaded behind its kernel.  Synchronous.
static int host_pipeline(const float *src_host, float *dst_host, int H, int W, size_t src_pitch_host,
                         size_t dst_pitch_host, int nbands, int nstages, const PipeStage *stages) {
    REQUIRE(src_host != nullptr && dst_host != nullptr, "null image pointer");
    REQUIRE(H >= 1 && W >= 1, "image must be at least 1x1 (got %dx%d)", H, W);
    REQUIRE(src_pitch_host >= (size_t)W * 4 && dst_pitch_host >= (size_t)W * 4, "bad host pitch");
    const size_t pitch = ((size_t)W * 4 + 15) / 16 * 16;  // device rows: 16-byte pitch for TMA
    int rc = pipe_prepare(pitch * (size_t)H, pitch * (size_t)H, nstages > 1 ? pitch * (size_t)H : 0);
    if (rc) return rc;
    HostPipe &hp = g_pipe;
    int edge[kMaxBands + 1];
    nbands = band_edges(H, W, nbands, edge);
    auto band_of = [&](int row) -> int {   // the band holding row `row`
        int b = 0;
        while (b < nbands - 1 && edge[b + 1] <= row) ++b;
        return b;
    };
    char *dsrc = (char *)hp.dsrc, *ddst = (char *)hp.ddst, *dmid = (char *)hp.dmid;
    // pageable source: bands go through a pinned staging buffer filled by the copy pool
    const bool stage = is_pageable(src_host);
    const size_t wbytes = (size_t)W * 4;
    if (stage && hp.stage_cap < wbytes * (size_t)H) {
        if (hp.hstage) CUDA_TRY(cudaFreeHost(hp.hstage));
        hp.hstage = nullptr;
        hp.stage_cap = 0;
        CUDA_TRY(cudaHostAlloc(&hp.hstage, wbytes * (size_t)H, cudaHostAllocDefault));
        hp.stage_cap = wbytes * (size_t)H;
    }
    // last input band an output band of stage s needs (the last row it can touch); evaluated on
    // first use, i.e. after the first uploads are already under way (8 x 2049 polynomial samples
    // cost the host ~50 us)
    int need[2][kMaxBands];
    for (int st = 0; st < 2; ++st)
        for (int b = 0; b < nbands; ++b) need[st][b] = -1;
    auto need_of = [&](int st, int k) -> int {
        if (need[st][k] < 0) {
            int lo = 0, hi = H - 1;
            stages[st].rows(edge[k], edge[k + 1], &lo, &hi);
            need[st][k] = band_of(std::max(0, std::min(H - 1, hi)));
        }
        return need[st][k];
    };
    // DCB_PIPE_TRACE=1 (diagnostics): device-side time of every band's upload, first-stage kernel
    // start, last-stage kernel end and download, relative to the first upload's start, on stderr
    const bool trace = getenv("DCB_PIPE_TRACE") != nullptr && getenv("DCB_PIPE_TRACE")[0] == '1';
    cudaEvent_t tr0 = nullptr, tr[4][kMaxBands];
    if (trace) {
        CUDA_TRY(cudaEventCreate(&tr0));
        for (int k = 0; k < 4; ++k)
            for (int b = 0; b < nbands; ++b) CUDA_TRY(cudaEventCreate(&tr[k][b]));
        CUDA_TRY(cudaEventRecord(tr0, hp.up));
    }
    // DCB_PIPE_NOKERNEL (diagnostics): copies only -- what the two PCIe directions allow
    const bool no_kernel = getenv("DCB_PIPE_NOKERNEL") != nullptr;
    // DCB_PIPE_DIRECT=1: the last stage's kernels store straight into the caller's page-locked
    // destination (posted PCIe writes, one full 128-byte line per warp instruction) -- no device
    // copy of the result, no download stream, no kernel -> copy hand-over per band.  Measured on
    // the round's boxes: 41-46 GB/s against the copy engine's 55, the call 1.76 ms against 1.78
    // (profiles/r2/e2e_edges_r2z8.txt) -- opt-in.
    char *dst_dev = nullptr;
    if (getenv("DCB_PIPE_DIRECT") != nullptr && getenv("DCB_PIPE_DIRECT")[0] == '1' && !is_pageable(dst_host)) {
        void *q = nullptr;
        if (cudaHostGetDevicePointer(&q, dst_host, 0) == cudaSuccess && q != nullptr)
            dst_dev = (char *)q;
        else
            cudaGetLastError();
    }
    const int last = nstages - 1;
    int next[2] = {0, 0}, waited = -1;
    // the final stage's output band `k`: kernel into the device result (or the caller's buffer),
    // then its download
    auto launch_band = [&](int st, int k) -> int {
        const int q0 = edge[k], qn = edge[k + 1] - q0;
        const char *in = st == 0 ? dsrc : dmid;
        char *out = st < last ? dmid + (size_t)q0 * pitch
                              : (dst_dev != nullptr ? dst_dev + (size_t)q0 * dst_pitch_host : ddst + (size_t)q0 * pitch);
        const size_t out_pitch = (st == last && dst_dev != nullptr) ? dst_pitch_host : pitch;
        if (trace && st == 0) CUDA_TRY(cudaEventRecord(tr[3][k], hp.run));
        if (!no_kernel) {
            const int r = stages[st].launch((const float *)in, pitch, (float *)out, out_pitch, q0, qn, hp.run);
            if (r) {
                cudaDeviceSynchronize();
                return r;
            }
        }
        if (st < last) return DCB_OK;
        if (trace) CUDA_TRY(cudaEventRecord(tr[1][k], hp.run));
        if (dst_dev != nullptr) {
            if (trace) CUDA_TRY(cudaEventRecord(tr[2][k], hp.run));
            return DCB_OK;
        }
        CUDA_TRY(cudaEventRecord(hp.ev_run[k], hp.run));
        CUDA_TRY(cudaStreamWaitEvent(hp.down, hp.ev_run[k], 0));
        if (dst_pitch_host == wbytes && pitch == wbytes)
            CUDA_TRY(cudaMemcpyAsync((char *)dst_host + (size_t)q0 * dst_pitch_host, ddst + (size_t)q0 * pitch,
                                     wbytes * (size_t)qn, cudaMemcpyDeviceToHost, hp.down));
        else
            CUDA_TRY(cudaMemcpy2DAsync((char *)dst_host + (size_t)q0 * dst_pitch_host, dst_pitch_host,
                                       ddst + (size_t)q0 * pitch, pitch, wbytes, qn, cudaMemcpyDeviceToHost,
                                       hp.down));
        if (trace) CUDA_TRY(cudaEventRecord(tr[2][k], hp.down));
        return DCB_OK;
    };
    // uploads in row order
    for (int b = 0; b < nbands; ++b) {
        const int r0 = edge[b], nr = edge[b + 1] - r0;
        const char *from = (const char *)src_host + (size_t)r0 * src_pitch_host;
        size_t from_pitch = src_pitch_host;
        if (stage) {
            char *to = (char *)hp.hstage + (size_t)r0 * wbytes;
            CopyPool::get().run({from, to, src_pitch_host, wbytes, wbytes, nr});
            from = to;
            from_pitch = wbytes;
        }
        // (rows that are contiguous on both sides go as ONE linear copy)
        if (from_pitch == wbytes && pitch == wbytes)
            CUDA_TRY(cudaMemcpyAsync(dsrc + (size_t)r0 * pitch, from, wbytes * (size_t)nr,
                                     cudaMemcpyHostToDevice, hp.up));
        else
            CUDA_TRY(cudaMemcpy2DAsync(dsrc + (size_t)r0 * pitch, pitch, from, from_pitch, wbytes, nr,
                                       cudaMemcpyHostToDevice, hp.up));
        CUDA_TRY(cudaEventRecord(hp.ev_up[b], hp.up));
        if (trace) CUDA_TRY(cudaEventRecord(tr[0][b], hp.up));
        // (pinned source: every upload is enqueued before the first launch, as measured best)
        if (!stage && b < nbands - 1) continue;
        const bool all_up = b == nbands - 1;
        for (; next[0] < nbands && (all_up || need_of(0, next[0]) <= b); ++next[0]) {
            const int nd = need_of(0, next[0]);
            if (nd > waited) {
                CUDA_TRY(cudaStreamWaitEvent(hp.run, hp.ev_up[nd], 0));
                waited = nd;
            }
            rc = launch_band(0, next[0]);
            if (rc) return rc;
            // second stage: every band whose last input row has now been launched
            for (; nstages > 1 && next[1] < nbands &&
                   (next[0] == nbands - 1 || need_of(1, next[1]) <= next[0]); ++next[1]) {
                rc = launch_band(1, next[1]);
                if (rc) return rc;
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(hp.down));
    CUDA_TRY(cudaStreamSynchronize(hp.run));
    CUDA_TRY(cudaStreamSynchronize(hp.up));
    if (trace) {
        fprintf(stderr, "[dcb] pipe trace (%d bands, %d stage(s); rows, needs, us: upload done / kernel start / kernel done / download done)\n", nbands, nstages);
        for (int b = 0; b < nbands; ++b) {
            float t[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], tr0, tr[k][b]);
            fprintf(stderr, "[dcb]   band %2d rows %5d need %2d  %8.1f %8.1f %8.1f %8.1f\n", b, edge[b + 1] - edge[b],
                    need_of(0, b), t[0] * 1e3, t[3] * 1e3, t[1] * 1e3, t[2] * 1e3);
        }
        cudaEventDestroy(tr0);
        for (int k = 0; k < 4; ++k)
            for (int b = 0; b < nbands; ++b) cudaEventDestroy(tr[k][b]);
    }
    return DCB_OK;
}

static PipeStage radial_stage(const dcb_radial *model, const dcb_options *opt, int H, int W) {
    PipeStage st;
    st.rows = [=](int r0, int r1, int *lo, int *hi) { radial_row_range(*model, H, W, r0, r1, lo, hi); };
    st.launch = [=](const float *in, size_t in_pitch, float *out, size_t out_pitch, int q0, int qn,
                    cudaStream_t stream) {
        return dcb_unwarp_stack_backward_f32(in, out, 1, H, W, 0, H, in_pitch, in_pitch * (size_t)H, out_pitch,
                                             out_pitch * (size_t)qn, q0, qn, 1, model, opt, stream);
    };
    return st;
}

static PipeStage persp_stage(const dcb_persp *model, const dcb_options *opt, int H, int W) {
    PipeStage st;
    st.rows = [=](int r0, int r1, int *lo, int *hi) { persp_row_range(*model, H, W, r0, r1, lo, hi); };
    st.launch = [=](const float *in, size_t in_pitch, float *out, size_t out_pitch, int q0, int qn,
                    cudaStream_t stream) {
        return persp_rows_f32(in, out, H, W, in_pitch, out_pitch, q0, qn, model, opt, stream);
    };
    return st;
}

int dcb_host_copy_2d(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes,
                     int rows) {
    REQUIRE(dst != nullptr && src != nullptr, "null pointer");
    REQUIRE(rows >= 0 && dst_pitch >= width_bytes && src_pitch >= width_bytes, "bad pitch");
    if (rows == 0 || width_bytes == 0) return DCB_OK;
    CopyPool::get().run({(const char *)src, (char *)dst, src_pitch, dst_pitch, width_bytes, rows});
    return DCB_OK;
}

int dcb_unwarp_image_backward_host_f32(const float *src_host, float *dst_host, int H, int W,
                                       size_t src_pitch_host, size_t dst_pitch_host,
                                       const dcb_radial *model, const dcb_options *opt, int nbands) {
    REQUIRE(model != nullptr, "radial model is NULL");
    const PipeStage st = radial_stage(model, opt, H, W);
    return host_pipeline(src_host, dst_host, H, W, src_pitch_host, dst_pitch_host, nbands, 1, &st);
}

int dcb_correct_perspective_image_host_f32(const float *src_host, float *dst_host, int H, int W,
                                           size_t src_pitch_host, size_t dst_pitch_host,
                                           const dcb_persp *model, const dcb_options *opt, int nbands) {
    REQUIRE(model != nullptr, "perspective model is NULL");
    const PipeStage st = persp_stage(model, opt, H, W);
    return host_pipeline(src_host, dst_host, H, W, src_pitch_host, dst_pitch_host, nbands, 1, &st);
}

int dcb_unwarp_image_backward_perspective_host_f32(const float *src_host, float *dst_host, int H, int W,
                                                   size_t src_pitch_host, size_t dst_pitch_host,
                                                   const dcb_radial *radial, const dcb_persp *persp,
                                                   const dcb_options *opt, int nbands) {
    REQUIRE(radial != nullptr, "radial model is NULL");
    REQUIRE(persp != nullptr, "perspective model is NULL");
    const PipeStage st[2] = {radial_stage(radial, opt, H, W), persp_stage(persp, opt, H, W)};
    return host_pipeline(src_host, dst_host, H, W, src_pitch_host, dst_pitch_host, nbands, 2, st);
}

}  // extern "C"

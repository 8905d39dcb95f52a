/*
 * oracle_c.c -- plain-C restatement of Discorpy's backward-unwarp hot path.
 *
 * *** TEST INFRASTRUCTURE, NOT PRODUCT ***  Only tests/, __graft_entry__.smoke()
 * and bench.py's CPU legs may load liboracle_c.so; discorpy_b200 never does.
 *
 * Parity status: PINNED -- tests/test_oracle.py checks this file against the
 * NumPy oracle (itself bit-identical to the real reference on the committed
 * golden vectors) and against tests/golden/ directly.  The only tolerated
 * difference is the one SURVEY.md section 7 (hard part 6) describes: libm `pow` here
 * vs NumPy's `power` loop may differ in the last fp64 bit, which can flip a
 * float32-rounded coordinate with probability ~1e-8 per pixel.
 *
 * Restates (paths relative to /root/reference/):
 *   discorpy/post/postprocessing.py:138-147   unwarp_image_backward
 *   discorpy/post/postprocessing.py:214-228   unwarp_slice_backward
 *   discorpy/post/postprocessing.py:302-312   unwarp_chunk_slices_backward
 *   discorpy/post/postprocessing.py:450-457   _generate_perspective_map
 * and the order-0/1 arithmetic of scipy.ndimage.map_coordinates (third-party,
 * source not in the reference tree; formula in SURVEY.md section 8 a1).
 *
 * Build: make -C oracle   (gcc -O2 -pthread -ffp-contract=off)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <pthread.h>

static double clipd(double v, double lo, double hi) {
    /* np.clip: NaN propagates */
    if (v != v) return v;
    return v < lo ? lo : (v > hi ? hi : v);
}

/* sum_i a[i] * r**i, terms added first to last (postprocessing.py:142-143) */
static double radial_factor(double r, const double *a, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        double p;
        if (i == 0) p = 1.0;
        else if (i == 1) p = r;
        else if (i == 2) p = r * r;
        else p = pow(r, (double)i);
        double term = a[i] * p;
        acc = (i == 0) ? term : acc + term;
    }
    return acc;
}

/* order-1 / order-0 sample of one slice at a coordinate inside the image;
 * rows are additionally folded into [ylo, yhi] (the reference's cropped window).  Domain: the
 * coordinate lies within one row of the window, where folding equals SciPy's 'reflect'; chunk rows
 * that sample further outside (possible because postprocessing.py:289-301 takes the window from the
 * first and last row only) are covered by oracle_np.reflect_coordinate, not by this restatement */
static float sample(const float *img, int64_t pitch, int W, int ylo, int yhi, double y, double x,
                    int order) {
    if (order == 0) {
        int64_t yi = (int64_t)floor(y + 0.5), xi = (int64_t)floor(x + 0.5);
        if (yi < ylo) yi = ylo;
        if (yi > yhi) yi = yhi;
        return img[yi * pitch + xi];
    }
    double y0f = floor(y), x0f = floor(x);
    double ty = y - y0f, tx = x - x0f;
    int64_t y0 = (int64_t)y0f, x0 = (int64_t)x0f;
    int64_t y1 = y0 + 1, x1 = x0 + 1;
    if (x1 > W - 1) x1 = W - 1;
    if (y1 > yhi) y1 = yhi;
    if (y1 < ylo) y1 = ylo;
    if (y0 > yhi) y0 = yhi;
    if (y0 < ylo) y0 = ylo;
    double wy0 = 1.0 - ty, wx0 = 1.0 - tx;
    double t = 0.0;
    t += ((double)img[y0 * pitch + x0] * wy0) * wx0;
    t += ((double)img[y0 * pitch + x1] * wy0) * tx;
    t += ((double)img[y1 * pitch + x0] * ty) * wx0;
    t += ((double)img[y1 * pitch + x1] * ty) * tx;
    return (float)t;
}

/* ---- row-range workers, run on `nthreads` pthreads (no OpenMP runtime in this image) ---- */
typedef struct {
    const float *src;
    float *dst;
    int D, H, W, row0, nrows, ylo, yhi, coord_round, n, order;
    double xc, yc;
    const double *a; /* radial coefficients or the 8 perspective coefficients */
    int j0, j1;      /* this worker's rows [j0, j1) of the output */
    int persp;
} job_t;

static void radial_rows(const job_t *q) {
    const int W = q->W, H = q->H;
    for (int j = q->j0; j < q->j1; ++j) {
        const int y = q->row0 + j;
        const double yu = (double)y - q->yc;
        for (int x = 0; x < W; ++x) {
            const double xu = (double)x - q->xc;
            const double r = sqrt(xu * xu + yu * yu);
            const double f = radial_factor(r, q->a, q->n);
            double xd = clipd(q->xc + f * xu, 0.0, (double)(W - 1));
            double yd = clipd(q->yc + f * yu, 0.0, (double)(H - 1));
            if (q->coord_round) {
                xd = (double)(float)xd;
                yd = (double)(float)yd;
            }
            for (int z = 0; z < q->D; ++z)
                q->dst[((int64_t)z * q->nrows + j) * W + x] = sample(
                    q->src + (int64_t)z * H * W, W, W, q->ylo, q->yhi, yd, xd, q->order);
        }
    }
}

static void persp_rows(const job_t *q) {
    const int W = q->W, H = q->H;
    const double *c = q->a;
    for (int y = q->j0; y < q->j1; ++y)
        for (int x = 0; x < W; ++x) {
            const double den = c[6] * x + c[7] * y + 1.0;
            double xd = (c[0] * x + c[1] * y + c[2]) / den;
            double yd = (c[3] * x + c[4] * y + c[5]) / den;
            xd = (double)(float)clipd(xd, 0.0, (double)(W - 1));
            yd = (double)(float)clipd(yd, 0.0, (double)(H - 1));
            q->dst[(int64_t)y * W + x] = sample(q->src, W, W, 0, H - 1, yd, xd, q->order);
        }
}

static void *worker(void *arg) {
    const job_t *q = (const job_t *)arg;
    if (q->persp) persp_rows(q);
    else radial_rows(q);
    return NULL;
}

static void run_jobs(job_t base, int total_rows, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > total_rows) nthreads = total_rows > 0 ? total_rows : 1;
    pthread_t tid[256];
    job_t jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = base;
        jobs[t].j0 = (int)((int64_t)total_rows * t / nthreads);
        jobs[t].j1 = (int)((int64_t)total_rows * (t + 1) / nthreads);
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&tid[t], NULL, worker, &jobs[t]);
    worker(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(tid[t], NULL);
}

/* Rows row0..row0+nrows-1 of every slice of a (D,H,W) stack.  src points at
 * image row 0 of slice 0 (whole slices, dense).  coord_round = 1: coordinates
 * rounded to float32 (image / chunk semantics); 0: kept in double (slice).
 * ylo..yhi: rows the taps may touch (0..H-1 for whole images). */
void orc_unwarp_stack_backward_f32(const float *src, float *dst, int D, int H, int W, int row0,
                                   int nrows, int ylo, int yhi, int coord_round, double xc,
                                   double yc, const double *a, int n, int order, int nthreads) {
    job_t q = {src, dst, D, H, W, row0, nrows, ylo, yhi, coord_round, n, order, xc, yc, a, 0, 0, 0};
    run_jobs(q, nrows, nthreads);
}

void orc_correct_perspective_f32(const float *src, float *dst, int H, int W, const double *c,
                                 int order, int nthreads) {
    job_t q = {src, dst, 1, H, W, 0, H, 0, H - 1, 1, 8, order, 0.0, 0.0, c, 0, 0, 1};
    run_jobs(q, H, nthreads);
}

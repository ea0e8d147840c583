/*
 * TEST INFRASTRUCTURE -- CPU oracle #2: plain-C restatement of the reference's
 * match -> triangulate -> normal-optimise -> patch-extract path.
 *
 * Not part of the product: only tests/, bench.py's cpu_baseline / --impl reference legs and
 * __graft_entry__.smoke() may load the library built from this file (oracle/Makefile ->
 * oracle/_build/libfm3d_oracle.so).  The product (libfm3d.so) has no CPU path.
 *
 * Why it exists: the reference cannot be compiled in this image (needs OpenCV 2.4 C++ with
 * `nonfree`, PCL, Boost, lmfit), and oracle/oracle_cv.py (Python on real cv2 calls) is too
 * slow to check thousands of features or to serve as the same-run CPU baseline.  This file
 * restates the same arithmetic without OpenCV, function by function; tests/test_oracle_pins.py and tests/test_golden_oracle.py
 * pins it against oracle_cv.py (i.e. against cv2.undistortPoints / triangulatePoints /
 * projectPoints / pyrDown / BFMatcher) and against the committed tests/golden vectors.
 *
 * Reference lines followed (paths under the reference tree):
 *   orc_knn2_f32 / orc_knn2_hamming / orc_nndr_filter   DescriptorsMatcher/descriptorsmatcher.cpp:117-129
 *                                                       (exact brute force as cv::BFMatcher; ties -> lower index)
 *   orc_undistort_points                                cv::undistortPoints as called at
 *                                                       Triangulator/singlecameratriangulator.cpp:169-170,542
 *   orc_triangulate                                     singlecameratriangulator.cpp:145-230 (cv::triangulatePoints, 4x4 DLT)
 *   orc_pyrdown                                         Triangulator/normaloptimizer.cpp:206-221 (cv::pyrDown, u8)
 *   orc_disc_pixels                                     singlecameratriangulator.cpp:341-397
 *   eval_normal                                         normaloptimizer.cpp:65-149 + singlecameratriangulator.cpp:421-470,
 *                                                       530-665 + tools.cpp:129-142,767-777
 *   orc_lmmin                                           lmfit lmmin as called at normaloptimizer.cpp:269-287
 *                                                       (restated; see oracle/lmmin_py.py for the provenance note)
 *   orc_optimize_normals                                normaloptimizer.cpp:223-292,321-452
 *   orc_feature_frames                                  normaloptimizer.cpp:454-505
 *   orc_extract_patches                                 Triangulator/neighborhoodsgenerator.cpp:134-158 +
 *                                                       singlecameratriangulator.cpp:805-849
 * Deviations D1/D2/D12 are the ones listed in oracle/oracle_cv.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC   (no FMA contraction: the
 * reference's float sampler rounds every multiply and add separately).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_EPS 2.220446049250313e-16
#define ORC_DWARF 2.2250738585072014e-308

enum { FEAT_OK = 0, FEAT_NO_PIXELS = 1, FEAT_ABORT_BBOX = 2, FEAT_ABORT_PIXEL = 3, FEAT_ABORT_NAN = 4 };
enum { PENALTY_FABS = 0, PENALTY_INT_ABS = 1, PENALTY_OFF = 2 };
/* COST_SSD: the reference's residual, I1 - I2 (normaloptimizer.cpp:145-148).  COST_NCC: the zero-mean normalised residual
 * a_i/|a| - b_i/|b| (a = I1 - mean I1, b = I2 - mean I2), whose squared sum is 2 - 2 NCC(I1, I2): not in the reference
 * (SURVEY fact 4), an option of the new build the north star names; restated here so the kernel has something to equal. */
enum { COST_SSD = 0, COST_NCC = 1 };

typedef struct {
    double fx, fy, cx, cy;
    double k1, k2, p1, p2, k3;
    double R[9], t[3]; /* g12: X2 = R X1 + t */
    double zmin, zmax;
} orc_camera;

static void cam_init(orc_camera* c, const double* K, const double* dist, const double* g12,
                     double zmin, double zmax) {
    c->fx = K[0]; c->fy = K[4]; c->cx = K[2]; c->cy = K[5];
    c->k1 = dist[0]; c->k2 = dist[1]; c->p1 = dist[2]; c->p2 = dist[3]; c->k3 = dist[4];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) c->R[i * 3 + j] = g12 ? g12[i * 4 + j] : (i == j);
        c->t[i] = g12 ? g12[i * 4 + 3] : 0.0;
    }
    c->zmin = zmin; c->zmax = zmax;
}

/* ------------------------------------------------------------------ matcher ---- */

int orc_knn2_f32(const float* q, int nq, const float* t, int nt, int dim, int32_t* idx,
                 float* dist, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int i = 0; i < nq; i++) {
        float b0 = INFINITY, b1 = INFINITY;
        int i0 = -1, i1 = -1;
        const float* a = q + (size_t)i * dim;
        for (int j = 0; j < nt; j++) {
            const float* b = t + (size_t)j * dim;
            float s = 0.f;
            for (int k = 0; k < dim; k++) { float d = a[k] - b[k]; s += d * d; }
            if (s < b0) { b1 = b0; i1 = i0; b0 = s; i0 = j; }
            else if (s < b1) { b1 = s; i1 = j; }
        }
        idx[2 * i] = i0; idx[2 * i + 1] = i1;
        dist[2 * i] = i0 >= 0 ? sqrtf(b0) : INFINITY;
        dist[2 * i + 1] = i1 >= 0 ? sqrtf(b1) : INFINITY;
    }
    return 0;
}

int orc_knn2_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int nbytes,
                     int32_t* idx, float* dist, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int i = 0; i < nq; i++) {
        int b0 = 1 << 30, b1 = 1 << 30, i0 = -1, i1 = -1;
        const uint8_t* a = q + (size_t)i * nbytes;
        for (int j = 0; j < nt; j++) {
            const uint8_t* b = t + (size_t)j * nbytes;
            int s = 0;
            for (int k = 0; k < nbytes; k++) s += __builtin_popcount((unsigned)(a[k] ^ b[k]));
            if (s < b0) { b1 = b0; i1 = i0; b0 = s; i0 = j; }
            else if (s < b1) { b1 = s; i1 = j; }
        }
        idx[2 * i] = i0; idx[2 * i + 1] = i1;
        dist[2 * i] = i0 >= 0 ? (float)b0 : INFINITY;
        dist[2 * i + 1] = i1 >= 0 ? (float)b1 : INFINITY;
    }
    return 0;
}

/* descriptorsmatcher.cpp:119-129 */
int orc_nndr_filter(const int32_t* idx, const float* dist, int nq, double eps, int32_t* qidx,
                    int32_t* tidx, float* dout) {
    int n = 0;
    for (int i = 0; i < nq; i++) {
        if (idx[2 * i + 1] >= 0 && idx[2 * i] >= 0) {
            if ((double)dist[2 * i] <= eps * (double)dist[2 * i + 1]) {
                qidx[n] = i; tidx[n] = idx[2 * i]; dout[n] = dist[2 * i]; n++;
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------- lens model ---- */

static inline void undistort1(const orc_camera* c, double u, double v, double* xo, double* yo) {
    double x0 = (u - c->cx) / c->fx, y0 = (v - c->cy) / c->fy;
    double x = x0, y = y0;
    for (int it = 0; it < 5; it++) {
        double r2 = x * x + y * y;
        double icd = 1.0 / (1.0 + ((c->k3 * r2 + c->k2) * r2 + c->k1) * r2);
        if (icd < 0) { x = x0; y = y0; break; }
        double dx = 2 * c->p1 * x * y + c->p2 * (r2 + 2 * x * x);
        double dy = c->p1 * (r2 + 2 * y * y) + 2 * c->p2 * x * y;
        x = (x0 - dx) * icd;
        y = (y0 - dy) * icd;
    }
    *xo = x; *yo = y;
}

static inline void distort_project(const orc_camera* c, double X, double Y, double Z, double* u,
                                   double* v) {
    double z = Z ? 1.0 / Z : 1.0;
    double x = X * z, y = Y * z;
    double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    double cdist = 1 + c->k1 * r2 + c->k2 * r4 + c->k3 * r6;
    double xd = x * cdist + c->p1 * a1 + c->p2 * a2;
    double yd = y * cdist + c->p1 * a3 + c->p2 * a1;
    *u = xd * c->fx + c->cx;
    *v = yd * c->fy + c->cy;
}

int orc_undistort_points(const double* K, const double* dist, const double* pts, int n,
                         double* out) {
    orc_camera c; cam_init(&c, K, dist, 0, 0, 0);
    for (int i = 0; i < n; i++) undistort1(&c, pts[2 * i], pts[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
    return 0;
}

/* project camera-`view` (1: identity pose, 2: g12) with distortion; as cv::projectPoints */
int orc_project_points(const double* K, const double* dist, const double* g12, int view,
                       const double* X, int n, double* out) {
    orc_camera c; cam_init(&c, K, dist, view == 2 ? g12 : 0, 0, 0);
    for (int i = 0; i < n; i++) {
        const double* p = X + 3 * i;
        double x = c.R[0] * p[0] + c.R[1] * p[1] + c.R[2] * p[2] + c.t[0];
        double y = c.R[3] * p[0] + c.R[4] * p[1] + c.R[5] * p[2] + c.t[1];
        double z = c.R[6] * p[0] + c.R[7] * p[1] + c.R[8] * p[2] + c.t[2];
        distort_project(&c, x, y, z, &out[2 * i], &out[2 * i + 1]);
    }
    return 0;
}

/* ----------------------------------------------------------- triangulation ---- */

/* right singular vector of the smallest singular value of a 4x4 (one-sided Jacobi) */
static void null_vector4(const double A[16], double x[4]) {
    double U[16], V[16];
    memcpy(U, A, sizeof(U));
    memset(V, 0, sizeof(V));
    for (int i = 0; i < 4; i++) V[i * 4 + i] = 1;
    for (int sweep = 0; sweep < 60; sweep++) {
        int changed = 0;
        for (int p = 0; p < 3; p++)
            for (int q = p + 1; q < 4; q++) {
                double a = 0, b = 0, g = 0;
                for (int k = 0; k < 4; k++) {
                    a += U[k * 4 + p] * U[k * 4 + p];
                    b += U[k * 4 + q] * U[k * 4 + q];
                    g += U[k * 4 + p] * U[k * 4 + q];
                }
                if (fabs(g) <= 1e-300 || fabs(g) <= ORC_EPS * sqrt(a * b)) continue;
                changed = 1;
                double zeta = (b - a) / (2 * g);
                double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
                double cs = 1 / sqrt(1 + tt * tt), sn = cs * tt;
                for (int k = 0; k < 4; k++) {
                    double up = U[k * 4 + p], uq = U[k * 4 + q];
                    U[k * 4 + p] = cs * up - sn * uq; U[k * 4 + q] = sn * up + cs * uq;
                    double vp = V[k * 4 + p], vq = V[k * 4 + q];
                    V[k * 4 + p] = cs * vp - sn * vq; V[k * 4 + q] = sn * vp + cs * vq;
                }
            }
        if (!changed) break;
    }
    int jm = 0; double sm = INFINITY;
    for (int j = 0; j < 4; j++) {
        double s = 0;
        for (int k = 0; k < 4; k++) s += U[k * 4 + j] * U[k * 4 + j];
        if (s < sm) { sm = s; jm = j; }
    }
    for (int k = 0; k < 4; k++) x[k] = V[k * 4 + jm];
}

/* setKeypoints + triangulate. kp1/kp2 float (x,y); qidx/tidx may be NULL. */
int orc_triangulate(const double* K, const double* dist, const double* g12, double zmin,
                    double zmax, const float* kp1, const float* kp2, const int32_t* qidx,
                    const int32_t* tidx, int n, double* xyz_all, uint8_t* mask, double* xyz) {
    orc_camera c; cam_init(&c, K, dist, g12, zmin, zmax);
    double P2[12];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) P2[i * 4 + j] = c.R[i * 3 + j]; P2[i * 4 + 3] = c.t[i]; }
    int ninl = 0;
    for (int i = 0; i < n; i++) {
        int a = qidx ? qidx[i] : i, b = tidx ? tidx[i] : i;
        double u1x, u1y, u2x, u2y;
        undistort1(&c, (double)kp1[2 * a], (double)kp1[2 * a + 1], &u1x, &u1y);
        undistort1(&c, (double)kp2[2 * b], (double)kp2[2 * b + 1], &u2x, &u2y);
        double A[16];
        /* rows: x*P[2]-P[0], y*P[2]-P[1] for P1=[I|0] then P2 */
        A[0] = -1; A[1] = 0; A[2] = u1x; A[3] = 0;
        A[4] = 0; A[5] = -1; A[6] = u1y; A[7] = 0;
        for (int k = 0; k < 4; k++) {
            A[8 + k] = u2x * P2[8 + k] - P2[k];
            A[12 + k] = u2y * P2[8 + k] - P2[4 + k];
        }
        double h[4];
        null_vector4(A, h);
        double X = h[0] / h[3], Y = h[1] / h[3], Z = h[2] / h[3];
        xyz_all[3 * i] = X; xyz_all[3 * i + 1] = Y; xyz_all[3 * i + 2] = Z;
        int out = (Z < zmin) || (Z >= zmax) || isnan(Z);
        mask[i] = !out;
        if (!out) { xyz[3 * ninl] = X; xyz[3 * ninl + 1] = Y; xyz[3 * ninl + 2] = Z; ninl++; }
    }
    return ninl;
}

/* ----------------------------------------------------------------- pyrDown ---- */

static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}

/* dst is ((w+1)/2) x ((h+1)/2) */
int orc_pyrdown(const uint8_t* src, int w, int h, uint8_t* dst) {
    static const int kk[5] = {1, 4, 6, 4, 1};
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            int s = 0;
            for (int a = -2; a <= 2; a++) {
                int yy = reflect101(2 * y + a, h);
                int rs = 0;
                for (int b = -2; b <= 2; b++) rs += kk[b + 2] * src[(size_t)yy * w + reflect101(2 * x + b, w)];
                s += kk[a + 2] * rs;
            }
            dst[(size_t)y * dw + x] = (uint8_t)((s + 128) >> 8);
        }
    return 0;
}

/* ------------------------------------------------------------------ sampler ---- */

typedef struct { const uint8_t* p; int w, h; } orc_image;

/* getBilinearInterpPix32f (tools.cpp:129-142); flat addressing like a continuous cv::Mat (D2) */
static inline float bilinear32f(const orc_image* im, float x, float y) {
    int x0 = (int)floor((double)x), y0 = (int)floor((double)y);
    int x1 = x0 + 1, y1 = y0 + 1;
    long long n = (long long)im->w * im->h;
    long long i00 = (long long)y0 * im->w + x0, i01 = (long long)y1 * im->w + x0;
    long long i10 = (long long)y0 * im->w + x1, i11 = (long long)y1 * im->w + x1;
    float b0 = (i00 >= 0 && i00 < n) ? (float)im->p[i00] : 0.f;
    float b1 = (i01 >= 0 && i01 < n) ? (float)im->p[i01] : 0.f;
    float b2 = (i10 >= 0 && i10 < n) ? (float)im->p[i10] : 0.f;
    float b3 = (i11 >= 0 && i11 < n) ? (float)im->p[i11] : 0.f;
    float xm0 = 1.0f - (x - (float)x0), xm1 = (x - (float)x0);
    float ym0 = 1.0f - (y - (float)y0), ym1 = (y - (float)y0);
    return xm0 * (b0 * ym0 + b1 * ym1) + xm1 * (b2 * ym0 + b3 * ym1);
}

/* isPixelGood (singlecameratriangulator.cpp:657-665) */
static inline int pixel_good(double x, double y, double scale, int cols, int rows) {
    if ((x < 0) || (x > ((1 / scale) * cols)) || (y < 0) || (y > ((1 / scale) * rows))) return 0;
    return 1;
}

/* ------------------------------------------------------------ disc + evaluate ---- */

/* extractPixelsContour: returns m; pix (x,y) pairs, capacity (2r+1)^2 */
int orc_disc_pixels(const double* K, const double* dist, const double* P, int r, int width,
                    int height, double* pix) {
    orc_camera c; cam_init(&c, K, dist, 0, 0, 0);
    double cu, cv;
    distort_project(&c, P[0], P[1], P[2], &cu, &cv);
    int m = 0;
    for (int i = -r; i <= r; i++)
        for (int j = -r; j <= r; j++)
            if (i * i + j * j <= r * r) {
                double px = cu + i, py = cv + j;
                if (px < 0 || py < 0 || px >= width || py >= height) continue;
                pix[2 * m] = px; pix[2 * m + 1] = py; m++;
            }
    return m;
}

typedef struct {
    const orc_camera* cam;
    double P[3];
    int m;
    const double* pix;   /* m x 2 image-1 pixels */
    double* rays;        /* m x 2 undistorted (cached unless as_written) */
    float* i1;           /* m image-1 intensities at this level */
    int i1_valid;
    orc_image img1, img2;
    double scale;
    int penalty_mode;
    int as_written;      /* 1: redo undistort + image-1 sampling at every evaluation like the reference */
    int cost_mode;       /* COST_SSD / COST_NCC */
    float* i2;           /* COST_NCC: m image-2 intensities of the evaluation */
    int npenalty;
    int abort_code;
    long long pixel_evals;
} eval_ctx;

static double penalty_weight(double phi, double theta, int mode, int* entered) {
    *entered = 0;
    if (mode == PENALTY_OFF) return 1.0;
    double at, ap;
    if (mode == PENALTY_INT_ABS) { at = (double)abs((int)theta); ap = (double)abs((int)phi); }
    else { at = fabs(theta); ap = fabs(phi); }
    if (at - M_PI / 2 > 0 || ap - M_PI > 0) {
        double wt = exp(at - M_PI / 2) + 1;
        double wp = exp(ap - M_PI + 1) + 1;
        *entered = 1;
        return wp * wt;
    }
    return 1.0;
}

/* evaluateNormal (normaloptimizer.cpp:65-149). returns info (0 / -1) */
static int eval_normal(const double* par, int m_dat, void* data, double* fvec) {
    eval_ctx* D = (eval_ctx*)data;
    const orc_camera* c = D->cam;
    double phi = par[0], theta = par[1];
    double n[3] = {cos(theta) * cos(phi), cos(theta) * sin(phi), sin(theta)};
    if (isnan(n[0]) || isnan(n[1]) || isnan(n[2])) { D->abort_code = FEAT_ABORT_NAN; return -1; }
    int m = D->m;
    D->pixel_evals += m;
    if (D->as_written) {
        for (int i = 0; i < m; i++) undistort1(c, D->pix[2 * i], D->pix[2 * i + 1], &D->rays[2 * i], &D->rays[2 * i + 1]);
        D->i1_valid = 0;
    }
    int cmax = (int)(2 * c->zmax);
    double mnum = n[0] * D->P[0] + n[1] * D->P[1] + n[2] * D->P[2];
    /* A) get3dPointsFromImage1Pixels: all bounding-box tests come first */
    for (int i = 0; i < m; i++) {
        double vx = D->rays[2 * i], vy = D->rays[2 * i + 1];
        double k = mnum / (n[0] * vx + n[1] * vy + n[2]);
        double X = k * vx, Y = k * vy, Z = k;
        if (isnan(X) || isnan(Y) || isnan(Z)) { D->abort_code = FEAT_ABORT_NAN; return -1; }
        if (!((X > -cmax && X < cmax) && (Y > -cmax && Y < cmax) && (Z > 0 && Z < cmax))) {
            D->abort_code = FEAT_ABORT_BBOX; return -1;
        }
    }
    /* B) updateImage1PixelsIntensity */
    if (!D->i1_valid) {
        for (int i = 0; i < m; i++) {
            double x = D->pix[2 * i], y = D->pix[2 * i + 1];
            if (!pixel_good(x, y, D->scale, D->img1.w, D->img1.h)) { D->abort_code = FEAT_ABORT_PIXEL; return -1; }
            D->i1[i] = bilinear32f(&D->img1, (float)(D->scale * x), (float)(D->scale * y));
        }
        D->i1_valid = 1;
    }
    /* C) projectPointsToImage2 + residual */
    int entered;
    double w = penalty_weight(phi, theta, D->penalty_mode, &entered);
    if (entered) D->npenalty++;
    for (int i = 0; i < m; i++) {
        double vx = D->rays[2 * i], vy = D->rays[2 * i + 1];
        double k = mnum / (n[0] * vx + n[1] * vy + n[2]);
        double X = k * vx, Y = k * vy, Z = k;
        double x2 = c->R[0] * X + c->R[1] * Y + c->R[2] * Z + c->t[0];
        double y2 = c->R[3] * X + c->R[4] * Y + c->R[5] * Z + c->t[1];
        double z2 = c->R[6] * X + c->R[7] * Y + c->R[8] * Z + c->t[2];
        double u, v;
        distort_project(c, x2, y2, z2, &u, &v);
        if (!pixel_good(u, v, D->scale, D->img1.w, D->img1.h)) { D->abort_code = FEAT_ABORT_PIXEL; return -1; }
        float i2 = bilinear32f(&D->img2, (float)(D->scale * u), (float)(D->scale * v));
        if (D->cost_mode == COST_NCC) D->i2[i] = i2;
        else fvec[i] = w * (double)(D->i1[i] - i2);
    }
    if (D->cost_mode == COST_NCC) {
        double m1 = 0, m2 = 0, A = 0, B = 0;
        for (int i = 0; i < m; i++) { m1 += D->i1[i]; m2 += D->i2[i]; }
        m1 /= m; m2 /= m;
        for (int i = 0; i < m; i++) { double a = D->i1[i] - m1, b = D->i2[i] - m2; A += a * a; B += b * b; }
        A = sqrt(A); B = sqrt(B);
        if (!(A > 0) || !(B > 0)) { D->abort_code = FEAT_ABORT_NAN; return -1; }   /* a flat patch has no normalised residual */
        for (int i = 0; i < m; i++) fvec[i] = w * ((D->i1[i] - m1) / A - (D->i2[i] - m2) / B);
    }
    (void)m_dat;
    return 0;
}

/* ----------------------------------------------------------------- lmmin ---- */

static double enorm(int n, const double* x) {
    double s = 0;
    for (int i = 0; i < n; i++) s += x[i] * x[i];
    return sqrt(s);
}

#define NMAX 8

/* a: column-major m x n (a[j*m+i]) */
static void qrfac(int m, int n, double* a, int* ipvt, double* rdiag, double* acnorm, double* wa) {
    for (int j = 0; j < n; j++) {
        acnorm[j] = enorm(m, a + (size_t)j * m);
        rdiag[j] = acnorm[j]; wa[j] = rdiag[j]; ipvt[j] = j;
    }
    int minmn = m < n ? m : n;
    for (int j = 0; j < minmn; j++) {
        int kmax = j;
        for (int k = j; k < n; k++) if (rdiag[k] > rdiag[kmax]) kmax = k;
        if (kmax != j) {
            for (int i = 0; i < m; i++) { double t = a[(size_t)j * m + i]; a[(size_t)j * m + i] = a[(size_t)kmax * m + i]; a[(size_t)kmax * m + i] = t; }
            rdiag[kmax] = rdiag[j]; wa[kmax] = wa[j];
            int t = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = t;
        }
        double ajnorm = enorm(m - j, a + (size_t)j * m + j);
        if (ajnorm == 0.0) { rdiag[j] = 0; continue; }
        if (a[(size_t)j * m + j] < 0.0) ajnorm = -ajnorm;
        for (int i = j; i < m; i++) a[(size_t)j * m + i] /= ajnorm;
        a[(size_t)j * m + j] += 1.0;
        for (int k = j + 1; k < n; k++) {
            double sum = 0;
            for (int i = j; i < m; i++) sum += a[(size_t)j * m + i] * a[(size_t)k * m + i];
            double temp = sum / a[(size_t)j * m + j];
            for (int i = j; i < m; i++) a[(size_t)k * m + i] -= temp * a[(size_t)j * m + i];
            if (rdiag[k] != 0.0) {
                temp = a[(size_t)k * m + j] / rdiag[k];
                temp = fmax(0.0, 1.0 - temp * temp);
                rdiag[k] *= sqrt(temp);
                temp = rdiag[k] / wa[k];
                if (0.05 * temp * temp <= ORC_EPS) {
                    rdiag[k] = enorm(m - j - 1, a + (size_t)k * m + j + 1);
                    wa[k] = rdiag[k];
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

/* r: n x n row-major r[i*n+j]; upper triangle R, lower overwritten */
static void qrsolv(int n, double* r, const int* ipvt, const double* diag, const double* qtb,
                   double* x, double* sdiag) {
    double wa[NMAX];
    for (int j = 0; j < n; j++) {
        for (int i = j; i < n; i++) r[i * n + j] = r[j * n + i];
        x[j] = r[j * n + j]; wa[j] = qtb[j];
    }
    for (int j = 0; j < n; j++) {
        int l = ipvt[j];
        if (diag[l] != 0.0) {
            for (int k = j; k < n; k++) sdiag[k] = 0;
            sdiag[j] = diag[l];
            double qtbpj = 0;
            for (int k = j; k < n; k++) {
                if (sdiag[k] == 0.0) continue;
                double sn, cs;
                if (fabs(r[k * n + k]) < fabs(sdiag[k])) {
                    double cotan = r[k * n + k] / sdiag[k];
                    sn = 0.5 / sqrt(0.25 + 0.25 * cotan * cotan); cs = sn * cotan;
                } else {
                    double tn = sdiag[k] / r[k * n + k];
                    cs = 0.5 / sqrt(0.25 + 0.25 * tn * tn); sn = cs * tn;
                }
                r[k * n + k] = cs * r[k * n + k] + sn * sdiag[k];
                double temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < n; i++) {
                    temp = cs * r[i * n + k] + sn * sdiag[i];
                    sdiag[i] = -sn * r[i * n + k] + cs * sdiag[i];
                    r[i * n + k] = temp;
                }
            }
        }
        sdiag[j] = r[j * n + j];
        r[j * n + j] = x[j];
    }
    int nsing = n;
    for (int j = 0; j < n; j++) {
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0;
    }
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        double sum = 0;
        for (int i = j + 1; i < nsing; i++) sum += r[i * n + j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < n; j++) x[ipvt[j]] = wa[j];
}

static double lmpar(int n, double* r, const int* ipvt, const double* diag, const double* qtb,
                    double delta, double par, double* x, double* sdiag) {
    double wa1[NMAX], wa2[NMAX];
    int nsing = n;
    for (int j = 0; j < n; j++) {
        wa1[j] = qtb[j];
        if (r[j * n + j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0;
    }
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        wa1[j] /= r[j * n + j];
        double temp = wa1[j];
        for (int i = 0; i < j; i++) wa1[i] -= r[i * n + j] * temp;
    }
    for (int j = 0; j < n; j++) x[ipvt[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
    double dxnorm = enorm(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) return 0.0;
    double parl = 0;
    if (nsing >= n) {
        for (int j = 0; j < n; j++) { int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < n; j++) {
            double sum = 0;
            for (int i = 0; i < j; i++) sum += r[i * n + j] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j * n + j];
        }
        double temp = enorm(n, wa1);
        parl = fp / delta / temp / temp;
    }
    for (int j = 0; j < n; j++) {
        double sum = 0;
        for (int i = 0; i <= j; i++) sum += r[i * n + j] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    double gnorm = enorm(n, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = ORC_DWARF / fmin(delta, 0.1);
    par = fmax(par, parl);
    par = fmin(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;
    for (;;) {
        iter++;
        if (par == 0.0) par = fmax(ORC_DWARF, 0.001 * paru);
        double temp = sqrt(par);
        for (int j = 0; j < n; j++) wa1[j] = temp * diag[j];
        qrsolv(n, r, ipvt, wa1, qtb, x, sdiag);
        for (int j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
        dxnorm = enorm(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        for (int j = 0; j < n; j++) { int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < n; j++) {
            wa1[j] /= sdiag[j];
            double t2 = wa1[j];
            for (int i = j + 1; i < n; i++) wa1[i] -= r[i * n + j] * t2;
        }
        temp = enorm(n, wa1);
        double parc = fp / delta / temp / temp;
        if (fp > 0) parl = fmax(parl, par);
        if (fp < 0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
    return par;
}

typedef int (*orc_eval_fn)(const double* par, int m, void* data, double* fvec);

typedef struct { double ftol, xtol, gtol, epsilon, stepbound; int patience, scale_diag, minpack_mode; } orc_lm_control;
typedef struct { int info, nfev; double fnorm; } orc_lm_status;

/* lmmin restated (see header). work: caller-provided m*(n+2) doubles */
static void orc_lmmin_impl(int n, double* x, int m, void* data, orc_eval_fn evaluate,
                           const orc_lm_control* c, orc_lm_status* st, double* work) {
    double* fvec = work;
    double* wa4 = work + m;
    double* fjac = work + 2 * (size_t)m;
    double diag[NMAX], qtf[NMAX], wa1[NMAX], wa2[NMAX], wa3[NMAX], rdiag[NMAX], acnorm[NMAX], sdiag[NMAX];
    double r[NMAX * NMAX];
    int ipvt[NMAX];
    const double p1 = 0.1, p0001 = 1.0e-4;
    int maxfev = c->patience * (n + 1);
    double eps = sqrt(fmax(c->epsilon, ORC_EPS));
    st->info = 0; st->nfev = 0; st->fnorm = 0;
    if (n <= 0 || n > NMAX || m < n || c->ftol < 0 || c->xtol < 0 || c->gtol < 0 || maxfev <= 0 || c->stepbound <= 0) { st->info = 10; return; }
    int iter = 0;
    double par = 0, delta = 0, xnorm = 0;
    int info = evaluate(x, m, data, fvec);
    st->nfev++;
    if (info < 0) { st->info = 11; return; }
    double fnorm = enorm(m, fvec);
    for (;;) {
        for (int j = 0; j < n; j++) {
            double temp = x[j], step;
            if (c->minpack_mode) { step = eps * fabs(temp); if (step == 0.0) step = eps; }
            else step = fmax(eps * eps, eps * fabs(temp));
            x[j] = temp + step;
            info = evaluate(x, m, data, wa4);
            st->nfev++;
            if (info < 0) { x[j] = temp; st->info = 11; st->fnorm = fnorm; return; }
            for (int i = 0; i < m; i++) fjac[(size_t)j * m + i] = (wa4[i] - fvec[i]) / step;
            x[j] = temp;
        }
        qrfac(m, n, fjac, ipvt, rdiag, acnorm, wa3);
        if (!iter) {
            for (int j = 0; j < n; j++) diag[j] = c->scale_diag ? (acnorm[j] != 0.0 ? acnorm[j] : 1.0) : 1.0;
            for (int j = 0; j < n; j++) wa3[j] = diag[j] * x[j];
            xnorm = enorm(n, wa3);
            delta = c->stepbound * xnorm;
            if (delta == 0.0) delta = c->stepbound;
        } else if (c->scale_diag) {
            for (int j = 0; j < n; j++) diag[j] = fmax(diag[j], acnorm[j]);
        }
        for (int i = 0; i < m; i++) wa4[i] = fvec[i];
        for (int j = 0; j < n; j++) {
            double temp3 = fjac[(size_t)j * m + j];
            if (temp3 != 0.0) {
                double sum = 0;
                for (int i = j; i < m; i++) sum += fjac[(size_t)j * m + i] * wa4[i];
                double temp = -sum / temp3;
                for (int i = j; i < m; i++) wa4[i] += fjac[(size_t)j * m + i] * temp;
            }
            fjac[(size_t)j * m + j] = rdiag[j];
            qtf[j] = wa4[j];
        }
        double gnorm = 0;
        if (fnorm != 0.0) {
            for (int j = 0; j < n; j++) {
                if (acnorm[ipvt[j]] == 0.0) continue;
                double sum = 0;
                for (int i = 0; i <= j; i++) sum += fjac[(size_t)j * m + i] * qtf[i] / fnorm;
                gnorm = fmax(gnorm, fabs(sum / acnorm[ipvt[j]]));
            }
        }
        if (gnorm <= c->gtol) { st->info = 4; st->fnorm = fnorm; return; }
        for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) r[i * n + j] = (i <= j) ? fjac[(size_t)j * m + i] : 0.0;
        double ratio;
        do {
            par = lmpar(n, r, ipvt, diag, qtf, delta, par, wa1, sdiag);
            for (int j = 0; j < n; j++) { wa2[j] = x[j] - wa1[j]; wa3[j] = diag[j] * wa1[j]; }
            double pnorm = enorm(n, wa3);
            if (c->minpack_mode ? (iter == 0) : (st->nfev <= 1 + n)) delta = fmin(delta, pnorm);
            info = evaluate(wa2, m, data, wa4);
            st->nfev++;
            if (info < 0) { st->info = 11; st->fnorm = fnorm; return; }
            double fnorm1 = enorm(m, wa4);
            double actred = (p1 * fnorm1 < fnorm) ? 1 - (fnorm1 / fnorm) * (fnorm1 / fnorm) : -1;
            for (int j = 0; j < n; j++) {
                wa3[j] = 0;
                for (int i = 0; i <= j; i++) wa3[i] -= r[i * n + j] * wa1[ipvt[j]];
            }
            double temp1 = enorm(n, wa3) / fnorm;
            double temp2 = sqrt(par) * pnorm / fnorm;
            double prered = temp1 * temp1 + 2 * temp2 * temp2;
            double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = prered != 0 ? actred / prered : 0;
            if (ratio <= 0.25) {
                double temp;
                if (actred >= 0.0) temp = 0.5;
                else temp = 0.5 * dirder / (dirder + (c->minpack_mode ? 0.5 : 0.55) * actred);
                if (p1 * fnorm1 >= fnorm || temp < p1) temp = p1;
                delta = temp * fmin(delta, pnorm / p1);
                par /= temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par *= 0.5;
            }
            if (ratio >= p0001) {
                for (int j = 0; j < n; j++) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
                for (int i = 0; i < m; i++) fvec[i] = wa4[i];
                xnorm = enorm(n, wa2);
                fnorm = fnorm1;
                iter++;
            }
            info = 0;
            if (fabs(actred) <= c->ftol && prered <= c->ftol && 0.5 * ratio <= 1) info = 1;
            if (delta <= c->xtol * xnorm) info += 2;
            if (info != 0) { st->info = info; st->fnorm = fnorm; return; }
            if (st->nfev >= maxfev) info = 5;
            if (fabs(actred) <= ORC_EPS && prered <= ORC_EPS && 0.5 * ratio <= 1) info = 6;
            if (delta <= ORC_EPS * xnorm) info = 7;
            if (gnorm <= ORC_EPS) info = 8;
            if (info != 0) { st->info = info; st->fnorm = fnorm; return; }
        } while (ratio < p0001);
    }
}

/* Exposed for pinning against scipy.optimize.leastsq: f_i = x0*exp(x1*t_i) + x2*sin(x3*t_i) - y_i */
typedef struct { const double* t; const double* y; } orc_expsin;
static int expsin_eval(const double* p, int m, void* data, double* f) {
    orc_expsin* d = (orc_expsin*)data;
    for (int i = 0; i < m; i++) f[i] = p[0] * exp(p[1] * d->t[i]) + p[2] * sin(p[3] * d->t[i]) - d->y[i];
    return 0;
}
int orc_lmmin_expsin(const double* t, const double* y, int m, double* x, double epsilon,
                     int patience, int minpack_mode, int* nfev, int* info) {
    orc_expsin d = {t, y};
    orc_lm_control c = {30 * ORC_EPS, 30 * ORC_EPS, 30 * ORC_EPS, epsilon, 100.0, patience, 1, minpack_mode};
    orc_lm_status st;
    double* work = (double*)malloc(sizeof(double) * (size_t)m * 6);
    orc_lmmin_impl(4, x, m, &d, expsin_eval, &c, &st, work);
    free(work);
    *nfev = st.nfev; *info = st.info;
    return 0;
}

/* Two-parameter problem f_i = x0*exp(x1*t_i) - y_i: lets tests compare the product's n=2 LM
 * state machine (csrc/fm3d_lm2.h, driven by tests/cpp/lm2_harness.cpp) with this generic one. */
static int exp2_eval(const double* p, int m, void* data, double* f) {
    orc_expsin* d = (orc_expsin*)data;
    for (int i = 0; i < m; i++) f[i] = p[0] * exp(p[1] * d->t[i]) - d->y[i];
    return 0;
}
int orc_lmmin_exp2(const double* t, const double* y, int m, double* x, double epsilon,
                   int patience, int minpack_mode, int* nfev, int* info) {
    orc_expsin d = {t, y};
    orc_lm_control c = {30 * ORC_EPS, 30 * ORC_EPS, 30 * ORC_EPS, epsilon, 100.0, patience, 1, minpack_mode};
    orc_lm_status st;
    double* work = (double*)malloc(sizeof(double) * (size_t)m * 4);
    orc_lmmin_impl(2, x, m, &d, exp2_eval, &c, &st, work);
    free(work);
    *nfev = st.nfev; *info = st.info;
    return 0;
}

/* ------------------------------------------------------- normal optimisation ---- */

typedef struct {
    const uint8_t* data; /* all levels of one image, concatenated */
    int w[8], h[8];
    size_t off[8];
} orc_pyramid;

static void pyramid_layout(orc_pyramid* p, const uint8_t* data, int w, int h, int levels) {
    p->data = data;
    size_t off = 0;
    for (int l = 0; l <= levels; l++) {
        p->w[l] = w; p->h[l] = h; p->off[l] = off;
        off += (size_t)w * h;
        w = (w + 1) / 2; h = (h + 1) / 2;
    }
}

size_t orc_pyramid_bytes(int w, int h, int levels) {
    size_t off = 0;
    for (int l = 0; l <= levels; l++) { off += (size_t)w * h; w = (w + 1) / 2; h = (h + 1) / 2; }
    return off;
}

/* builds the concatenated pyramid (level 0 copied from img) */
int orc_build_pyramid(const uint8_t* img, int w, int h, int levels, uint8_t* out) {
    memcpy(out, img, (size_t)w * h);
    size_t off = 0;
    for (int l = 1; l <= levels; l++) {
        size_t next = off + (size_t)w * h;
        orc_pyrdown(out + off, w, h, out + next);
        off = next; w = (w + 1) / 2; h = (h + 1) / 2;
    }
    return 0;
}

static int optimize_one(const orc_camera* cam, const double* P, const orc_pyramid* p1,
                        const orc_pyramid* p2, int levels, int r, double eps_lmmin,
                        int penalty_mode, int cost_mode, int patience, int as_written, double* normal,
                        int32_t* nfev, int32_t* npenalty, double* cost, int32_t* m_out,
                        long long* pixel_evals) {
    int W = p1->w[0], H = p1->h[0];
    size_t cap = (size_t)(2 * r + 1) * (2 * r + 1);
    double* pix = (double*)malloc(sizeof(double) * cap * 2);
    double K[9] = {cam->fx, 0, cam->cx, 0, cam->fy, cam->cy, 0, 0, 1};
    double dist[5] = {cam->k1, cam->k2, cam->p1, cam->p2, cam->k3};
    int m = orc_disc_pixels(K, dist, P, r, W, H, pix);
    double nrm = sqrt(P[0] * P[0] + P[1] * P[1] + P[2] * P[2]);
    normal[0] = P[0] / nrm; normal[1] = P[1] / nrm; normal[2] = P[2] / nrm;
    if (m_out) *m_out = m;
    if (npenalty) *npenalty = 0;
    if (cost) *cost = NAN;
    for (int l = 0; l <= levels; l++) if (nfev) nfev[l] = 0;
    if (m <= 0) { free(pix); return FEAT_NO_PIXELS; }
    double* rays = (double*)malloc(sizeof(double) * 2 * m);
    float* i1 = (float*)malloc(sizeof(float) * m);
    float* i2 = (float*)malloc(sizeof(float) * m);
    double* work = (double*)malloc(sizeof(double) * (size_t)m * 4);
    for (int i = 0; i < m; i++) undistort1(cam, pix[2 * i], pix[2 * i + 1], &rays[2 * i], &rays[2 * i + 1]);
    int status = FEAT_OK;
    double n[3] = {normal[0], normal[1], normal[2]};
    float img_scale = (float)pow(2.0, (double)levels);
    for (int lvl = levels; lvl >= 0; lvl--) {
        eval_ctx D;
        memset(&D, 0, sizeof(D));
        D.cam = cam; D.P[0] = P[0]; D.P[1] = P[1]; D.P[2] = P[2];
        D.m = m; D.pix = pix; D.rays = rays; D.i1 = i1; D.i1_valid = 0;
        D.img1.p = p1->data + p1->off[lvl]; D.img1.w = p1->w[lvl]; D.img1.h = p1->h[lvl];
        D.img2.p = p2->data + p2->off[lvl]; D.img2.w = p2->w[lvl]; D.img2.h = p2->h[lvl];
        D.scale = 1.0 / img_scale;
        D.penalty_mode = penalty_mode; D.as_written = as_written;
        D.cost_mode = cost_mode; D.i2 = i2;
        double par[2];
        par[1] = atan2(n[2], sqrt(n[0] * n[0] + n[1] * n[1])); /* theta, car2sph tools.cpp:767-771 */
        par[0] = atan2(n[1], n[0]);                            /* phi */
        orc_lm_control c = {30 * ORC_EPS, 30 * ORC_EPS, 30 * ORC_EPS, eps_lmmin, 100.0, patience, 1, 0};
        orc_lm_status st;
        orc_lmmin_impl(2, par, m, &D, eval_normal, &c, &st, work);
        if (nfev) nfev[lvl] = st.nfev;
        if (npenalty) *npenalty += D.npenalty;
        if (pixel_evals) *pixel_evals += D.pixel_evals;
        if (st.info == 11) { status = D.abort_code; break; }
        n[0] = cos(par[1]) * cos(par[0]); n[1] = cos(par[1]) * sin(par[0]); n[2] = sin(par[1]);
        if (cost) *cost = st.fnorm * st.fnorm;
        img_scale /= 2.0f;
    }
    if (status == FEAT_OK) { normal[0] = n[0]; normal[1] = n[1]; normal[2] = n[2]; }
    free(pix); free(rays); free(i1); free(i2); free(work);
    return status;
}

/* pyr1/pyr2: concatenated pyramids from orc_build_pyramid. Returns total pixel evaluations. */
long long orc_optimize_normals2(const double* K, const double* dist, const double* g12,
                                double zmin, double zmax, const uint8_t* pyr1,
                                const uint8_t* pyr2, int w, int h, int levels,
                                const double* xyz, int n, int pixels_ray, double eps_lmmin,
                                int penalty_mode, int cost_mode, int patience, int as_written, int threads,
                                double* normals, int32_t* status, int32_t* nfev,
                                int32_t* npenalty, double* cost, int32_t* m_out) {
    orc_camera cam; cam_init(&cam, K, dist, g12, zmin, zmax);
    orc_pyramid p1, p2;
    pyramid_layout(&p1, pyr1, w, h, levels);
    pyramid_layout(&p2, pyr2, w, h, levels);
    long long total = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1) reduction(+ : total)
    for (int i = 0; i < n; i++) {
        long long pe = 0;
        status[i] = optimize_one(&cam, xyz + 3 * i, &p1, &p2, levels, pixels_ray, eps_lmmin,
                                 penalty_mode, cost_mode, patience, as_written, normals + 3 * i,
                                 nfev ? nfev + (size_t)i * (levels + 1) : 0,
                                 npenalty ? npenalty + i : 0, cost ? cost + i : 0,
                                 m_out ? m_out + i : 0, &pe);
        total += pe;
    }
    return total;
}

long long orc_optimize_normals(const double* K, const double* dist, const double* g12,
                               double zmin, double zmax, const uint8_t* pyr1,
                               const uint8_t* pyr2, int w, int h, int levels,
                               const double* xyz, int n, int pixels_ray, double eps_lmmin,
                               int penalty_mode, int patience, int as_written, int threads,
                               double* normals, int32_t* status, int32_t* nfev,
                               int32_t* npenalty, double* cost, int32_t* m_out) {
    return orc_optimize_normals2(K, dist, g12, zmin, zmax, pyr1, pyr2, w, h, levels, xyz, n, pixels_ray, eps_lmmin, penalty_mode,
                                 COST_SSD, patience, as_written, threads, normals, status, nfev, npenalty, cost, m_out);
}

/* one evaluateNormal call: cost = sum fvec^2 at (phi,theta), level `level` */
int orc_evaluate_cost2(const double* K, const double* dist, const double* g12, double zmin,
                       double zmax, const uint8_t* pyr1, const uint8_t* pyr2, int w, int h,
                       int levels, const double* xyz, const double* phi_theta, int n,
                       int pixels_ray, int level, int penalty_mode, int cost_mode, double* cost, int32_t* m_out,
                       int32_t* status) {
    orc_camera cam; cam_init(&cam, K, dist, g12, zmin, zmax);
    orc_pyramid p1, p2;
    pyramid_layout(&p1, pyr1, w, h, levels);
    pyramid_layout(&p2, pyr2, w, h, levels);
    size_t cap = (size_t)(2 * pixels_ray + 1) * (2 * pixels_ray + 1);
    double* pix = (double*)malloc(sizeof(double) * cap * 2);
    double* rays = (double*)malloc(sizeof(double) * cap * 2);
    float* i1 = (float*)malloc(sizeof(float) * cap);
    float* i2 = (float*)malloc(sizeof(float) * cap);
    double* fvec = (double*)malloc(sizeof(double) * cap);
    for (int f = 0; f < n; f++) {
        const double* P = xyz + 3 * f;
        int m = orc_disc_pixels(K, dist, P, pixels_ray, w, h, pix);
        m_out[f] = m; cost[f] = NAN;
        if (m <= 0) { status[f] = FEAT_NO_PIXELS; continue; }
        for (int i = 0; i < m; i++) undistort1(&cam, pix[2 * i], pix[2 * i + 1], &rays[2 * i], &rays[2 * i + 1]);
        eval_ctx D;
        memset(&D, 0, sizeof(D));
        D.cam = &cam; D.P[0] = P[0]; D.P[1] = P[1]; D.P[2] = P[2];
        D.m = m; D.pix = pix; D.rays = rays; D.i1 = i1;
        D.img1.p = p1.data + p1.off[level]; D.img1.w = p1.w[level]; D.img1.h = p1.h[level];
        D.img2.p = p2.data + p2.off[level]; D.img2.w = p2.w[level]; D.img2.h = p2.h[level];
        D.scale = 1.0 / pow(2.0, level);
        D.penalty_mode = penalty_mode;
        D.cost_mode = cost_mode; D.i2 = i2;
        int info = eval_normal(phi_theta + 2 * f, m, &D, fvec);
        if (info < 0) { status[f] = D.abort_code; continue; }
        status[f] = FEAT_OK;
        double s = 0;
        for (int i = 0; i < m; i++) s += fvec[i] * fvec[i];
        cost[f] = s;
    }
    free(pix); free(rays); free(i1); free(i2); free(fvec);
    return 0;
}

int orc_evaluate_cost(const double* K, const double* dist, const double* g12, double zmin,
                      double zmax, const uint8_t* pyr1, const uint8_t* pyr2, int w, int h,
                      int levels, const double* xyz, const double* phi_theta, int n,
                      int pixels_ray, int level, int penalty_mode, double* cost, int32_t* m_out,
                      int32_t* status) {
    return orc_evaluate_cost2(K, dist, g12, zmin, zmax, pyr1, pyr2, w, h, levels, xyz, phi_theta, n, pixels_ray, level,
                              penalty_mode, COST_SSD, cost, m_out, status);
}

/* ------------------------------------------------------------ frames + patches ---- */

int orc_feature_frames(const double* xyz, const double* normals, int n, const double* g,
                       double* frames) {
    for (int i = 0; i < n; i++) {
        const double* z = normals + 3 * i;
        double x[3] = {g[1] * z[2] - g[2] * z[1], g[2] * z[0] - g[0] * z[2], g[0] * z[1] - g[1] * z[0]};
        double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
        double nx = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        double ny = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
        double* F = frames + 16 * i;
        for (int k = 0; k < 3; k++) {
            F[k * 4 + 0] = x[k] / nx; F[k * 4 + 1] = y[k] / ny; F[k * 4 + 2] = z[k]; F[k * 4 + 3] = xyz[3 * i + k];
        }
        F[12] = 0; F[13] = 0; F[14] = 0; F[15] = 1;
    }
    return 0;
}

int orc_patch_size(double epsilon_m, double cm_per_pixel) {
    return 2 * ((int)floor(epsilon_m / (0.01 * cm_per_pixel)));
}

/* patches n x S x S, image_points n x S*S x 2 (nullable). img1 full-resolution w x h. */
int orc_extract_patches(const double* K, const double* dist, const uint8_t* img1, int w, int h,
                        const double* frames, int n, double epsilon_m, double cm_per_pixel,
                        uint8_t* patches, double* image_points, int threads) {
    orc_camera cam; cam_init(&cam, K, dist, 0, 0, 0);
    int S = orc_patch_size(epsilon_m, cm_per_pixel);
    double inc = cm_per_pixel * 0.01;
    orc_image im = {img1, w, h};
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int f = 0; f < n; f++) {
        const double* F = frames + 16 * f;
        uint8_t* patch = patches + (size_t)f * S * S;
        for (int i = 0; i < S; i++)
            for (int j = 0; j < S; j++) {
                double rx = -epsilon_m + inc * i, ry = -epsilon_m + inc * j;
                double X = F[0] * rx + F[1] * ry + F[3];
                double Y = F[4] * rx + F[5] * ry + F[7];
                double Z = F[8] * rx + F[9] * ry + F[11];
                double u, v;
                distort_project(&cam, X, Y, Z, &u, &v);
                if (image_points) {
                    image_points[((size_t)f * S * S + (size_t)i * S + j) * 2] = u;
                    image_points[((size_t)f * S * S + (size_t)i * S + j) * 2 + 1] = v;
                }
                uint8_t val = 0;
                if (pixel_good(u, v, 1.0, w, h)) val = (uint8_t)bilinear32f(&im, (float)u, (float)v);
                patch[(size_t)j * S + i] = val; /* patch.at<uchar>(col=j, row=i) */
            }
    }
    return 0;
}

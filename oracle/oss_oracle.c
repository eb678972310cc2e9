/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.  See oss_oracle.h for what this
 * is, who may call it, and how it is pinned against cv2 and the reference's own
 * object code.  All arithmetic is IEEE fp32/fp64, round-to-nearest, evaluated
 * left to right exactly as written; build with -ffp-contract=off (no FMA).
 * Citations are relative to /root/reference/. */
#include "oss_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ========================================================================== */
/* a-1  convertAndScaleImage, libstacker/src/util.cpp:441-450                  */
/* convertTo(CV_32F, alpha) computes float(v) * float(alpha).                  */

void orc_u16_to_f32(const uint16_t *src, float *dst, size_t n)
{
    const float s = (float)(1 / 65535.0);
    for (size_t i = 0; i < n; i++) dst[i] = (float)src[i] * s;
}

void orc_u8_to_f32(const uint8_t *src, float *dst, size_t n)
{
    const float s = (float)(1 / 255.0);
    for (size_t i = 0; i < n; i++) dst[i] = (float)src[i] * s;
}

/* ========================================================================== */
/* a-2  stackBias/stackDarks/stackDarkFlats/stackFlats, util.cpp:584-722       */
/* result = f0 - subs; result += (fi - subs) for i>=1; result *= float(1/n).   */

void orc_stack_mean(const float *const *frames, int n, size_t count, const float *sub1,
                    const float *sub2, float *out)
{
    for (int f = 0; f < n; f++) {
        const float *src = frames[f];
        for (size_t i = 0; i < count; i++) {
            float v = src[i];
            if (sub1) v = v - sub1[i]; /* util.cpp:593,604,627,638,662,675: bias   */
            if (sub2) v = v - sub2[i]; /* util.cpp:663,676: dark flat              */
            out[i] = f == 0 ? v : out[i] + v;
        }
    }
    /* `result /= n` is convertTo(.., 1.0/n): util.cpp:610,644,682,716 */
    const float r = (float)(1.0 / (double)n);
    for (size_t i = 0; i < count; i++) out[i] = out[i] * r;
}

void orc_flat_normalize(float *flat, size_t npix, int channels)
{
    /* util.cpp:686-688: float avg = cv::mean(result)[0]; result /= avg;        */
    double s = 0.0;
    for (size_t i = 0; i < npix; i++) s += (double)flat[i * channels];
    float avg = (float)(s * (1.0 / (double)npix));
    const float r = (float)(1.0 / (double)avg);
    for (size_t i = 0; i < npix * (size_t)channels; i++) flat[i] = flat[i] * r;
}

/* ========================================================================== */
/* a-3  getCalibratedImage, util.cpp:269-271                                   */

void orc_calibrate(float *img, size_t count, const float *bias, const float *dark,
                   const float *flat)
{
    for (size_t i = 0; i < count; i++) {
        float v = img[i];
        if (bias) v = v - bias[i];
        if (dark) v = v - dark[i];
        if (flat) v = flat[i] == 0.0f ? 0.0f : v / flat[i]; /* cv::divide of OpenCV 3.3.0 (Dockerfile:7): x / 0 = 0 */
        img[i] = v;
    }
}

/* ========================================================================== */
/* a-4  cvtColor BGR2GRAY, stardetector.cpp:110-116                            */

void orc_gray(const float *img, int channels, float *gray, size_t npix)
{
    if (channels == 1) {
        memcpy(gray, img, npix * sizeof(float));
        return;
    }
    for (size_t i = 0; i < npix; i++) {
        const float *p = img + i * channels;
        gray[i] = (p[0] * 0.114f + p[1] * 0.587f) + p[2] * 0.299f;
    }
}

/* ========================================================================== */
/* a-5  generateSkyBackground, stardetector.cpp:158-171                        */

/* cv::resize INTER_LINEAR on CV_32FC1 (stardetector.cpp:161,165). */
void orc_resize_linear(const float *S, int sw, int sh, float *D, int dw, int dh)
{
    int *xo = (int *)malloc(sizeof(int) * 2 * (size_t)dw);
    float *xa = (float *)malloc(sizeof(float) * (size_t)dw);
    float *r0 = (float *)malloc(sizeof(float) * (size_t)dw);
    float *r1 = (float *)malloc(sizeof(float) * (size_t)dw);
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= (float)sx;
        if (sx < 0) { fx = 0.f; sx = 0; }
        if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
        xo[2 * dx] = sx;
        xo[2 * dx + 1] = sx + 1 < sw ? sx + 1 : sw - 1;
        xa[dx] = fx;
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= (float)sy;
        int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        int y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        const float *s0 = S + (size_t)y0 * sw, *s1 = S + (size_t)y1 * sw;
        for (int dx = 0; dx < dw; dx++) {
            float a = xa[dx], ia = 1.f - a;
            r0[dx] = s0[xo[2 * dx]] * ia + s0[xo[2 * dx + 1]] * a;
            r1[dx] = s1[xo[2 * dx]] * ia + s1[xo[2 * dx + 1]] * a;
        }
        float b1 = fy, b0 = 1.f - fy;
        float *d = D + (size_t)dy * dw;
        for (int dx = 0; dx < dw; dx++) d[dx] = r0[dx] * b0 + r1[dx] * b1;
    }
    free(xo); free(xa); free(r0); free(r1);
}

/* cv::medianBlur ksize 5 on CV_32F, BORDER_REPLICATE (stardetector.cpp:164). */
void orc_median5(const float *S, float *D, int w, int h)
{
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            float v[25];
            int n = 0;
            for (int dy = -2; dy <= 2; dy++) {
                int yy = y + dy; yy = yy < 0 ? 0 : (yy >= h ? h - 1 : yy);
                for (int dx = -2; dx <= 2; dx++) {
                    int xx = x + dx; xx = xx < 0 ? 0 : (xx >= w ? w - 1 : xx);
                    v[n++] = S[(size_t)yy * w + xx];
                }
            }
            /* selection up to the 13th smallest */
            for (int i = 0; i <= 12; i++) {
                int m = i;
                for (int j = i + 1; j < 25; j++) if (v[j] < v[m]) m = j;
                float t = v[i]; v[i] = v[m]; v[m] = t;
            }
            D[(size_t)y * w + x] = v[12];
        }
    }
}

/* cv::getGaussianKernel(31, sigma=0 -> 5.0, CV_32F) as cv2 4.13 returns it
 * (bit patterns; symmetric, only taps 0..15 listed). */
static const uint32_t kGaussBits[16] = {
    0x3a68cc99u, 0x3acfe4e8u, 0x3b325fb9u, 0x3b930b60u, 0x3be8ede4u, 0x3c314133u,
    0x3c819931u, 0x3cb61437u, 0x3cf5c7e8u, 0x3d1f615cu, 0x3d4699a0u, 0x3d6dc47au,
    0x3d88bfb5u, 0x3d972180u, 0x3da079edu, 0x3da3b7d6u};

static void gauss_taps(float *k)
{
    for (int i = 0; i < 16; i++) {
        float f;
        memcpy(&f, &kGaussBits[i], 4);
        k[i] = f;
        k[30 - i] = f;
    }
}

static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

/* cv::GaussianBlur 31x31 (stardetector.cpp:168): separable, row pass summed
 * k=0..30 in order, column pass as centre + symmetric pairs, BORDER_REFLECT_101. */
void orc_gauss31(const float *S, float *D, int w, int h)
{
    float k[31];
    gauss_taps(k);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)w * h);
    float *line = (float *)malloc(sizeof(float) * (size_t)(w + 30));
    for (int y = 0; y < h; y++) {
        const float *s = S + (size_t)y * w;
        for (int x = -15; x < w + 15; x++) line[x + 15] = s[reflect101(x, w)];
        float *t = tmp + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float acc = k[0] * line[x];
            for (int j = 1; j < 31; j++) acc = acc + k[j] * line[x + j];
            t[x] = acc;
        }
    }
    const float *rows[31];
    for (int y = 0; y < h; y++) {
        for (int j = -15; j <= 15; j++) rows[j + 15] = tmp + (size_t)reflect101(y + j, h) * w;
        float *d = D + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float acc = k[15] * rows[15][x];
            for (int j = 1; j <= 15; j++) acc = acc + k[15 + j] * (rows[15 + j][x] + rows[15 - j][x]);
            d[x] = acc;
        }
    }
    free(tmp); free(line);
}

void orc_sky_background(const float *gray, float *sky, int w, int h)
{
    int lw = w / 8, lh = h / 8; /* stardetector.cpp:161 integer division */
    float *lo = (float *)malloc(sizeof(float) * (size_t)lw * lh);
    float *med = (float *)malloc(sizeof(float) * (size_t)lw * lh);
    float *up = (float *)malloc(sizeof(float) * (size_t)w * h);
    orc_resize_linear(gray, w, h, lo, lw, lh);
    orc_median5(lo, med, lw, lh);
    orc_resize_linear(med, lw, lh, up, w, h);
    orc_gauss31(up, sky, w, h);
    free(lo); free(med); free(up);
}

/* cv::meanStdDev over Rect(W/20, H/20, int(W*0.9), int(H*0.9)), stardetector.cpp:121-126 */
void orc_roi_stats(const float *img, int w, int h, double *mean, double *sigma)
{
    int x0 = w / 20, y0 = h / 20, rw = (int)(w * 0.9), rh = (int)(h * 0.9);
    double s = 0.0, sq = 0.0;
    for (int y = y0; y < y0 + rh; y++) {
        const float *p = img + (size_t)y * w;
        for (int x = x0; x < x0 + rw; x++) {
            double v = (double)p[x];
            s += v;
            sq += v * v;
        }
    }
    double scale = 1.0 / ((double)rw * (double)rh);
    double m = s * scale;
    double var = sq * scale - m * m;
    if (var < 0.0) var = 0.0;
    *mean = m;
    *sigma = sqrt(var);
}

/* ========================================================================== */
/* a-8/a-9  AdjoiningPixel::deblend / createStar, adjoiningpixel.cpp           */

typedef struct { int x, y; float v; } pix_t;

typedef struct { float v; int i; } sortkey_t;

static int cmp_sortkey(const void *a, const void *b)
{
    const sortkey_t *p = (const sortkey_t *)a, *q = (const sortkey_t *)b;
    if (p->v < q->v) return -1;
    if (p->v > q->v) return 1;
    return p->i < q->i ? -1 : (p->i > q->i ? 1 : 0); /* ties: list order (std::sort is unspecified) */
}

/* adjoiningpixel.cpp:120-150 */
static orc_star create_star(const pix_t *p, int n)
{
    orc_star s;
    s.area = n;
    float peak = 0.f, value = 0.f;
    for (int i = 0; i < n; i++) {
        if (i == 0 || peak < p[i].v) peak = p[i].v; /* adjoiningpixel.cpp:25-31 */
        value += p[i].v;
    }
    s.peak = peak;
    s.value = value;
    sortkey_t *key = (sortkey_t *)malloc(sizeof(sortkey_t) * (size_t)n);
    for (int i = 0; i < n; i++) { key[i].v = p[i].v; key[i].i = i; }
    qsort(key, (size_t)n, sizeof(sortkey_t), cmp_sortkey);
    float xa = 0.f, ya = 0.f, wt = 0.f;
    for (int i = 0; i < n; i++) {
        const pix_t *q = &p[key[i].i];
        xa = (float)((double)xa + (q->x + 0.5) * (double)q->v);
        ya = (float)((double)ya + (q->y + 0.5) * (double)q->v);
        wt = wt + q->v;
    }
    s.x = (int)(xa / wt);
    s.y = (int)(ya / wt);
    free(key);
    return s;
}

/* adjoiningpixel.cpp:104-118: fp32 accumulation in list order, truncated */
static void gravity(const pix_t *p, int n, int *gx, int *gy)
{
    float x = 0.f, y = 0.f, w = 0.f;
    for (int i = 0; i < n; i++) {
        x += p[i].v * (float)p[i].x;
        y += p[i].v * (float)p[i].y;
        w += p[i].v;
    }
    *gx = (int)(x / w);
    *gy = (int)(y / w);
}

/* adjoiningpixel.cpp:41-102.  Net effect of the loop (the "foot" branches only
 * touch a by-value copy): a peeled region becomes a star iff no accepted region
 * touches it (Manhattan distance <= 1) and none shares its truncated centre. */
static int deblend(const pix_t *pix, int n, float step, orc_star *out, int cap)
{
    if (n >= 10000) { /* adjoiningpixel.cpp:45-48 */
        if (cap > 0) out[0] = create_star(pix, n);
        return 1;
    }
    int x0 = pix[0].x, x1 = pix[0].x, y0 = pix[0].y, y1 = pix[0].y;
    for (int i = 1; i < n; i++) {
        if (pix[i].x < x0) x0 = pix[i].x;
        if (pix[i].x > x1) x1 = pix[i].x;
        if (pix[i].y < y0) y0 = pix[i].y;
        if (pix[i].y > y1) y1 = pix[i].y;
    }
    /* one-cell apron so neighbour probes need no bounds tests */
    int bw = x1 - x0 + 3, bh = y1 - y0 + 3;
    int *cell = (int *)malloc(sizeof(int) * (size_t)bw * bh);
    for (size_t i = 0; i < (size_t)bw * bh; i++) cell[i] = -1;
#define CELL(X, Y) cell[(size_t)((Y)-y0 + 1) * bw + ((X)-x0 + 1)]
    for (int i = 0; i < n; i++) CELL(pix[i].x, pix[i].y) = i;

    enum { REMAINING = -1, DROPPED = -2 };
    int *owner = (int *)malloc(sizeof(int) * (size_t)n); /* REMAINING, DROPPED or accepted id */
    for (int i = 0; i < n; i++) owner[i] = REMAINING;
    pix_t *S = (pix_t *)malloc(sizeof(pix_t) * (size_t)n);
    int *Sidx = (int *)malloc(sizeof(int) * (size_t)n);
    int *stack = (int *)malloc(sizeof(int) * ((size_t)4 * n + 4));
    int *acc_gx = (int *)malloc(sizeof(int) * (size_t)n);
    int *acc_gy = (int *)malloc(sizeof(int) * (size_t)n);
    char *touch = (char *)malloc((size_t)n);
    int nacc = 0, left = n, nout = 0;

    while (left > 0) {
        int pk = -1;
        float mx = 0.f;
        for (int i = 0; i < n; i++) { /* getPeak: first maximum among the remaining */
            if (owner[i] != REMAINING) continue;
            if (pk < 0 || mx < pix[i].v) { pk = i; mx = pix[i].v; }
        }
        /* adjoiningpixel.cpp:57: float / float, then double arithmetic, stored to float */
        float T = (float)((double)pix[pk].v - (double)step * sqrt((double)(pix[pk].v / step) + 1.0));
        /* extract(): recursive pre-order up, left, right, down (adjoiningpixel.cpp:172-189) */
        int ns = 0, sp = 0;
        stack[sp++] = pk;
        while (sp > 0) {
            int i = stack[--sp];
            if (i < 0 || owner[i] != REMAINING || !(pix[i].v > T)) continue;
            owner[i] = DROPPED; /* provisional; fixed below when accepted */
            S[ns] = pix[i];
            Sidx[ns++] = i;
            int x = pix[i].x, y = pix[i].y;
            stack[sp++] = CELL(x, y + 1);
            stack[sp++] = CELL(x + 1, y);
            stack[sp++] = CELL(x - 1, y);
            stack[sp++] = CELL(x, y - 1);
        }
        if (ns == 0) break; /* step <= 0 or NaN: the reference would spin forever */
        left -= ns;
        int gx, gy;
        gravity(S, ns, &gx, &gy);
        memset(touch, 0, (size_t)nacc);
        for (int s = 0; s < ns; s++) {
            int x = S[s].x, y = S[s].y;
            int nb[4] = {CELL(x, y - 1), CELL(x - 1, y), CELL(x + 1, y), CELL(x, y + 1)};
            for (int q = 0; q < 4; q++)
                if (nb[q] >= 0 && owner[nb[q]] >= 0) touch[owner[nb[q]]] = 1;
        }
        int blending = 0;
        for (int a = 0; a < nacc; a++)
            if (touch[a] || (acc_gx[a] == gx && acc_gy[a] == gy)) blending++;
        if (blending == 0) {
            for (int s = 0; s < ns; s++) owner[Sidx[s]] = nacc;
            acc_gx[nacc] = gx;
            acc_gy[nacc] = gy;
            nacc++;
            if (nout < cap) out[nout] = create_star(S, ns);
            nout++;
        }
    }
#undef CELL
    free(cell); free(owner); free(S); free(Sidx); free(stack);
    free(acc_gx); free(acc_gy); free(touch);
    return nout;
}

int orc_deblend(const int *px, const int *py, const float *pv, int npix, float step,
                orc_star *out, int cap)
{
    pix_t *p = (pix_t *)malloc(sizeof(pix_t) * (size_t)(npix > 0 ? npix : 1));
    for (int i = 0; i < npix; i++) { p[i].x = px[i]; p[i].y = py[i]; p[i].v = pv[i]; }
    int n = npix > 0 ? deblend(p, npix, step, out, cap) : 0;
    free(p);
    return n;
}

/* a-6  getAdjoiningPixels / detectAdjoiningPixel, stardetector.cpp:192-246 */
int orc_detect(float *img, int w, int h, float threshold, float minpeak, orc_star *out, int cap,
               int *ncomponents)
{
    size_t cap_pix = 1024, cap_stack = 4096;
    pix_t *pix = (pix_t *)malloc(sizeof(pix_t) * cap_pix);
    int *stack = (int *)malloc(sizeof(int) * 2 * cap_stack);
    int nstars = 0, ncomp = 0;
    const float erased = threshold - 1;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            if (!(img[(size_t)y * w + x] > threshold)) continue;
            size_t np = 0, sp = 0;
            stack[0] = x; stack[1] = y; sp = 1;
            float peak = 0.f;
            while (sp > 0) {
                sp--;
                int cx = stack[2 * sp], cy = stack[2 * sp + 1];
                float v = img[(size_t)cy * w + cx];
                if (!(v > threshold)) continue;
                if (np == cap_pix) { cap_pix *= 2; pix = (pix_t *)realloc(pix, sizeof(pix_t) * cap_pix); }
                pix[np].x = cx; pix[np].y = cy; pix[np].v = v;
                if (np == 0 || peak < v) peak = v;
                np++;
                img[(size_t)cy * w + cx] = erased;
                if (sp + 4 > cap_stack) { cap_stack *= 2; stack = (int *)realloc(stack, sizeof(int) * 2 * cap_stack); }
                if (cy - 1 >= 0) { stack[2 * sp] = cx; stack[2 * sp + 1] = cy - 1; sp++; }
                if (cx - 1 >= 0) { stack[2 * sp] = cx - 1; stack[2 * sp + 1] = cy; sp++; }
                if (cx + 1 < w) { stack[2 * sp] = cx + 1; stack[2 * sp + 1] = cy; sp++; }
                if (cy + 1 < h) { stack[2 * sp] = cx; stack[2 * sp + 1] = cy + 1; sp++; }
            }
            if (peak > minpeak) { /* stardetector.cpp:201 */
                ncomp++;
                int room = cap - nstars;
                nstars += deblend(pix, (int)np, threshold, out + (room > 0 ? nstars : 0), room > 0 ? room : 0);
            }
        }
    }
    free(pix); free(stack);
    if (ncomponents) *ncomponents = ncomp;
    return nstars;
}

/* a-4  StarDetectorImpl::getStars, stardetector.cpp:108-147 */
int orc_get_stars(const float *img, int w, int h, int channels, int threshold_coeff,
                  orc_star *out, int cap, orc_detect_info *info)
{
    size_t npix = (size_t)w * h;
    float *gray = (float *)malloc(sizeof(float) * npix);
    float *sky = (float *)malloc(sizeof(float) * npix);
    orc_gray(img, channels, gray, npix);
    orc_sky_background(gray, sky, w, h);
    for (size_t i = 0; i < npix; i++) gray[i] = gray[i] - sky[i];
    free(sky);
    double mean, sigma;
    orc_roi_stats(gray, w, h, &mean, &sigma);
    float threshold = (float)(sigma * threshold_coeff * 1.5);
    float minpeak = (float)(sigma * threshold_coeff * 2.0);
    int ncomp = 0;
    int n = orc_detect(gray, w, h, threshold, minpeak, out, cap, &ncomp);
    free(gray);
    if (info) {
        info->mean = mean; info->sigma = sigma;
        info->threshold = threshold; info->minpeak = minpeak;
        info->ncomponents = ncomp; info->nstars = n;
    }
    return n;
}

/* ========================================================================== */
/* a-10  generateTriangleList / sidesPos, focas.cpp:5-84                       */

static int side_slot(int i, int j, int n)
{
    int lo = i < j ? i : j, hi = i < j ? j : i;
    return lo * (2 * n - lo - 3) / 2 + hi;
}

int orc_triangles(const int *xy, int nstars, int *tri_idx, float *tri_xy)
{
    int n = nstars < ORC_NOBJS ? nstars : ORC_NOBJS;
    float side[ORC_NOBJS * (ORC_NOBJS - 1) / 2 + 1];
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            float dx = (float)(xy[2 * i] - xy[2 * j]);
            float dy = (float)(xy[2 * i + 1] - xy[2 * j + 1]);
            side[side_slot(i, j, n)] = sqrtf(dx * dx + dy * dy);
        }
    int nt = 0;
    for (int i = 0; i < n - 2; i++)
        for (int j = i + 1; j < n - 1; j++)
            for (int k = j + 1; k < n; k++) {
                float dij = side[side_slot(i, j, n)];
                float djk = side[side_slot(j, k, n)];
                float dki = side[side_slot(k, i, n)];
                /* longest/middle/shortest with the reference's exact tie behaviour
                 * (focas.cpp:36-60): the three strict comparisons pick one of six
                 * role assignments (s1 = vertex joining a,b; s2 = a,c; s3 = b,c). */
                float a, b, c;
                int s1, s2, s3;
                if (dki > djk) {
                    if (dki > dij) {
                        a = dki;
                        if (djk > dij) { b = djk; c = dij; s1 = k; s2 = i; s3 = j; }
                        else           { b = dij; c = djk; s1 = i; s2 = k; s3 = j; }
                    } else { a = dij; b = dki; c = djk; s1 = i; s2 = j; s3 = k; }
                } else if (djk > dij) {
                    a = djk;
                    if (dij > dki) { b = dij; c = dki; s1 = j; s2 = k; s3 = i; }
                    else           { b = dki; c = dij; s1 = k; s2 = j; s3 = i; }
                } else { a = dij; b = djk; c = dki; s1 = j; s2 = i; s3 = k; }
                float r = b / a;
                if ((double)r < 0.9) { /* focas.cpp:63 */
                    tri_idx[3 * nt] = s1; tri_idx[3 * nt + 1] = s2; tri_idx[3 * nt + 2] = s3;
                    tri_xy[2 * nt] = r;
                    tri_xy[2 * nt + 1] = c / a;
                    nt++;
                }
            }
    return nt;
}

/* ========================================================================== */
/* a-11  sortTriangles, focas.cpp:146-210: quicksort whose down-scan is guarded
 * by (j > 1) instead of (j > l), so slot 0 is never visited by it.  Emulated
 * literally on (key, payload index). */

static void quirk_sort(float *x, int *id, int l, int r)
{
    while (r > l) {
        float v = x[r];
        int i = l - 1, j = r;
        for (;;) {
            while (x[++i] < v) {}
            while (j > 1 && x[--j] > v) {}
            if (i >= j) break;
            float tx = x[i]; x[i] = x[j]; x[j] = tx;
            int ti = id[i]; id[i] = id[j]; id[j] = ti;
        }
        float tx = x[i]; x[i] = x[r]; x[r] = tx;
        int ti = id[i]; id[i] = id[r]; id[r] = ti;
        quirk_sort(x, id, l, i - 1);
        l = i + 1; /* tail call for the right part */
    }
}

void orc_sort_keys(float *x, int *id, int n)
{
    int *tmp = NULL;
    if (!id) {
        tmp = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
        for (int i = 0; i < n; i++) tmp[i] = i;
        id = tmp;
    }
    quirk_sort(x, id, 0, n - 1);
    free(tmp);
}

/* findMatches, focas.cpp:86-144 (+ binSearchTriangles :212-249, checkTolerance :251-279).
 * B is sorted in place.  matches is 3 rows of 121 (the reference's rows are 120
 * long and slot 120 can be written when k == 120: focas.cpp:125-133). */
void orc_find_matches(const int *idxA, const float *xyA, int nA, int *idxB, float *xyB, int nB,
                      int *table, int *matches, int *k_out)
{
    const int N = ORC_NOBJS, LD = ORC_MAX_MATCH + 1;
    const double tol = 0.002, tol2 = 0.002 * 0.002;
    memset(table, 0, sizeof(int) * N * N);
    memset(matches, 0, sizeof(int) * 3 * LD);

    /* sort B by x, records move as a whole */
    {
        float *kx = (float *)malloc(sizeof(float) * (size_t)(nB > 0 ? nB : 1));
        int *perm = (int *)malloc(sizeof(int) * (size_t)(nB > 0 ? nB : 1));
        int *idx2 = (int *)malloc(sizeof(int) * 3 * (size_t)(nB > 0 ? nB : 1));
        float *xy2 = (float *)malloc(sizeof(float) * 2 * (size_t)(nB > 0 ? nB : 1));
        for (int i = 0; i < nB; i++) { kx[i] = xyB[2 * i]; perm[i] = i; }
        quirk_sort(kx, perm, 0, nB - 1);
        for (int i = 0; i < nB; i++) {
            memcpy(idx2 + 3 * i, idxB + 3 * perm[i], 3 * sizeof(int));
            memcpy(xy2 + 2 * i, xyB + 2 * perm[i], 2 * sizeof(float));
        }
        memcpy(idxB, idx2, sizeof(int) * 3 * (size_t)nB);
        memcpy(xyB, xy2, sizeof(float) * 2 * (size_t)nB);
        free(kx); free(perm); free(idx2); free(xy2);
    }

    for (int t = 0; t < nA; t++) {
        float ax = xyA[2 * t], ay = xyA[2 * t + 1];
        int lo = 0, hi = nB - 1, mid = 0, found = 0, guard = 0;
        while (!found && hi - lo > 1 && guard++ < 64) {
            mid = (lo + hi) / 2;
            float bx = xyB[2 * mid];
            if (fabs((double)(bx - ax)) < tol) found = 1;
            else if ((double)ax < (double)bx - tol) hi = mid;
            else if ((double)ax > (double)bx + tol) lo = mid;
        }
        if (!found) continue;
        int first, last, i;
        for (i = mid; i > 0; i--)
            if (fabs((double)(xyB[2 * i] - ax)) > tol) break;
        first = i;
        for (i = mid; i < nB - 1; i++)
            if (fabs((double)(xyB[2 * i] - ax)) > tol) break;
        last = i;
        for (i = first; i <= last; i++) {
            float dy = ay - xyB[2 * i + 1];
            if ((double)dy < tol) { /* signed test, focas.cpp:261 */
                float dx = ax - xyB[2 * i];
                float d2 = dx * dx + dy * dy;
                if ((double)d2 < tol2) {
                    table[idxA[3 * t] * N + idxB[3 * i]]++;
                    table[idxA[3 * t + 1] * N + idxB[3 * i + 1]]++;
                    table[idxA[3 * t + 2] * N + idxB[3 * i + 2]]++;
                }
            }
        }
    }

    /* focas.cpp:107-139: keep the N best-voted pairs plus ties with the N-th */
    int *mi = matches, *mj = matches + LD, *mn = matches + 2 * LD;
    int k = 0;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            int n = table[i * N + j];
            if (n <= 0) continue;
            int grow = k < N;
            if (!grow && n < mn[N - 1]) continue;
            int l;
            for (l = k; l > 0 && n > mn[l - 1]; l--) { mi[l] = mi[l - 1]; mj[l] = mj[l - 1]; mn[l] = mn[l - 1]; }
            mi[l] = i; mj[l] = j; mn[l] = n;
            if (grow) k++;
            else {
                int lim = k < ORC_MAX_MATCH ? k + 1 : k;
                int nth = mn[N - 1];
                for (k = N; k < lim && nth == mn[k]; k++) {}
            }
        }
    *k_out = k;
}

/* ========================================================================== */
/* a-12  HFTI / H12 / DIFF, 3rdparty/focas/{hfti,h12,diff}.f (Lawson & Hanson).
 * REAL = float, DOUBLE PRECISION = double; indices kept 1-based via macros.    */

/* volatile forces the rounded fp32 sum to exist exactly like the Fortran
 * function call diff(x,y) does (diff.f:4). */
static float lh_diff(float x, float y)
{
    volatile float a = x, b = y;
    return a - b;
}

/* h12.f:29-80.  u(1,j) -> u[(j-1)*iue]; c() is 1-based linear. */
static void lh_h12(int mode, int lpivot, int l1, int m, float *u, int iue, float *up, float *c,
                   int ice, int icv, int ncv)
{
#define U(j) u[(size_t)((j)-1) * iue]
#define C(i) c[(i)-1]
    if (0 >= lpivot || lpivot >= l1 || l1 > m) return;
    float cl = fabsf(U(lpivot));
    if (mode != 2) {
        for (int j = l1; j <= m; j++) { float t = fabsf(U(j)); if (t > cl) cl = t; }
        if (cl <= 0.f) return;
        float clinv = 1.f / cl;
        double t = (double)U(lpivot) * (double)clinv;
        double sm = t * t;
        for (int j = l1; j <= m; j++) { t = (double)U(j) * (double)clinv; sm = sm + t * t; }
        float sm1 = (float)sm;
        cl = cl * sqrtf(sm1);
        if (U(lpivot) > 0.f) cl = -cl;
        *up = U(lpivot) - cl;
        U(lpivot) = cl;
    } else if (cl <= 0.f) return;
    if (ncv <= 0) return;
    double b = (double)*up * (double)U(lpivot);
    if (b >= 0.0) return;
    b = 1.0 / b;
    int i2 = 1 - icv + ice * (lpivot - 1);
    int incr = ice * (l1 - lpivot);
    for (int j = 1; j <= ncv; j++) {
        i2 += icv;
        int i3 = i2 + incr, i4 = i3;
        double sm = (double)C(i2) * (double)*up;
        for (int i = l1; i <= m; i++) { sm = sm + (double)C(i3) * (double)U(i); i3 += ice; }
        if (sm != 0.0) {
            sm = sm * b;
            C(i2) = (float)((double)C(i2) + sm * (double)*up);
            for (int i = l1; i <= m; i++) { C(i4) = (float)((double)C(i4) + sm * (double)U(i)); i4 += ice; }
        }
    }
#undef U
#undef C
}

void orc_hfti(float *a, int mda, int m, int n, float *b, int mdb, int nb, float tau, int *krank,
              float *rnorm, float *h, float *g, int *ip)
{
#define A(i, j) a[(size_t)((j)-1) * mda + ((i)-1)]
#define B(i, j) b[(size_t)((j)-1) * mdb + ((i)-1)]
#define H(i) h[(i)-1]
#define G(i) g[(i)-1]
#define IP(i) ip[(i)-1]
    const float factor = 0.001f;
    int k = 0;
    int ldiag = m < n ? m : n;
    float hmax = 0.f;
    if (ldiag <= 0) { *krank = 0; return; }
    for (int j = 1; j <= ldiag; j++) {
        int lmax = j, recompute = 1;
        if (j != 1) { /* hfti.f:23-30 */
            for (int l = j; l <= n; l++) {
                H(l) = H(l) - A(j - 1, l) * A(j - 1, l);
                if (H(l) > H(lmax)) lmax = l;
            }
            if (lh_diff(hmax + factor * H(lmax), hmax) > 0.f) recompute = 0;
        }
        if (recompute) { /* hfti.f:34-42 */
            lmax = j;
            for (int l = j; l <= n; l++) {
                H(l) = 0.f;
                for (int i = j; i <= m; i++) H(l) = H(l) + A(i, l) * A(i, l);
                if (H(l) > H(lmax)) lmax = l;
            }
            hmax = H(lmax);
        }
        IP(j) = lmax; /* hfti.f:49-57 */
        if (IP(j) != j) {
            for (int i = 1; i <= m; i++) { float t = A(i, j); A(i, j) = A(i, lmax); A(i, lmax) = t; }
            H(lmax) = H(j);
        }
        lh_h12(1, j, j + 1, m, &A(1, j), 1, &H(j), &A(1, j + 1 <= n ? j + 1 : j), 1, mda, n - j);
        lh_h12(2, j, j + 1, m, &A(1, j), 1, &H(j), b, 1, mdb, nb);
    }
    int j;
    for (j = 1; j <= ldiag; j++) /* hfti.f:88-93 */
        if (fabsf(A(j, j)) <= tau) break;
    k = j <= ldiag ? j - 1 : ldiag;
    int kp1 = k + 1;
    for (int jb = 1; jb <= nb; jb++) { /* hfti.f:98-104 */
        float t = 0.f;
        for (int i = kp1; i <= m; i++) t = t + B(i, jb) * B(i, jb);
        rnorm[jb - 1] = sqrtf(t);
    }
    if (k <= 0) {
        for (int jb = 1; jb <= nb; jb++)
            for (int i = 1; i <= n; i++) B(i, jb) = 0.f;
        *krank = k;
        return;
    }
    if (k != n) /* hfti.f:121-128 */
        for (int ii = 1; ii <= k; ii++) {
            int i = kp1 - ii;
            lh_h12(1, i, kp1, n, &A(i, 1), mda, &G(i), a, mda, 1, i - 1);
        }
    for (int jb = 1; jb <= nb; jb++) {
        for (int l = 1; l <= k; l++) { /* hfti.f:136-144 */
            double sm = 0.0;
            int i = kp1 - l;
            for (int jj = i + 1; jj <= k; jj++) sm = sm + (double)A(i, jj) * (double)B(jj, jb);
            float sm1 = (float)sm;
            B(i, jb) = (B(i, jb) - sm1) / A(i, i);
        }
        if (k != n) { /* hfti.f:148-154 */
            for (int jj = kp1; jj <= n; jj++) B(jj, jb) = 0.f;
            for (int i = 1; i <= k; i++)
                lh_h12(2, i, kp1, n, &A(i, 1), mda, &G(i), &B(1, jb), 1, mdb, 1);
        }
        for (int jj = 1; jj <= ldiag; jj++) { /* hfti.f:159-166 */
            int jx = ldiag + 1 - jj;
            if (IP(jx) != jx) {
                int l = IP(jx);
                float t = B(l, jb); B(l, jb) = B(jx, jb); B(jx, jb) = t;
            }
        }
    }
    *krank = k;
#undef A
#undef B
#undef H
#undef G
#undef IP
}

/* findTransform, focas.cpp:281-379.  matches: 3 rows of 121; rows 0/1 are
 * compacted in place by the clipping loop like the reference does. */
int orc_find_transform(int *matches, int m, const int *xy1, const int *xy2, float *M)
{
    enum { LD = ORC_MAX_MATCH + 1, MD = ORC_MAX_MATCH };
    int *mi = matches, *mj = matches + LD;
    float a[3 * MD], b[2 * MD], h[3], g[3], rnorm[2];
    int ip[3], krank;
    memset(M, 0, 6 * sizeof(float));
    if (m < 6) return -1;
    int j = m < 12 ? m : 12;
    for (int i = 0; i < j; i++) {
        a[i] = (float)xy1[2 * mi[i]];
        a[MD + i] = (float)xy1[2 * mi[i] + 1];
        a[2 * MD + i] = 1.f;
        b[i] = (float)xy2[2 * mj[i]];
        b[MD + i] = (float)xy2[2 * mj[i] + 1];
    }
    orc_hfti(a, MD, j, 3, b, MD, 2, 0.1f, &krank, rnorm, h, g, ip);
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++) M[r * 3 + c] = b[r * MD + c];
    for (;;) {
        float *res = a, *sorted = a + MD; /* the reference reuses a[0][], a[1][] */
        for (int i = 0; i < m; i++) {
            float px = (float)xy1[2 * mi[i]], py = (float)xy1[2 * mi[i] + 1];
            float x = M[0] * px + M[1] * py + M[2];
            float y = M[3] * px + M[4] * py + M[5];
            x -= (float)xy2[2 * mj[i]];
            y -= (float)xy2[2 * mj[i] + 1];
            float r2 = x * x + y * y;
            int q;
            for (q = i; q > 0 && r2 < sorted[q - 1]; q--) sorted[q] = sorted[q - 1];
            res[i] = r2;
            sorted[q] = r2;
        }
        int cut = (int)(0.6 * m);
        float lim = 3 * sorted[cut];
        if (lim >= sorted[m - 1]) break;
        int kept = 0;
        for (int i = 0; i < m; i++)
            if (res[i] < lim) {
                int i1 = mi[i], i2 = mj[i];
                mi[kept] = i1; mj[kept] = i2;
                a[kept] = (float)xy1[2 * i1];
                a[MD + kept] = (float)xy1[2 * i1 + 1];
                a[2 * MD + kept] = 1.f;
                b[kept] = (float)xy2[2 * i2];
                b[MD + kept] = (float)xy2[2 * i2 + 1];
                kept++;
            }
        m = kept;
        if (m < 6) return -1;
        orc_hfti(a, MD, m, 3, b, MD, 2, 0.1f, &krank, rnorm, h, g, ip);
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++) M[r * 3 + c] = b[r * MD + c];
    }
    return 0;
}

/* util.cpp:562-569 */
int orc_align(const int *xy1, int n1, const int *xy2, int n2, int *k_out, int *matches,
              int *table, float *M)
{
    int *idxA = (int *)malloc(sizeof(int) * 3 * ORC_MAX_TRI);
    int *idxB = (int *)malloc(sizeof(int) * 3 * ORC_MAX_TRI);
    float *xyA = (float *)malloc(sizeof(float) * 2 * ORC_MAX_TRI);
    float *xyB = (float *)malloc(sizeof(float) * 2 * ORC_MAX_TRI);
    int local_table[ORC_NOBJS * ORC_NOBJS];
    int local_matches[3 * (ORC_MAX_MATCH + 1)];
    if (!table) table = local_table;
    if (!matches) matches = local_matches;
    int nA = orc_triangles(xy1, n1, idxA, xyA);
    int nB = orc_triangles(xy2, n2, idxB, xyB);
    int k = 0;
    orc_find_matches(idxA, xyA, nA, idxB, xyB, nB, table, matches, &k);
    if (k_out) *k_out = k;
    /* findTransform receives `matches` by value (focas.cpp:281): keep the caller's copy intact */
    int work[3 * (ORC_MAX_MATCH + 1)];
    memcpy(work, matches, sizeof(work));
    int err = orc_find_transform(work, k, xy1, xy2, M);
    free(idxA); free(idxB); free(xyA); free(xyB);
    return err;
}

/* ========================================================================== */
/* a-14  cv::warpAffine(INTER_LINEAR | WARP_INVERSE_MAP), BORDER_CONSTANT 0,
 * util.cpp:579: fixed-point source coordinates (10 fractional bits, rounded to
 * 1/32 px), fp32 blend.                                                        */

static inline int round_half_even(double v) { return (int)lrint(v); }
static inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

void orc_warp_affine(const float *src, float *dst, int w, int h, int channels, const float *Mf)
{
    double M[6];
    for (int i = 0; i < 6; i++) M[i] = (double)Mf[i];
    int *adelta = (int *)malloc(sizeof(int) * (size_t)w);
    int *bdelta = (int *)malloc(sizeof(int) * (size_t)w);
    for (int x = 0; x < w; x++) {
        adelta[x] = round_half_even(M[0] * x * 1024);
        bdelta[x] = round_half_even(M[3] * x * 1024);
    }
    const int C = channels;
    for (int y = 0; y < h; y++) {
        int X0 = round_half_even((M[1] * y + M[2]) * 1024) + 16;
        int Y0 = round_half_even((M[4] * y + M[5]) * 1024) + 16;
        float *d = dst + (size_t)y * w * C;
        for (int x = 0; x < w; x++) {
            int X = (X0 + adelta[x]) >> 5, Y = (Y0 + bdelta[x]) >> 5;
            int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
            float fx = (float)(X & 31) * (1.f / 32), fy = (float)(Y & 31) * (1.f / 32);
            float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx;
            float w10 = fy * (1.f - fx), w11 = fy * fx;
            int in_x0 = sx >= 0 && sx < w, in_x1 = sx + 1 >= 0 && sx + 1 < w;
            int in_y0 = sy >= 0 && sy < h, in_y1 = sy + 1 >= 0 && sy + 1 < h;
            for (int c = 0; c < C; c++) {
                float s00 = in_x0 && in_y0 ? src[((size_t)sy * w + sx) * C + c] : 0.f;
                float s01 = in_x1 && in_y0 ? src[((size_t)sy * w + sx + 1) * C + c] : 0.f;
                float s10 = in_x0 && in_y1 ? src[((size_t)(sy + 1) * w + sx) * C + c] : 0.f;
                float s11 = in_x1 && in_y1 ? src[((size_t)(sy + 1) * w + sx + 1) * C + c] : 0.f;
                d[(size_t)x * C + c] = ((s00 * w00 + s01 * w01) + s10 * w10) + s11 * w11;
            }
        }
    }
    free(adelta); free(bdelta);
}

/* ========================================================================== */
/* a-13/a-15  one pass of processConcurrent's loop body, util.cpp:738-751      */

int orc_process_light(const uint16_t *raw, int w, int h, int channels, const float *bias,
                      const float *dark, const float *flat, int threshold_coeff,
                      const float *ref_cal, const int *ref_xy, int ref_n, float *sum,
                      orc_frame_report *rep)
{
    size_t count = (size_t)w * h * channels;
    float *cal = (float *)malloc(sizeof(float) * count);
    orc_u16_to_f32(raw, cal, count);
    orc_calibrate(cal, count, bias, dark, flat);

    const int cap = 1 << 16;
    orc_star *stars = (orc_star *)malloc(sizeof(orc_star) * cap);
    int *xy1_own = NULL;
    if (!ref_xy) { /* util.cpp:559: the reference re-detects the ref stars for every light */
        ref_n = orc_get_stars(ref_cal, w, h, channels, threshold_coeff, stars, cap, NULL);
        if (ref_n > cap) ref_n = cap;
        xy1_own = (int *)malloc(sizeof(int) * 2 * (size_t)(ref_n > 0 ? ref_n : 1));
        for (int i = 0; i < ref_n; i++) { xy1_own[2 * i] = stars[i].x; xy1_own[2 * i + 1] = stars[i].y; }
        ref_xy = xy1_own;
    }
    orc_detect_info info;
    int n2 = orc_get_stars(cal, w, h, channels, threshold_coeff, stars, cap, &info);
    if (n2 > cap) n2 = cap;
    int *xy2 = (int *)malloc(sizeof(int) * 2 * (size_t)(n2 > 0 ? n2 : 1));
    for (int i = 0; i < n2; i++) { xy2[2 * i] = stars[i].x; xy2[2 * i + 1] = stars[i].y; }

    float M[6];
    int k = 0;
    int err = orc_align(ref_xy, ref_n, xy2, n2, &k, NULL, NULL, M);
    if (rep) {
        rep->nstars = n2; rep->k = k; rep->err = err; rep->threshold = info.threshold;
        memcpy(rep->M, M, sizeof(M));
    }
    if (!err) {
        float *warped = (float *)malloc(sizeof(float) * count);
        orc_warp_affine(cal, warped, w, h, channels, M);
        for (size_t i = 0; i < count; i++) sum[i] = sum[i] + warped[i]; /* util.cpp:749 */
        free(warped);
    }
    free(cal); free(stars); free(xy2); free(xy1_own);
    return err;
}

/* a-16  workingImage /= totalValidImages, imagestacker.cpp:361 */
void orc_scale(float *img, size_t count, int total_valid)
{
    const float r = (float)(1.0 / (double)total_valid);
    for (size_t i = 0; i < count; i++) img[i] = img[i] * r;
}

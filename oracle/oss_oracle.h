/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
 *
 * oss_oracle: a plain-C, CPU-only restatement of the arithmetic of libstacker's
 * per-frame pipeline (calibrate -> detect stars -> FOCAS align -> warp -> mean
 * stack).  It exists to CHECK the CUDA path; nothing under openskystacker_b200/
 * may call, link or import it.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Pinning (how we know this restatement is right):
 *   - the OpenCV primitives (convertTo, cvtColor, resize, medianBlur,
 *     GaussianBlur, meanStdDev, warpAffine, subtract/divide/add) are compared
 *     bit-for-bit with cv2 4.13 in plain mode (cv2.setUseOptimized(False)) by
 *     tests/test_oracle_vs_cv2.py, and against committed cv2 outputs in
 *     tests/golden/ (made by tests/golden/make_golden.py);
 *   - deblend / createStar / triangles / quicksort / vote / findTransform / HFTI
 *     are compared with the reference's OWN object code (oracle/_ref/, built from
 *     /root/reference by oracle/Makefile) by tests/test_oracle_vs_ref.py and the
 *     committed golden vectors.
 * The reference's own test-suite holds no numeric vectors for this path
 * (SURVEY.md §4), so these two are the pins.
 *
 * Every function cites the reference file:line it follows; all paths are
 * relative to /root/reference/.
 */
#ifndef OSS_ORACLE_H
#define OSS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NOBJS 40      /* focas.h:50 */
#define ORC_MAX_MATCH 120 /* focas.h:52 */
#define ORC_MAX_TRI 9880  /* C(40,3) */

typedef struct {
    int x, y, area; /* model.h:54-56 (radius is never written; left out) */
    float peak, value;
} orc_star;

typedef struct {
    double mean, sigma;      /* stardetector.cpp:125-126 */
    float threshold, minpeak; /* stardetector.cpp:128-129 */
    int ncomponents;          /* components with peak > minPeak */
    int nstars;
} orc_detect_info;

typedef struct {
    int nstars;      /* stars found in the light */
    int k;           /* matched pairs handed to findTransform */
    int err;         /* 0 ok, -1 alignment failed (focas.cpp:295-299,365-369) */
    float M[6];      /* ref -> target, row-major 2x3 */
    float threshold; /* detection threshold of the light */
} orc_frame_report;

/* ---- a-1  util.cpp:429-463 ------------------------------------------------ */
void orc_u16_to_f32(const uint16_t *src, float *dst, size_t n);
void orc_u8_to_f32(const uint8_t *src, float *dst, size_t n);

/* ---- a-2  util.cpp:584-722 (non-RAW branches) ------------------------------
 * frames: n pointers to f32 frames of `count` values (already through a-1).
 * sub1/sub2: masters subtracted from every frame first (NULL = absent).      */
void orc_stack_mean(const float *const *frames, int n, size_t count, const float *sub1,
                    const float *sub2, float *out);
/* util.cpp:686-688: divide by the mean of channel 0 */
void orc_flat_normalize(float *flat, size_t npix, int channels);

/* ---- a-3  util.cpp:266-274 ------------------------------------------------ */
void orc_calibrate(float *img, size_t count, const float *bias, const float *dark,
                   const float *flat);

/* ---- a-4/a-5  stardetector.cpp:108-171 ------------------------------------ */
void orc_gray(const float *img, int channels, float *gray, size_t npix);
void orc_resize_linear(const float *S, int sw, int sh, float *D, int dw, int dh);
void orc_median5(const float *S, float *D, int w, int h);
void orc_gauss31(const float *S, float *D, int w, int h);
void orc_sky_background(const float *gray, float *sky, int w, int h);
void orc_roi_stats(const float *img, int w, int h, double *mean, double *sigma);

/* ---- a-6/a-8/a-9  stardetector.cpp:192-246, adjoiningpixel.cpp ------------- */
int orc_detect(float *stars_img, int w, int h, float threshold, float minpeak, orc_star *out,
               int cap, int *ncomponents);
int orc_deblend(const int *px, const int *py, const float *pv, int npix, float step,
                orc_star *out, int cap);
/* full getStars: img is f32 HWC with `channels` interleaved channels */
int orc_get_stars(const float *img, int w, int h, int channels, int threshold_coeff,
                  orc_star *out, int cap, orc_detect_info *info);

/* ---- a-10..a-12  focas.cpp, 3rdparty/focas/{hfti,h12,diff}.f ----------------------------- */
int orc_triangles(const int *xy, int nstars, int *tri_idx, float *tri_xy);
void orc_sort_keys(float *x, int *id, int n);
void orc_find_matches(const int *idxA, const float *xyA, int nA, int *idxB, float *xyB, int nB,
                      int *table /*1600*/, int *matches /*3*121*/, int *k);
void orc_hfti(float *a, int mda, int m, int n, float *b, int mdb, int nb, float tau, int *krank,
              float *rnorm, float *h, float *g, int *ip);
int orc_find_transform(int *matches /*3*121*/, int m, const int *xy1, const int *xy2, float *M);
int orc_align(const int *xy1, int n1, const int *xy2, int n2, int *k, int *matches /*3*121*/,
              int *table /*1600 or NULL*/, float *M);

/* ---- a-14  util.cpp:579 (cv::warpAffine INTER_LINEAR|WARP_INVERSE_MAP) ----- */
void orc_warp_affine(const float *src, float *dst, int w, int h, int channels, const float *M);

/* ---- a-13/a-15  util.cpp:557-582, 724-758: one light frame ------------------
 * raw: u16 HWC.  Calibrates, detects, aligns against ref (ref stars either
 * given, or re-detected from ref_cal when ref_xy == NULL as the reference does
 * for every light, util.cpp:559), warps, adds into sum.  Returns err.        */
int orc_process_light(const uint16_t *raw, int w, int h, int channels, const float *bias,
                      const float *dark, const float *flat, int threshold_coeff,
                      const float *ref_cal, const int *ref_xy, int ref_n, float *sum,
                      orc_frame_report *rep);

/* ---- a-16  imagestacker.cpp:361 -------------------------------------------- */
void orc_scale(float *img, size_t count, int total_valid);

#ifdef __cplusplus
}
#endif
#endif

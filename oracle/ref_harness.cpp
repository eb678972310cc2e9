// TEST INFRASTRUCTURE — not part of the product.
//
// extern "C" harness around the REFERENCE'S OWN object code: focas.cpp and
// adjoiningpixel.cpp are compiled unmodified from /root/reference (see
// oracle/Makefile), hfti.o/h12.o/diff.o are the reference's prebuilt Fortran
// objects.  Only the glue that the reference keeps inside OpenCV-dependent files
// is restated here, on a raw float* image:
//   getAdjoiningPixels / detectAdjoiningPixel   libstacker/src/stardetector.cpp:192-246
//   getStars component loop                       libstacker/src/stardetector.cpp:131-146
//   generateAlignedImage's alignment half         libstacker/src/util.cpp:557-572
// Its job is to pin oracle/oss_oracle.c (the plain-C restatement) against the
// reference itself.  Output lands in oracle/_ref/ only.
#include <cstring>
#include <stack>
#include <vector>

#include "adjoiningpixel.h"
#include "focas.h"

using namespace openskystacker;

namespace {

struct RefStar {  // wire format shared with oss_oracle.h's orc_star
    int x, y, area;
    float peak, value;
};

// stardetector.cpp:212-246 restated on a raw buffer (push order up, left, right, down).
AdjoiningPixel floodComponent(float *img, int w, int h, int x, int y, float threshold)
{
    AdjoiningPixel ap;
    std::stack<Pixel> todo;
    todo.push(Pixel(x, y, img[(size_t)y * w + x]));
    while (!todo.empty()) {
        Pixel p = todo.top();
        todo.pop();
        float &cell = img[(size_t)p.y * w + p.x];
        if (cell > threshold) {
            ap.addPixel(p);
            cell = threshold - 1;
            if (p.y - 1 >= 0) todo.push(Pixel(p.x, p.y - 1, img[(size_t)(p.y - 1) * w + p.x]));
            if (p.x - 1 >= 0) todo.push(Pixel(p.x - 1, p.y, img[(size_t)p.y * w + p.x - 1]));
            if (p.x + 1 < w) todo.push(Pixel(p.x + 1, p.y, img[(size_t)p.y * w + p.x + 1]));
            if (p.y + 1 < h) todo.push(Pixel(p.x, p.y + 1, img[(size_t)(p.y + 1) * w + p.x]));
        }
    }
    return ap;
}

std::vector<Star> toStars(const int *xy, int n)
{
    std::vector<Star> v(n);
    for (int i = 0; i < n; i++) {
        v[i].x = xy[2 * i];
        v[i].y = xy[2 * i + 1];
        v[i].area = 0;
        v[i].peak = v[i].radius = v[i].value = 0.f;
    }
    return v;
}

}  // namespace

extern "C" {

// Sky-subtracted image in (destroyed), star list out.  Returns the number of
// stars found (may exceed cap; only cap are written); *ncomponents = kept components.
int ref_detect(float *img, int w, int h, float threshold, float minPeak, RefStar *out, int cap,
               int *ncomponents)
{
    int n = 0, nc = 0;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            if (img[(size_t)y * w + x] > threshold) {
                AdjoiningPixel ap = floodComponent(img, w, h, x, y, threshold);
                if (ap.getPeakValue() > minPeak) {
                    nc++;
                    std::vector<AdjoiningPixel> parts = ap.deblend(threshold);
                    for (size_t j = 0; j < parts.size(); j++) {
                        Star s = parts[j].createStar();
                        if (n < cap) {
                            out[n].x = s.x;
                            out[n].y = s.y;
                            out[n].area = s.area;
                            out[n].peak = s.peak;
                            out[n].value = s.value;
                        }
                        n++;
                    }
                }
            }
        }
    }
    if (ncomponents) *ncomponents = nc;
    return n;
}

// One component given as an explicit pixel list (list order = DFS order).
int ref_deblend(const int *px, const int *py, const float *pv, int npix, float step, RefStar *out,
                int cap)
{
    AdjoiningPixel ap;
    for (int i = 0; i < npix; i++) ap.addPixel(Pixel(px[i], py[i], pv[i]));
    std::vector<AdjoiningPixel> parts = ap.deblend(step);
    int n = 0;
    for (size_t j = 0; j < parts.size(); j++) {
        Star s = parts[j].createStar();
        if (n < cap) {
            out[n].x = s.x;
            out[n].y = s.y;
            out[n].area = s.area;
            out[n].peak = s.peak;
            out[n].value = s.value;
        }
        n++;
    }
    return n;
}

// generateTriangleList (focas.cpp:5-76).  tri_idx: 3 ints per triangle, tri_xy: 2 floats.
int ref_triangles(const int *xy, int nstars, int *tri_idx, float *tri_xy, int cap)
{
    std::vector<Triangle> t = generateTriangleList(toStars(xy, nstars));
    int n = (int)t.size();
    for (int i = 0; i < n && i < cap; i++) {
        tri_idx[3 * i] = t[i].s1;
        tri_idx[3 * i + 1] = t[i].s2;
        tri_idx[3 * i + 2] = t[i].s3;
        tri_xy[2 * i] = t[i].x;
        tri_xy[2 * i + 1] = t[i].y;
    }
    return n;
}

// sortTriangles (focas.cpp:146-210) on bare x keys; ids ride along in s1.
void ref_sort_keys(float *x, int *id, int n)
{
    std::vector<Triangle> t(n);
    for (int i = 0; i < n; i++) {
        t[i].x = x[i];
        t[i].y = 0.f;
        t[i].s1 = id ? id[i] : i;
        t[i].s2 = t[i].s3 = 0;
    }
    sortTriangles(&t, 0, n - 1);
    for (int i = 0; i < n; i++) {
        x[i] = t[i].x;
        if (id) id[i] = t[i].s1;
    }
}

// Alignment half of generateAlignedImage (util.cpp:562-569).
// matches_out: 3*120 ints (rows: ref index, target index, votes).  M: 6 floats row-major.
// Returns err (0 ok, -1 failed) exactly as findTransform sets it.
int ref_align(const int *xy1, int n1, const int *xy2, int n2, int *k_out, int *matches_out,
              float *M)
{
    std::vector<Star> L1 = toStars(xy1, n1), L2 = toStars(xy2, n2);
    std::vector<Triangle> A = generateTriangleList(L1);
    std::vector<Triangle> B = generateTriangleList(L2);
    int k = 0;
    std::vector<std::vector<int> > matches = findMatches(40, &k, A, B);
    if (matches_out)
        for (int r = 0; r < 3; r++)
            for (int i = 0; i < MAX_MATCH; i++) matches_out[r * MAX_MATCH + i] = matches[r][i];
    int err = 0;
    std::vector<std::vector<float> > xf = findTransform(matches, k, L1, L2, &err);
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++) M[i * 3 + j] = xf[i][j];
    if (k_out) *k_out = k;
    return err;
}

// Direct door to the prebuilt Fortran HFTI (hfti.h:11-13).  a: 3 columns of 120,
// b: 2 columns of 120 (column-major, leading dimension 120), both overwritten.
void ref_hfti(float *a, int m, float *b, float tau, int *krank, float *rnorm, int *ip)
{
    int mda = MAX_MATCH, mdb = MAX_MATCH, n = 3, nb = 2;
    float h[3], g[3];
    hfti_((float(*)[MAX_MATCH])a, &mda, &m, &n, (float(*)[MAX_MATCH])b, &mdb, &nb, &tau, krank,
          rnorm, h, g, ip);
}

}  // extern "C"

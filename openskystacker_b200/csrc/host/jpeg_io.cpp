// JPEG input for the ImageStacker facade: the reference's RGB_IMAGE branch reads whatever cv::imread(ANYDEPTH | ANYCOLOR) reads
// (libstacker/src/util.cpp:291-294) and its own test suite stacks a directory of JPEGs (test/testoss.cpp:48).  No JPEG library
// is part of the build image, so this is a decoder of its own for what cameras and converters write: 8-bit Huffman-coded
// baseline / extended-sequential (SOF0, SOF1) and progressive (SOF2) files, grey or three components, any of the usual chroma
// subsamplings, restart intervals.  Pixels have to equal cv::imread's, i.e. libjpeg(-turbo)'s defaults, bit for bit:
//   * the "slow integer" inverse DCT (13-bit constants, two passes with 2 extra bits between them, range-limit table with its
//     1024-entry wrap);
//   * "fancy" (triangle) chroma upsampling for 2:1 ratios — (3 near + far + 1 or 2) >> 2 horizontally, 3:1 vertical blend then the
//     same horizontally with (+ 8 / + 7) >> 4 for 2 x 2 — with the component's edge samples and first / last rows repeated; plain
//     replication for other integral ratios or components of fewer than 3 samples per row;
//   * YCbCr -> RGB with the 16-bit fixed-point tables (1.402, 0.34414, 0.71414, 1.772) and rounding of the library;
//   * the EXIF orientation applied afterwards like cv::imread does (flags without IMREAD_IGNORE_ORIENTATION).
// Output layout is cv::imread's: grey -> 1 channel, colour -> interleaved B, G, R, uint8.
// Not handled (reported as errors): arithmetic coding, lossless / hierarchical processes, 12-bit samples, CMYK / YCCK.
#include <cstring>
#include <fstream>

#include "image_io.hpp"

namespace openskystacker {
namespace {

bool jfail(std::string *err, const std::string &m)
{
    if (err) *err = m;
    return false;
}

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    bool present = false;
    uint8_t bits[17] = {}, vals[256] = {};
    // canonical decoding: codes of length l are mincode[l] .. maxcode[l], their values start at valptr[l]
    int mincode[17] = {}, maxcode[18] = {}, valptr[17] = {};
    // 9-bit lookahead: (length << 8) | value, 0 = longer code
    uint16_t look[512] = {};
    bool build()
    {
        int code = 0, k = 0;
        memset(look, 0, sizeof look);
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k;
            mincode[l] = code;
            for (int i = 0; i < bits[l]; i++, k++, code++) {
                if (k >= 256) return false;
                if (l <= 9) {
                    const int first = code << (9 - l), n = 1 << (9 - l);
                    if (first + n > 512) return false;
                    for (int j = 0; j < n; j++) look[first + j] = (uint16_t)((l << 8) | vals[k]);
                }
            }
            maxcode[l] = bits[l] ? code - 1 : -1;
            if (code > (1 << l)) return false;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        return true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int td = 0, ta = 0;            // tables of the current scan
    int bw = 0, bh = 0;            // blocks per row / column, padded to whole MCUs of an interleaved scan
    int dw = 0, dh = 0;            // real ("downsampled") size in samples
    int pred = 0;                  // DC predictor
    std::vector<int16_t> coef;     // progressive: bw * bh * 64 quantised coefficients, natural order
    std::vector<uint8_t> plane;    // bw * 8 x bh * 8 samples after the inverse DCT
};

struct Decoder {
    const uint8_t *p = nullptr, *end = nullptr;
    uint16_t quant[4][64] = {};
    bool have_q[4] = {};
    HuffTable dc[4], ac[4];
    Component comp[3];
    int ncomp = 0, width = 0, height = 0, hmax = 1, vmax = 1;
    bool progressive = false, have_frame = false, headers_only = false;
    int restart_interval = 0, orientation = 1;
    bool adobe = false;
    int adobe_transform = 0;
    // bit reader
    uint32_t bitbuf = 0;
    int bitcnt = 0;
    bool hit_marker = false;
    int eobrun = 0;
    std::string *err = nullptr;

    // ---- entropy-coded segment: bytes with 0xFF00 un-stuffed, zeros after the segment's end (like the library's "insert EOI")
    void fill()
    {
        while (bitcnt <= 24) {
            uint32_t b = 0;
            if (!hit_marker && p < end) {
                b = *p;
                if (b == 0xff) {
                    if (p + 1 < end && p[1] == 0) p += 2;
                    else { hit_marker = true; b = 0; }
                } else p++;
            }
            bitbuf |= b << (24 - bitcnt);
            bitcnt += 8;
        }
    }
    int getbits(int n)
    {
        if (n == 0) return 0;
        if (bitcnt < n) fill();
        const int v = (int)(bitbuf >> (32 - n));
        bitbuf <<= n;
        bitcnt -= n;
        return v;
    }
    int getbit() { return getbits(1); }
    static int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }
    int decode(const HuffTable &t)
    {
        if (bitcnt < 16) fill();
        const uint16_t e = t.look[bitbuf >> 23];
        if (e) {
            const int l = e >> 8;
            bitbuf <<= l;
            bitcnt -= l;
            return e & 255;
        }
        int code = (int)(bitbuf >> 22), l = 10; // the first 10 bits
        while (l <= 16 && code > t.maxcode[l]) { l++; code = (int)(bitbuf >> (32 - l)); }
        if (l > 16) { bitbuf <<= 16; bitcnt -= 16; return 0; } // corrupt data: the library warns and returns 0
        bitbuf <<= l;
        bitcnt -= l;
        return t.vals[t.valptr[l] + code - t.mincode[l]];
    }
    void reset_entropy()
    {
        bitbuf = 0; bitcnt = 0; hit_marker = false; eobrun = 0;
        for (int i = 0; i < ncomp; i++) comp[i].pred = 0;
    }
    // restart marker between MCUs: skip to it (bits left in the buffer are padding)
    void restart()
    {
        while (p + 1 < end && !(p[0] == 0xff && p[1] >= 0xd0 && p[1] <= 0xd7)) {
            if (p[0] == 0xff && p[1] != 0 && p[1] != 0xff) return reset_entropy(); // some other marker: leave it to the caller
            p++;
        }
        if (p + 1 < end) p += 2;
        reset_entropy();
    }

    // ---- inverse DCT, libjpeg's jpeg_idct_islow arithmetic: dequantised coefficients in natural order -> 8x8 samples
    static uint8_t range_limit(int x)
    {
        const int t = x & 1023; // the library's post-IDCT table and its wrap for wild values
        return (uint8_t)(t < 128 ? t + 128 : t < 512 ? 255 : t < 896 ? 0 : t - 896);
    }
    static void idct(const int16_t *in, const uint16_t *q, uint8_t *out, int stride)
    {
        constexpr int CB = 13, P1 = 2;
        constexpr long F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137,
                       F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
        int ws[64];
        auto descale = [](long x, int n) { return (int)((x + (1L << (n - 1))) >> n); };
        for (int c = 0; c < 8; c++) {
            // the library's shortcut for a column without AC terms (most columns of a dark frame): the full formula gives
            // descale(d0 << CB, CB - P1) = d0 << P1 in every row
            if ((in[8 + c] | in[16 + c] | in[24 + c] | in[32 + c] | in[40 + c] | in[48 + c] | in[56 + c]) == 0) {
                const int dc = (int)((long)in[c] * q[c] * (1L << P1));
                for (int r = 0; r < 8; r++) ws[8 * r + c] = dc;
                continue;
            }
            long d[8];
            for (int r = 0; r < 8; r++) d[r] = (long)in[8 * r + c] * q[8 * r + c];
            long z2 = d[2], z3 = d[6];
            long z1 = (z2 + z3) * F0541;
            long tmp2 = z1 + z3 * (-F1847), tmp3 = z1 + z2 * F0765;
            long tmp0 = (d[0] + d[4]) * (1L << CB), tmp1 = (d[0] - d[4]) * (1L << CB);
            const long t10 = tmp0 + tmp3, t13 = tmp0 - tmp3, t11 = tmp1 + tmp2, t12 = tmp1 - tmp2;
            tmp0 = d[7]; tmp1 = d[5]; tmp2 = d[3]; tmp3 = d[1];
            z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
            long z4 = tmp1 + tmp3;
            const long z5 = (z3 + z4) * F1175;
            tmp0 *= F0298; tmp1 *= F2053; tmp2 *= F3072; tmp3 *= F1501;
            z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            ws[c] = descale(t10 + tmp3, CB - P1);      ws[56 + c] = descale(t10 - tmp3, CB - P1);
            ws[8 + c] = descale(t11 + tmp2, CB - P1);  ws[48 + c] = descale(t11 - tmp2, CB - P1);
            ws[16 + c] = descale(t12 + tmp1, CB - P1); ws[40 + c] = descale(t12 - tmp1, CB - P1);
            ws[24 + c] = descale(t13 + tmp0, CB - P1); ws[32 + c] = descale(t13 - tmp0, CB - P1);
        }
        for (int r = 0; r < 8; r++) {
            const int *w = ws + 8 * r;
            if ((w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) == 0) { // likewise for a row: descale(w0 << CB, CB + P1 + 3) = descale(w0, P1 + 3)
                const uint8_t v = range_limit(descale((long)w[0], P1 + 3));
                memset(out + (size_t)r * stride, v, 8);
                continue;
            }
            long z2 = w[2], z3 = w[6];
            long z1 = (z2 + z3) * F0541;
            long tmp2 = z1 + z3 * (-F1847), tmp3 = z1 + z2 * F0765;
            long tmp0 = ((long)w[0] + w[4]) * (1L << CB), tmp1 = ((long)w[0] - w[4]) * (1L << CB);
            const long t10 = tmp0 + tmp3, t13 = tmp0 - tmp3, t11 = tmp1 + tmp2, t12 = tmp1 - tmp2;
            tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
            z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
            long z4 = tmp1 + tmp3;
            const long z5 = (z3 + z4) * F1175;
            tmp0 *= F0298; tmp1 *= F2053; tmp2 *= F3072; tmp3 *= F1501;
            z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            uint8_t *o = out + (size_t)r * stride;
            constexpr int S = CB + P1 + 3;
            o[0] = range_limit(descale(t10 + tmp3, S)); o[7] = range_limit(descale(t10 - tmp3, S));
            o[1] = range_limit(descale(t11 + tmp2, S)); o[6] = range_limit(descale(t11 - tmp2, S));
            o[2] = range_limit(descale(t12 + tmp1, S)); o[5] = range_limit(descale(t12 - tmp1, S));
            o[3] = range_limit(descale(t13 + tmp0, S)); o[4] = range_limit(descale(t13 - tmp0, S));
        }
    }

    // ---- one block of a sequential scan: quantised coefficients in natural order
    void block_sequential(Component &c, int16_t *blk)
    {
        memset(blk, 0, 64 * sizeof(int16_t));
        const int s = decode(dc[c.td]);
        c.pred += s ? extend(getbits(s), s) : 0;
        blk[0] = (int16_t)c.pred;
        const HuffTable &t = ac[c.ta];
        for (int k = 1; k < 64;) {
            const int rs = decode(t), r = rs >> 4, sz = rs & 15;
            if (sz == 0) {
                if (r != 15) break;
                k += 16;
            } else {
                k += r;
                if (k > 63) break;
                blk[kZigzag[k++]] = (int16_t)extend(getbits(sz), sz);
            }
        }
    }
    // ---- progressive scans (spectral selection ss..se, successive approximation ah -> al), coefficients accumulate in c.coef
    void block_dc_progressive(Component &c, int16_t *blk, int ah, int al)
    {
        if (ah == 0) {
            const int s = decode(dc[c.td]);
            c.pred += s ? extend(getbits(s), s) : 0;
            blk[0] = (int16_t)(c.pred * (1 << al));
        } else if (getbit()) blk[0] = (int16_t)(blk[0] | (1 << al));
    }
    void block_ac_progressive(Component &c, int16_t *blk, int ss, int se, int ah, int al)
    {
        const HuffTable &t = ac[c.ta];
        if (ah == 0) {
            if (eobrun > 0) { eobrun--; return; }
            for (int k = ss; k <= se;) {
                const int rs = decode(t), r = rs >> 4, sz = rs & 15;
                if (sz == 0) {
                    if (r < 15) {
                        eobrun = (1 << r) - 1;
                        if (r) eobrun += getbits(r);
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    if (k > 63) break;
                    blk[kZigzag[k++]] = (int16_t)(extend(getbits(sz), sz) * (1 << al));
                }
            }
            return;
        }
        // refinement: one more bit for the coefficients that are already non-zero, new +-1 << al coefficients between them
        const int p1 = 1 << al, m1 = -(1 << al);
        int k = ss;
        if (eobrun == 0) {
            for (; k <= se;) {
                const int rs = decode(t);
                int r = rs >> 4;
                const int sz = rs & 15;
                int val = 0;
                if (sz) val = getbit() ? p1 : m1;
                else if (r < 15) {
                    eobrun = 1 << r;
                    if (r) eobrun += getbits(r);
                    break;
                }
                for (; k <= se; k++) {
                    int16_t &co = blk[kZigzag[k]];
                    if (co) {
                        if (getbit() && (co & p1) == 0) co = (int16_t)(co + (co >= 0 ? p1 : m1));
                    } else if (--r < 0) break;
                }
                if (val && k <= se) blk[kZigzag[k]] = (int16_t)val;
                k++;
            }
        }
        if (eobrun > 0) {
            for (; k <= se; k++) {
                int16_t &co = blk[kZigzag[k]];
                if (co && getbit() && (co & p1) == 0) co = (int16_t)(co + (co >= 0 ? p1 : m1));
            }
            eobrun--;
        }
    }

    bool fail(const std::string &m) { return jfail(err, m); }

    // ---- markers
    bool parse_dqt(const uint8_t *s, int len)
    {
        while (len > 0) {
            const int pq = s[0] >> 4, tq = s[0] & 15, n = 1 + 64 * (pq ? 2 : 1);
            if (tq > 3 || pq > 1 || len < n) return fail("bad quantisation table");
            for (int i = 0; i < 64; i++) quant[tq][kZigzag[i]] = pq ? (uint16_t)(s[1 + 2 * i] << 8 | s[2 + 2 * i]) : s[1 + i];
            have_q[tq] = true;
            s += n; len -= n;
        }
        return true;
    }
    bool parse_dht(const uint8_t *s, int len)
    {
        while (len > 0) {
            if (len < 17) return fail("bad Huffman table");
            const int tc = s[0] >> 4, th = s[0] & 15;
            if (tc > 1 || th > 3) return fail("bad Huffman table");
            HuffTable &t = tc ? ac[th] : dc[th];
            int n = 0;
            for (int l = 1; l <= 16; l++) { t.bits[l] = s[l]; n += s[l]; }
            if (n > 256 || len < 17 + n) return fail("bad Huffman table");
            memcpy(t.vals, s + 17, n);
            if (!t.build()) return fail("bad Huffman table");
            t.present = true;
            s += 17 + n; len -= 17 + n;
        }
        return true;
    }
    bool parse_sof(const uint8_t *s, int len, int marker)
    {
        if (have_frame) return fail("more than one frame (hierarchical JPEG is not supported)");
        if (len < 6) return fail("bad frame header");
        if (s[0] != 8) return fail("only 8-bit JPEG samples are supported");
        height = s[1] << 8 | s[2]; width = s[3] << 8 | s[4]; ncomp = s[5];
        if (ncomp == 4) return fail("CMYK / YCCK JPEG files are not supported");
        if ((ncomp != 1 && ncomp != 3) || len < 6 + 3 * ncomp || !width || !height) return fail("unsupported frame header");
        progressive = marker == 0xc2;
        hmax = vmax = 1;
        for (int i = 0; i < ncomp; i++) {
            Component &c = comp[i];
            c.id = s[6 + 3 * i]; c.h = s[7 + 3 * i] >> 4; c.v = s[7 + 3 * i] & 15; c.tq = s[8 + 3 * i];
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail("bad sampling factors");
            if (c.h > hmax) hmax = c.h;
            if (c.v > vmax) vmax = c.v;
        }
        if (ncomp == 1) comp[0].h = comp[0].v = hmax = vmax = 1; // a single component is never interleaved
        const int mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
        for (int i = 0; i < ncomp; i++) {
            Component &c = comp[i];
            if (hmax % c.h || vmax % c.v) return fail("non-integral chroma sampling ratios are not supported");
            c.bw = mcux * c.h; c.bh = mcuy * c.v;
            c.dw = (width * c.h + hmax - 1) / hmax; c.dh = (height * c.v + vmax - 1) / vmax;
            if (headers_only) continue; // a probe needs the geometry, not 36 MB of cleared planes per 24 MP file
            c.plane.assign((size_t)c.bw * 8 * c.bh * 8, 0);
            if (progressive) c.coef.assign((size_t)c.bw * c.bh * 64, 0);
        }
        have_frame = true;
        return true;
    }
    void parse_exif(const uint8_t *s, int len)
    {
        if (len < 14 || memcmp(s, "Exif\0\0", 6)) return;
        const uint8_t *t = s + 6;
        const int n = len - 6;
        const bool big = t[0] == 'M';
        if (!((t[0] == 'I' && t[1] == 'I') || (t[0] == 'M' && t[1] == 'M'))) return;
        auto u16 = [&](int o) { return big ? t[o] << 8 | t[o + 1] : t[o] | t[o + 1] << 8; };
        auto u32 = [&](int o) { return big ? (uint32_t)t[o] << 24 | t[o + 1] << 16 | t[o + 2] << 8 | t[o + 3]
                                           : (uint32_t)t[o] | t[o + 1] << 8 | t[o + 2] << 16 | (uint32_t)t[o + 3] << 24; };
        if (u16(2) != 42) return;
        const uint32_t ifd = u32(4);
        if (ifd + 2 > (uint32_t)n) return;
        const int cnt = u16((int)ifd);
        for (int i = 0; i < cnt; i++) {
            const int o = (int)ifd + 2 + 12 * i;
            if (o + 12 > n) return;
            if (u16(o) == 0x0112) { // Orientation, SHORT
                const int v = u16(o + 8);
                if (v >= 1 && v <= 8) orientation = v;
                return;
            }
        }
    }

    // ---- one scan
    bool decode_scan(const uint8_t *s, int len)
    {
        if (!have_frame) return fail("scan before the frame header");
        const int ns = len > 0 ? s[0] : 0;
        if (ns < 1 || ns > ncomp || len < 4 + 2 * ns) return fail("bad scan header");
        Component *sc[3];
        for (int i = 0; i < ns; i++) {
            sc[i] = nullptr;
            for (int j = 0; j < ncomp; j++)
                if (comp[j].id == s[1 + 2 * i]) sc[i] = &comp[j];
            if (!sc[i]) return fail("scan names an unknown component");
            sc[i]->td = s[2 + 2 * i] >> 4; sc[i]->ta = s[2 + 2 * i] & 15;
            if (sc[i]->td > 3 || sc[i]->ta > 3) return fail("bad table selector");
        }
        const int ss = s[1 + 2 * ns], se = s[2 + 2 * ns], ah = s[3 + 2 * ns] >> 4, al = s[3 + 2 * ns] & 15;
        if (progressive) {
            if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss != 0 && ns != 1) || al > 13) return fail("bad progressive scan parameters");
        } else if (ss != 0 || se != 63 || ah != 0 || al != 0) return fail("bad sequential scan parameters");
        for (int i = 0; i < ns; i++) {
            if ((!progressive || ss == 0) && ah == 0 && !dc[sc[i]->td].present) return fail("missing DC Huffman table");
            if ((!progressive || ss != 0) && !ac[sc[i]->ta].present) return fail("missing AC Huffman table");
            if (!have_q[sc[i]->tq]) return fail("missing quantisation table");
        }
        reset_entropy();
        int16_t blk[64];
        // a scan of ONE component covers its real blocks only (ceil(dw / 8) x ceil(dh / 8)); an interleaved scan walks MCUs
        const bool inter = ns > 1;
        const int nx = inter ? (width + 8 * hmax - 1) / (8 * hmax) : (sc[0]->dw + 7) / 8;
        const int ny = inter ? (height + 8 * vmax - 1) / (8 * vmax) : (sc[0]->dh + 7) / 8;
        int todo = restart_interval;
        for (int my = 0; my < ny; my++)
            for (int mx = 0; mx < nx; mx++) {
                if (restart_interval && todo == 0) { restart(); todo = restart_interval; }
                for (int i = 0; i < ns; i++) {
                    Component &c = *sc[i];
                    const int bh = inter ? c.h : 1, bv = inter ? c.v : 1;
                    for (int by = 0; by < bv; by++)
                        for (int bx = 0; bx < bh; bx++) {
                            const int X = mx * bh + bx, Y = my * bv + by;
                            if (progressive) {
                                int16_t *b = c.coef.data() + ((size_t)Y * c.bw + X) * 64;
                                if (ss == 0) block_dc_progressive(c, b, ah, al);
                                else block_ac_progressive(c, b, ss, se, ah, al);
                            } else {
                                block_sequential(c, blk);
                                idct(blk, quant[c.tq], c.plane.data() + ((size_t)Y * 8 * c.bw + X) * 8, c.bw * 8);
                            }
                        }
                }
                todo--;
            }
        return true;
    }

    bool run(const uint8_t *data, size_t n, bool header_only)
    {
        headers_only = header_only;
        p = data; end = data + n;
        if (n < 4 || p[0] != 0xff || p[1] != 0xd8) return fail("not a JPEG file");
        p += 2;
        bool done = false, any_scan = false;
        while (!done) {
            while (p < end && *p != 0xff) p++; // (garbage or the tail of an entropy-coded segment)
            while (p < end && *p == 0xff) p++;
            if (p >= end) break;
            const int m = *p++;
            if (m == 0xd9) break;
            if (m == 0x01 || (m >= 0xd0 && m <= 0xd7) || m == 0) continue;
            if (p + 2 > end) return fail("truncated JPEG file");
            const int len = (p[0] << 8 | p[1]) - 2;
            const uint8_t *s = p + 2;
            if (len < 0 || s + len > end) return fail("truncated JPEG file");
            p = s + len;
            switch (m) {
            case 0xdb: if (!parse_dqt(s, len)) return false; break;
            case 0xc4: if (!parse_dht(s, len)) return false; break;
            case 0xc0: case 0xc1: case 0xc2:
                if (!parse_sof(s, len, m)) return false;
                break;
            case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf:
                return fail("unsupported JPEG process (lossless, hierarchical or arithmetic coding)");
            case 0xdd: if (len >= 2) restart_interval = s[0] << 8 | s[1]; break;
            case 0xe1: parse_exif(s, len); break;
            case 0xee:
                if (len >= 12 && !memcmp(s, "Adobe", 5)) { adobe = true; adobe_transform = s[11]; }
                break;
            case 0xda:
                if (header_only) { done = true; break; }
                if (!decode_scan(s, len)) return false;
                any_scan = true;
                break;
            default: break;
            }
        }
        if (!have_frame) return fail("no frame header in the JPEG file");
        if (header_only) return true;
        if (!any_scan) return fail("no scan in the JPEG file");
        if (progressive)
            for (int i = 0; i < ncomp; i++) {
                Component &c = comp[i];
                if (!have_q[c.tq]) return fail("missing quantisation table");
                for (int Y = 0; Y < c.bh; Y++)
                    for (int X = 0; X < c.bw; X++)
                        idct(c.coef.data() + ((size_t)Y * c.bw + X) * 64, quant[c.tq], c.plane.data() + ((size_t)Y * 8 * c.bw + X) * 8, c.bw * 8);
            }
        return true;
    }

    // ---- upsampling of one component to the full frame: rows of `width` samples
    // fancy = libjpeg's triangle filters (2:1 only, components of at least 3 samples per row); else replication
    void upsample(const Component &c, std::vector<uint8_t> *out) const
    {
        out->resize((size_t)width * height);
        const int hx = hmax / c.h, vx = vmax / c.v, stride = c.bw * 8;
        const uint8_t *src = c.plane.data();
        if (hx == 1 && vx == 1) {
            for (int y = 0; y < height; y++) memcpy(out->data() + (size_t)y * width, src + (size_t)y * stride, width);
            return;
        }
        const bool fancy = c.dw > 2;
        std::vector<uint8_t> line((size_t)c.dw * hx + 8);
        std::vector<int> sums(c.dw);
        auto emit = [&](int y, const uint8_t *l) { // one full-resolution row (the component may be a sample wider than the frame)
            if (y < height) memcpy(out->data() + (size_t)y * width, l, width);
        };
        for (int r = 0; r < c.dh; r++) {
            const uint8_t *in0 = src + (size_t)r * stride;
            if (hx == 2 && vx == 1 && fancy) { // h2v1
                const int n = c.dw;
                line[0] = in0[0];
                line[1] = (uint8_t)((in0[0] * 3 + in0[1] + 2) >> 2);
                for (int i = 1; i < n - 1; i++) {
                    const int v = in0[i] * 3;
                    line[2 * i] = (uint8_t)((v + in0[i - 1] + 1) >> 2);
                    line[2 * i + 1] = (uint8_t)((v + in0[i + 1] + 2) >> 2);
                }
                line[2 * n - 2] = (uint8_t)((in0[n - 1] * 3 + in0[n - 2] + 1) >> 2);
                line[2 * n - 1] = in0[n - 1];
                emit(r, line.data());
            } else if (hx == 2 && vx == 2 && fancy) { // h2v2: 3:1 vertical blend with the nearer neighbour row, then the same horizontally
                for (int v = 0; v < 2; v++) {
                    const int rn = v == 0 ? (r > 0 ? r - 1 : 0) : (r + 1 < c.dh ? r + 1 : c.dh - 1);
                    const uint8_t *in1 = src + (size_t)rn * stride;
                    const int n = c.dw;
                    for (int i = 0; i < n; i++) sums[i] = in0[i] * 3 + in1[i];
                    line[0] = (uint8_t)((sums[0] * 4 + 8) >> 4);
                    line[1] = (uint8_t)((sums[0] * 3 + sums[1] + 7) >> 4);
                    for (int i = 1; i < n - 1; i++) {
                        line[2 * i] = (uint8_t)((sums[i] * 3 + sums[i - 1] + 8) >> 4);
                        line[2 * i + 1] = (uint8_t)((sums[i] * 3 + sums[i + 1] + 7) >> 4);
                    }
                    line[2 * n - 2] = (uint8_t)((sums[n - 1] * 3 + sums[n - 2] + 8) >> 4);
                    line[2 * n - 1] = (uint8_t)((sums[n - 1] * 4 + 7) >> 4);
                    emit(2 * r + v, line.data());
                }
            } else if (hx == 1 && vx == 2) { // h1v2 (the library has no replicating special case for it)
                for (int v = 0; v < 2; v++) {
                    const int rn = v == 0 ? (r > 0 ? r - 1 : 0) : (r + 1 < c.dh ? r + 1 : c.dh - 1);
                    const uint8_t *in1 = src + (size_t)rn * stride;
                    const int bias = v == 0 ? 1 : 2;
                    for (int i = 0; i < c.dw; i++) line[i] = (uint8_t)((in0[i] * 3 + in1[i] + bias) >> 2);
                    emit(2 * r + v, line.data());
                }
            } else { // replication
                for (int i = 0; i < c.dw; i++)
                    for (int k = 0; k < hx; k++) line[(size_t)i * hx + k] = in0[i];
                for (int v = 0; v < vx; v++) emit(r * vx + v, line.data());
            }
        }
    }
};

bool loadWhole(const std::string &path, std::vector<uint8_t> *buf, size_t limit, std::string *err)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return jfail(err, "cannot open " + path);
    f.seekg(0, std::ios::end);
    size_t n = (size_t)f.tellg();
    if (limit && n > limit) n = limit;
    f.seekg(0);
    buf->resize(n);
    f.read((char *)buf->data(), (std::streamsize)n);
    return f.gcount() == (std::streamsize)n ? true : jfail(err, "cannot read " + path);
}

void describe(const Decoder &d, ImageInfo *info)
{
    const bool swap = d.orientation >= 5;
    info->width = swap ? d.height : d.width;
    info->height = swap ? d.width : d.height;
    info->channels = d.ncomp == 1 ? 1 : 3;
    info->sample = SAMPLE_U8;
}

}  // namespace

bool isJpegFile(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    unsigned char b[3] = {0, 0, 0};
    f.read((char *)b, 3);
    return f.gcount() == 3 && b[0] == 0xff && b[1] == 0xd8 && b[2] == 0xff;
}

bool probeJpeg(const std::string &path, ImageInfo *info, std::string *err)
{
    // the frame header follows the APPn / table segments: usually within the first KBs, behind an EXIF thumbnail within the first
    // MB, rarely later (a very large APPn block) — read 64 KB, then 1 MB, then the whole file
    std::vector<uint8_t> buf;
    Decoder d;
    for (size_t limit : {(size_t)1 << 16, (size_t)1 << 20, (size_t)0}) {
        if (!loadWhole(path, &buf, limit, err)) return false;
        d = Decoder();
        d.err = err;
        if (d.run(buf.data(), buf.size(), true)) break;
        if (!limit || buf.size() < limit) return false; // the whole file has been seen
    }
    describe(d, info);
    return true;
}

bool readJpegInto(const std::string &path, const ImageInfo *expect, ImageInfo *info_out, void *dst, std::vector<uint8_t> *alloc,
                  std::string *err)
{
    std::vector<uint8_t> buf;
    if (!loadWhole(path, &buf, 0, err)) return false;
    Decoder d;
    d.err = err;
    if (!d.run(buf.data(), buf.size(), false)) {
        if (err) *err = path + ": " + *err;
        return false;
    }
    ImageInfo info;
    describe(d, &info);
    if (expect && (info.width != expect->width || info.height != expect->height || info.channels != expect->channels ||
                   info.sample != expect->sample))
        return jfail(err, path + ": geometry differs from the reference frame");
    if (info_out) *info_out = info;
    if (alloc) { alloc->resize(info.bytes()); dst = alloc->data(); }
    if (!dst) return jfail(err, "no destination buffer");
    const int W = d.width, H = d.height, C = info.channels;
    // decoded frame in file orientation, then the EXIF transform
    std::vector<uint8_t> tmp;
    uint8_t *img = d.orientation == 1 ? (uint8_t *)dst : (tmp.resize((size_t)W * H * C), tmp.data());
    std::vector<uint8_t> pl[3];
    for (int i = 0; i < d.ncomp; i++) d.upsample(d.comp[i], &pl[i]);
    if (C == 1) memcpy(img, pl[0].data(), (size_t)W * H);
    else {
        const bool rgb = d.adobe ? d.adobe_transform == 0 : (d.comp[0].id == 'R' && d.comp[1].id == 'G' && d.comp[2].id == 'B');
        if (rgb) {
            for (size_t i = 0; i < (size_t)W * H; i++) { img[3 * i] = pl[2][i]; img[3 * i + 1] = pl[1][i]; img[3 * i + 2] = pl[0][i]; }
        } else {
            int crr[256], cbb[256], crg[256], cbg[256];
            for (int i = 0; i < 256; i++) {
                const int x = i - 128;
                crr[i] = (int)((91881L * x + 32768) >> 16);   // 1.40200
                cbb[i] = (int)((116130L * x + 32768) >> 16);  // 1.77200
                crg[i] = (int)(-46802L * x);                  // 0.71414
                cbg[i] = (int)(-22554L * x + 32768);          // 0.34414 (+ the rounding half)
            }
            auto clamp = [](int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); };
            const uint8_t *Y = pl[0].data(), *Cb = pl[1].data(), *Cr = pl[2].data();
            for (size_t i = 0; i < (size_t)W * H; i++) {
                const int y = Y[i], cb = Cb[i], cr = Cr[i];
                img[3 * i + 2] = clamp(y + crr[cr]);
                img[3 * i + 1] = clamp(y + ((cbg[cb] + crg[cr]) >> 16));
                img[3 * i] = clamp(y + cbb[cb]);
            }
        }
    }
    if (d.orientation != 1) { // cv::imread's ExifTransform: flips and a transpose
        uint8_t *o = (uint8_t *)dst;
        const int ow = info.width, oh = info.height;
        for (int y = 0; y < oh; y++)
            for (int x = 0; x < ow; x++) {
                int sx, sy;
                switch (d.orientation) {
                case 2: sx = W - 1 - x; sy = y; break;
                case 3: sx = W - 1 - x; sy = H - 1 - y; break;
                case 4: sx = x; sy = H - 1 - y; break;
                case 5: sx = y; sy = x; break;
                case 6: sx = y; sy = H - 1 - x; break;
                case 7: sx = W - 1 - y; sy = H - 1 - x; break;
                default: sx = W - 1 - y; sy = x; break; // 8
                }
                memcpy(o + ((size_t)y * ow + x) * C, img + ((size_t)sy * W + sx) * C, C);
            }
    }
    return true;
}

}  // namespace openskystacker

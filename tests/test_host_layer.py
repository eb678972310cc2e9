"""Host C++ layer above the C ABI (openskystacker_b200/csrc/host): TIFF codecs, image-list JSON,
openskystacker-cl argument handling (CPU), and the CLI end to end against the oracle (GPU)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import oracle
import openskystacker_b200 as ob
from openskystacker_b200 import synth

cv2 = pytest.importorskip("cv2")
CSRC = ob.binding.CSRC
CLI = os.path.join(CSRC, "openskystacker-cl")


@pytest.fixture(scope="module")
def host():
    ob.build()
    L = C.CDLL(os.path.join(CSRC, "liboss_stacker.so"))
    return L


def read_tiff(L, path):
    whcs = (C.c_int * 4)()
    err = C.create_string_buffer(256)
    assert L.oss_host_probe_image(path.encode(), whcs, err, 256) == 0, err.value
    w, h, c, s = list(whcs)
    dt = [np.uint8, np.uint16, np.float32][s]
    out = np.empty((h, w) if c == 1 else (h, w, c), dt)
    assert L.oss_host_read_image(path.encode(), whcs, C.c_void_p(out.ctypes.data), err, 256) == 0, err.value
    return out


@pytest.mark.parametrize("dtype,channels,compression", [(np.uint16, 3, 5), (np.uint16, 1, 5), (np.uint8, 3, 5),
                                                        (np.uint16, 3, 1), (np.float32, 1, 1), (np.uint16, 3, 32773),
                                                        (np.uint16, 3, 8), (np.uint16, 1, 32946), (np.float32, 3, 8)])
def test_tiff_reader_matches_imread(host, tmp_path, dtype, channels, compression):
    rng = np.random.default_rng(channels + compression)
    shape = (57, 83) if channels == 1 else (57, 83, channels)
    if dtype == np.float32:
        img = rng.random(shape, dtype=np.float32)
    else:
        # smooth + noise so that LZW actually builds long strings
        base = np.linspace(0, np.iinfo(dtype).max * 0.8, shape[1])[None, :, None] if channels == 3 else np.linspace(0, np.iinfo(dtype).max * 0.8, shape[1])[None, :]
        img = (np.broadcast_to(base, shape) + rng.integers(0, 4, shape)).astype(dtype)
    path = str(tmp_path / "a.tif")
    assert cv2.imwrite(path, img, [cv2.IMWRITE_TIFF_COMPRESSION, compression])
    want = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)   # util.cpp:292
    if compression == 32773:  # PackBits: unsupported, must fail loudly rather than decode garbage
        whcs = (C.c_int * 4)()
        err = C.create_string_buffer(256)
        assert host.oss_host_probe_image(path.encode(), whcs, err, 256) != 0 and b"compression" in err.value
        return
    got = read_tiff(host, path)
    assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want)


def test_tiff_reader_randomised_against_imread(host, tmp_path):
    """Shapes from 1x1 up, noise / ramps / constants / few-valued content (short and maximal LZW strings, table resets, KwKwK
    codes), uncompressed / LZW / Deflate, grey and RGB, 8 and 16 bits, files whose directory sits behind the pixel data (libtiff)
    and far beyond the probe's first window: every pixel equals cv::imread's."""
    rng = np.random.default_rng(2024)
    kinds = ("noise", "ramp", "const", "few", "runs")
    for it in range(60):
        h, w = (int(rng.integers(1, 40)), int(rng.integers(1, 40))) if it % 3 else (int(rng.integers(200, 420)), int(rng.integers(200, 420)))
        c = int(rng.choice([1, 3]))
        dtype = [np.uint8, np.uint16][int(rng.integers(0, 2))]
        top = np.iinfo(dtype).max
        shape = (h, w) if c == 1 else (h, w, c)
        kind = kinds[it % len(kinds)]
        if kind == "noise":
            img = rng.integers(0, top + 1, shape)
        elif kind == "ramp":
            img = (np.arange(h)[:, None] * 7 + np.arange(w)[None, :] * 3)
            img = img if c == 1 else np.stack([img, img + 1, img * 2], 2)
        elif kind == "const":
            img = np.full(shape, int(rng.integers(0, top + 1)))
        elif kind == "few":
            img = rng.integers(0, 3, shape) * (top // 2)
        else:
            img = np.repeat(rng.integers(0, top + 1, (h, (w + 15) // 16) + ((c,) if c == 3 else ())), 16, axis=1)[:, :w]
        img = (np.asarray(img) % (top + 1)).astype(dtype)
        comp = int(rng.choice([1, 5, 8]))
        path = str(tmp_path / f"r{it}.tif")
        assert cv2.imwrite(path, img, [cv2.IMWRITE_TIFF_COMPRESSION, comp])
        want = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
        got = read_tiff(host, path)
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), (it, shape, dtype, kind, comp)


def test_tiff_strips_can_be_shared_between_decode_workers(tmp_path):
    """readImageBandInto: band k of n decodes its share of the strips (they are independent: compression and predictor restart
    per strip) — the bands together are the whole frame, for every band count, compressed or not; a PNG is decoded by band 0."""
    src = tmp_path / "bands.cpp"
    src.write_text('''#include "image_io.hpp"
#include <cstdio>
#include <cstring>
#include <vector>
using namespace openskystacker;
int main(int argc, char **argv)
{
    for (int a = 1; a < argc; a++) {
        ImageInfo info; std::string err;
        if (!probeImage(argv[a], &info, &err)) { printf("probe %s: %s\\n", argv[a], err.c_str()); return 1; }
        std::vector<unsigned char> whole(info.bytes()), parts(info.bytes());
        if (!readImageInto(argv[a], info, whole.data(), &err)) { printf("read: %s\\n", err.c_str()); return 1; }
        for (int n : {1, 2, 3, 7, 16, 1000}) {
            memset(parts.data(), 0xA5, parts.size());
            for (int k = n - 1; k >= 0; k--)
                if (!readImageBandInto(argv[a], info, parts.data(), k, n, &err)) { printf("band: %s\\n", err.c_str()); return 1; }
            if (memcmp(whole.data(), parts.data(), whole.size())) { printf("%s: %d bands differ from the whole frame\\n", argv[a], n); return 1; }
        }
    }
    printf("ok\\n");
    return 0;
}
''')
    rng = np.random.default_rng(9)
    files = []
    for i, (shape, comp) in enumerate([((301, 417, 3), 1), ((301, 417, 3), 5), ((97, 64), 8), ((5, 9, 3), 5), ((1, 1), 1)]):
        img = (rng.integers(0, 50, shape) + np.arange(shape[1])[None, :, None][..., 0 if len(shape) == 2 else slice(None)] * 11).astype(np.uint16)
        path = str(tmp_path / f"b{i}.tif")
        assert cv2.imwrite(path, img, [cv2.IMWRITE_TIFF_COMPRESSION, comp])
        files.append(path)
    png = str(tmp_path / "b.png")
    assert cv2.imwrite(png, rng.integers(0, 60000, (33, 47, 3)).astype(np.uint16))
    files.append(png)
    exe = str(tmp_path / "bands")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(CSRC, "host"), "-I", os.path.join(os.path.dirname(CSRC), "..", "include"),
                        str(src), "-o", exe, "-L", CSRC, "-loss_stacker", "-loss_b200", "-Wl,-rpath," + CSRC], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe] + files, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_float_tiff_writer_round_trips_through_cv2(host, tmp_path):
    rng = np.random.default_rng(3)
    for shape in [(31, 45, 3), (31, 45)]:
        img = rng.random(shape, dtype=np.float32)
        path = str(tmp_path / "o.tif")
        c = 1 if img.ndim == 2 else 3
        assert host.oss_host_write_tiff_f32(path.encode(), C.c_void_p(img.ctypes.data), shape[1], shape[0], c) == 0
        back = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
        assert back.dtype == np.float32 and np.array_equal(back, img)
        assert np.array_equal(read_tiff(host, path), img)


def test_png_reader_matches_imread(host, tmp_path):
    """The reference's RGB_IMAGE branch is cv::imread(ANYDEPTH | ANYCOLOR) (util.cpp:291-294): 16-bit PNG is the other
    lossless input besides TIFF.  Every colour type / depth cv2 or PIL can write, against cv2.imread on the same file."""
    rng = np.random.default_rng(17)
    h, w = 41, 67
    ramp = np.linspace(0, 1, w)[None, :]

    def img(dtype, ch):
        top = np.iinfo(dtype).max
        base = np.broadcast_to(ramp * top * 0.7, (h, w))
        if ch > 1:
            base = np.stack([base * g for g in (1.0, 0.8, 0.6, 0.9)[:ch]], 2)
        return (base + rng.integers(0, 40, base.shape)).astype(dtype)
    cases = {"g8": img(np.uint8, 1), "g16": img(np.uint16, 1), "rgb8": img(np.uint8, 3), "rgb16": img(np.uint16, 3),
             "rgba8": img(np.uint8, 4), "rgba16": img(np.uint16, 4)}
    for name, a in cases.items():
        path = str(tmp_path / f"{name}.png")
        assert cv2.imwrite(path, a, [cv2.IMWRITE_PNG_COMPRESSION, 6])
        want = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
        got = read_tiff(host, path)
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), name
    try:
        from PIL import Image
    except Exception:
        return
    # palette, 1-bit, grey + alpha: colour types cv2 does not write
    rgb = cases["rgb8"][:, :, ::-1].copy()
    pil = {"pal": Image.fromarray(rgb).quantize(37), "bit1": Image.fromarray((cases["g8"] > 90).astype(np.uint8) * 255).convert("1"),
           "ga8": Image.fromarray(np.stack([cases["g8"], 255 - cases["g8"]], 2), "LA")}
    for name, im in pil.items():
        path = str(tmp_path / f"{name}.png")
        im.save(path)
        want = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
        got = read_tiff(host, path)
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), name
    # Adam7: interlace a file by hand (zlib + the seven passes, filter 0) and read it back
    import struct
    import zlib
    a = cases["rgb16"]
    be = a[:, :, ::-1].astype(">u2")            # file order RGB, big-endian
    raw = b""
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        sub = be[y0::dy, x0::dx]
        if sub.size:
            raw += b"".join(b"\0" + sub[r].tobytes() for r in range(sub.shape[0]))
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 2, 0, 0, 1)) + \
        chunk(b"IDAT", zlib.compress(raw)[:1000]) + chunk(b"IDAT", zlib.compress(raw)[1000:]) + chunk(b"IEND", b"")
    path = str(tmp_path / "adam7.png")
    open(path, "wb").write(png)
    want = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    assert np.array_equal(want, a)              # cv2 agrees that this is the image
    assert np.array_equal(read_tiff(host, path), a)


def _jpeg_test_image(rng, h, w, ch):
    yy, xx = np.mgrid[0:h, 0:w]
    base = 90 + 70 * np.sin(xx / 7.0) * np.cos(yy / 5.0)
    if ch == 3:
        base = np.stack([base, base[::-1] * 0.8 + 20, 255 - base], 2)
    a = base + rng.normal(0, 12, base.shape)
    for _ in range(6):   # hard edges: large coefficients, values the range limiter has to clip
        y, x = rng.integers(0, h), rng.integers(0, w)
        a[max(y - 1, 0):y + 2, max(x - 1, 0):x + 2] = rng.integers(0, 256)
    return np.clip(a, 0, 255).astype(np.uint8)


def test_jpeg_reader_matches_imread(host, tmp_path):
    """The reference's own test stacks a directory of JPEGs through cv::imread (test/testoss.cpp:48, util.cpp:291-294).  The
    facade's decoder (jpeg_io.cpp: Huffman baseline / progressive, islow inverse DCT, fancy upsampling, fixed-point YCbCr) must
    return cv2.imread's pixels exactly: every chroma subsampling cv2 writes, grey, qualities 5..100, progressive, optimised
    Huffman tables, restart intervals, frame sizes that are not MCU multiples (down to 1 x 1 and 2-sample chroma rows)."""
    rng = np.random.default_rng(1)
    sf = {"411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411, "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
          "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440,
          "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444}
    n = 0
    for (h, w) in [(41, 67), (1, 1), (7, 9), (16, 16), (3, 40), (17, 2), (33, 5), (131, 70)]:
        for ch in (1, 3):
            for q, (prog, opt, rst) in [(5, (0, 0, 0)), (50, (1, 0, 0)), (92, (0, 1, 3)), (100, (1, 1, 2))]:
                for name in (["444"] if ch == 1 else list(sf)):
                    a = _jpeg_test_image(rng, h, w, ch)
                    path = str(tmp_path / f"t{n}.jpg")
                    params = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_PROGRESSIVE, prog, cv2.IMWRITE_JPEG_OPTIMIZE, opt,
                              cv2.IMWRITE_JPEG_RST_INTERVAL, rst] + ([cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf[name]] if ch == 3 else [])
                    assert cv2.imwrite(path, a, params)
                    want = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)   # util.cpp:292
                    got = read_tiff(host, path)
                    assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), (h, w, ch, q, name, prog, opt, rst)
                    n += 1
    assert n == 8 * (4 + 4 * 5)


def test_jpeg_exif_orientation_and_refusals(host, tmp_path):
    """cv::imread applies the EXIF orientation (the reference's flags do not include IMREAD_IGNORE_ORIENTATION): all eight
    values, both TIFF byte orders, width / height swapped in the probe for the transposing ones.  CMYK files and files that
    only look like JPEG are refused with a message instead of decoded wrongly."""
    import struct
    rng = np.random.default_rng(2)
    for ch in (1, 3):
        a = _jpeg_test_image(rng, 37, 58, ch)
        path = str(tmp_path / f"e{ch}.jpg")
        assert cv2.imwrite(path, a, [cv2.IMWRITE_JPEG_QUALITY, 90])
        plain = open(path, "rb").read()
        for o in range(1, 9):
            for big in (False, True):
                e = ">" if big else "<"
                tiff = (b"MM\0*" if big else b"II*\0") + struct.pack(e + "I", 8) + struct.pack(e + "H", 1) + \
                    struct.pack(e + "HHIHH", 0x0112, 3, 1, o, 0) + struct.pack(e + "I", 0)
                seg = b"Exif\0\0" + tiff
                p2 = str(tmp_path / f"e{ch}_{o}{int(big)}.jpg")
                open(p2, "wb").write(plain[:2] + b"\xff\xe1" + struct.pack(">H", len(seg) + 2) + seg + plain[2:])
                want = cv2.imread(p2, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
                got = read_tiff(host, p2)
                assert got.shape == want.shape and np.array_equal(got, want), (ch, o, big)
                assert (got.shape[0] == 58) == (o >= 5)
    whcs = (C.c_int * 4)()
    err = C.create_string_buffer(256)
    try:
        from PIL import Image
        Image.fromarray(_jpeg_test_image(rng, 40, 40, 3)).convert("CMYK").save(str(tmp_path / "cmyk.jpg"))
        assert host.oss_host_probe_image(str(tmp_path / "cmyk.jpg").encode(), whcs, err, 256) != 0 and b"CMYK" in err.value
    except ImportError:
        pass
    open(str(tmp_path / "trunc.jpg"), "wb").write(plain[:len(plain) // 2])   # half a file: decodes (zeros after the cut), like libjpeg
    assert host.oss_host_probe_image(str(tmp_path / "trunc.jpg").encode(), whcs, err, 256) == 0
    open(str(tmp_path / "junk.jpg"), "wb").write(b"\xff\xd8\xff\xe0" + b"\0" * 40)
    assert host.oss_host_probe_image(str(tmp_path / "junk.jpg").encode(), whcs, err, 256) != 0


def test_gui_output_path_writes_mono_as_16_bit(host, tmp_path):
    """MainWindow::finishedStacking (openskystacker/ui/mainwindow.cpp:91-101): a single-channel stack goes through
    convertTo(CV_16U, 65535) before imwrite, a colour one is written as float32."""
    rng = np.random.default_rng(4)
    mono = rng.random((33, 47), dtype=np.float32) * 1.2 - 0.1       # values below 0 and above 1 saturate
    mono[0, :4] = [0.5 / 65535, 1.5 / 65535, 2.5 / 65535, np.nan]    # ties round to even; NaN -> 0
    path = str(tmp_path / "m.tif")
    assert host.oss_host_save_like_gui(path.encode(), C.c_void_p(mono.ctypes.data), 47, 33, 1) == 0
    back = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    v = np.nan_to_num(mono * np.float32(65535.0), nan=0.0)
    want = np.clip(np.rint(v), 0, 65535).astype(np.uint16)
    assert back.dtype == np.uint16 and np.array_equal(back, want) and list(back[0, :3]) == [0, 2, 2]
    rgb = rng.random((12, 9, 3), dtype=np.float32)
    assert host.oss_host_save_like_gui(path.encode(), C.c_void_p(rgb.ctypes.data), 9, 12, 3) == 0
    back = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    assert back.dtype == np.float32 and np.array_equal(back, rgb)


def write_list(tmp_path, entries):
    p = tmp_path / "list.json"
    p.write_text(json.dumps(entries))
    return str(p)


def test_image_list_loader(host, tmp_path):
    img = np.zeros((10, 12, 3), np.uint16)
    for n in ("r.tif", "l1.tif", "d1.tif"):
        cv2.imwrite(str(tmp_path / n), img)
    path = write_list(tmp_path, [{"filename": "r.tif", "type": "ref", "checked": True},
                                 {"filename": "l1.tif", "type": "light", "checked": True},
                                 {"filename": str(tmp_path / "d1.tif"), "type": "dark", "checked": False}])
    buf = C.create_string_buffer(4096)
    assert host.oss_host_load_image_list(path.encode(), buf, 4096) == 0
    rows = [r.split(" ", 5) for r in buf.value.decode().strip().split("\n")]
    assert [r[:5] for r in rows] == [["0", "1", "1", "12", "10"], ["0", "0", "1", "12", "10"], ["1", "0", "0", "12", "10"]]
    assert rows[0][5] == str(tmp_path / "r.tif") and rows[2][5] == str(tmp_path / "d1.tif")   # relative -> JSON dir
    # error codes of util.cpp:480-520
    for text, code in [('{"a": 1}', -1), ('[1]', -2), ('[{"type": "light"}]', -3),
                       ('[{"filename": "r.tif", "type": "lite"}]', -4), ('not json', -1)]:
        (tmp_path / "bad.json").write_text(text)
        assert host.oss_host_load_image_list(str(tmp_path / "bad.json").encode(), buf, 4096) == code


def test_time_string(host):
    buf = C.create_string_buffer(128)
    for s, want in [(1, "1 second"), (42, "42 seconds"), (61, "1 minute and 1 second"), (3725, "1 hour, 2 minutes, and 5 seconds"),
                    (90061, "1 day, 1 hour, 1 minute, and 1 second")]:
        host.oss_host_time_string(s, buf, 128)
        assert buf.value.decode() == want


def test_cli_argument_errors(host, tmp_path):
    run = lambda *a: subprocess.run([CLI, *a], capture_output=True, text=True)
    r = run()
    assert r.returncode == 1 and "Error: Must specify an image list" in r.stdout
    (tmp_path / "bad.json").write_text("{}")
    r = run("-f", str(tmp_path / "bad.json"), "-o", "x")
    assert r.returncode == 1 and "Error -1: Couldn't load image list" in r.stdout
    r = run("-f", str(tmp_path / "bad.json"), "-t", "500")
    assert r.returncode == 1 and "must be an integer between 1 and 100" in r.stdout


def test_cli_reports_the_reference_errors_before_touching_the_gpu(host, tmp_path):
    """Type and size validation (imagestacker.cpp:236-258) happen on file headers, before any device work."""
    a = np.zeros((20, 30, 3), np.uint16)
    cv2.imwrite(str(tmp_path / "ref.tif"), a)
    cv2.imwrite(str(tmp_path / "ok.tif"), a)
    cv2.imwrite(str(tmp_path / "small.tif"), a[:10])
    (tmp_path / "l.fits").write_bytes(b"SIMPLE")
    run = lambda lst: subprocess.run([CLI, "-f", lst, "-o", str(tmp_path / "o")], capture_output=True, text=True)
    lst = write_list(tmp_path, [{"filename": "ref.tif", "type": "ref", "checked": True},
                                {"filename": "l.fits", "type": "light", "checked": True}])
    r = run(lst)
    assert r.returncode == 1 and "Error: Images must be the same type." in r.stdout
    lst = write_list(tmp_path, [{"filename": "ref.tif", "type": "ref", "checked": True},
                                {"filename": "ok.tif", "type": "light", "checked": True},
                                {"filename": "small.tif", "type": "light", "checked": True}])
    r = run(lst)
    assert r.returncode == 1 and "Error: Images must all be the same size." in r.stdout
    # unchecked frames are ignored (cl/cl.cpp:134-135): the bad one no longer matters, the run proceeds to the GPU stage
    lst = write_list(tmp_path, [{"filename": "ref.tif", "type": "ref", "checked": True},
                                {"filename": "ok.tif", "type": "light", "checked": True},
                                {"filename": "small.tif", "type": "light", "checked": False}])
    r = run(lst)
    assert "same size" not in r.stdout


@pytest.mark.gpu
def test_cli_end_to_end_matches_oracle(host, tmp_path):
    w, h = 1000, 700
    ds = synth.dataset(w, h, 4, 3, nstars=120, ndarks=2, nflats=2, nbias=2, seed=1234)
    entries = []

    def put(name, frame, kind):
        cv2.imwrite(str(tmp_path / name), frame)          # cv2 default: LZW-compressed 16-bit TIFF
        entries.append({"filename": name, "type": kind, "checked": True})
    # files hold what imread will return: the dataset arrays ARE the decoded (BGR) frames
    put("ref.tif", ds["ref"], "ref")
    for i, f in enumerate(ds["lights"]):
        put(f"light{i}.tif", f, "light")
    for i, f in enumerate(ds["darks"]):
        put(f"dark{i}.tif", f, "dark")
    for i, f in enumerate(ds["flats"]):
        put(f"flat{i}.tif", f, "flat")
    for i, f in enumerate(ds["bias"]):
        put(f"bias{i}.tif", f, "bias")
    lst = write_list(tmp_path, entries)
    out = str(tmp_path / "stack")
    r = subprocess.run([CLI, "-f", lst, "-o", out, "-t", "20", "-j", "4", "-v"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"File saved to {out}.tif" in r.stdout and "Stacking light frames..." in r.stdout and "Stacking completed in" in r.stdout
    got = cv2.imread(out + ".tif", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    frames = lambda key: [cv2.imread(str(tmp_path / f"{key}{i}.tif"), cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR) for i in range(len(ds[{"light": "lights", "dark": "darks", "flat": "flats", "bias": "bias"}[key]]))]
    want = oracle.stack(cv2.imread(str(tmp_path / "ref.tif"), cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR), frames("light"), 20,
                        bias_frames=frames("bias"), dark_frames=frames("dark"), flat_frames=frames("flat"))
    assert want["total_valid"] == 5
    assert got.shape == want["image"].shape and np.array_equal(got.view(np.uint32), want["image"].view(np.uint32))
    # -s: star count of the (uncalibrated) reference, one integer line (cl/cl.cpp:215-219)
    r = subprocess.run([CLI, "-f", lst, "-s", "-t", "20"], capture_output=True, text=True)
    nref = len(oracle.get_stars(oracle.u16_to_f32(ds["ref"]), 20)[0])
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1] == str(nref)


def _write_stack(tmp_path, ds, ext, put_frame):
    entries = []

    def put(name, frame, kind):
        put_frame(str(tmp_path / (name + ext)), frame)
        entries.append({"filename": name + ext, "type": kind, "checked": True})
    put("ref", ds["ref"], "ref")
    for i, f in enumerate(ds["lights"]):
        put(f"light{i}", f, "light")
    for i, f in enumerate(ds["darks"]):
        put(f"dark{i}", f, "dark")
    for i, f in enumerate(ds["flats"]):
        put(f"flat{i}", f, "flat")
    for i, f in enumerate(ds["bias"]):
        put(f"bias{i}", f, "bias")
    return write_list(tmp_path, entries)


@pytest.mark.gpu
def test_cli_many_png_lights_pipelined_mono16_and_capacity_retry(host, tmp_path):
    """40 mono 16-bit PNG lights of 640x480 = 3 batches of 16 through the two pinned sets (decode of batch b + 1 overlaps copy
    and compute of batch b, handed back by oss_wait_host_frames), 7 darks streamed 16 at a time through oss_master_add;
    --mono16 writes the GUI's 16-bit file.  Then the same stack starting from a detection work space that is too small
    (test hook OSS_STACKER_CAPACITY=64 at threshold 2): the facade must notice OSS_E_CAPACITY, enlarge and re-run — same file."""
    w, h = 640, 480
    ds = synth.dataset(w, h, 40, 1, nstars=80, ndarks=7, seed=5)
    lst = _write_stack(tmp_path, ds, ".png", lambda p, f: cv2.imwrite(p, f))
    want = oracle.stack(ds["ref"], ds["lights"], 20, dark_frames=ds["darks"])
    assert 30 <= want["total_valid"] < 41      # a few lights fail in the reference's own clipping loop: skipped, in order
    out = str(tmp_path / "stack")
    r = subprocess.run([CLI, "-f", lst, "-o", out, "-t", "20", "-j", "6"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = cv2.imread(out + ".tif", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want["image"].view(np.uint32))
    r = subprocess.run([CLI, "-f", lst, "-o", out + "16", "-t", "20", "-j", "2", "--mono16"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got16 = cv2.imread(out + "16.tif", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    assert got16.dtype == np.uint16 and np.array_equal(got16, np.clip(np.rint(want["image"] * np.float32(65535)), 0, 65535).astype(np.uint16))
    # threshold 2 on this sparse field: 3 sigma -> ~0.2 % of the pixels, above 1/64 only with the hook ... use threshold 1 (6.7 %)
    want1 = oracle.stack(ds["ref"], ds["lights"][:6], 1, dark_frames=ds["darks"])
    lst1 = _write_stack(tmp_path, dict(ds, lights=ds["lights"][:6]), ".png", lambda p, f: cv2.imwrite(p, f))
    env = dict(os.environ, OSS_STACKER_CAPACITY="64")
    r = subprocess.run([CLI, "-f", lst1, "-o", out + "t1", "-t", "1", "-j", "2"], capture_output=True, text=True, env=env)
    if want1["total_valid"] < 2:
        assert r.returncode == 1 and "No images could be aligned" in r.stdout
    else:
        assert r.returncode == 0, r.stdout + r.stderr
        got1 = cv2.imread(out + "t1.tif", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
        assert np.array_equal(got1.view(np.uint32), want1["image"].view(np.uint32))


@pytest.mark.gpu
def test_cli_shards_the_lights_over_two_gpus(host, tmp_path):
    """openskystacker-cl --devices 0,1: one host thread per GPU, light k -> GPU k % 2 (util.cpp:738), masters built in row
    slices + NVLink exchange, one NCCL reduce.  Against the oracle within float reassociation (1e-6); identical reports."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    w, h = 1000, 700
    ds = synth.dataset(w, h, 21, 3, nstars=120, ndarks=3, nflats=3, nbias=2, seed=1234)
    lst = _write_stack(tmp_path, ds, ".tif", lambda p, f: cv2.imwrite(p, f))
    want = oracle.stack(ds["ref"], ds["lights"], 20, ds["bias"], ds["darks"], None, ds["flats"])
    out = str(tmp_path / "stack2")
    r = subprocess.run([CLI, "-f", lst, "-o", out, "-t", "20", "-j", "8", "--devices", "0,1", "-v"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r0_stdout = r.stdout
    got = cv2.imread(out + ".tif", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    assert got.shape == want["image"].shape and float(np.max(np.abs(got - want["image"]))) <= 1e-6
    one = str(tmp_path / "stack1")
    r = subprocess.run([CLI, "-f", lst, "-o", one, "-t", "20", "-j", "8", "--devices", "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got_all = cv2.imread(one + ".tif", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    assert float(np.max(np.abs(got_all - want["image"]))) <= 1e-6
    # --balance: the lights come from one shared queue (which GPU takes which batch depends on timing): same stack up to
    # the order of the float additions
    bal = str(tmp_path / "stack_bal")
    rb = subprocess.run([CLI, "-f", lst, "-o", bal, "-t", "20", "-j", "8", "--devices", "0,1", "--balance", "-v"], capture_output=True, text=True)
    assert rb.returncode == 0, rb.stdout + rb.stderr
    got_bal = cv2.imread(bal + ".tif", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    assert float(np.max(np.abs(got_bal - want["image"]))) <= 1e-6
    assert rb.stdout.splitlines()[-1] == r0_stdout.splitlines()[-1].replace("stack2", "stack_bal")  # "File saved to ..."


def test_qt_adapter_compiles_against_stub_headers(tmp_path):
    """csrc/host/imagestacker_qt.hpp (the QObject face of the facade for a Qt build: signals finished / updateProgress /
    processingError / doneDetectingStars, slots process / detectStars, the reference's setters with QString / QStringList,
    imagestacker.h:20-114) is only a header here — Qt and C++ OpenCV are absent from the image.  Syntax- and type-check it
    against minimal stand-ins of the Qt / OpenCV classes it touches (tests/qt_stub), so that it cannot rot unnoticed."""
    import shutil
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "qt_check.cpp"
    src.write_text('''#include "imagestacker_qt.hpp"
int main()
{
    openskystacker::QImageStacker s;
    s.setUseDarks(true); s.setUseFlats(false); s.setUseBias(true); s.setUseDarkFlats(false);
    s.setRefImageFileName(QString::fromStdString("ref.tif"));
    QStringList l; l.push_back(QString::fromStdString("a.tif"));
    s.setTargetImageFileNames(l); s.setDarkFrameFileNames(l); s.setDarkFlatFrameFileNames(l); s.setFlatFrameFileNames(l); s.setBiasFrameFileNames(l);
    s.setSaveFilePath(s.getSaveFilePath());
    s.setWorkingImage(s.getWorkingImage()); s.setRefImage(s.getRefImage()); s.setFinalImage(s.getFinalImage());
    QStringList all = s.getTargetImageFileNames();
    (void)s.getDarkFrameFileNames(); (void)s.getDarkFlatFrameFileNames(); (void)s.getFlatFrameFileNames(); (void)s.getBiasFrameFileNames();
    bool any = s.getUseDarks() || s.getUseDarkFlats() || s.getUseFlats() || s.getUseBias() || s.cancel || all.empty();
    (void)any; (void)s.getRefImageFileName();
    s.engine().setDynamicSharding(true);
    s.process(20, 4);
    s.detectStars(QString::fromStdString("ref.tif"), 20);
    return 0;
}
''')
    r = subprocess.run([cxx, "-std=c++17", "-fsyntax-only", "-Wall", "-DOSS_WITH_QT", "-I", os.path.join(root, "tests", "qt_stub"),
                        "-I", os.path.join(root, "openskystacker_b200", "csrc", "host"), "-I", os.path.join(root, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


# ---- FITS (fitsToMat, libstacker/src/util.cpp:300-346) -----------------------------------------------------------------
def write_fits(path, data, bitpix, bzero=None, extra_axis=None):
    """Minimal primary HDU: 80-byte cards, 2880-byte blocks, big-endian samples."""
    h, w = data.shape[-2:]
    cards = ["SIMPLE  = %20s" % "T", "BITPIX  = %20d" % bitpix, "NAXIS   = %20d" % (3 if extra_axis else 2),
             "NAXIS1  = %20d" % w, "NAXIS2  = %20d" % h]
    if extra_axis:
        cards.append("NAXIS3  = %20d" % extra_axis)
    if bzero is not None:
        cards += ["BZERO   = %20.1f" % bzero, "BSCALE  = %20.1f" % 1.0]
    cards += ["COMMENT made by the test-suite", "EXPTIME = %20.1f / seconds" % 30.0, "END"]
    head = "".join(c.ljust(80) for c in cards).encode()
    head += b" " * (-len(head) % 2880)
    body = np.ascontiguousarray(data).astype({8: ">u1", 16: ">i2", 32: ">i4", -32: ">f4", -64: ">f8"}[bitpix]).tobytes()
    body += b"\0" * (-len(body) % 2880)
    with open(path, "wb") as f:
        f.write(head + body)


@pytest.mark.gpu
def test_cli_stacks_jpeg_lights_like_the_same_pixels_from_tiff(host, tmp_path):
    """A stack whose frames are 8-bit JPEG files (the reference's test data, test/testoss.cpp:48) against the stack of the very
    same pixels — cv2.imread of those files — stored as lossless 8-bit TIFF: identical output file, bit for bit."""
    w, h = 800, 600
    ds = synth.dataset(w, h, 4, 3, nstars=100, ndarks=2, seed=9)
    to8 = lambda f: (f >> 8).astype(np.uint8)
    dj, dt = tmp_path / "jpg", tmp_path / "tif"
    dj.mkdir(); dt.mkdir()
    lists = []
    for d, ext in ((dj, ".jpg"), (dt, ".tif")):
        entries = []
        for name, frame, kind in [("ref", ds["ref"], "ref")] + [(f"light{i}", f, "light") for i, f in enumerate(ds["lights"])] + \
                [(f"dark{i}", f, "dark") for i, f in enumerate(ds["darks"])]:
            if ext == ".jpg":
                assert cv2.imwrite(str(d / (name + ext)), to8(frame), [cv2.IMWRITE_JPEG_QUALITY, 97])
            else:
                px = cv2.imread(str(dj / (name + ".jpg")), cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
                assert px.dtype == np.uint8 and cv2.imwrite(str(d / (name + ext)), px)
            entries.append({"filename": name + ext, "type": kind, "checked": True})
        lists.append(write_list(d, entries))
    outs = []
    for lst, d in zip(lists, (dj, dt)):
        out = str(d / "stack")
        r = subprocess.run([CLI, "-f", lst, "-o", out, "-t", "20", "-j", "3"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(cv2.imread(out + ".tif", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR))
    assert outs[0].dtype == np.float32 and outs[0].shape == (h, w, 3)
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


def test_fits_reader_follows_fitsToMat(host, tmp_path):
    rng = np.random.default_rng(7)
    h, w = 37, 53
    u16 = rng.integers(0, 65536, (h, w)).astype(np.uint16)
    # unsigned 16-bit the FITS way: BITPIX 16, BZERO 32768 (CFITSIO hands back unsigned shorts, util.cpp:316-321)
    write_fits(tmp_path / "u16.fits", u16.astype(np.int32) - 32768, 16, bzero=32768.0)
    got = read_tiff(host, str(tmp_path / "u16.fits"))
    assert got.dtype == np.uint16 and np.array_equal(got, u16)          # no vertical flip: FITS row 0 = Mat row 0
    # plain signed 16-bit: negative values saturate at 0 like CFITSIO's conversion to unsigned short
    s16 = rng.integers(-200, 32768, (h, w)).astype(np.int16)
    write_fits(tmp_path / "s16.fit", s16, 16)
    assert np.array_equal(read_tiff(host, str(tmp_path / "s16.fit")), np.clip(s16, 0, None).astype(np.uint16))
    u8 = rng.integers(0, 256, (h, w)).astype(np.uint8)
    write_fits(tmp_path / "u8.fts", u8, 8)
    got = read_tiff(host, str(tmp_path / "u8.fts"))
    assert got.dtype == np.uint8 and np.array_equal(got, u8)
    f32 = rng.random((h, w), dtype=np.float32)
    write_fits(tmp_path / "f32.fits", f32, -32)
    got = read_tiff(host, str(tmp_path / "f32.fits"))
    assert got.dtype == np.float32 and np.array_equal(got, f32)
    f64 = rng.random((h, w))
    write_fits(tmp_path / "f64.fits", f64, -64)
    assert np.array_equal(read_tiff(host, str(tmp_path / "f64.fits")), f64.astype(np.float32))   # convertTo(CV_32F), util.cpp:338
    # a cube: the reference wraps only the first plane (rows x cols elements of the valarray)
    cube = rng.integers(0, 65536, (3, h, w)).astype(np.uint16)
    write_fits(tmp_path / "cube.fits", cube.astype(np.int32) - 32768, 16, bzero=32768.0, extra_axis=3)
    assert np.array_equal(read_tiff(host, str(tmp_path / "cube.fits")), cube[0])
    # BITPIX 32 is mis-read by the reference (8-byte unsigned long as CV_32SC1): refused, not imitated
    write_fits(tmp_path / "i32.fits", rng.integers(0, 1000, (h, w)), 32)
    whcs, err = (C.c_int * 4)(), C.create_string_buffer(256)
    assert host.oss_host_probe_image(str(tmp_path / "i32.fits").encode(), whcs, err, 256) != 0 and b"BITPIX 32" in err.value
    # truncated data must fail loudly
    blob = (tmp_path / "u16.fits").read_bytes()
    (tmp_path / "cut.fits").write_bytes(blob[:2880 + 100])
    assert host.oss_host_probe_image(str(tmp_path / "cut.fits").encode(), whcs, err, 256) == 0
    out = np.empty((h, w), np.uint16)
    assert host.oss_host_read_image(str(tmp_path / "cut.fits").encode(), whcs, C.c_void_p(out.ctypes.data), err, 256) != 0
    assert b"truncated" in err.value
    # image list records carry the FITS geometry (util.cpp:112-113)
    lst = write_list(tmp_path, [{"filename": "u16.fits", "type": "ref", "checked": True}])
    buf = C.create_string_buffer(4096)
    assert host.oss_host_load_image_list(lst.encode(), buf, 4096) == 0
    assert f" {w} {h} " in buf.value.decode()


@pytest.mark.gpu
def test_cli_stacks_fits_like_the_oracle(host, tmp_path):
    """Mono 16-bit FITS lights + darks through openskystacker-cl: same float32 stack as the oracle on the same arrays,
    and -s counts the stars of the FITS reference (the reference's own testDetectStars case, test/testoss.cpp:37-39)."""
    w, h = 1000, 700
    ds = synth.dataset(w, h, 3, 1, nstars=120, ndarks=2, seed=99)
    entries = []

    def put(name, frame, kind):
        write_fits(tmp_path / name, frame.astype(np.int32) - 32768, 16, bzero=32768.0)
        entries.append({"filename": name, "type": kind, "checked": True})
    put("ref.fits", ds["ref"], "ref")
    for i, f in enumerate(ds["lights"]):
        put(f"light{i}.fits", f, "light")
    for i, f in enumerate(ds["darks"]):
        put(f"dark{i}.fits", f, "dark")
    lst = write_list(tmp_path, entries)
    out = str(tmp_path / "stack")
    r = subprocess.run([CLI, "-f", lst, "-o", out, "-t", "20", "-j", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = cv2.imread(out + ".tif", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    want = oracle.stack(ds["ref"], ds["lights"], 20, dark_frames=ds["darks"])
    assert want["total_valid"] == 4
    assert got.shape == want["image"].shape and np.array_equal(got.view(np.uint32), want["image"].view(np.uint32))
    r = subprocess.run([CLI, "-f", lst, "-s", "-t", "20"], capture_output=True, text=True)
    nref = len(oracle.get_stars(oracle.u16_to_f32(ds["ref"]), 20)[0])
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1] == str(nref)

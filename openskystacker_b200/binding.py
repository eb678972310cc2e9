"""ctypes binding of include/oss_b200.h (tests + bench only; no compute happens here)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
NOBJS, MAX_MATCH = 40, 120

U8, U16, F32 = 0, 1, 2
WARP_AUTO, WARP_TMA, WARP_TILED = 0, 1, 2
HOST, DEVICE = 0, 1
BIAS, DARK, DARKFLAT, FLAT = 0, 1, 2, 3
_SAMPLE = {np.dtype(np.uint8): U8, np.dtype(np.uint16): U16, np.dtype(np.float32): F32}

STAR_DTYPE = np.dtype([("x", "i4"), ("y", "i4"), ("area", "i4"), ("peak", "f4"), ("value", "f4")])


class OssError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"oss_b200 status {status}: {text}")
        self.status = status


class DetectInfo(C.Structure):
    _fields_ = [("mean", C.c_double), ("sigma", C.c_double), ("threshold", C.c_float), ("minpeak", C.c_float),
                ("ncandidates", C.c_int32), ("ncomponents", C.c_int32), ("nstars", C.c_int32), ("overflow", C.c_int32)]


class FrameReport(C.Structure):
    _fields_ = [("nstars", C.c_int32), ("k", C.c_int32), ("err", C.c_int32), ("M", C.c_float * 6),
                ("threshold", C.c_float), ("overflow", C.c_int32)]


def lib_path():
    return os.path.join(CSRC, "liboss_b200.so")


def build(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (csrc/Makefile); cross-compiles without a GPU."""
    so = lib_path()
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".hpp"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "oss_b200.h"))
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", CSRC, "--no-print-directory", "all"], check=True, stdout=subprocess.DEVNULL)
    return so


_lib = None

_SIGS = {
    "oss_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "oss_destroy": (None, [C.c_void_p]),
    "oss_last_error": (C.c_char_p, [C.c_void_p]),
    "oss_status_string": (C.c_char_p, [C.c_int]),
    "oss_version": (C.c_char_p, []),
    "oss_device_count": (C.c_int, []),
    "oss_host_alloc": (C.c_void_p, [C.c_size_t]),
    "oss_host_free": (None, [C.c_void_p]),
    "oss_device_alloc": (C.c_void_p, [C.c_void_p, C.c_size_t]),
    "oss_device_free": (None, [C.c_void_p, C.c_void_p]),
    "oss_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "oss_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "oss_set_geometry": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "oss_set_pipelines": (C.c_int, [C.c_void_p, C.c_int]),
    "oss_build_master": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "oss_set_master": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "oss_get_master": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "oss_clear_masters": (C.c_int, [C.c_void_p]),
    "oss_set_reference": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "oss_add_lights": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "oss_sync": (C.c_int, [C.c_void_p]),
    "oss_wait_host_frames": (C.c_int, [C.c_void_p, C.c_int]),
    "oss_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "oss_reset_stack": (C.c_int, [C.c_void_p]),
    "oss_num_reports": (C.c_int, [C.c_void_p]),
    "oss_get_frame_report": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(FrameReport)]),
    "oss_get_reference_stars": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(DetectInfo)]),
    "oss_get_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "oss_result_device_ptr": (C.c_void_p, [C.c_void_p]),
    "oss_detect_stars": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(DetectInfo)]),
    "oss_detect_stars_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "oss_detect_sweep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "oss_align_lists": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "oss_warp_affine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oss_warp_accumulate": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "oss_set_candidate_capacity": (C.c_int, [C.c_void_p, C.c_int]),
    "oss_set_sm_split": (C.c_int, [C.c_void_p, C.c_int]),
    "oss_get_sm_split": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "oss_comm_unique_id": (C.c_int, [C.c_void_p]),
    "oss_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "oss_comm_broadcast_masters": (C.c_int, [C.c_void_p, C.c_int]),
    "oss_comm_broadcast_master": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "oss_comm_broadcast_reference": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "oss_comm_master_slice": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "oss_comm_allgather_master": (C.c_int, [C.c_void_p, C.c_int]),
    "oss_master_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "oss_master_add": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "oss_master_end": (C.c_int, [C.c_void_p]),
    "oss_comm_reduce": (C.c_int, [C.c_void_p, C.c_int]),
    "oss_mark": (C.c_int, [C.c_void_p, C.c_int]),
    "oss_elapsed_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "oss_synth_light": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int,
                                  C.c_uint, C.c_void_p]),
    "oss_synth_calib": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float, C.c_uint]),
    "oss_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "oss_profile_reset": (C.c_int, [C.c_void_p]),
    "oss_profile_count": (C.c_int, [C.c_void_p]),
    "oss_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "oss_launch_count": (C.c_longlong, [C.c_void_p]),
    "oss_debug_copy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
}
EXPORTS = tuple(_SIGS)


def lib():
    """Load liboss_b200.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        so = lib_path()
        if not os.path.exists(so):
            raise OssError(-1, f"{so} is missing: run __graft_entry__.build() / make -C openskystacker_b200/csrc; "
                               "there is no CPU fallback")
        L = C.CDLL(so)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class Engine:
    """One oss_ctx (one GPU).  Thin, 1:1 with the C ABI."""

    def __init__(self, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        st = self._L.oss_create(C.byref(self._h), device)
        if st:
            raise OssError(st, self._L.oss_status_string(st).decode())
        self.geom = None

    def close(self):
        if self._h:
            self._L.oss_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st:
            raise OssError(st, self._L.oss_last_error(self._h).decode())

    # -- geometry / memory
    def set_geometry(self, width, height, channels, dtype=np.uint16, frames_in_flight=4):
        self._ck(self._L.oss_set_geometry(self._h, width, height, channels, _SAMPLE[np.dtype(dtype)], frames_in_flight))
        self.geom = (height, width, channels, np.dtype(dtype))

    def set_pipelines(self, n):
        self._ck(self._L.oss_set_pipelines(self._h, n))

    def _shape(self):
        h, w, c, _ = self.geom
        return (h, w) if c == 1 else (h, w, c)

    def _frame(self, a):
        h, w, c, dt = self.geom
        a = np.ascontiguousarray(a, dtype=dt)
        assert a.size == h * w * c, f"frame has {a.size} samples, geometry wants {h * w * c}"
        return a

    def device_alloc(self, nbytes):
        p = self._L.oss_device_alloc(self._h, nbytes)
        if not p:
            raise OssError(-8, "device allocation failed")
        return p

    def device_free(self, p):
        self._L.oss_device_free(self._h, p)

    def to_device(self, a):
        a = np.ascontiguousarray(a)
        p = self.device_alloc(a.nbytes)
        self._ck(self._L.oss_memcpy_h2d(self._h, p, _ptr(a), a.nbytes))
        return p

    def from_device(self, p, shape, dtype):
        out = np.empty(shape, dtype)
        self._ck(self._L.oss_memcpy_d2h(self._h, _ptr(out), p, out.nbytes))
        return out

    def _ptrs(self, frames, mem):
        if mem == DEVICE:
            return (C.c_void_p * len(frames))(*frames), None
        keep = [self._frame(f) for f in frames]
        return (C.c_void_p * len(keep))(*[f.ctypes.data for f in keep]), keep

    # -- masters
    def build_master(self, kind, frames, mem=HOST):
        ptrs, keep = self._ptrs(frames, mem)
        self._ck(self._L.oss_build_master(self._h, kind, ptrs, len(frames), mem))

    def set_master(self, kind, data):
        data = np.ascontiguousarray(data, np.float32)
        self._ck(self._L.oss_set_master(self._h, kind, _ptr(data), HOST))

    def get_master(self, kind):
        out = np.empty(self._shape(), np.float32)
        self._ck(self._L.oss_get_master(self._h, kind, _ptr(out)))
        return out

    def clear_masters(self):
        self._ck(self._L.oss_clear_masters(self._h))

    # -- stacking
    def set_reference(self, frame, threshold, mem=HOST):
        if mem == HOST:
            frame = self._frame(frame)
            self._ck(self._L.oss_set_reference(self._h, _ptr(frame), HOST, threshold))
        else:
            self._ck(self._L.oss_set_reference(self._h, C.c_void_p(frame), DEVICE, threshold))

    def add_lights(self, frames, mem=HOST, sync=True):
        ptrs, keep = self._ptrs(frames, mem)
        self._ck(self._L.oss_add_lights(self._h, ptrs, len(frames), mem))
        if sync or keep is not None:
            self.sync()

    def add_light_ptrs(self, ptr_array, n, mem):
        """Pre-built (c_void_p * n) array: the bench's timed call (no per-call Python marshalling)."""
        self._ck(self._L.oss_add_lights(self._h, ptr_array, n, mem))

    def sync(self):
        self._ck(self._L.oss_sync(self._h))

    def wait_host_frames(self, keep_recent=0):
        self._ck(self._L.oss_wait_host_frames(self._h, keep_recent))

    def finish(self, want_image=True):
        out = np.empty(self._shape(), np.float32) if want_image else None
        tv = C.c_int()
        self._ck(self._L.oss_finish(self._h, _ptr(out) if want_image else None, C.byref(tv)))
        return out, tv.value

    def reset_stack(self):
        self._ck(self._L.oss_reset_stack(self._h))

    def get_sum(self):
        out = np.empty(self._shape(), np.float32)
        tv = C.c_int()
        self._ck(self._L.oss_get_sum(self._h, _ptr(out), C.byref(tv)))
        return out, tv.value

    def reports(self):
        n = self._L.oss_num_reports(self._h)
        out = []
        for i in range(n):
            r = FrameReport()
            self._ck(self._L.oss_get_frame_report(self._h, i, C.byref(r)))
            out.append(r)
        return out

    def reference_stars(self, cap=1 << 18):
        buf = np.zeros(cap, STAR_DTYPE)
        n = C.c_int()
        info = DetectInfo()
        self._ck(self._L.oss_get_reference_stars(self._h, _ptr(buf), cap, C.byref(n), C.byref(info)))
        return buf[:min(n.value, cap)].copy(), info

    # -- single operators
    def detect_stars(self, frame, threshold, cap=1 << 18, mem=HOST):
        buf = np.zeros(cap, STAR_DTYPE)
        n = C.c_int()
        info = DetectInfo()
        if mem == HOST:
            frame = self._frame(frame)
            p = _ptr(frame)
        else:
            p = C.c_void_p(frame)
        self._ck(self._L.oss_detect_stars(self._h, p, mem, threshold, _ptr(buf), cap, C.byref(n), C.byref(info)))
        return buf[:min(n.value, cap)].copy(), info

    def detect_stars_batch(self, frames, threshold, mem=HOST):
        ptrs, keep = self._ptrs(frames, mem)
        counts = np.zeros(len(frames), np.int32)
        self._ck(self._L.oss_detect_stars_batch(self._h, ptrs, len(frames), mem, threshold, _ptr(counts)))
        return counts

    def detect_sweep(self, frame, t_first, t_last):
        frame = self._frame(frame)
        counts = np.zeros(t_last - t_first + 1, np.int32)
        self._ck(self._L.oss_detect_sweep(self._h, _ptr(frame), HOST, t_first, t_last, _ptr(counts)))
        return counts

    def align_lists(self, xy_ref, xy_tgt):
        a = np.ascontiguousarray(xy_ref, np.int32).reshape(-1, 2)
        b = np.ascontiguousarray(xy_tgt, np.int32).reshape(-1, 2)
        k, err = C.c_int32(), C.c_int32()
        M = np.zeros(6, np.float32)
        matches = np.zeros((3, MAX_MATCH), np.int32)
        votes = np.zeros((NOBJS, NOBJS), np.int32)
        self._ck(self._L.oss_align_lists(self._h, _ptr(a), len(a), _ptr(b), len(b), C.byref(k), C.byref(err),
                                         _ptr(M), _ptr(matches), _ptr(votes)))
        return dict(err=err.value, k=k.value, M=M.reshape(2, 3), matches=matches, table=votes)

    def warp_affine(self, src, M):
        src = np.ascontiguousarray(src, np.float32)
        M = np.ascontiguousarray(M, np.float32).reshape(6)
        out = np.empty_like(src)
        self._ck(self._L.oss_warp_affine(self._h, _ptr(src), _ptr(out), _ptr(M)))
        return out

    def warp_accumulate(self, srcs, Ms, err=None, which=WARP_AUTO, bands=1):
        """sum over the frames (in order, err == 0 only) of warpAffine(srcs[i], Ms[i]) through the stacking path's kernels."""
        keep = [np.ascontiguousarray(s, np.float32) for s in srcs]
        ptrs = (C.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
        Ms = np.ascontiguousarray(Ms, np.float32).reshape(len(keep), 6)
        e = np.ascontiguousarray(err, np.int32) if err is not None else None
        out = np.empty_like(keep[0])
        self._ck(self._L.oss_warp_accumulate(self._h, ptrs, _ptr(Ms), _ptr(e) if e is not None else None, len(keep),
                                             _ptr(out), which, bands))
        return out

    def set_sm_split(self, fp_sms):
        self._ck(self._L.oss_set_sm_split(self._h, fp_sms))

    def sm_split(self):
        a, b = C.c_int(), C.c_int()
        self._ck(self._L.oss_get_sm_split(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_candidate_capacity(self, one_in):
        self._ck(self._L.oss_set_candidate_capacity(self._h, one_in))

    def debug_plane(self, what, slot=0):
        h, w, c, _ = self.geom
        shape = {0: self._shape(), 1: (h, w), 2: (h, w), 3: (h // 8, w // 8), 4: (h // 8, w // 8)}[what]
        out = np.empty(shape, np.float32)
        self._ck(self._L.oss_debug_copy(self._h, what, slot, _ptr(out)))
        return out

    # -- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        st = lib().oss_comm_unique_id(buf)
        if st:
            raise OssError(st, "ncclGetUniqueId failed")
        return buf.raw

    def comm_init(self, uid, rank, nranks):
        self._ck(self._L.oss_comm_init(self._h, C.c_char_p(uid), rank, nranks))

    def comm_broadcast_masters(self, root=0):
        self._ck(self._L.oss_comm_broadcast_masters(self._h, root))

    def comm_broadcast_master(self, kind, root=0):
        self._ck(self._L.oss_comm_broadcast_master(self._h, kind, root))

    def comm_broadcast_reference(self, threshold, root=0):
        self._ck(self._L.oss_comm_broadcast_reference(self._h, root, threshold))

    def comm_master_slice(self):
        a, b = C.c_int(), C.c_int()
        self._ck(self._L.oss_comm_master_slice(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def comm_allgather_master(self, kind):
        self._ck(self._L.oss_comm_allgather_master(self._h, kind))

    def build_master_slice(self, kind, frames, row0, rows, mem=HOST, chunk=0):
        """oss_master_begin / add / end over rows [row0, row0 + rows) (rows <= 0: whole), `chunk` frames per add (0: all)."""
        ptrs, keep = self._ptrs(frames, mem)
        n = len(frames)
        self._ck(self._L.oss_master_begin(self._h, kind, n, row0, rows))
        step = chunk or n
        for i in range(0, n, step):
            m = min(step, n - i)
            sub = (C.c_void_p * m)(*[ptrs[j] for j in range(i, i + m)])
            self._ck(self._L.oss_master_add(self._h, sub, m, mem))
        self._ck(self._L.oss_master_end(self._h))

    def comm_reduce(self, root=0):
        self._ck(self._L.oss_comm_reduce(self._h, root))

    # -- bench utilities
    def mark(self, which):
        self._ck(self._L.oss_mark(self._h, which))

    def elapsed_ms(self):
        ms = C.c_double()
        self._ck(self._L.oss_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def synth_light(self, dst_dev, stars_xyas, background=0.05, noise=0.002, pedestal_adu=0.0, vignette=False, seed=0,
                    gains=(1.0, 0.9, 0.8)):
        st = np.ascontiguousarray(stars_xyas, np.float32).reshape(-1, 4)
        g = np.asarray(gains, np.float32)
        self._ck(self._L.oss_synth_light(self._h, C.c_void_p(dst_dev), _ptr(st), len(st), background, noise, pedestal_adu,
                                         int(vignette), seed, _ptr(g)))

    def synth_calib(self, dst_dev, level_adu, sigma_adu, flat_profile=False, pedestal_adu=0.0, seed=0):
        self._ck(self._L.oss_synth_calib(self._h, C.c_void_p(dst_dev), level_adu, sigma_adu, int(flat_profile), pedestal_adu, seed))

    def frame_bytes(self):
        h, w, c, dt = self.geom
        return h * w * c * dt.itemsize

    # -- instrumentation
    def profile(self, on):
        self._ck(self._L.oss_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._ck(self._L.oss_profile_reset(self._h))

    def profile_table(self):
        out = {}
        for i in range(self._L.oss_profile_count(self._h)):
            name, ms, n = C.c_char_p(), C.c_double(), C.c_int()
            self._L.oss_profile_get(self._h, i, C.byref(name), C.byref(ms), C.byref(n))
            out[name.value.decode()] = (ms.value, n.value)
        return out

    def launch_count(self):
        return int(self._L.oss_launch_count(self._h))

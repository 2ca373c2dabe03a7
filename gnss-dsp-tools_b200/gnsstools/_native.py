"""ctypes binding of libgnssacq.so (include/gnssacq.h).

There is no CPU fallback: if the CUDA library is missing or no device is present,
every entry point raises. ``Engine`` takes an optional already-loaded ``CDLL`` so the
test-suite can inject its own build; the product path always goes through
:func:`library`, which loads only the in-tree sm_100a library.
"""

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgnssacq.so')

_lib = None
_lock = threading.Lock()


class NativeError(RuntimeError):
    pass


def _declare(lib):
    p, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.gnssacq_last_error.restype = C.c_char_p
    lib.gnssacq_last_error.argtypes = []
    sigs = {
        'gnssacq_create': [C.c_int, C.POINTER(p)],
        'gnssacq_destroy': [p],
        'gnssacq_set_stream': [p, p],
        'gnssacq_set_nco_table': [p, p],
        'gnssacq_set_signal': [p, p, i64],
        'gnssacq_set_signal_device': [p, p, i64],
        'gnssacq_set_replicas': [p, p, i32, i32],
        'gnssacq_set_replicas_device': [p, p, i32, i32],
        'gnssacq_set_replicas_i8': [p, p, i32, i32],
        'gnssacq_set_replicas_i8_device': [p, p, i32, i32],
        'gnssacq_set_profiling': [p, i32],
        'gnssacq_set_option': [p, C.c_char_p, i32],
        'gnssacq_set_schedule': [p, i32, p, i32],
        'gnssacq_get_stage_times': [p, p, p, i32],
        'gnssacq_search': [p, p, i32, i32, i32, i32, i32, p, p, p, p],
        'gnssacq_search_device': [p, p, i32, i32, i32, i32, i32, p],
        'gnssacq_search_grouped': [p, p, i32, i32, i32, i32, i32, i32, p, p, p],
        'gnssacq_nccl_unique_id': [p],
        'gnssacq_nccl_init': [p, p, i32, i32],
        'gnssacq_search_sharded': [p, p, i32, i32, i32, i32, i32, p, p, p],
        'gnssacq_mix': [p, p, i64, dbl, dbl],
        'gnssacq_preprocess': [p, p, i64, dbl, dbl, p, i32, dbl, i64, p],
        'gnssacq_set_replicas_from_chips': [p, p, i32, i32, i32, i32, dbl, dbl, i32, dbl],
        'gnssacq_correlate_bank': [p, p, i32, dbl, i32, i32, i32, p, i32, dbl, p],
        'gnssacq_correlate_epl': [p, p, i32, i32, p, i32, i32, i32, p, i32, p, p, p, p, p],
        'gnssacq_plan_info': [p, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
        'gnssacq_synchronize': [p],
        'gnssacq_kernel_variant': [p],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.gnssacq_launch_count.argtypes = [p]
    lib.gnssacq_launch_count.restype = i64
    return lib


def library():
    """The in-tree CUDA library; raises NativeError if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise NativeError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                                  '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
            _warn_if_stale()
            _lib = _declare(C.CDLL(LIB_PATH))
    return _lib


def _warn_if_stale():
    """A library older than the CUDA sources next to it is a build that was not redone after an
    edit: say so (the sources do not travel with every installation, so this is only a warning)."""
    csrc = os.path.join(os.path.dirname(_HERE), 'csrc')
    if not os.path.isdir(csrc):
        return
    built = os.path.getmtime(LIB_PATH)
    newer = [f for f in os.listdir(csrc) if os.path.getmtime(os.path.join(csrc, f)) > built + 1.0]
    if newer:
        import warnings
        warnings.warn('%s is older than %s: rebuild with `python __graft_entry__.py`' % (LIB_PATH, ', '.join(sorted(newer)[:3])),
                      RuntimeWarning, stacklevel=3)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


RECORD_DTYPE = np.dtype([('metric', np.float32), ('lag', np.int32), ('dbin', np.int32), ('pad', np.int32)])


class Engine:
    """One acquisition engine bound to one GPU (a gnssacq_t handle)."""

    def __init__(self, device=0, lib=None):
        self._lib = _declare(lib) if lib is not None else library()
        self._h = C.c_void_p()
        self._check(self._lib.gnssacq_create(int(device), C.byref(self._h)))
        self.device = device
        from . import nco as _nco
        tab = np.ascontiguousarray(_nco.nco_table, dtype=np.complex128)
        self._check(self._lib.gnssacq_set_nco_table(self._h, _ptr(tab)))
        self.R = 0
        self.N = 0

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.gnssacq_last_error().decode('utf-8', 'replace')
            if rc == -1:
                raise ValueError(msg)
            raise NativeError('gnssacq error %d: %s' % (rc, msg))

    def close(self):
        if self._h:
            self._lib.gnssacq_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration
    def set_stream(self, cuda_stream):
        self._check(self._lib.gnssacq_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def set_signal(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        self._check(self._lib.gnssacq_set_signal(self._h, _ptr(x), x.size))
        self.n_samples = x.size

    def set_signal_device(self, device_ptr, n_samples):
        self._check(self._lib.gnssacq_set_signal_device(self._h, C.c_void_p(int(device_ptr)), int(n_samples)))
        self.n_samples = int(n_samples)

    def set_replicas(self, replicas):
        """R x N time-domain replicas. int8 arrays (+-1 / 0) travel as int8, anything else as float32."""
        rep = np.asarray(replicas)
        if rep.ndim != 2:
            raise ValueError('replicas must be R x N')
        if rep.dtype == np.int8:
            rep = np.ascontiguousarray(rep)
            self._check(self._lib.gnssacq_set_replicas_i8(self._h, _ptr(rep), rep.shape[0], rep.shape[1]))
        else:
            rep = np.ascontiguousarray(rep, dtype=np.float32)
            self._check(self._lib.gnssacq_set_replicas(self._h, _ptr(rep), rep.shape[0], rep.shape[1]))
        self.R, self.N = rep.shape

    def set_replicas_from_chips(self, chips01, n, N, incr, boc=False, chips=0, frac=0):
        """Build the R replicas on the device from R x L chip tables (0/1): what
        `<sig>.code(prn, chips, frac, incr, n)` [* nco.boc11(chips, frac, incr, n)] followed by
        N - n zeros would give (acquire-gps-l1.py:22-24, acquire-gps-l1cd.py:22-26)."""
        c = np.ascontiguousarray(chips01, dtype=np.int8)
        if c.ndim != 2:
            raise ValueError('chips must be R x L')
        R, L = c.shape
        base = float((chips % L) + frac)              # gnsstools/gps/ca.py:108, float64 like numpy
        base2 = float((chips % 2) + frac)             # gnsstools/nco.py:15
        self._check(self._lib.gnssacq_set_replicas_from_chips(self._h, _ptr(c), R, L, int(n), int(N), base, float(incr),
                                                              1 if boc else 0, base2))
        self.R, self.N = R, int(N)

    def correlate_bank(self, chips01, nco_freq, n, n_blocks, block_stride, base, incr):
        """out[h, b] = sum_i x[b*stride+i] * nco(nco_freq,0,n)[i] * (1-2*chips01[floor(base[h,b]+incr*i) mod L])
        on the resident capture; base = (chips % L) + frac per (hypothesis, block), float64."""
        c = np.ascontiguousarray(chips01, dtype=np.int8).ravel()
        base = np.ascontiguousarray(base, dtype=np.float64)
        if base.ndim != 2 or base.shape[1] != n_blocks:
            raise ValueError('base must be H x n_blocks')
        out = np.empty(base.shape, dtype=np.complex128)
        self._check(self._lib.gnssacq_correlate_bank(self._h, _ptr(c), c.size, float(nco_freq), int(n), int(n_blocks),
                                                     int(block_stride), _ptr(base), base.shape[0], float(incr), _ptr(out)))
        return out

    def correlate_epl(self, x, chips01, start, incr, xsel=None, csel=None, mode=0, params=None):
        """Batched tracking correlators: out[h] = <sig>.correlate(x[xsel[h]], ., start[h], 0, incr[h], chips01[csel[h]], ...)
        of the reference (mode 0 plain, 1 two-level sub-chip pattern, 2 CBOC, 3 TMBOC; see gnssacq.h).
        x: (nx, n) or (n,) complex64 blocks, chips01: (ncodes, L) or (L,) 0/1 chips."""
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.complex64)
        c = np.ascontiguousarray(np.atleast_2d(chips01), dtype=np.int8)
        start = np.ascontiguousarray(np.atleast_1d(start), dtype=np.float64)
        H = start.size
        incr = np.ascontiguousarray(np.broadcast_to(np.asarray(incr, dtype=np.float64), (H,)))
        xsel = np.zeros(H, np.int32) if xsel is None else np.ascontiguousarray(xsel, dtype=np.int32)
        csel = np.zeros(H, np.int32) if csel is None else np.ascontiguousarray(csel, dtype=np.int32)
        prm = None if params is None else np.ascontiguousarray(params, dtype=np.float64)
        if prm is not None and prm.size < (37 if mode == 3 else 4):
            raise ValueError('params: {sub0, sub1, a1, a6[, pattern x 33]}')
        out = np.empty(H, np.complex128)
        self._check(self._lib.gnssacq_correlate_epl(self._h, _ptr(x), x.shape[0], x.shape[1], _ptr(c), c.shape[0], c.shape[1], int(mode),
                                                    _ptr(prm) if prm is not None else None, H, _ptr(xsel), _ptr(csel), _ptr(start), _ptr(incr),
                                                    _ptr(out)))
        return out

    def set_replicas_i8_device(self, device_ptr, R, N):
        self._check(self._lib.gnssacq_set_replicas_i8_device(self._h, C.c_void_p(int(device_ptr)), int(R), int(N)))
        self.R, self.N = int(R), int(N)

    def set_replicas_device(self, device_ptr, R, N):
        self._check(self._lib.gnssacq_set_replicas_device(self._h, C.c_void_p(int(device_ptr)), int(R), int(N)))
        self.R, self.N = int(R), int(N)

    def set_option(self, name, value):
        self._check(self._lib.gnssacq_set_option(self._h, name.encode(), int(value)))

    def set_schedule(self, which, radices):
        r = np.ascontiguousarray(radices, dtype=np.int32)
        self._check(self._lib.gnssacq_set_schedule(self._h, int(which), _ptr(r), r.size))

    def set_profiling(self, on):
        self._check(self._lib.gnssacq_set_profiling(self._h, int(bool(on))))

    def stage_times(self, reset=True):
        """{'fwd','corr_rows','corr','finalize'} -> (milliseconds, launches) since the last reset."""
        ms = np.zeros(4, np.float64)
        nl = np.zeros(4, np.int64)
        self._check(self._lib.gnssacq_get_stage_times(self._h, _ptr(ms), _ptr(nl), int(bool(reset))))
        names = ('fwd', 'corr_rows', 'corr', 'finalize')
        return {k: (float(ms[i]), int(nl[i])) for i, k in enumerate(names)}

    # -- search
    def search(self, nco_freq, block_stride, n_blocks, normalize, n_lags=0, dump=False):
        f = np.ascontiguousarray(nco_freq, dtype=np.float64)
        D = f.size
        metric = np.empty(self.R, np.float32)
        lag = np.empty(self.R, np.int32)
        dbin = np.empty(self.R, np.int32)
        q = np.empty((self.R, D, self.N), np.float32) if dump else None
        self._check(self._lib.gnssacq_search(self._h, _ptr(f), D, int(block_stride), int(n_blocks), int(bool(normalize)),
                                             int(n_lags), _ptr(metric), _ptr(lag), _ptr(dbin),
                                             _ptr(q) if dump else None))
        return (metric, lag, dbin, q) if dump else (metric, lag, dbin)

    def search_grouped(self, nco_freq, group_len, block_stride, n_blocks, normalize, n_lags=0):
        """nco_freq = G consecutive groups of group_len entries; best per (group, replica).
        Returns (metric, lag, dbin) of shape (G, R), dbin relative to its group."""
        f = np.ascontiguousarray(nco_freq, dtype=np.float64)
        G = f.size // int(group_len)
        metric = np.empty((G, self.R), np.float32)
        lag = np.empty((G, self.R), np.int32)
        dbin = np.empty((G, self.R), np.int32)
        self._check(self._lib.gnssacq_search_grouped(self._h, _ptr(f), f.size, int(group_len), int(block_stride), int(n_blocks),
                                                     int(bool(normalize)), int(n_lags), _ptr(metric), _ptr(lag), _ptr(dbin)))
        return metric, lag, dbin

    # -- multi-GPU through the C ABI (NCCL inside the library; no torch needed)
    def nccl_unique_id(self):
        """128-byte ncclUniqueId: create on one rank, pass to every rank's nccl_init."""
        buf = C.create_string_buffer(128)
        self._check(self._lib.gnssacq_nccl_unique_id(buf))
        return buf.raw

    def nccl_init(self, unique_id, rank, world):
        if len(unique_id) != 128:
            raise ValueError('ncclUniqueId is 128 bytes')
        self._check(self._lib.gnssacq_nccl_init(self._h, C.c_char_p(bytes(unique_id)), int(rank), int(world)))

    def search_sharded(self, nco_freq, block_stride, n_blocks, normalize, n_lags=0):
        """Collective: every rank passes the full Doppler list and gets the single-GPU answer."""
        f = np.ascontiguousarray(nco_freq, dtype=np.float64)
        metric = np.empty(self.R, np.float32)
        lag = np.empty(self.R, np.int32)
        dbin = np.empty(self.R, np.int32)
        self._check(self._lib.gnssacq_search_sharded(self._h, _ptr(f), f.size, int(block_stride), int(n_blocks), int(bool(normalize)),
                                                     int(n_lags), _ptr(metric), _ptr(lag), _ptr(dbin)))
        return metric, lag, dbin

    def search_device(self, nco_freq, block_stride, n_blocks, normalize, n_lags, device_records_ptr):
        f = np.ascontiguousarray(nco_freq, dtype=np.float64)
        self._check(self._lib.gnssacq_search_device(self._h, _ptr(f), f.size, int(block_stride), int(n_blocks),
                                                    int(bool(normalize)), int(n_lags),
                                                    C.c_void_p(int(device_records_ptr))))

    def mix(self, x, f, p):
        if not (isinstance(x, np.ndarray) and x.dtype == np.complex64 and x.flags['C_CONTIGUOUS']):
            raise TypeError('mix needs a contiguous complex64 array (as io.get_samples_complex returns)')
        self._check(self._lib.gnssacq_mix(self._h, _ptr(x), x.size, float(f), float(p)))

    def preprocess(self, raw_iq, mix_f, mix_p, fir, step, n_out, return_c128=False):
        """int8 I/Q recording -> mix -> filtfilt(fir) -> np.interp resample, all on the device; the
        result becomes the resident capture. raw_iq: int8 array (or bytes) of interleaved I,Q."""
        raw = np.frombuffer(raw_iq, dtype=np.int8) if isinstance(raw_iq, (bytes, bytearray, memoryview)) else \
            np.ascontiguousarray(raw_iq, dtype=np.int8)
        if raw.size % 2:
            raise ValueError('raw I/Q needs an even number of bytes')
        fir = np.ascontiguousarray(fir, dtype=np.float64)
        out = np.empty(int(n_out), np.complex128) if return_c128 else None
        self._check(self._lib.gnssacq_preprocess(self._h, _ptr(raw), raw.size // 2, float(mix_f), float(mix_p), _ptr(fir),
                                                 fir.size, float(step), int(n_out), _ptr(out) if return_c128 else None))
        self.n_samples = int(n_out)
        return out

    def plan_info(self):
        v = [C.c_int32() for _ in range(4)]
        self._check(self._lib.gnssacq_plan_info(self._h, *[C.byref(a) for a in v]))
        return dict(N=v[0].value, N1=v[1].value, N2=v[2].value, large=bool(v[3].value))

    def kernel_variant(self):
        v = self._lib.gnssacq_kernel_variant(self._h)
        if v < 0:
            self._check(v)
        return v

    def launch_count(self):
        return int(self._lib.gnssacq_launch_count(self._h))

    def synchronize(self):
        self._check(self._lib.gnssacq_synchronize(self._h))


_default_engine = None


def default_engine():
    """Process-wide engine on the current device (LOCAL_RANK under torchrun, else 0)."""
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(int(os.environ.get('LOCAL_RANK', '0')))
    return _default_engine


def mix_inplace(x, f, p):
    default_engine().mix(x, f, p)

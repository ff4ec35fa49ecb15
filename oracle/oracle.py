"""ctypes front-ends for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``StructuredOracle``  -> oracle/libekforacle.so  (oracle/ekf_oracle.cpp), the runtime-capacity
  restatement of slam_ros/Robot.cpp:126-943 (full, non-symmetrised covariance, GSL loop order).
* ``LiteralReference``  -> oracle/_ref/libslamref.so, the reference's own Robot.cpp (Q1-patched on a
  pipe) compiled over oracle/gsl_shim; fixed LINESIZE=100 (Robot.h:13).

Neither may be used by the product path (slam_ros_b200/); see the header of ekf_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libekforacle.so")
_ORACLE_O0_SO = os.path.join(_HERE, "libekforacle_O0.so")
_REF_SO = os.path.join(_HERE, "_ref", "libslamref.so")
_REF_O0_SO = os.path.join(_HERE, "_ref", "libslamref_O0.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(quiet=True):
    """Compile the structured oracle and, where /root/reference exists, the literal reference."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def have_literal():
    return os.path.exists(_REF_SO)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class StructuredOracle:
    def __init__(self, capacity_lines, gate=0.4, encoder_noise=0.024, reset_headroom=10, threads=1, opt0=False):
        """opt0: the -O0 build of the same source (the reference's CMakeLists sets no optimisation flag)."""
        so = _ORACLE_O0_SO if opt0 else _ORACLE_SO
        if not os.path.exists(so):
            build()
        L = self._lib = C.CDLL(so)
        L.ekfo_create.restype = C.c_void_p
        L.ekfo_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int]
        L.ekfo_y_ptr.restype = _dp
        L.ekfo_P_ptr.restype = _dp
        for f in ("ekfo_destroy", "ekfo_set_threads", "ekfo_set_pose", "ekfo_predict", "ekfo_associate",
                  "ekfo_gate_pair", "ekfo_update", "ekfo_last_gain", "ekfo_queue", "ekfo_end", "ekfo_scan", "ekfo_localize", "ekfo_n",
                  "ekfo_lines", "ekfo_y_ptr", "ekfo_P_ptr", "ekfo_get_pose", "ekfo_get_xpre", "ekfo_stats",
                  "ekfo_get_live", "ekfo_get_ellipse", "ekfo_get_threads", "ekfo_upper_stats", "ekfo_set_lines"):
            getattr(L, f).argtypes = None
        self._h = C.c_void_p(L.ekfo_create(int(capacity_lines), float(gate), float(encoder_noise), int(reset_headroom)))
        if not self._h:
            raise MemoryError("ekfo_create failed")
        L.ekfo_set_threads(self._h, C.c_int(int(threads)))
        self.capacity = int(capacity_lines)
        self.n = 3 + 2 * self.capacity

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ekfo_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def threads(self):
        return int(self._lib.ekfo_get_threads(self._h))

    @property
    def lines(self):
        return int(self._lib.ekfo_lines(self._h))

    @property
    def pose(self):
        p = np.zeros(3)
        self._lib.ekfo_get_pose(self._h, p.ctypes.data_as(_dp))
        return p

    @property
    def x_pre(self):
        p = np.zeros(3)
        self._lib.ekfo_get_xpre(self._h, p.ctypes.data_as(_dp))
        return p

    def adopt(self, other):
        """Copy the whole state (y, P, savedLineCount, pose) of another oracle of the same capacity."""
        assert other.n == self.n
        np.ctypeslib.as_array(self._lib.ekfo_y_ptr(self._h), shape=(self.n,))[:] = other.y_full()
        nl = 3 + 2 * other.lines
        self.P_view()[:nl] = other.P_view()[:nl]
        self._lib.ekfo_set_lines(self._h, C.c_int(other.lines))
        self.set_pose(other.pose)

    def set_pose(self, pose):
        a, p = _d(pose)
        self._lib.ekfo_set_pose(self._h, p)

    def y_full(self):
        return np.ctypeslib.as_array(self._lib.ekfo_y_ptr(self._h), shape=(self.n,)).copy()

    def P_full(self):
        return np.ctypeslib.as_array(self._lib.ekfo_P_ptr(self._h), shape=(self.n, self.n)).copy()

    def P_view(self):
        return np.ctypeslib.as_array(self._lib.ekfo_P_ptr(self._h), shape=(self.n, self.n))

    def live(self):
        nl = 3 + 2 * self.lines
        y = np.zeros(nl)
        P = np.zeros((nl, nl))
        self._lib.ekfo_get_live(self._h, y.ctypes.data_as(_dp), P.ctypes.data_as(_dp))
        return y, P

    def predict(self, u):
        a, p = _d(u)
        x = np.zeros(3)
        self._lib.ekfo_predict(self._h, p, x.ctypes.data_as(_dp))
        return x

    def associate(self, z, R):
        za, zp = _d(z)
        Ra, Rp = _d(R)
        innov = np.zeros(2)
        d2 = C.c_double(0)
        j = int(self._lib.ekfo_associate(self._h, zp, Rp, innov.ctypes.data_as(_dp), C.byref(d2)))
        return j, innov, d2.value

    def gate_pair(self, j, z, R):
        za, zp = _d(z)
        Ra, Rp = _d(R)
        S = np.zeros(4); Si = np.zeros(4); v = np.zeros(2); d2 = C.c_double(0)
        self._lib.ekfo_gate_pair(self._h, C.c_int(int(j)), zp, Rp, S.ctypes.data_as(_dp), Si.ctypes.data_as(_dp),
                                 v.ctypes.data_as(_dp), C.byref(d2))
        return S.reshape(2, 2), Si.reshape(2, 2), v, d2.value

    def update(self, j, z, R):
        za, zp = _d(z)
        Ra, Rp = _d(R)
        self._lib.ekfo_update(self._h, C.c_int(int(j)), zp, Rp)

    def queue(self, line_idx):
        self._lib.ekfo_queue(self._h, C.c_int(int(line_idx)))

    def end(self, z, R):
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 2)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 4)
        return int(self._lib.ekfo_end(self._h, C.c_int(z.shape[0]), z.ctypes.data_as(_dp), R.ctypes.data_as(_dp)))

    def last_gain(self):
        K = np.zeros((self.n, 2)); KS = np.zeros((self.n, 2))
        self._lib.ekfo_last_gain(self._h, K.ctypes.data_as(_dp), KS.ctypes.data_as(_dp))
        return K, KS

    def scan(self, u, z, R):
        """One Robot::localize with odometry u; returns (status, j_out)."""
        ua, up = _d(u)
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 2)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 4)
        m = z.shape[0]
        j = np.full(max(m, 1), -1, dtype=np.int32)
        st = self._lib.ekfo_scan(self._h, up, C.c_int(m), z.ctypes.data_as(_dp), R.ctypes.data_as(_dp),
                                 j.ctypes.data_as(_ip))
        return int(st), j[:m]

    def localize(self, z, R, encoder):
        ea, ep = _d(encoder)
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 2)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 4)
        m = z.shape[0]
        j = np.full(max(m, 1), -1, dtype=np.int32)
        st = self._lib.ekfo_localize(self._h, C.c_int(m), z.ctypes.data_as(_dp), R.ctypes.data_as(_dp), ep,
                                     j.ctypes.data_as(_ip))
        return int(st), j[:m]

    def stats(self):
        mm = C.c_double(0); g = C.c_longlong(0); mt = C.c_longlong(0); rs = C.c_longlong(0)
        self._lib.ekfo_stats(self._h, C.byref(mm), C.byref(g), C.byref(mt), C.byref(rs))
        return {"min_margin": mm.value, "gates": g.value, "matches": mt.value, "resets": rs.value}

    def upper_stats(self):
        """(trace, sum, sumsq) of the live covariance read as libekfcuda reads it: upper triangle mirrored."""
        t = C.c_double(0); s = C.c_double(0); q = C.c_double(0)
        self._lib.ekfo_upper_stats(self._h, C.byref(t), C.byref(s), C.byref(q))
        return t.value, s.value, q.value

    def get_ellipse(self):
        ax = (C.c_float * 2)(); ang = C.c_float(0)
        ok = self._lib.ekfo_get_ellipse(self._h, ax, C.byref(ang))
        return bool(ok), (float(ax[0]), float(ax[1])), float(ang.value)


_REF1K_SO = os.path.join(_HERE, "_ref", "libslamref1k.so")


def have_literal_1k():
    return os.path.exists(_REF1K_SO)


class LiteralReference:
    """The reference's own Robot, driven through oracle/ref_harness.cpp.  big=False: LINESIZE = 100 as shipped
    (Robot.h:13); big=True: the same sources with LINESIZE = 1000 / SLAMSIZE = 2003 (oracle/_ref/libslamref1k.so,
    BASELINE configs[1]) -- its localize needs ~300 MB of stack and runs on a dedicated thread."""

    def __init__(self, big=False, opt0=False, so_path=None):
        """so_path: another library with the same harness ABI (oracle/_ref/libslamdropin.so: the drop-in built on the
        reference's headers)."""
        so = so_path or (_REF1K_SO if big else (_REF_O0_SO if opt0 else _REF_SO))
        self._big = bool(big)
        if not os.path.exists(so):
            build()
        if not os.path.exists(so):
            raise FileNotFoundError(so + " (needs /root/reference to build)")
        L = self._lib = C.CDLL(so)
        L.ref_create.restype = C.c_void_p
        L.ref_gate.restype = C.c_double
        L.ref_encoder_noise.restype = C.c_double
        L.ref_range_errors.restype = C.c_ulong
        L.ref_badlen_errors.restype = C.c_ulong
        self._h = C.c_void_p(L.ref_create())
        self.capacity = int(L.ref_linesize())
        self.n = int(L.ref_slamsize())
        self.gate = float(L.ref_gate())
        self.encoder_noise = float(L.ref_encoder_noise())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ref_destroy(self._h)
            self._h = None

    __del__ = close

    def localize_intervals(self, z, R, encoder, intervals):
        """Robot::localize with line::lineInterval set (intervals: m x 4 = two (alfa, r) end points per line);
        returns Robot::lineIntervals.data as the call left it (Robot.cpp:868-879), float32."""
        ea, ep = _d(encoder)
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 2)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 4)
        iv = np.ascontiguousarray(intervals, dtype=np.float64).reshape(-1, 4)
        assert iv.shape[0] == z.shape[0] and not self._big
        out = np.zeros(4 * z.shape[0] + 4, dtype=np.float32)
        self._lib.ref_localize_iv.restype = C.c_int
        n = self._lib.ref_localize_iv(self._h, C.c_int(z.shape[0]), z.ctypes.data_as(_dp), R.ctypes.data_as(_dp), ep,
                                      iv.ctypes.data_as(_dp), out.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(out.size))
        return out[:n].copy()

    def localize(self, z, R, encoder):
        ea, ep = _d(encoder)
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 2)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 4)
        if self._big:
            rc = self._lib.ref_localize_bigstack(self._h, C.c_int(z.shape[0]), z.ctypes.data_as(_dp), R.ctypes.data_as(_dp), ep, C.c_int(1024))
            if rc != 0:
                raise RuntimeError("ref_localize_bigstack failed: %d" % rc)
        else:
            self._lib.ref_localize(self._h, C.c_int(z.shape[0]), z.ctypes.data_as(_dp), R.ctypes.data_as(_dp), ep)

    def state(self, want_cov=True):
        L = C.c_int(0); pose = np.zeros(3)
        if not want_cov:
            self._lib.ref_get(self._h, None, None, C.byref(L), pose.ctypes.data_as(_dp))
            return None, None, int(L.value), pose
        y = np.zeros(self.n); P = np.zeros((self.n, self.n))
        self._lib.ref_get(self._h, y.ctypes.data_as(_dp), P.ctypes.data_as(_dp), C.byref(L), pose.ctypes.data_as(_dp))
        return y, P, int(L.value), pose

    def get_ellipse(self):
        ax = (C.c_float * 2)(); ang = C.c_float(0)
        ok = self._lib.ref_get_ellipse(self._h, ax, C.byref(ang))
        return bool(ok), (float(ax[0]), float(ax[1])), float(ang.value)

    def range_errors(self):
        return int(self._lib.ref_range_errors())


# ------------------------------------------------------------------------------------------------
# line extraction: the restatement and the deterministic build of the reference itself
# ------------------------------------------------------------------------------------------------
_LINES_SO = os.path.join(_HERE, "liblinesoracle.so")
_REF_LINES_SO = os.path.join(_HERE, "_ref", "libslamlines.so")
_fp = C.POINTER(C.c_float)


def have_literal_lines():
    return os.path.exists(_REF_LINES_SO)


def _extract(fn, payload, max_lines):
    d = np.ascontiguousarray(np.asarray(payload, dtype=np.float32).reshape(-1))
    out = np.zeros((max_lines, 10))
    n = fn(d.size // 2, d.ctypes.data_as(_fp), max_lines, out.ctypes.data_as(_dp))
    if n < 0:
        raise RuntimeError("line extraction failed")
    return out[:min(n, max_lines)].copy(), n


class LinesOracle:
    """oracle/lines_oracle.cpp: mapping_cb + LineExtraction (slam_ros/main.cpp:37-71, lineFitting.cpp:640-702).
    extract(payload) -> (rows, n): rows[i] = alfa, r, C_AR[4], interval0 (alfa, r), interval1 (alfa, r)."""

    def __init__(self):
        if not os.path.exists(_LINES_SO):
            build()
        self._lib = C.CDLL(_LINES_SO)
        self._lib.lxo_extract.argtypes = [C.c_int, _fp, C.c_int, _dp]
        self._lib.lxo_fit.argtypes = [C.c_int, _dp, _dp, _dp]

    def extract(self, payload, max_lines=128):
        return _extract(self._lib.lxo_extract, payload, max_lines)

    def fit(self, alfa, r):
        a, ap = _d(alfa); rr, rp = _d(r); out = np.zeros(2)
        self._lib.lxo_fit(int(a.size), ap, rp, out.ctypes.data_as(_dp))
        return out


class LiteralLineExtraction:
    """oracle/_ref/libslamlines.so: the reference's own lineFitting.cpp / simplifyPath.cpp / vec2.cpp, built with
    zero-initialised automatic variables and gsl_matrix_alloc storage (one deterministic instance of it)."""

    def __init__(self):
        if not have_literal_lines():
            raise RuntimeError("oracle/_ref/libslamlines.so is not built (needs /root/reference: make -C oracle ref)")
        self._lib = C.CDLL(_REF_LINES_SO)
        self._lib.ref_extract_lines.argtypes = [C.c_int, _fp, C.c_int, _dp]

    def extract(self, payload, max_lines=128):
        return _extract(self._lib.ref_extract_lines, payload, max_lines)


class LineProviderTransform:
    """The reference's own end-point transform: lineprovider/main.cpp's Transform() (lines 60-84) compiled unmodified into
    oracle/_ref/libslamlineprov.so (oracle/lineprov_harness.cpp).  Test infrastructure."""
    PATH = os.path.join(_HERE, "_ref", "libslamlineprov.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        self._lib = C.CDLL(self.PATH, mode=C.RTLD_LOCAL)
        self._lib.lp_transform.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_float)]

    def transform(self, intervals, pose):
        """intervals: (n, 4) = (alfa0, r0, alfa1, r1) per line; pose = (x, y, theta).  Returns (n, 4) float32."""
        iv = np.ascontiguousarray(np.asarray(intervals, dtype=np.float64).reshape(-1, 4))
        ps = np.ascontiguousarray(np.asarray(pose, dtype=np.float64).reshape(3))
        out = np.zeros((iv.shape[0], 4), dtype=np.float32)
        k = self._lib.lp_transform(iv.shape[0], iv.ctypes.data_as(C.POINTER(C.c_double)), ps.ctypes.data_as(C.POINTER(C.c_double)),
                                   out.ctypes.data_as(C.POINTER(C.c_float)))
        assert k == 4 * iv.shape[0]
        return out

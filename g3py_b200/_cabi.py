"""ctypes binding of libg3b.so (include/g3b.h).  PyTorch-free; NumPy arrays in and out.

There is no CPU fallback: if the shared library is missing or no sm_100 device is present,
`load()` / `Context()` raise.  Contexts are created lazily per process (fork-safe) and are
dropped on pickling, like the compiled functions of the reference's `makefn`
(g3py/libs/tensors.py:35-74; pickling at g3py/processes/stochastic.py:107-119).
"""
import ctypes as C
import os

import numpy as np

G3_MAX_NODES = 32
G3_MAX_THETA = 64
G3_MAX_DIM = 16

# leaf / node opcodes (include/g3b.h)
K_SE, K_OU, K_MAT32, K_MAT52, K_RQ, K_SIN, K_NOISE, K_WN = 1, 2, 3, 4, 5, 6, 7, 8
K_COS, K_SINC, K_SM = 9, 10, 11
K_DOT, K_BW, K_VAR, K_EQ = 12, 13, 14, 15
K_SUM, K_PROD, K_SCALE, K_SHIFT, K_MAX = 16, 17, 18, 19, 20
KF_PROCESS_NOISE = 1
KF_NN, KF_EQ2 = 0x10000, 0x20000

ST_NONFINITE_INPUT, ST_DIAG_SHIFT, ST_JITTER, ST_POTRF_FAILED, ST_NONFINITE_RESULT = 1, 2, 4, 8, 16
KIND_GAUSS, KIND_STUDENT = 0, 1
POST_NOISE, POST_COV = 1, 2


class KNode(C.Structure):
    _fields_ = [("op", C.c_int32), ("dim0", C.c_int32), ("dim1", C.c_int32), ("var_idx", C.c_int32),
                ("p0_idx", C.c_int32), ("p1_idx", C.c_int32), ("flags", C.c_int32), ("value", C.c_double)]


class KernelDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_theta", C.c_int32), ("nodes", KNode * G3_MAX_NODES)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_ctxp = C.c_void_p

SIGNATURES = {
    "g3_ctx_create": (C.c_int, [C.c_int, C.POINTER(_ctxp)]),
    "g3_ctx_destroy": (C.c_int, [_ctxp]),
    "g3_ctx_trim": (C.c_int, [_ctxp]),
    "g3_last_error": (C.c_char_p, [_ctxp]),
    "g3_sync": (C.c_int, [_ctxp]),
    "g3_set_jitter": (C.c_int, [_ctxp, C.c_double, C.c_int]),
    "g3_set_potrf_block": (C.c_int, [_ctxp, C.c_int]),
    "g3_set_lookahead": (C.c_int, [_ctxp, C.c_int]),
    "g3_set_splitk": (C.c_int, [_ctxp, C.c_int]),
    "g3_set_trtri_pipeline": (C.c_int, [_ctxp, C.c_int]),
    "g3_set_groups": (C.c_int, [_ctxp, C.c_int]),
    "g3_set_graphs": (C.c_int, [_ctxp, C.c_int]),
    "g3_graph_replays": (C.c_int64, [_ctxp]),
    "g3_set_gemm_mode": (C.c_int, [_ctxp, C.c_int, C.c_int]),
    "g3_ozaki_launch_count": (C.c_int64, [_ctxp]),
    "g3_timer_begin": (C.c_int, [_ctxp]),
    "g3_timer_end": (C.c_int, [_ctxp, C.POINTER(C.c_float)]),
    "g3_launch_count": (C.c_int64, [_ctxp]),
    "g3_prof_enable": (C.c_int, [_ctxp, C.c_int]),
    "g3_prof_read": (C.c_int, [_ctxp, _dp, C.POINTER(C.c_int64)]),
    "g3_debug_read": (C.c_int, [_ctxp, C.c_char_p, C.c_void_p, C.c_size_t]),
    "g3_debug_potrf_stress": (C.c_int, [_ctxp, C.POINTER(KernelDesc), _dp, C.c_int, C.c_int, _ip, _dp]),
    "g3_debug_gemm_stress": (C.c_int, [_ctxp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]),
    "g3_debug_fp64_peak": (C.c_int, [_ctxp, C.c_double, _dp, _dp, _dp]),
    "g3_set_diag_variant": (C.c_int, [_ctxp, C.c_int]),
    "g3_set_tile_split": (C.c_int, [_ctxp, C.c_int]),
    "g3_set_trsv_fused": (C.c_int, [_ctxp, C.c_int]),
    "g3_set_speculate_grad": (C.c_int, [_ctxp, C.c_int]),
    "g3_debug_diag_time": (C.c_int, [_ctxp, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_float),
                                     C.POINTER(C.c_longlong), _dp]),
    "g3_set_stream": (C.c_int, [_ctxp, C.c_void_p]),
    "g3_dev_gram_block": (C.c_int, [_ctxp, C.POINTER(KernelDesc), _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                    C.c_void_p, C.c_longlong]),
    "g3_dev_potrf_panel": (C.c_int, [_ctxp, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "g3_dev_trsv_panel": (C.c_int, [_ctxp, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "g3_dev_syrk_panel": (C.c_int, [_ctxp, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "g3_comm_get_unique_id": (C.c_int, [C.c_char_p]),
    "g3_comm_init": (C.c_int, [_ctxp, C.c_int, C.c_int, C.c_char_p]),
    "g3_comm_destroy": (C.c_int, [_ctxp]),
    "g3_comm_size": (C.c_int, [_ctxp]),
    "g3_comm_rank": (C.c_int, [_ctxp]),
    "g3_comm_barrier": (C.c_int, [_ctxp]),
    "g3_comm_allgather": (C.c_int, [_ctxp, C.c_void_p, C.c_void_p, C.c_size_t]),
    "g3_comm_allreduce": (C.c_int, [_ctxp, _dp, C.c_int, C.c_int]),
    "g3_dist_factor": (C.c_int, [_ctxp, C.POINTER(KernelDesc), _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _ip,
                                 C.POINTER(C.c_float), C.POINTER(C.c_float), _dp]),
    "g3_dist_solve": (C.c_int, [_ctxp, _dp, _dp, _dp, C.POINTER(C.c_float)]),
    "g3_dist_posterior": (C.c_int, [_ctxp, _dp, C.c_int, C.c_int, _dp, _dp]),
    "g3_dist_grad": (C.c_int, [_ctxp, C.c_double, _dp, _dp, C.POINTER(C.c_float)]),
    "g3_dist_residual": (C.c_int, [_ctxp, C.c_int, C.c_uint, _dp]),
    "g3_dist_read_piece": (C.c_int, [_ctxp, C.c_int, _dp, _ip]),
    "g3_dist_free": (C.c_int, [_ctxp]),
    "g3_dist_layout": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip]),
    "g3_potrf_2d": (C.c_int, [_ctxp, C.POINTER(KernelDesc), _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _ip,
                              C.POINTER(C.c_float)]),
    "g3_set_data": (C.c_int, [_ctxp, _dp, C.c_int, C.c_int]),
    "g3_gram": (C.c_int, [_ctxp, C.POINTER(KernelDesc), _dp, C.c_int, _dp, C.c_int, C.c_int, _dp, C.c_int, _dp, _ip]),
    "g3_gram_vjp": (C.c_int, [_ctxp, C.POINTER(KernelDesc), _dp, C.c_int, _dp, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp]),
    "g3_potrf_robust": (C.c_int, [_ctxp, _dp, C.c_int, C.c_int, C.c_int, _ip, _dp]),
    "g3_potrf_robust_solve": (C.c_int, [_ctxp, _dp, C.c_int, C.c_int, _dp, _dp, _ip, _dp]),
    "g3_gp_logp_grad": (C.c_int, [_ctxp, C.POINTER(KernelDesc), C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, _dp, _dp,
                                  _dp, _dp, _ip]),
    "g3_gp_grad_resume": (C.c_int, [_ctxp, _dp, _dp]),
    "g3_gp_upload": (C.c_int, [_ctxp, C.POINTER(KernelDesc), C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int]),
    "g3_gp_run": (C.c_int, [_ctxp]),
    "g3_gp_download": (C.c_int, [_ctxp, _dp, _dp, _dp, _dp, _ip]),
    "g3_gp_posterior": (C.c_int, [_ctxp, C.POINTER(KernelDesc), _dp, C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp, _ip]),
    "g3_gram_potrf_device": (C.c_int, [_ctxp, C.POINTER(KernelDesc), _dp, _dp, _ip, C.POINTER(C.c_float),
                                       C.POINTER(C.c_float)]),
}

_LIB = None


def lib_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libg3b.so")


def load():
    """Load libg3b.so and set the prototypes.  Raises if the library was not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError("libg3b.so not found at %s: build it with `make -C g3py_b200/csrc` "
                           "(or __graft_entry__.build()); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing: loud by design
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def comm_unique_id():
    """128-byte NCCL id (rank 0 creates it, the host distributes it: g3py_b200/comm.py)."""
    buf = C.create_string_buffer(128)
    rc = load().g3_comm_get_unique_id(buf)
    if rc != 0:
        raise G3Error("g3_comm_get_unique_id failed (%d): libnccl.so.2 not loadable?" % rc)
    return buf.raw


def dist_layout(N, nb, Pr, Pc, I, J, p):
    """{owner, first, count, index} of block (I, J) / piece (J, p): pure index arithmetic of the C side."""
    out = (C.c_int * 4)()
    rc = load().g3_dist_layout(int(N), int(nb), int(Pr), int(Pc), int(I), int(J), int(p), out)
    if rc != 0:
        raise ValueError("g3_dist_layout: bad arguments")
    return {"owner": out[0], "first": out[1], "count": out[2], "index": out[3]}


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class G3Error(RuntimeError):
    pass


def _theta2d(theta, desc, what):
    """(B, n_theta) float64, validated against the descriptor: the C side copies B * n_theta doubles from the
    caller's buffer, so a mismatched array would be read out of bounds instead of raising."""
    theta = np.atleast_2d(_f64(theta))
    if theta.ndim != 2 or theta.shape[1] != max(desc.n_theta, 0):
        if not (desc.n_theta == 0 and theta.size == 0):
            raise ValueError("%s: theta has shape %s, the kernel descriptor takes %d hypers per row"
                             % (what, theta.shape, desc.n_theta))
    return theta


class Context:
    """One device context (one stream, cached workspaces)."""

    def __init__(self, device=0):
        self._lib = load()
        h = _ctxp()
        rc = self._lib.g3_ctx_create(int(device), C.byref(h))
        if rc != 0:
            why = {-3: "no CUDA device visible", -4: "device is not sm_100 (B200)", -1: "bad device index"}.get(rc, "CUDA error")
            raise G3Error("g3_ctx_create(device=%d) failed (%d): %s; there is no CPU fallback" % (device, rc, why))
        self._h = h
        self.device = device
        self.N = 0
        self.D = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.g3_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            msg = self._lib.g3_last_error(self._h)
            raise G3Error("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))

    # ---- configuration / timing
    def set_jitter(self, jitter_rel, max_tries=20):
        self._ck(self._lib.g3_set_jitter(self._h, float(jitter_rel), int(max_tries)), "g3_set_jitter")

    def set_potrf_block(self, w):
        self._ck(self._lib.g3_set_potrf_block(self._h, int(w)), "g3_set_potrf_block")

    def set_lookahead(self, on):
        self._ck(self._lib.g3_set_lookahead(self._h, int(bool(on))), "g3_set_lookahead")

    def set_trtri_pipeline(self, on):
        self._ck(self._lib.g3_set_trtri_pipeline(self._h, int(bool(on))), "g3_set_trtri_pipeline")

    def set_splitk(self, on):
        """0 off, 1 (default) at least 128 of the contraction per share, 2 shares down to 32 and triangular solves too (measured slower)."""
        self._ck(self._lib.g3_set_splitk(self._h, int(on)), "g3_set_splitk")

    def set_speculate_grad(self, on):
        """logp-only evaluations also prepare U = L^-T for a gradient that follows (gp_grad_resume); default off."""
        self._ck(self._lib.g3_set_speculate_grad(self._h, int(bool(on))), "g3_set_speculate_grad")

    def set_trsv_fused(self, on):
        """Whole triangular solves in one launch (default on); off = one launch per 128-row block."""
        self._ck(self._lib.g3_set_trsv_fused(self._h, int(bool(on))), "g3_set_trsv_fused")

    def set_tile_split(self, on):
        """Spread each tile of a few-tile GEMM launch over 2 / 4 CTAs by columns (default on)."""
        self._ck(self._lib.g3_set_tile_split(self._h, int(bool(on))), "g3_set_tile_split")

    def set_graphs(self, on):
        """CUDA-graph replay of small-batch evaluations (default on)."""
        self._ck(self._lib.g3_set_graphs(self._h, int(bool(on))), "g3_set_graphs")

    def graph_replays(self):
        return int(self._lib.g3_graph_replays(self._h))

    def set_gemm_mode(self, mode, min_k=0):
        """'dmma' (default) or 'ozaki': deep panel updates of the batched Cholesky on the int8 tensor cores (include/g3b.h)."""
        m = {"dmma": 0, "ozaki": 1}.get(mode, mode)
        self._ck(self._lib.g3_set_gemm_mode(self._h, int(m), int(min_k)), "g3_set_gemm_mode")

    def ozaki_launch_count(self):
        return int(self._lib.g3_ozaki_launch_count(self._h))

    def set_groups(self, n):
        self._ck(self._lib.g3_set_groups(self._h, int(n)), "g3_set_groups")

    # ---- device-pointer building blocks (multi-GPU Cholesky); pointers are integers (tensor.data_ptr())
    def set_stream(self, cuda_stream):
        self._ck(self._lib.g3_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None), "g3_set_stream")

    def dev_gram_block(self, desc, theta, row0, col0, rows, cols, diag_shift, out_ptr, ld):
        theta = _f64(theta).ravel()
        if theta.size != desc.n_theta:
            raise ValueError("dev_gram_block: theta has %d entries, the descriptor takes %d" % (theta.size, desc.n_theta))
        self._ck(self._lib.g3_dev_gram_block(self._h, C.byref(desc), _d(theta), row0, col0, rows, cols, float(diag_shift),
                                             C.c_void_p(out_ptr), ld), "g3_dev_gram_block")

    def dev_potrf_panel(self, p_ptr, rows, nb, logdet_ptr, info_ptr, dinv_ptr=0):
        self._ck(self._lib.g3_dev_potrf_panel(self._h, C.c_void_p(p_ptr), rows, nb, C.c_void_p(dinv_ptr) if dinv_ptr else None,
                                              C.c_void_p(logdet_ptr), C.c_void_p(info_ptr)), "g3_dev_potrf_panel")

    def dev_trsv_panel(self, p_ptr, rows, nb, dinv_ptr, r_ptr, u_ptr, beta_ptr):
        self._ck(self._lib.g3_dev_trsv_panel(self._h, C.c_void_p(p_ptr), rows, nb, C.c_void_p(dinv_ptr), C.c_void_p(r_ptr),
                                             C.c_void_p(u_ptr), C.c_void_p(beta_ptr)), "g3_dev_trsv_panel")

    def dev_syrk_panel(self, p_ptr, rows_p, nb, row_off, d_ptr, rows_d):
        self._ck(self._lib.g3_dev_syrk_panel(self._h, C.c_void_p(p_ptr), rows_p, nb, row_off, C.c_void_p(d_ptr), rows_d),
                 "g3_dev_syrk_panel")

    def trim(self):
        """Free all cached device / pinned workspaces (re-created on demand)."""
        self._resident = None
        self._ck(self._lib.g3_ctx_trim(self._h), "g3_ctx_trim")

    def sync(self):
        self._ck(self._lib.g3_sync(self._h), "g3_sync")

    def timer_begin(self):
        self._ck(self._lib.g3_timer_begin(self._h), "g3_timer_begin")

    def timer_end(self):
        ms = C.c_float()
        self._ck(self._lib.g3_timer_end(self._h, C.byref(ms)), "g3_timer_end")
        return float(ms.value)

    PROF_CLASSES = ("dgemm_nt", "potrf_diag", "gram_fwd", "gram_vjp", "trsv", "other")

    def prof_enable(self, on=True):
        self._ck(self._lib.g3_prof_enable(self._h, 1 if on else 0), "g3_prof_enable")

    def prof_read(self):
        ms = np.zeros(6)
        n = np.zeros(6, dtype=np.int64)
        self._ck(self._lib.g3_prof_read(self._h, _d(ms), n.ctypes.data_as(C.POINTER(C.c_int64))), "g3_prof_read")
        return {k: {"ms": float(ms[i]), "launches": int(n[i])} for i, k in enumerate(self.PROF_CLASSES)}

    def debug_read(self, name, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        self._ck(self._lib.g3_debug_read(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes), "g3_debug_read")
        return out

    def debug_potrf_stress(self, desc, theta, iters):
        theta = np.atleast_2d(_f64(theta))
        out = np.zeros((iters, 4), dtype=np.int32)
        tiles = np.zeros((2, 128, 128))
        self._resident = None
        self._ck(self._lib.g3_debug_potrf_stress(self._h, C.byref(desc), _d(theta), theta.shape[0], iters, _i(out), _d(tiles)),
                 "g3_debug_potrf_stress")
        return out, tiles

    def debug_gemm_stress(self, rows, B, launches, inplace, kdepth=128):
        out = (C.c_longlong * 4)()
        self._ck(self._lib.g3_debug_gemm_stress(self._h, rows, B, launches, inplace, kdepth, out), "g3_debug_gemm_stress")
        return list(out)

    def fp64_peak(self, seconds=0.5, copy=True):
        """In-run roofline denominators: {dmma_tflops, dfma_tflops, copy_gbs} measured on this context's GPU."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._ck(self._lib.g3_debug_fp64_peak(self._h, float(seconds), C.cast(C.byref(a), _dp), C.cast(C.byref(b), _dp),
                                              C.cast(C.byref(c), _dp) if copy else None), "g3_debug_fp64_peak")
        return {"dmma_tflops": float(a.value), "dfma_tflops": float(b.value), "copy_gbs": float(c.value) if copy else None}

    def set_diag_variant(self, variant):
        """2 = low-latency diagonal-tile kernel (default), 1 = the first kernel (A/B timing)."""
        self._ck(self._lib.g3_set_diag_variant(self._h, int(variant)), "g3_set_diag_variant")

    def debug_diag_time(self, variant=2, B=1, reps=20, cond_shift=0.05):
        """Microseconds per launch of the 128x128 factor+inverse kernel alone, its phase clocks and host-checked errors."""
        us = C.c_float()
        stamps = (C.c_longlong * 64)()
        err = np.zeros(4)
        self._ck(self._lib.g3_debug_diag_time(self._h, int(variant), int(B), int(reps), float(cond_shift), C.byref(us), stamps,
                                              _d(err)), "g3_debug_diag_time")
        return float(us.value), np.array(list(stamps), dtype=np.int64), err

    def launch_count(self):
        return int(self._lib.g3_launch_count(self._h))

    # ---- data
    def set_data(self, X):
        X = _f64(X)
        if X.ndim == 1:
            X = X[:, None]
        self._resident = None
        self._ck(self._lib.g3_set_data(self._h, _d(X), X.shape[0], X.shape[1]), "g3_set_data")
        self.N, self.D = X.shape
        self._data_serial = getattr(self, "_data_serial", 0) + 1
        self._X_host = X.copy()

    def set_data_if_changed(self, X):
        """Upload X only when it differs from the resident copy (Theano `perform` hands the same array every call)."""
        X = _f64(X)
        if X.ndim == 1:
            X = X[:, None]
        cur = getattr(self, "_X_host", None)
        if cur is None or cur.shape != X.shape or not np.array_equal(cur, X):
            self.set_data(X)
            self._data_tag = None

    # ---- gram
    def gram(self, desc, X1, X2, theta):
        X1 = _f64(X1)
        theta = _theta2d(theta, desc, "gram")
        B = theta.shape[0]
        n1, D = X1.shape
        if X2 is None:
            n2, x2p = n1, None
        else:
            X2 = _f64(X2)
            n2, x2p = X2.shape[0], _d(X2)
        K = np.empty((B, n1, n2))
        st = np.zeros(B, dtype=np.int32)
        self._ck(self._lib.g3_gram(self._h, C.byref(desc), _d(X1), n1, x2p, n2, D, _d(theta), B, _d(K), _i(st)), "g3_gram")
        return K, st

    def gram_vjp(self, desc, X1, X2, theta, W):
        X1 = _f64(X1)
        theta = _theta2d(theta, desc, "gram_vjp")
        B = theta.shape[0]
        n1, D = X1.shape
        if X2 is None:
            n2, x2p = n1, None
        else:
            X2 = _f64(X2)
            n2, x2p = X2.shape[0], _d(X2)
        W = _f64(W).reshape(B, n1, n2)
        g = np.zeros((B, desc.n_theta))
        self._ck(self._lib.g3_gram_vjp(self._h, C.byref(desc), _d(X1), n1, x2p, n2, D, _d(theta), B, _d(W), _d(g)),
                 "g3_gram_vjp")
        return g

    # ---- cholesky
    def potrf_robust(self, A):
        """A: (B, n, n) or (n, n) symmetric; returns (L, info, jitter).  Input is not modified."""
        A = np.array(A, dtype=np.float64, order="C", copy=True)
        single = A.ndim == 2
        if single:
            A = A[None]
        B, n, _ = A.shape
        info = np.zeros(B, dtype=np.int32)
        jit = np.zeros(B)
        self._ck(self._lib.g3_potrf_robust(self._h, _d(A), n, n, B, _i(info), _d(jit)), "g3_potrf_robust")
        if single:
            return A[0], int(info[0]), float(jit[0])
        return A, info, jit

    def potrf_robust_solve(self, A, rhs):
        """One matrix: (L, info, jitter, u = L^-1 rhs), factorisation and substitution both on the device."""
        A = np.array(A, dtype=np.float64, order="C", copy=True)
        n = A.shape[0]
        rhs = _f64(rhs).ravel()
        if rhs.shape[0] != n:
            raise ValueError("rhs must have n entries")
        u = np.empty(n)
        info = np.zeros(1, dtype=np.int32)
        jit = np.zeros(1)
        self._ck(self._lib.g3_potrf_robust_solve(self._h, _d(A), n, n, _d(rhs), _d(u), _i(info), _d(jit)),
                 "g3_potrf_robust_solve")
        return A, int(info[0]), float(jit[0]), u

    # ---- fused logp + grad
    def gp_logp_grad(self, desc, kind, delta, theta, nu=None, want_grad=True):
        """delta: (N,) shared or (B, N); theta: (B, P_kernel) natural space.
        Returns dict(beta, logdet, dtheta, ddelta, status)."""
        theta = _theta2d(theta, desc, "gp_logp_grad")
        B = theta.shape[0]
        delta = _f64(delta)
        stride = 0 if delta.ndim == 1 else self.N
        if delta.ndim == 2 and delta.shape[0] != B:
            raise ValueError("delta must be (N,) or (B, N)")
        if delta.shape[-1] != self.N:
            raise ValueError("delta length %d != N %d" % (delta.shape[-1], self.N))
        nu_a = _f64(np.broadcast_to(nu, (B,))) if nu is not None else None
        beta = np.empty(B)
        logdet = np.empty(B)
        st = np.zeros(B, dtype=np.int32)
        dth = np.zeros((B, max(desc.n_theta, 1))) if want_grad else None
        ddl = np.zeros((B, self.N)) if want_grad else None
        self._resident = None
        self._ck(self._lib.g3_gp_logp_grad(self._h, C.byref(desc), int(kind), _d(delta), stride, _d(theta), B, _d(nu_a),
                                           _d(beta), _d(logdet), _d(dth), _d(ddl), _i(st)), "g3_gp_logp_grad")
        if not want_grad:        # the factor of this evaluation stays on the device: see gp_grad_resume
            self._resident = (self.eval_key(desc, kind, delta, theta, nu_a), B, desc.n_theta)
        return {"beta": beta, "logdet": logdet, "dtheta": None if dth is None else dth[:, :desc.n_theta],
                "ddelta": ddl, "status": st}

    def eval_key(self, desc, kind, delta, theta, nu):
        """Identity of one evaluation on the resident data: bytes of every input of g3_gp_logp_grad."""
        return (getattr(self, "_data_serial", 0), bytes(desc), int(kind), _f64(delta).tobytes(),
                np.atleast_2d(_f64(theta)).tobytes(), None if nu is None else _f64(nu).tobytes())

    def resident_matches(self, desc, kind, delta, theta, nu):
        r = getattr(self, "_resident", None)
        return r is not None and r[0] == self.eval_key(desc, kind, delta, theta, nu)

    def gp_grad_resume(self):
        """(dtheta (B, P), ddelta (B, N)) of the last logp-only gp_logp_grad call, from its resident factor."""
        r = getattr(self, "_resident", None)
        if r is None:
            raise G3Error("gp_grad_resume: no resident factor")
        _, B, P = r
        self._resident = None
        dth = np.zeros((B, max(P, 1)))
        ddl = np.zeros((B, self.N))
        self._ck(self._lib.g3_gp_grad_resume(self._h, _d(dth), _d(ddl)), "g3_gp_grad_resume")
        return dth[:, :P], ddl

    def gp_upload(self, desc, kind, delta, theta, nu=None, want_grad=True):
        self._resident = None
        theta = _theta2d(theta, desc, "gp_upload")
        B = theta.shape[0]
        delta = _f64(delta)
        stride = 0 if delta.ndim == 1 else self.N
        if self.N and (delta.shape[-1] != self.N or (delta.ndim == 2 and delta.shape[0] != B) or delta.ndim > 2):
            raise ValueError("gp_upload: delta must be (N,) or (B, N) with N = %d, B = %d; got %s" % (self.N, B, delta.shape))
        nu_a = _f64(np.broadcast_to(nu, (B,))) if nu is not None else None
        self._ck(self._lib.g3_gp_upload(self._h, C.byref(desc), int(kind), _d(delta), stride, _d(theta), B, _d(nu_a),
                                        1 if want_grad else 0), "g3_gp_upload")
        self._up = (B, desc.n_theta, bool(want_grad))

    def gp_run(self):
        self._ck(self._lib.g3_gp_run(self._h), "g3_gp_run")

    def gp_download(self):
        B, P, want_grad = self._up
        beta = np.empty(B)
        logdet = np.empty(B)
        st = np.zeros(B, dtype=np.int32)
        dth = np.zeros((B, max(P, 1))) if want_grad else None
        ddl = np.zeros((B, self.N)) if want_grad else None
        self._ck(self._lib.g3_gp_download(self._h, _d(beta), _d(logdet), _d(dth), _d(ddl), _i(st)), "g3_gp_download")
        return {"beta": beta, "logdet": logdet, "dtheta": None if dth is None else dth[:, :P], "ddelta": ddl, "status": st}

    # ---- posterior
    def gp_posterior(self, desc, Xs, delta, theta, noise=False, cov=False):
        Xs = _f64(Xs)
        if Xs.ndim == 1:
            Xs = Xs[:, None]
        M = Xs.shape[0]
        delta = _f64(delta).ravel()
        theta = _f64(theta).ravel()
        if theta.size != desc.n_theta:
            raise ValueError("gp_posterior: theta has %d entries, the kernel descriptor takes %d" % (theta.size, desc.n_theta))
        if delta.size != self.N or Xs.shape[1] != self.D:
            raise ValueError("gp_posterior: delta must have N = %d entries and Xs D = %d columns" % (self.N, self.D))
        mean = np.empty(M)
        var = np.empty(M)
        covm = np.empty((M, M)) if cov else None
        beta = C.c_double()
        st = C.c_int()
        flags = (POST_NOISE if noise else 0) | (POST_COV if cov else 0)
        self._resident = None
        self._ck(self._lib.g3_gp_posterior(self._h, C.byref(desc), _d(Xs), M, _d(delta), _d(theta), flags, _d(mean),
                                           _d(var), _d(covm), C.cast(C.byref(beta), _dp), C.cast(C.byref(st), _ip)),
                 "g3_gp_posterior")
        return {"mean": mean, "var": var, "cov": covm, "beta": float(beta.value), "status": int(st.value)}

    # ---- multi-GPU: NCCL communicator owned by the library, block-cyclic exact GP (include/g3b.h)
    def comm_init(self, nranks, rank, uid):
        self._ck(self._lib.g3_comm_init(self._h, int(nranks), int(rank), uid), "g3_comm_init")
        self.nranks, self.rank = int(nranks), int(rank)

    def comm_destroy(self):
        self._ck(self._lib.g3_comm_destroy(self._h), "g3_comm_destroy")

    def comm_size(self):
        return int(self._lib.g3_comm_size(self._h))

    def comm_rank(self):
        return int(self._lib.g3_comm_rank(self._h))

    def comm_barrier(self):
        self._ck(self._lib.g3_comm_barrier(self._h), "g3_comm_barrier")

    def comm_allgather(self, arr):
        """arr: this rank's contiguous array; returns (nranks, *arr.shape), rank order."""
        a = np.ascontiguousarray(arr)
        out = np.empty((self.comm_size(),) + a.shape, dtype=a.dtype)
        self._ck(self._lib.g3_comm_allgather(self._h, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), a.nbytes),
                 "g3_comm_allgather")
        return out

    def comm_allreduce(self, vals, op="sum"):
        v = np.array(vals, dtype=np.float64, copy=True).ravel()
        self._ck(self._lib.g3_comm_allreduce(self._h, _d(v), v.size, {"sum": 0, "max": 1, "min": 2}[op]), "g3_comm_allreduce")
        return v

    def dist_factor(self, desc, theta, nb, Pr, Pc, lookahead=True, ring=2):
        theta = _f64(theta).ravel()
        if theta.size != desc.n_theta:
            raise ValueError("dist_factor: theta has %d entries, the descriptor takes %d" % (theta.size, desc.n_theta))
        ld, info, mg, mp, gib = C.c_double(), C.c_int(), C.c_float(), C.c_float(), C.c_double()
        flags = (0 if lookahead else 1) | (4 if ring == 3 else 0)
        self._resident = None
        self._ck(self._lib.g3_dist_factor(self._h, C.byref(desc), _d(theta), int(nb), int(Pr), int(Pc), flags,
                                          C.cast(C.byref(ld), _dp), C.cast(C.byref(info), _ip), C.byref(mg), C.byref(mp),
                                          C.cast(C.byref(gib), _dp)), "g3_dist_factor")
        return {"logdet": float(ld.value), "info": int(info.value), "ms_gram": float(mg.value), "ms_potrf": float(mp.value),
                "local_gib": float(gib.value)}

    def dist_solve(self, delta, want_u=False):
        delta = _f64(delta).ravel()
        if delta.size != self.N:
            raise ValueError("dist_solve: delta must have N = %d entries" % self.N)
        beta, ms = C.c_double(), C.c_float()
        u = np.empty(self.N) if want_u else None
        self._ck(self._lib.g3_dist_solve(self._h, _d(delta), C.cast(C.byref(beta), _dp), _d(u), C.byref(ms)), "g3_dist_solve")
        return {"beta": float(beta.value), "u": u, "ms_solve": float(ms.value)}

    def dist_posterior(self, Xs, noise=False):
        Xs = _f64(Xs)
        if Xs.ndim == 1:
            Xs = Xs[:, None]
        if Xs.shape[1] != self.D:
            raise ValueError("dist_posterior: Xs must have D = %d columns" % self.D)
        M = Xs.shape[0]
        mean, var = np.empty(M), np.empty(M)
        self._ck(self._lib.g3_dist_posterior(self._h, _d(Xs), M, POST_NOISE if noise else 0, _d(mean), _d(var)), "g3_dist_posterior")
        return mean, var

    def dist_grad(self, n_theta, cfac=1.0, want_ddelta=True):
        dth = np.zeros(max(int(n_theta), 1))
        ddl = np.empty(self.N) if want_ddelta else None
        ms = (C.c_float * 3)()
        self._ck(self._lib.g3_dist_grad(self._h, float(cfac), _d(dth), _d(ddl), ms), "g3_dist_grad")
        return {"dtheta": dth[:n_theta], "ddelta": ddl, "ms_alpha": float(ms[0]), "ms_inverse": float(ms[1]), "ms_contract": float(ms[2])}

    def dist_residual(self, nvec=4, seed=1234):
        out = np.zeros(4)
        self._ck(self._lib.g3_dist_residual(self._h, int(nvec), int(seed), _d(out)), "g3_dist_residual")
        return out[:nvec]

    def dist_read_piece(self, J, nb):
        cnt = C.c_int()
        self._ck(self._lib.g3_dist_read_piece(self._h, int(J), None, C.cast(C.byref(cnt), _ip)), "g3_dist_read_piece")
        if cnt.value == 0:
            return None
        out = np.empty((cnt.value * nb, nb))
        self._ck(self._lib.g3_dist_read_piece(self._h, int(J), _d(out), C.cast(C.byref(cnt), _ip)), "g3_dist_read_piece")
        return out

    def dist_free(self):
        self._ck(self._lib.g3_dist_free(self._h), "g3_dist_free")

    # ---- big-matrix Cholesky
    def gram_potrf_device(self, desc, theta):
        theta = _f64(theta).ravel()
        if theta.size != desc.n_theta:
            raise ValueError("gram_potrf_device: theta has %d entries, the descriptor takes %d" % (theta.size, desc.n_theta))
        ld = C.c_double()
        info = C.c_int()
        mg = C.c_float()
        mp = C.c_float()
        self._resident = None
        self._ck(self._lib.g3_gram_potrf_device(self._h, C.byref(desc), _d(theta), C.cast(C.byref(ld), _dp),
                                                C.cast(C.byref(info), _ip), C.byref(mg), C.byref(mp)),
                 "g3_gram_potrf_device")
        return {"logdet": float(ld.value), "info": int(info.value), "ms_gram": float(mg.value), "ms_potrf": float(mp.value)}

"""Process API: GaussianProcess / WarpedGaussianProcess / StudentTProcess / WarpedStudentTProcess.

Host-side mirror of g3py/processes/{stochastic,elliptical,gaussian,studentT}.py for the exact-GP hot
path.  The reference compiles one Theano function per method (`_method_name`,
stochastic.py:385-430); here every method assembles O(N) host terms (mean, warping, scalar
constants) around ONE call into libg3b.so, which does the Gram matrix, Cholesky, solves and the
analytic gradient on the GPU for a whole batch of hyper samples.  No CPU fallback exists: without
the library or a B200 the calls raise.
"""
import math
import warnings

import numpy as np
from scipy import stats
from scipy.special import gammaln, digamma

from . import _cabi as cabi
from .hypers import Registry, HyperVar, Freedom
from .hypers.kernels import KernelSum, KernelNoise, DescBuilder, Kernel
from .hypers.mappings import Identity

__all__ = ["StochasticProcess", "EllipticalProcess", "GaussianProcess", "WarpedGaussianProcess", "StudentTProcess",
           "WarpedStudentTProcess", "GP", "WGP", "TP", "WTP", "Consts", "kernel_cov", "get_context"]

f32 = np.float32
_CONTEXTS = {}
_CONTEXTS_LOCK = __import__("threading").Lock()
_TOKENS = __import__("itertools").count(1)     # unique per process object (id() can be recycled after GC)
_TLS = __import__("threading").local()         # per-thread scratch (the (B, N) delta buffer handed to the device call)
_VERSIONS = __import__("itertools").count(1)   # unique per data set ever installed (never reused after a restore)


def get_context(device=0):
    """Per-process, per-THREAD, per-device context, created on first use.  Fork-safe (no CUDA state before the first
    call in a process; emcee / PyMC3 fork workers, stochastic.py:775-783) and thread-safe: a g3_ctx owns its streams and
    workspaces and is not re-entrant, while ctypes releases the GIL during a call, so every thread of a threaded caller
    (emcee `threads > 1`, bayesian/average.py:29,36) gets its own context instead of racing on a shared one."""
    import os
    import threading
    key = (os.getpid(), threading.get_ident(), int(device))
    ctx = _CONTEXTS.get(key)
    if ctx is None:
        with _CONTEXTS_LOCK:
            ctx = _CONTEXTS.get(key)
            if ctx is None:
                ctx = cabi.Context(device)
                _CONTEXTS[key] = ctx
    return ctx


class Consts:
    """Scalar constants of the reference graph.  strict=True keeps the float32-rounded literals that
    survive in an fp64 run of g3py (gaussian.py:218, studentT.py:122-128, tensors.py:98,204,221)."""

    def __init__(self, strict=True):
        self.strict = strict
        if strict:
            self.log_2pi = float(f32(math.log(float(f32(2.0 * np.pi)))))
            self.pi = float(f32(np.pi))
            self.log_2pi_student = float(f32(math.log(float(f32(2.0) * f32(np.pi)))))
            self.jitter = float(f32(1e-6))
            self.guard = float(f32(-1e30))
            self.fallback = float(f32(1e-10))
        else:
            self.log_2pi = math.log(2.0 * math.pi)
            self.pi = math.pi
            self.log_2pi_student = math.log(2.0 * math.pi)
            self.jitter = 1e-6
            self.guard = -1e30
            self.fallback = 1e-10


class DictObj(dict):
    """libs/__init__.py:17-44 — dict with attribute access."""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__

    def clone(self):
        return DictObj(self.copy())


_HERMGAUSS10 = np.polynomial.hermite.hermgauss(10)     # gaussian.py:127-174 uses 10 nodes; computing them costs 0.45 ms per call


def _gauss_hermite_moments(T, mu, sd, block=8192):
    """E[T(z)] and E[T(z)^2] for z ~ N(mu, sd^2) by the reference's 10-node Gauss-Hermite rule (gaussian.py:127-174:
    grid mu + sd sqrt(2) a_k, weights w_k / sqrt(pi)).  The grid is walked node by node in blocks of points, so every
    temporary stays below glibc's mmap threshold: the one-shot (10, M) formulation spends most of its time in page faults
    of fresh 800 KB arrays (3.6 ms for M = 10 000 against 1 ms)."""
    a, w = _HERMGAUSS10
    M = len(mu)
    m1, m2 = np.empty(M), np.empty(M)
    c = 1.0 / np.sqrt(np.pi)
    r2 = np.sqrt(2.0)
    for i0 in range(0, M, block):
        sl = slice(i0, min(i0 + block, M))
        mu_b, sd_b = mu[sl], sd[sl] * r2
        s1 = np.zeros(len(mu_b))
        s2 = np.zeros(len(mu_b))
        for k in range(len(a)):
            tg = np.asarray(T(mu_b + sd_b * a[k]), dtype=np.float64)
            s1 += w[k] * tg
            tg *= tg
            s2 += w[k] * tg
        m1[sl] = s1 * c
        m2[sl] = s2 * c
    return m1, m2


def tt_to_num(r, nan=0.0, inf=1e10):
    """libs/tensors.py:90-92 applied to host vectors (the gradient scrub of stochastic.py:308-309)."""
    r = np.asarray(r, dtype=np.float64)
    if np.isfinite(r).all():                 # nothing to scrub (the usual case; saves two np.where per call on the hot path)
        return r
    return np.where(np.isnan(r), nan, np.where(np.isinf(r), inf, r))


class StochasticProcess:
    """stochastic.py:20-201 — data handling, parameter dictionaries, bijection."""

    def __init__(self, space=None, order=None, inputs=None, outputs=None, hidden=None, index=None, name="SP",
                 device=0, strict_constants=True, reference_nan_quirk=False, **kwargs):
        ndim = 1
        if space is not None:
            if hasattr(space, "shape"):
                if len(space.shape) > 1:
                    ndim = space.shape[1]
            else:
                ndim = int(space)
                space = None
        self.nspace = ndim
        self.name = name
        self.device = device
        self.consts = Consts(strict_constants)
        # True: return 0 for the gradient components the reference's autodiff loses to NaN (Matern rates, SINC
        # freq) instead of their analytic value -- literal drop-in behaviour for dlogp
        self.reference_nan_quirk = bool(reference_nan_quirk)
        # the reference's 2-point dummy dataset (stochastic.py:46-56)
        self.space = np.array([[0.0, 1.0]] * ndim).T
        self.inputs = np.array([[0.0, 1.0]] * ndim).T
        self.outputs = np.array([0.0, 1.0])
        self.order = np.array([0.0, 1.0])
        self.index = np.array([0.0, 1.0])
        self.hidden = None
        self.is_observed = False
        self.executed = {"logp": 0, "dlogp": 0, "predict": 0}
        self._data_version = 0
        self._token = next(_TOKENS)
        self.registry = Registry()
        self._check_hypers()
        self._define_process()
        self.set_space(space=space, hidden=hidden, order=order, inputs=inputs, outputs=outputs, index=index)
        self._params = None

    # ---- pickling: the device context is per process and never pickled (stochastic.py:107-119)
    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("_token", None)
        d.pop("_delta_buf", None)
        d.pop("_affine_cache", None)
        return d

    @property
    def ctx(self):
        ctx = get_context(self.device)
        if "_token" not in self.__dict__:
            self._token = next(_TOKENS)          # unpickled object
        tag = (self._token, self._data_version)
        if getattr(ctx, "_data_tag", None) != tag:
            ctx.set_jitter(self.consts.jitter, 20)
            ctx.set_data(self.inputs)
            ctx._data_tag = tag
        return ctx

    def set_space(self, space=None, hidden=None, order=None, inputs=None, outputs=None, index=None):
        # stochastic.py:150-185
        if space is not None:
            space = np.asarray(space, dtype=np.float64)
            if space.ndim < 2:
                space = space.reshape(len(space), 1)
            self.space = space
        if hidden is not None:
            self.hidden = np.asarray(hidden, dtype=np.float64).reshape(-1)
        if order is not None:
            self.order = np.asarray(order).reshape(-1)
        elif self.nspace == 1:
            self.order = self.space.reshape(len(self.space))
        if inputs is not None:
            inputs = np.asarray(inputs, dtype=np.float64)
            if inputs.ndim < 2:
                inputs = inputs.reshape(len(inputs), 1)
            self.inputs = np.ascontiguousarray(inputs)
            self._data_version = next(_VERSIONS)
        if outputs is not None:
            self.outputs = np.asarray(outputs, dtype=np.float64).reshape(-1)
        if index is not None:
            self.index = np.asarray(index).reshape(-1)
        elif self.nspace == 1:
            self.index = self.inputs.reshape(len(self.inputs))
        if len(self.order) != len(self.space):
            self.order = np.arange(len(self.space))
        if len(self.index) != len(self.inputs):
            self.index = np.arange(len(self.inputs))

    def _substituted(self, inputs=None, outputs=None):
        """Per-call `inputs=` / `outputs=` (the reference's lambda_method only substitutes the values for that one
        call, stochastic.py:385-430, and never touches self.inputs / self.outputs): the observed data are swapped
        for the duration of the `with` block and restored afterwards, the device copy follows the data version."""
        proc = self

        class _Swap:
            def __enter__(self_):
                self_.saved = None
                if inputs is None and outputs is None:
                    return proc
                self_.saved = {k: proc.__dict__.get(k) for k in ("inputs", "outputs", "index", "_data_version",
                                                                  "_affine_cache")}
                proc.set_space(inputs=inputs, outputs=outputs)
                return proc

            def __exit__(self_, *exc):
                if self_.saved is not None:
                    for k, v in self_.saved.items():
                        if v is None:
                            proc.__dict__.pop(k, None)
                        else:
                            proc.__dict__[k] = v
                return False
        return _Swap()

    def observed(self, inputs=None, outputs=None, order=None, index=None, hidden=None):
        # stochastic.py:187-201
        self.set_space(inputs=inputs, outputs=outputs, order=order, index=index, hidden=hidden)
        self.is_observed = not (inputs is None and outputs is None)
        self._params = None

    # ---- theta layout / bijection (bayesian/models.py:143-203)
    def _finish_layout(self):
        off = 0
        for v in self.registry.vars:
            v.offset = off
            off += v.size
        self.ndim = off
        self.positive_mask = np.zeros(off, dtype=bool)
        for v in self.registry.vars:
            self.positive_mask[v.offset:v.offset + v.size] = v.positive
        self._pos_idx = np.flatnonzero(self.positive_mask)

    @property
    def layout(self):
        return [(v.name, v.size, v.positive) for v in self.registry.vars]

    def dict_to_array(self, params):
        """DictToArrayBijection.map: dict keyed by transformed names (`*_log__`, log values) -> flat theta.
        Bare names are accepted too and taken as natural-space values."""
        theta = self.dict_to_array_default()
        for v in self.registry.vars:
            sl = slice(v.offset, v.offset + v.size)
            for key, is_log in ((v.tname, True), (v.name + "_log_", True), (v.name, False)):
                if key in params:
                    val = np.asarray(params[key], dtype=np.float64).reshape(-1)
                    if v.positive and not is_log:
                        val = np.log(val)
                    theta[sl] = val
                    break
        return theta

    def dict_to_array_default(self):
        return np.zeros(self.ndim)

    def array_to_dict(self, theta):
        """DictToArrayBijection.rmap."""
        theta = np.asarray(theta, dtype=np.float64)
        out = DictObj()
        for v in self.registry.vars:
            val = theta[v.offset:v.offset + v.size]
            out[v.tname] = float(val[0]) if v.scalar else val.copy()
        return out

    def natural(self, theta):
        out = np.array(theta, dtype=np.float64)          # a copy: the positive entries are overwritten with their exp
        idx = self._pos_idx
        if idx.size:
            with np.errstate(over="ignore"):
                out[..., idx] = np.exp(out[..., idx])
        return out

    @property
    def params_test(self):
        """model.test_point: Flat -> 0, FlatExp -> 1 (log value 0) (hypers/__init__.py:116-126)."""
        return self.array_to_dict(np.zeros(self.ndim))

    @property
    def params_default(self):
        """bayesian/models.py:174-182: test point overwritten with the components' default_hypers."""
        theta = np.zeros(self.ndim)
        for h, val in self.default_hypers().items():
            if not isinstance(h, HyperVar):
                continue
            val = np.broadcast_to(np.asarray(val, dtype=np.float64).reshape(-1), (h.size,))
            theta[h.offset:h.offset + h.size] = np.log(val) if h.positive else val
        return self.array_to_dict(theta)

    @property
    def params(self):
        if self._params is None:
            self._params = self.params_default if self.is_observed else self.params_test
        return self._params

    @params.setter
    def params(self, p):
        self._params = DictObj(p)

    def _theta(self, params, array):
        if params is None:
            return self.dict_to_array(self.params)
        if array or isinstance(params, np.ndarray):
            return np.asarray(params, dtype=np.float64)
        return self.dict_to_array(params)

    def default_hypers(self):
        return {}

    def _check_hypers(self):
        pass

    def _define_process(self):
        pass


class EllipticalProcess(StochasticProcess):
    """elliptical.py:18-107 — location + kernel (+ auto Noise) + mapping (+ degree)."""
    KIND = cabi.KIND_GAUSS
    WARPED = False

    def __init__(self, space=None, location=None, kernel=None, mapping=None, degree=None, noisy=True, var_noise=None,
                 *args, **kwargs):
        from .hypers.means import Zero
        self.f_location = location if location is not None else Zero()
        self.f_degree = degree
        self.f_mapping = mapping if mapping is not None else Identity()
        self.f_kernel = kernel
        self.noisy = bool(noisy)
        if noisy:                                               # elliptical.py:26-28
            self.f_kernel_noise = KernelSum(self.f_kernel, KernelNoise(name="Noise", var=var_noise))
        else:
            self.f_kernel_noise = self.f_kernel
        kwargs["space"] = space
        super().__init__(*args, **kwargs)

    def _check_hypers(self):
        # creation order = theta order (elliptical.py:35-52)
        x = self.inputs
        self.f_location.check_dims(x)
        self.f_kernel_noise.check_dims(x)
        self.f_mapping.check_dims(x)
        parent = self.name + "_"
        self.f_location.check_hypers(parent, self.registry)
        self.f_kernel_noise.check_hypers(parent, self.registry)
        self.f_mapping.check_hypers(parent, self.registry)
        for comp in (self.f_location, self.f_kernel_noise, self.f_mapping):      # elliptical.py:44-46
            comp.check_potential(self.registry)
        if self.f_degree is not None:
            self.f_degree.check_dims(None)
            self.f_degree.check_hypers(parent, self.registry)
            self.f_degree.check_potential(self.registry)
        self._finish_layout()

    def _define_process(self):
        b = DescBuilder(self.nspace)
        if self.noisy:
            self.f_kernel_noise.compile(b, process_noise=True)
        else:
            self.f_kernel_noise.compile(b)
        self.desc = b.finish()
        self._slots = b.slots
        bf = DescBuilder(self.nspace)                 # f_kernel alone: prior selectors with noise=False
        self.f_kernel.compile(bf)
        self.desc_f = bf.finish()
        self._slots_f = bf.slots

    def __getstate__(self):
        d = super().__getstate__()
        d.pop("desc", None)      # ctypes structures: rebuilt on load
        d.pop("desc_f", None)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._define_process()

    def default_hypers(self):
        x, y = self.inputs, self.outputs
        return {**self.f_location.default_hypers_dims(x, y), **self.f_kernel_noise.default_hypers_dims(x, y),
                **self.f_mapping.default_hypers_dims(x, y)}      # elliptical.py:54-58 (degree not included)

    # ---- pieces ----------------------------------------------------------------------------
    def _accessor(self, nat):
        def p(h):
            v = nat[h.offset:h.offset + h.size]
            return float(v[0]) if h.scalar else v
        return p

    def _slot_gather(self):
        """(theta indices, slot indices) when every free kernel hyper occupies exactly one slot of the compiled descriptor
        (no shared hypers): the slot <-> theta copies of the hot path become one fancy-indexed assignment each."""
        hit = self.__dict__.get("_slot_gather_cache")
        if hit is not None and hit[0] is self._slots:
            return hit[1]
        dst, src = [], []
        for h, off, size, const in self._slots:
            if h is not None:
                dst += list(range(h.offset, h.offset + h.size))
                src += list(range(off, off + size))
        out = (np.array(dst, dtype=np.intp), np.array(src, dtype=np.intp)) if len(set(dst)) == len(dst) else None
        self._slot_gather_cache = (self._slots, out)
        return out

    def _kernel_theta(self, nat2d, slots=None, n_theta=None):
        """(B, n_slots) natural-space kernel hypers in the slot order of the compiled descriptor."""
        if slots is None and n_theta is None:
            hit = self.__dict__.get("_theta_template")
            if hit is None or hit[0] is not self._slots:
                tmpl = np.zeros(max(self.desc.n_theta, 1))
                for h, off, size, const in self._slots:
                    if h is None:
                        tmpl[off:off + size] = const
                hit = self._theta_template = (self._slots, tmpl)
            gather = self._slot_gather()
            if gather is not None:
                th = np.tile(hit[1], (nat2d.shape[0], 1))
                th[:, gather[1]] = nat2d[:, gather[0]]
                return th[:, :self.desc.n_theta]
        slots = self._slots if slots is None else slots
        n_theta = self.desc.n_theta if n_theta is None else n_theta
        B = nat2d.shape[0]
        th = np.empty((B, max(n_theta, 1)))
        for h, off, size, const in slots:
            if h is None:
                th[:, off:off + size] = const
            else:
                th[:, off:off + size] = nat2d[:, h.offset:h.offset + h.size]
        return th[:, :n_theta]

    def _nu(self, nat2d):
        if self.f_degree is None:
            return None
        d = self.f_degree.degree
        deg = nat2d[:, d.offset] if isinstance(d, HyperVar) else np.full(nat2d.shape[0], float(d))
        return self.f_degree.bound + deg                                  # hypers/__init__.py:159-160

    def logprior_batch(self, Theta, nat=None):
        """Free-RV terms (Flat = 0, NonTransformLog barrier -inf at exp(theta) <= 1e-6) plus the `pm.Potential`
        regularisers (stochastic.py:300-306: logp = sum of RV terms + potentials, for prior and posterior alike)."""
        if nat is None:
            nat = self.natural(np.atleast_2d(Theta))
        bad = (nat[:, self._pos_idx] <= 1e-6).any(axis=1)
        lp = np.where(bad, -np.inf, 0.0)
        return lp + self._potentials(nat)[0] if self.registry.potentials else lp

    def _potentials(self, nat2d):
        """(sum of potentials (B,), d/d natural hypers (B, P)); hypers/__init__.py:94-109 on natural-space values."""
        val = np.zeros(nat2d.shape[0])
        g = np.zeros_like(nat2d)
        for _, kind, c, hs in self.registry.potentials:
            for h in hs:
                v = nat2d[:, h.offset:h.offset + h.size]
                if kind == "L1":
                    val += -c * np.abs(v).sum(axis=1)
                    g[:, h.offset:h.offset + h.size] += -c * np.sign(v)
                elif kind == "L2":
                    val += -c * (v ** 2).sum(axis=1)
                    g[:, h.offset:h.offset + h.size] += -2.0 * c * v
        return val, g

    def _affine_location(self, inputs):
        """(loc0 (N,), J (n_loc, N), idx (n_loc,)) when the location is affine in its free hypers (Zero, Bias, Linear
        and their sums / scalings are): m(X; p) = loc0 + p[idx] @ J, so a whole batch of locations is one GEMM and
        the gradient chain another.  None otherwise (MeanProd).  Checked numerically once per data set."""
        key = (self._data_version, id(self.f_location))
        hit = getattr(self, "_affine_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        loc_h = [h for h in self.f_location.hypers if isinstance(h, HyperVar)]
        idx = np.concatenate([np.arange(h.offset, h.offset + h.size) for h in loc_h]).astype(int) if loc_h else np.zeros(0, int)
        out = None
        try:
            rng = np.random.default_rng(0)
            z = np.zeros(self.ndim)
            loc0 = np.asarray(self.f_location(inputs, self._accessor(z)), dtype=np.float64)
            p1, p2 = z.copy(), z.copy()
            p1[idx], p2[idx] = rng.standard_normal(len(idx)), rng.standard_normal(len(idx))
            J = np.zeros((len(idx), len(loc0)))
            jl = self.f_location.jacobian(inputs, self._accessor(p1)) if loc_h else {}
            r = 0
            for h in loc_h:
                J[r:r + h.size] = np.atleast_2d(jl[h])
                r += h.size
            ok = True
            for p in (p1, p2, p1 + p2):
                direct = np.asarray(self.f_location(inputs, self._accessor(p)), dtype=np.float64)
                scale = max(np.max(np.abs(direct)), 1.0)
                ok = ok and np.max(np.abs(direct - (loc0 + p[idx] @ J))) <= 1e-13 * scale
            if ok:
                out = (loc0, J, idx)
        except Exception:
            out = None
        self._affine_cache = (key, out)
        return out

    def _host_terms(self, nat2d, inputs, outputs, want_grad):
        """delta (B,N) or (N,), det_m (B,), and the host Jacobians of the location / mapping hypers."""
        B = nat2d.shape[0]
        loc_h = [h for h in self.f_location.hypers if isinstance(h, HyperVar)]
        map_h = [h for h in self.f_mapping.hypers if isinstance(h, HyperVar)]
        varying = B > 1 and any(np.ptp(nat2d[:, h.offset:h.offset + h.size], axis=0).max() > 0 for h in loc_h + map_h)
        aff = self._affine_location(inputs) if not map_h else None
        if aff is not None:                        # no free warping hypers, affine location: no per-row Python work
            loc0, J, idx = aff
            key = (self._data_version, id(self.f_mapping), id(outputs))
            hit = self.__dict__.get("_minv_cache")
            if hit is not None and hit[0] == key:           # the warping has no free hypers: T^-1(y) and its log-Jacobian are constants
                minv, det = hit[1], hit[2]
            else:
                p0 = self._accessor(nat2d[0])
                with np.errstate(all="ignore"):
                    minv = np.asarray(self.f_mapping.inv(outputs, p0), dtype=np.float64)
                    det = float(self.f_mapping.logdet_dinv(outputs, p0))
                if not self.f_mapping.hypers:               # (a warping with FIXED numeric hypers is constant too, but keep it simple)
                    self._minv_cache = (key, minv, det)
            rows = nat2d if varying else nat2d[:1]
            tls = _TLS.__dict__.setdefault("delta_buf", {})   # reused per thread: a fresh (B, N) array costs more in page faults
            shp = (rows.shape[0], len(minv))                  # consumed by the device call before this thread's next use
            buf = tls.get(shp)
            if buf is None:
                if len(tls) > 4:
                    tls.clear()
                buf = tls[shp] = np.empty(shp)
            if 0 < len(idx) <= 4:                          # BLAS is slow on K <= 4: rank-1 broadcasts instead
                np.multiply(rows[:, idx[0]:idx[0] + 1], J[0], out=buf)
                for r in range(1, len(idx)):
                    buf += rows[:, idx[r]:idx[r] + 1] * J[r]
                buf += loc0
            elif len(idx):
                np.matmul(rows[:, idx], J, out=buf)
                buf += loc0
            else:
                buf[:] = loc0
            np.subtract(minv, buf, out=buf)
            return (buf if varying else buf[0]), np.full(B, det), ("affine", J, idx), varying
        if varying and map_h:
            # every row has its own warping hypers: one vectorised pass when the map offers it and the location is affine
            aff = self._affine_location(inputs)
            if aff is not None:
                loc0, J, idx = aff
                bt = self.f_mapping.batch_terms(outputs, lambda h: nat2d[:, h.offset], want_grad)
                if bt is not None:
                    inv, det_m, dinv, dld = bt
                    loc = loc0 + nat2d[:, idx] @ J if len(idx) else loc0
                    return inv - loc, det_m, ("batch", J, idx, dinv, dld), True
        rows = B if varying else 1
        delta = np.empty((rows, len(outputs)))
        det_m = np.empty(rows)
        jac = []
        for b in range(rows):
            p = self._accessor(nat2d[b])
            with np.errstate(all="ignore"):
                delta[b] = self.f_mapping.inv(outputs, p) - self.f_location(inputs, p)   # gaussian.py:208
                det_m[b] = self.f_mapping.logdet_dinv(outputs, p)
            if want_grad:
                with np.errstate(all="ignore"):
                    jl = self.f_location.jacobian(inputs, p) if loc_h else {}
                    dinv, dld = self.f_mapping.grads(outputs, p) if map_h else ({}, {})
                jac.append((jl, dinv, dld))
        if not varying:
            det_m = np.broadcast_to(det_m, (B,)).copy()
            return delta[0], det_m, jac, False
        return delta, det_m, jac, True

    def _eval_batch(self, Theta, inputs=None, outputs=None, want_grad=True, nan_quirk=None):
        """Core of logp/dlogp for a (B, P) array of theta (transformed space).  Returns
        (loglike (B,), dlogp (B,P) or None, info dict)."""
        Theta = np.atleast_2d(np.asarray(Theta, dtype=np.float64))
        if Theta.shape[1] != self.ndim:
            raise ValueError("theta has %d entries, the model has %d hypers" % (Theta.shape[1], self.ndim))
        if inputs is not None or outputs is not None:
            with self._substituted(inputs, outputs):
                return self._eval_batch(Theta, None, None, want_grad, nan_quirk)
        X, y = self.inputs, self.outputs
        N = len(y)
        B = Theta.shape[0]
        c = self.consts
        nat = self.natural(Theta)
        delta, det_m, jac, varying = self._host_terms(nat, X, y, want_grad)
        nu = self._nu(nat)
        thk = self._kernel_theta(nat)
        res = self.ctx.gp_logp_grad(self.desc, self.KIND, delta, thk, nu=nu, want_grad=want_grad)
        beta, logdet, st = res["beta"], res["logdet"], res["status"]
        clean = not st.any()                                              # no status bit anywhere: the usual case
        failed = None if clean else (st & cabi.ST_POTRF_FAILED) != 0
        if not clean and np.any(failed):      # CholeskyRobust.perform fallback L = 1e-10 * I (tensors.py:218-222)
            d2 = np.sum(np.atleast_2d(delta) ** 2, axis=1)
            d2 = np.broadcast_to(d2, (B,))
            beta = np.where(failed, d2 / c.fallback ** 2, beta)
            logdet = np.where(failed, N * math.log(c.fallback), logdet)
        n = float(N)
        with np.errstate(all="ignore"):
            if self.KIND == cabi.KIND_GAUSS:                              # gaussian.py:218-232
                ll = -0.5 * n * c.log_2pi + (-0.5 * beta) + (-logdet) + det_m
            else:                                                         # studentT.py:126-137
                r1 = -0.5 * (nu + n) * np.log1p(beta / (nu - 2.0))
                r2 = np.where(nu >= 1e6, -n * 0.5 * c.log_2pi_student,
                              gammaln((nu + n) * 0.5) - gammaln(nu * 0.5) - 0.5 * n * np.log((nu - 2.0) * c.pi))
                ll = r1 + r2 + (-logdet) + det_m
            # guards (gaussian.py:234-241): non-finite delta / det_m / L / lcho -> float32(-1e30)
            finite = np.isfinite(det_m + beta + logdet)                   # any non-finite term makes the sum non-finite
        if clean and finite.all():
            bad = None
        else:
            if failed is None:
                failed = np.zeros(B, dtype=bool)
            dev_bad = ((st & cabi.ST_NONFINITE_RESULT) != 0) & ~failed      # fallback items: L = 1e-10*I is finite
            bad = dev_bad | ~np.isfinite(det_m) | ~np.isfinite(beta) | ~np.isfinite(logdet)
            ll = np.where(bad, c.guard, ll)
        info = {"beta": beta, "logdet": logdet, "det_m": det_m, "status": st, "nu": nu, "nat": nat}
        self.executed["logp"] += B
        if not want_grad:
            return ll, None, info
        self.executed["dlogp"] += B
        if not clean and np.any(st & cabi.ST_DIAG_SHIFT):
            # tensors.py:95-98 differentiates through m = min(diag K); the shift enters the gradient here as a constant
            warnings.warn("dlogp: the tt_to_cov diagonal shift is active (min diag K <= 0) for %d item(s); the gradient "
                          "treats the shift as a constant (logp parity only in this regime)"
                          % int(np.count_nonzero(st & cabi.ST_DIAG_SHIFT)), RuntimeWarning, stacklevel=3)
        g_nat = np.zeros((B, self.ndim))
        dth, ddl = res["dtheta"], res["ddelta"]
        gather = self._slot_gather()
        if gather is not None:                                            # every free hyper sits in exactly one slot
            g_nat[:, gather[0]] = dth[:, gather[1]]
        else:
            for h, off, size, const in self._slots:
                if h is not None:
                    g_nat[:, h.offset:h.offset + h.size] += dth[:, off:off + size]
        if isinstance(jac, tuple):                                        # affine location: d delta / d loc = -J
            J, idx = jac[1], jac[2]
            if len(idx):
                g_nat[:, idx] += -(ddl @ J.T)
            if jac[0] == "batch":                                         # vectorised warping terms (scalar hypers)
                for h, Jm in jac[3].items():
                    g_nat[:, h.offset] += np.einsum("bn,bn->b", Jm, ddl) + jac[4][h]
            jac = []
        for b in range(B if jac else 0):
            jl, dinv, dld = jac[b if varying else 0]
            gd = ddl[b]
            for h, J in jl.items():                                       # d delta / d loc = -J
                g_nat[b, h.offset:h.offset + h.size] += -(np.atleast_2d(J) @ gd)
            for h, J in dinv.items():                                     # scalar hypers: J (N,); vector: (size, N)
                g_nat[b, h.offset:h.offset + h.size] += np.atleast_2d(J) @ gd + np.atleast_1d(dld[h])
        if self.KIND == cabi.KIND_STUDENT and isinstance(self.f_degree.degree, HyperVar):
            with np.errstate(all="ignore"):
                bn = beta / (nu - 2.0)
                d_r1 = -0.5 * np.log1p(bn) + 0.5 * (nu + n) * bn / ((nu - 2.0) * (1.0 + bn))
                d_r2 = np.where(nu >= 1e6, 0.0,
                                0.5 * digamma((nu + n) * 0.5) - 0.5 * digamma(nu * 0.5) - 0.5 * n / (nu - 2.0))
            g_nat[:, self.f_degree.degree.offset] += d_r1 + d_r2
        g = g_nat                                                         # chain rule through exp, in place
        g[:, self._pos_idx] *= nat[:, self._pos_idx]
        if bad is not None:
            g[failed | bad] = 0.0
        if self.registry.potentials:                                      # potentials do not depend on the data
            gp_nat = self._potentials(nat)[1]
            g = g + np.where(self.positive_mask[None, :], gp_nat * nat, gp_nat)
        if self.reference_nan_quirk if nan_quirk is None else nan_quirk:
            for h in self.f_kernel_noise.nan_quirk_hypers():
                g[:, h.offset:h.offset + h.size] = 0.0
        return ll, tt_to_num(g), info                                     # stochastic.py:308-309

    # ---- public methods (stochastic.py:365-366 binds th_logp/th_dlogp/th_loglike) -----------------
    def logp(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
             array=False):
        theta = self._theta(params, array)
        lp = float(self.logprior_batch(theta)[0])
        if prior or not self.is_observed and inputs is None:
            return lp
        ll, _, _ = self._eval_batch(theta, inputs, outputs, want_grad=False)
        return lp + float(ll[0])

    def loglike(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                array=False):
        ll, _, _ = self._eval_batch(self._theta(params, array), inputs, outputs, want_grad=False)
        return float(ll[0])

    def dlogp(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
              array=False, reference_nan_quirk=None):
        theta = self._theta(params, array)
        if prior:                       # free RVs are flat: only the potentials have a gradient
            nat = self.natural(np.atleast_2d(theta))
            gp_nat = self._potentials(nat)[1]
            return np.where(self.positive_mask[None, :], gp_nat * nat, gp_nat)[0]
        _, g, _ = self._eval_batch(theta, inputs, outputs, want_grad=True, nan_quirk=reference_nan_quirk)
        return g[0]

    def logp_dlogp(self, theta, reference_nan_quirk=None):
        ll, g, info = self._eval_batch(theta, want_grad=True, nan_quirk=reference_nan_quirk)
        return float(self.logprior_batch(theta, info["nat"])[0] + ll[0]), g[0]

    # batched entries replacing the per-theta loops of stochastic.py:515-564
    def logp_batch(self, Theta, prior=False):
        Theta = np.atleast_2d(Theta)
        lp = self.logprior_batch(Theta)
        if prior:
            return lp
        ll, _, _ = self._eval_batch(Theta, want_grad=False)
        return lp + ll

    def dlogp_batch(self, Theta, reference_nan_quirk=None):
        _, g, _ = self._eval_batch(np.atleast_2d(Theta), want_grad=True, nan_quirk=reference_nan_quirk)
        return g

    def logp_dlogp_batch(self, Theta, reference_nan_quirk=None):
        Theta = np.atleast_2d(Theta)
        ll, g, info = self._eval_batch(Theta, want_grad=True, nan_quirk=reference_nan_quirk)
        return self.logprior_batch(Theta, info["nat"]) + ll, g, info

    def logp_chain(self, chain, prior=False):
        """stochastic.py:515-520, vectorised (the reference's own TODO)."""
        return self.logp_batch(np.asarray(chain), prior=prior)

    # ---- posterior --------------------------------------------------------------------------
    def _posterior(self, theta, space, noise=False, cov=False, prior=False):
        nat = self.natural(theta)
        p = self._accessor(nat)
        X, y = self.inputs, self.outputs
        thk = self._kernel_theta(nat[None, :])[0]
        loc_s = self.f_location(space, p)
        out = {}
        if prior:
            if noise or not self.noisy:
                K, _ = self.ctx.gram(self.desc, space, None, thk[None, :])
            else:
                K, _ = self.ctx.gram(self.desc_f, space, None,
                                     self._kernel_theta(nat[None, :], self._slots_f, self.desc_f.n_theta))
            K = K[0]
            if noise:
                m = np.min(np.diag(K))
                if not m > 0.0:
                    K = K + (self.consts.jitter - m) * np.eye(len(K))        # tt_to_cov, elliptical.py:70
            out.update(location=loc_s, kernel_diag=np.maximum(np.diag(K), 0.0), kernel=K if cov else None)
            return out, nat, p
        from .hypers.kernels import kernel_leaves
        for leaf in kernel_leaves(self.f_kernel):
            if getattr(leaf, "TRAINING_GRAM_ONLY", False):
                raise NotImplementedError("%s.cov(x1, x2) does not exist in the reference either: its two-argument form multiplies an "
                                          "N1xN1 by an N2xN2 matrix (hypers/kernels.py:351); only the training Gram (logp, gradient, "
                                          "prior kernel) is defined" % type(leaf).__name__)
        with np.errstate(all="ignore"):
            delta = tt_to_num(self.f_mapping.inv(y, p)) - self.f_location(X, p)   # elliptical.py:63,83
        r = self.ctx.gp_posterior(self.desc, space, delta, thk, noise=noise, cov=cov)
        # The reference's posterior LU-solves the raw K (elliptical.py:78-92); here it goes through the robust
        # Cholesky.  The two agree while K factors cleanly; make every deviation visible instead of silent.
        if r["status"] & cabi.ST_POTRF_FAILED:
            warnings.warn("posterior: the Cholesky of K failed even after the jitter ladder (tensors.py:203-213); "
                          "the returned moments are not finite", RuntimeWarning, stacklevel=3)
        elif r["status"] & cabi.ST_JITTER:
            warnings.warn("posterior: K needed jitter (%d ladder tries) to factor; the reference LU-solves the raw K "
                          "here, so moments differ at the level of the added jitter" % ((r["status"] >> 8) & 0xff),
                          RuntimeWarning, stacklevel=3)
        out.update(location=loc_s + r["mean"], kernel_diag=r["var"], kernel=r["cov"], beta=r["beta"], status=r["status"])
        self.executed["predict"] += 1
        return out, nat, p

    def _scaling(self, post, nat):
        return 1.0

    def _quantile_z(self, q, nat):
        return stats.norm.ppf(q)                                            # gaussian.py:69-71

    def predict(self, params=None, space=None, inputs=None, outputs=None, mean=True, std=True, var=False, cov=False,
                median=False, quantiles=False, quantiles_noise=False, samples=0, distribution=False, prior=False,
                noise=False, simulations=None, array=False):
        """stochastic.py:444-513 (mean / variance / std / covariance / median / quantiles)."""
        if inputs is not None or outputs is not None:          # per-call substitution, self.inputs/outputs untouched
            with self._substituted(inputs, outputs):
                return self.predict(params, space, None, None, mean, std, var, cov, median, quantiles, quantiles_noise,
                                    samples, distribution, prior, noise, simulations, array)
        theta = self._theta(params, array)
        if not self.is_observed:
            prior = True
        if space is None:
            space = self.space
        space = np.asarray(space, dtype=np.float64)
        if space.ndim < 2:
            space = space.reshape(len(space), 1)
        post, nat, p = self._posterior(theta, space, noise=noise, cov=cov, prior=prior)
        mu, kd = post["location"], post["kernel_diag"]
        sd = np.sqrt(kd)
        T = lambda v: self.f_mapping(v, p)
        scaling = 1.0 if prior else self._scaling(post, nat)
        values = DictObj()
        if self.WARPED:                                                      # gaussian.py:127-174
            if getattr(self.f_mapping, "elementwise_forward", False):
                m1, m2 = _gauss_hermite_moments(T, mu, sd)
            else:          # Newton-inverse warpings: the whole (10, M) grid in ONE call, as the reference hands it over
                a, w = _HERMGAUSS10
                grille = mu[None, :] + sd[None, :] * np.sqrt(2.0) * a[:, None]
                tg = T(grille.ravel()).reshape(grille.shape)
                m1 = w.dot(tg) / np.sqrt(np.pi)
                m2 = w.dot(tg ** 2) / np.sqrt(np.pi)
            v_mean, v_var = m1, m2 - m1 ** 2
        else:                                                                # elliptical.py:194-200, studentT.py:45-46
            v_mean, v_var = T(mu), kd * scaling
        if mean:
            values["mean"] = v_mean
        if var:
            values["variance"] = v_var
        if std:
            values["std"] = np.sqrt(v_var)
        if cov:
            values["covariance"] = post["kernel"] * scaling
        if median:
            values["median"] = T(mu)                                         # elliptical.py:190-192
        if quantiles:
            z = self._quantile_z(0.975, nat) if not prior else self._quantile_z_prior(0.975, nat)
            values["quantile_up"] = T(mu + z * sd)
            values["quantile_down"] = T(mu - z * sd)
        if quantiles_noise:
            pn, _, _ = self._posterior(theta, space, noise=True, cov=False, prior=prior)
            sdn = np.sqrt(pn["kernel_diag"])
            z = self._quantile_z(0.975, nat) if not prior else self._quantile_z_prior(0.975, nat)
            values["noise_std"] = sdn * math.sqrt(scaling)
            values["noise_up"] = T(pn["location"] + z * sdn)
            values["noise_down"] = T(pn["location"] - z * sdn)
        if samples > 0:
            values["samples"] = self.sampler(theta, space, samples=samples, prior=prior, noise=noise)
        if distribution:                                                     # stochastic.py:509-512
            values["logpredictive"] = lambda x: self.logpredictive(theta, space, vector=x, prior=prior, array=True)
        return values

    # ---- selectors bound by EllipticalProcess._compile_methods (elliptical.py:206-215, stochastic.py:330-366):
    #      called as f(params, space, inputs, outputs, prior=, noise=, array=) and returning NumPy arrays
    def _selector(self, params, space, inputs, outputs, prior, noise, array, cov=False, _given=False):
        if inputs is not None or outputs is not None:
            with self._substituted(inputs, outputs):
                return self._selector(params, space, None, None, prior, noise, array, cov,
                                      _given=True)
        theta = self._theta(params, array)
        if not self.is_observed and not _given:
            prior = True
        space = self.space if space is None else np.asarray(space, dtype=np.float64)
        if space.ndim < 2:
            space = space.reshape(len(space), 1)
        post, nat, p = self._posterior(theta, space, noise=noise, cov=cov, prior=prior)
        return post, nat, p, theta, space, prior

    def location(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                 array=False):
        """th_location (elliptical.py:121-129)."""
        return self._selector(params, space, inputs, outputs, prior, noise, array)[0]["location"]

    def kernel(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
               array=False):
        """th_kernel (elliptical.py:131-141): prior Gram on `space` or posterior covariance."""
        return self._selector(params, space, inputs, outputs, prior, noise, array, cov=True)[0]["kernel"]

    def kernel_diag(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                    array=False):
        """th_kernel_diag (elliptical.py:154-164), clamped at 0 (`:94-97`)."""
        return self._selector(params, space, inputs, outputs, prior, noise, array)[0]["kernel_diag"]

    def kernel_sd(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                  array=False):
        """th_kernel_sd (elliptical.py:166-176)."""
        return np.sqrt(self.kernel_diag(params, space, inputs, outputs, vector, prior, noise, array))

    def cholesky(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                 array=False):
        """th_cholesky (elliptical.py:143-152): CholeskyRobust of the prior / posterior covariance on `space`."""
        K = self.kernel(params, space, inputs, outputs, vector, prior, noise, array)
        L, info, _ = self.ctx.potrf_robust(K)
        return self.consts.fallback * np.eye(len(K)) if info < 0 else L

    def _moment(self, key, params, space, inputs, outputs, prior, noise, array):
        flags = dict(mean=False, std=False, var=False, median=False, cov=False)
        flags[{"variance": "var", "covariance": "cov"}.get(key, key)] = True
        theta = self._theta(params, array)
        return self.predict(theta, space, inputs, outputs, prior=prior, noise=noise, array=True, **flags)[key]

    def mean(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
             array=False, simulations=None):
        return self._moment("mean", params, space, inputs, outputs, prior, noise, array)

    def median(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
               array=False, simulations=None):
        return self._moment("median", params, space, inputs, outputs, prior, noise, array)

    def variance(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                 array=False, simulations=None):
        return self._moment("variance", params, space, inputs, outputs, prior, noise, array)

    def std(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
            array=False, simulations=None):
        return self._moment("std", params, space, inputs, outputs, prior, noise, array)

    def covariance(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                   array=False):
        return self._moment("covariance", params, space, inputs, outputs, prior, noise, array)

    def quantiler(self, params=None, space=None, inputs=None, outputs=None, q=0.975, prior=False, noise=False,
                  simulations=None, array=False):
        """gaussian.py:56-73 / studentT.py:51-55: T(location + z_q * kernel_sd)."""
        post, nat, p, _, _, prior = self._selector(params, space, inputs, outputs, prior, noise, array)
        z = self._quantile_z_prior(q, nat) if prior else self._quantile_z(q, nat)
        return self.f_mapping(post["location"] + z * np.sqrt(post["kernel_diag"]), p)

    def mapping(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                array=False):
        """th_mapping (elliptical.py:118-119): T(outputs)."""
        nat = self.natural(self._theta(params, array))
        y = self.outputs if outputs is None else np.asarray(outputs, dtype=np.float64)
        return tt_to_num(self.f_mapping(y, self._accessor(nat)))

    def mapping_inv(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                    array=False):
        """th_mapping_inv (elliptical.py:115-116): T^-1(outputs), NaN/inf scrubbed (`:63`)."""
        nat = self.natural(self._theta(params, array))
        y = self.outputs if outputs is None else np.asarray(outputs, dtype=np.float64)
        with np.errstate(all="ignore"):
            return tt_to_num(self.f_mapping.inv(y, self._accessor(nat)))

    def freedom(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                array=False):
        """th_freedom (elliptical.py:109-113): nu, plus N for the posterior."""
        if self.f_degree is None:
            raise AttributeError("freedom is defined for Student-t processes")
        nat = self.natural(self._theta(params, array))
        nu = float(self._nu(nat[None, :])[0])
        n = len(self.inputs if inputs is None else inputs)
        return nu if prior else nu + n

    def logpredictive(self, params=None, space=None, vector=None, prior=False, noise=False, array=False):
        """GaussianProcess.th_logpredictive (gaussian.py:42-54): logp_cho of `vector` under the predictive
        location with the DIAGONAL Cholesky of the noisy predictive variance (cho = diag(sd))."""
        if self.KIND != cabi.KIND_GAUSS:
            raise NotImplementedError("the reference defines th_logpredictive for GaussianProcess only")
        theta = self._theta(params, array)
        space = self.space if space is None else np.asarray(space, dtype=np.float64).reshape(len(space), -1)
        x = np.asarray(vector, dtype=np.float64).reshape(-1)
        loc, nat, p = self._posterior(theta, space, noise=noise, prior=prior)
        sdn = np.sqrt(self._posterior(theta, space, noise=True, prior=prior)[0]["kernel_diag"])
        c = self.consts
        with np.errstate(all="ignore"):
            delta = self.f_mapping.inv(x, p) - loc["location"]
            lcho = delta / sdn
            det_m = self.f_mapping.logdet_dinv(x, p)
            r = -0.5 * len(x) * c.log_2pi - 0.5 * float(lcho @ lcho) - float(np.sum(np.log(sdn))) + det_m
        bad = not (np.all(np.isfinite(delta)) and np.isfinite(det_m) and np.all(np.isfinite(sdn)) and np.all(np.isfinite(lcho)))
        return c.guard if bad else float(r)

    def _quantile_z_prior(self, q, nat):
        return self._quantile_z(q, nat)

    def sampler(self, theta, space, samples=1, prior=False, noise=False, rng=None):
        """gaussian.py:75-97 / studentT.py:57-67: location + chol(posterior covariance) @ randn (Student-t: the
        normal draws are scaled by inverse-gamma draws), mapped through T.  The Cholesky runs on the device."""
        rng = rng or np.random.default_rng()
        post, nat, p = self._posterior(theta, space, noise=noise, cov=True, prior=prior)
        L, info, _ = self.ctx.potrf_robust(post["kernel"])
        if info < 0:
            L = self.consts.fallback * np.eye(len(L))
        z = rng.standard_normal((len(space), samples))
        if self.KIND == cabi.KIND_STUDENT:
            free = float(self._nu(nat[None, :])[0]) + (0 if prior else len(self.outputs))
            z = z * stats.invgamma.rvs(a=free / 2.0, scale=(free - 2.0) / 2.0, size=samples, random_state=rng)
        f = post["location"][:, None] + L.dot(z)
        return np.array([self.f_mapping(k, p) for k in f.T]).T

    # ---- MAP (stochastic.py:566-674; optimiser wrappers bayesian/selection.py:14-42) --------
    def find_MAP(self, start=None, points=1, display=False, powell=False, bfgs=True, max_time=None, return_points=False,
                 **kwargs):
        import scipy.optimize as spo
        if start is None:
            start = self.params
        x0 = self.dict_to_array(start) if isinstance(start, dict) else np.asarray(start, dtype=np.float64)

        # BFGS's Wolfe line search asks for f and f' at the same trial point: one fused logp+grad evaluation serves both
        # (the reference compiles and calls `logp` and `dlogp` separately, stochastic.py:591-596)
        last = {}

        def both(x):
            key = np.asarray(x, dtype=np.float64).tobytes()
            if last.get("key") != key:
                lp, g = self.logp_dlogp(np.asarray(x, dtype=np.float64))
                last.update(key=key, lp=lp, g=g)
            return last["lp"], last["g"]

        def f(x):
            try:
                v = -(both(x)[0] if bfgs else self.logp(x, array=True))
                return 1e100 if np.isnan(v) else v                          # libs/__init__.py:61-62 nan_to_high
            except Exception:
                return 1e32

        def df(x):
            try:
                return np.nan_to_num(-both(x)[1])
            except Exception:
                return np.full_like(x, 1e32)

        def f_only(x):                     # derivative-free Powell steps: no gradient work
            try:
                v = -self.logp(x, array=True)
                return 1e100 if np.isnan(v) else v
            except Exception:
                return 1e32

        pts = [("start", -f(x0), x0)]
        x = x0
        for i in range(max(points, 1)):
            if bfgs:
                x = spo.fmin_bfgs(f, x, fprime=df, disp=display, **kwargs)
                pts.append(("bfgs", -f(x), x))
            if powell:
                x = spo.fmin_powell(lambda z: f_only(z), x, disp=display)
                pts.append(("powell", -f(x), x))
        best = max(pts, key=lambda t: t[1])
        params = self.array_to_dict(best[2])
        if return_points:
            return params, pts
        return params


    # ---- batched drivers: the loops of stochastic.py:566-800 on top of logp_batch ---------------
    def find_MAP_multistart(self, starts, display=False, **kwargs):
        """Multi-start MAP: every start is scored in ONE batched launch, the best few are polished by BFGS
        (stochastic.py:606-613 scores the list of starts one at a time)."""
        S = np.array([self.dict_to_array(s) if isinstance(s, dict) else np.asarray(s, dtype=np.float64) for s in starts])
        lp = self.logp_batch(S)
        order = np.argsort(-lp)
        best, best_lp = None, -np.inf
        for i in order[:max(1, min(3, len(order)))]:
            p = self.find_MAP(start=S[i], display=display, **kwargs)
            v = self.logp(p)
            if v > best_lp:
                best, best_lp = p, v
        return best

    def sample_hypers(self, start=None, samples=200, chains=None, noise_mult=0.1, noise_sum=0.01, seed=None, a=2.0):
        """Affine-invariant ensemble sampler (Goodman & Weare stretch move, what emcee's EnsembleSampler runs for
        bayesian/average.py:20-54) with each half-ensemble proposal evaluated as one `logp_batch` call instead of
        chains/2 sequential Theano calls.  Returns (chain [samples, chains, ndim], logp [samples, chains])."""
        rng = np.random.default_rng(seed)
        nd = self.ndim
        chains = 2 * nd if chains is None else int(chains)
        chains += chains % 2
        x0 = self.dict_to_array(self.params if start is None else start) if not isinstance(start, np.ndarray) else start
        # stochastic.py:745-752: walkers start in a small ball around the start point
        pos = x0[None, :] * (1.0 + noise_mult * rng.standard_normal((chains, nd))) + noise_sum * rng.standard_normal((chains, nd))
        lp = self.logp_batch(pos)
        out = np.empty((samples, chains, nd))
        out_lp = np.empty((samples, chains))
        half = chains // 2
        for it in range(samples):
            for first in (True, False):
                S = slice(0, half) if first else slice(half, chains)
                C = slice(half, chains) if first else slice(0, half)
                z = ((a - 1.0) * rng.random(half) + 1.0) ** 2 / a
                partner = pos[C][rng.integers(0, half, size=half)]
                prop = partner + z[:, None] * (pos[S] - partner)
                lp_prop = self.logp_batch(prop)                       # ONE launch for the whole half-ensemble
                with np.errstate(all="ignore"):
                    logr = (nd - 1.0) * np.log(z) + lp_prop - lp[S]
                acc = np.log(rng.random(half)) < logr
                acc &= np.isfinite(lp_prop)
                newpos = pos[S].copy()
                newpos[acc] = prop[acc]
                pos[S] = newpos
                newlp = lp[S].copy()
                newlp[acc] = lp_prop[acc]
                lp[S] = newlp
            out[it] = pos
            out_lp[it] = lp
        return out, out_lp


    def sample_hmc(self, start=None, samples=100, chains=8, step=0.02, n_leapfrog=10, noise_sum=0.01, seed=None):
        """Hamiltonian Monte Carlo over `chains` independent chains advanced in lockstep: every leapfrog step is
        ONE batched logp+gradient launch (the gradient PyMC3's HMC/NUTS obtains through the Op boundary,
        SURVEY §3 F).  One chain per GPU = the same call with chains=1 in each process.
        Returns (chain [samples, chains, ndim], logp [samples, chains], accept_rate [chains])."""
        rng = np.random.default_rng(seed)
        nd = self.ndim
        x0 = self.dict_to_array(self.params if start is None else start) if not isinstance(start, np.ndarray) else start
        q = x0[None, :] + noise_sum * rng.standard_normal((chains, nd))
        lp, g, _ = self.logp_dlogp_batch(q)
        out = np.empty((samples, chains, nd))
        out_lp = np.empty((samples, chains))
        n_acc = np.zeros(chains)
        for it in range(samples):
            p = rng.standard_normal((chains, nd))
            h0 = -lp + 0.5 * np.sum(p * p, axis=1)
            qn, pn, gn, lpn = q.copy(), p.copy(), g.copy(), lp.copy()
            for _ in range(n_leapfrog):
                pn = pn + 0.5 * step * gn
                qn = qn + step * pn
                lpn, gn, _ = self.logp_dlogp_batch(qn)
                gn = np.nan_to_num(gn)
                pn = pn + 0.5 * step * gn
            h1 = -lpn + 0.5 * np.sum(pn * pn, axis=1)
            with np.errstate(all="ignore"):
                acc = (np.log(rng.random(chains)) < (h0 - h1)) & np.isfinite(lpn)
            q[acc], g[acc], lp[acc] = qn[acc], gn[acc], lpn[acc]
            n_acc += acc
            out[it], out_lp[it] = q, lp
        return out, out_lp, n_acc / max(samples, 1)


class GaussianProcess(EllipticalProcess):
    KIND = cabi.KIND_GAUSS

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("name", "GP")
        super().__init__(*args, **kwargs)


class WarpedGaussianProcess(GaussianProcess):
    WARPED = True

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("name", "WGP")
        super().__init__(*args, **kwargs)


class StudentTProcess(EllipticalProcess):
    KIND = cabi.KIND_STUDENT

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("name", "TP")
        if kwargs.get("degree") is None:
            kwargs["degree"] = Freedom()                                    # studentT.py:21-22
        super().__init__(*args, **kwargs)

    def _scaling(self, post, nat):
        # studentT.py:36-43: (nu + beta - 2) / (nu + N - 2)
        nu = float(self._nu(nat[None, :])[0])
        return (nu + post["beta"] - 2.0) / (nu + len(self.outputs) - 2.0)

    def _quantile_z(self, q, nat):
        nu = float(self._nu(nat[None, :])[0])
        return stats.t.ppf(q, df=nu + len(self.outputs))                    # studentT.py:51-55 (posterior freedom)

    def _quantile_z_prior(self, q, nat):
        return stats.t.ppf(q, df=float(self._nu(nat[None, :])[0]))


class WarpedStudentTProcess(StudentTProcess):
    WARPED = True

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("name", "WTP")
        super().__init__(*args, **kwargs)


GP = GaussianProcess
WGP = WarpedGaussianProcess
TP = StudentTProcess
WTP = WarpedStudentTProcess


def kernel_cov(kernel, x1, x2=None, hypers=None, device=0):
    """Numeric `Kernel.cov(x1, x2)` (kernels.py:106-110) for a free-standing kernel expression.
    `hypers`: {bare hyper name: natural value}, e.g. {"SE_var": 1.0, "SE_rate": [1, 2]}."""
    x1 = np.asarray(x1, dtype=np.float64)
    if x1.ndim < 2:
        x1 = x1[:, None]
    reg = Registry()
    kernel.check_dims(x1)
    if not getattr(kernel, "_standalone_reg", None):
        kernel.check_hypers("", reg)
        kernel._standalone_reg = reg
    reg = kernel._standalone_reg
    b = DescBuilder(x1.shape[1])
    kernel.compile(b)
    desc = b.finish()
    th = np.ones(max(desc.n_theta, 1))
    hypers = hypers or {}
    for h, off, size, const in b.slots:
        if h is None:
            th[off:off + size] = const
        elif h.name in hypers:
            th[off:off + size] = np.asarray(hypers[h.name], dtype=np.float64).reshape(-1)
    K, _ = get_context(device).gram(desc, x1, x2, th[None, :desc.n_theta] if desc.n_theta else th[None, :0])
    return K[0]

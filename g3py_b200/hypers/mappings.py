"""Warpings T — mirror of the closed-form maps of g3py/processes/hypers/mappings.py.

O(N) elementwise work that stays on the host: `inv(y)` and `logdet_dinv(y)` feed `delta` and
`det_m` to the device path; `dinv`/`dlogdet` are the hyper-derivatives Theano's autodiff would
produce for them.  Hyper values come through `p(h)` (natural space).
"""
import numpy as np

from . import Hypers, HyperVar

__all__ = ["Mapping", "Identity", "LinearMapping", "LogShifted", "BoxCoxShifted", "BoxCoxLinear", "ArcsinhLinear",
           "SinhArcsinh", "MappingComposed", "MappingInvSum", "BoxCoxLinear2", "Logistic", "WarpingTanh", "WarpingBoxCox"]

_F32_1EM32 = float(np.float32(1e-32))
_F32_1EM5 = float(np.float32(1e-5))


def _v(p, h):
    return float(p(h)) if isinstance(h, HyperVar) else float(h)


def _vec(p, h, n):
    return np.broadcast_to(np.asarray(p(h) if isinstance(h, HyperVar) else h, dtype=np.float64).reshape(-1), (n,))


def _signed_root(sc, power):
    """sign(sc) |sc|^(1/power) as exp(log|sc| / power): about a third of the time of NumPy's float pow on the 1e5-element
    grids of the Gauss-Hermite moments (same value to a few ulp; 0 stays 0, inf stays inf)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sign(sc) * np.exp(np.log(np.abs(sc)) / power)


def _tt_to_num(r):
    return np.where(np.isnan(r), 0.0, np.where(np.isinf(r), 1e10, r))


class Mapping(Hypers):
    # True when the forward map acts on every element independently (closed forms).  The Newton-inverse warpings iterate on the
    # WHOLE vector with one stopping test (libs/tensors.py:134-171), so their value at a point depends on what else is in the
    # vector: callers that want the reference's numbers must hand them the same vectors the reference does.
    elementwise_forward = True

    def __call__(self, z, p):
        raise NotImplementedError

    def inv(self, y, p):
        raise NotImplementedError

    def logdet_dinv(self, y, p):
        raise NotImplementedError

    def grads(self, y, p):
        """({HyperVar: d inv/d h (n,)}, {HyperVar: d logdet/d h}) in natural space."""
        return {}, {}

    def batch_terms(self, y, col, want_grad):
        """Optional vectorised form of (inv, logdet_dinv, grads) for a batch of hyper rows: `col(h)` is the (B,) column of
        natural-space values of hyper h.  Returns (inv (B, n), logdet (B,), {h: d inv/d h (B, n)}, {h: d logdet/d h (B,)})
        or None when the map has no such form (the caller then loops over the rows)."""
        return None

    def dinv_dy(self, y, p):
        """d inv / d y (n,), needed to chain composed maps."""
        raise NotImplementedError

    def dlog_dinv_dy(self, y, p):
        """d/dy log|d inv / d y| (n,): how the log-Jacobian of this map reacts to a shift of its argument (needed when
        another map feeds it).  Closed forms in the subclasses; this default is a 4th-order central difference."""
        h = 1e-4 * np.maximum(1.0, np.abs(y))
        f = lambda t: np.log(np.abs(self.dinv_dy(t, p)))
        return (-f(y + 2 * h) + 8 * f(y + h) - 8 * f(y - h) + f(y - 2 * h)) / (12 * h)

    def _reg(self, parent, reg, attr, positive):
        h = getattr(self, attr)
        if h is None:
            h = (reg.FlatExp if positive else reg.Flat)(parent + self.name + "_" + attr)
            setattr(self, attr, h)
        if isinstance(h, HyperVar) and h not in self.hypers:
            self.hypers += [h]

    def __matmul__(self, other):
        return MappingComposed(self, other)
    __imatmul__ = __matmul__


class MappingComposed(Mapping):      # mappings.py:56-71
    @property
    def elementwise_forward(self):
        return self.m1.elementwise_forward and self.m2.elementwise_forward

    def __init__(self, m1, m2):
        self.m1, self.m2 = m1, m2
        self.hypers = []
        self.name = m1.name + " " + m2.name
        self.dims = None
        self.shape = None

    def check_hypers(self, parent="", reg=None):
        self.m1.check_hypers(parent=parent, reg=reg)
        self.m2.check_hypers(parent=parent, reg=reg)
        self.hypers = self.m1.hypers + self.m2.hypers

    def check_dims(self, x=None):
        self.m1.check_dims(x)
        self.m2.check_dims(x)

    def check_potential(self, reg=None):
        Hypers.check_potential(self, reg)

    def default_hypers_dims(self, x=None, y=None):
        return {**self.m1.default_hypers_dims(x, y), **self.m2.default_hypers_dims(x, y)}

    def __call__(self, z, p):
        return self.m1(self.m2(z, p), p)

    def inv(self, y, p):
        return self.m2.inv(self.m1.inv(y, p), p)

    def logdet_dinv(self, y, p):
        return self.m2.logdet_dinv(self.m1.inv(y, p), p) + self.m1.logdet_dinv(y, p)

    def dinv_dy(self, y, p):
        w = self.m1.inv(y, p)
        return self.m2.dinv_dy(w, p) * self.m1.dinv_dy(y, p)

    def dlog_dinv_dy(self, y, p):
        w = self.m1.inv(y, p)
        return self.m2.dlog_dinv_dy(w, p) * self.m1.dinv_dy(y, p) + self.m1.dlog_dinv_dy(y, p)

    def grads(self, y, p):
        # chain rule on the pieces: inv = m2.inv(w), logdet = m2.logdet(w) + m1.logdet(y), w = m1.inv(y);
        # m1's hypers move w, which moves m2's value (d m2.inv / d w) and m2's log-Jacobian (d log|d m2.inv/dw| / d w)
        w = self.m1.inv(y, p)
        di1, dl1 = self.m1.grads(y, p)
        di2, dl2 = self.m2.grads(w, p)
        d2dw = self.m2.dinv_dy(w, p)
        dlog = self.m2.dlog_dinv_dy(w, p)
        dinv = {k: d2dw * v for k, v in di1.items()}
        dld = {k: v + np.sum(dlog * di1[k], axis=-1) for k, v in dl1.items()}
        for k, v in di2.items():
            dinv[k] = v
        for k, v in dl2.items():
            dld[k] = v
        return dinv, dld


class MappingInvSum(MappingComposed):   # mappings.py:73-85: inv(y) = m1.inv(y) + m2.inv(y)
    """The reference leaves `__call__` empty (`pass`) and `logdet_dinv` commented out, so the density uses the base-class
    numeric log-Jacobian sum(log(diag(jacobian(inv)))) (`mappings.py:17-22`): for element-wise maps that is
    sum(log(m1.inv'(y) + m2.inv'(y))).  The forward map has no closed form; it is obtained like the Newton warpings
    (`inverse_function`, libs/tensors.py:134-171: damped Newton, step 0.1, tolerance 1e-3 on the whole vector)."""

    elementwise_forward = False          # forward map by the whole-vector Newton iteration

    def __init__(self, m1, m2):
        super().__init__(m1, m2)
        self.name = m1.name + " +^ " + m2.name

    def inv(self, y, p):
        return self.m1.inv(y, p) + self.m2.inv(y, p)

    def dinv_dy(self, y, p):
        return self.m1.dinv_dy(y, p) + self.m2.dinv_dy(y, p)

    def logdet_dinv(self, y, p):
        with np.errstate(invalid="ignore", divide="ignore"):
            return float(np.sum(np.log(self.dinv_dy(y, p))))

    def dlog_dinv_dy(self, y, p):
        d1, d2 = self.m1.dinv_dy(y, p), self.m2.dinv_dy(y, p)
        return (d1 * self.m1.dlog_dinv_dy(y, p) + d2 * self.m2.dlog_dinv_dy(y, p)) / (d1 + d2)

    def __call__(self, z, p, tol=1e-3, n_steps=1024, alpha=0.1):
        z = np.asarray(z, dtype=np.float64)
        y = z.copy()
        for _ in range(n_steps):                                           # libs/tensors.py:134-171
            r = self.inv(y, p) - z
            if np.sqrt(np.sum(r * r)) < tol:
                break
            y = y - alpha * r / self.dinv_dy(y, p)
        return y

    def grads(self, y, p):
        """d inv / d h is the owning map's; d logdet / d h = sum_i (d (m.inv')(y_i) / d h) / (m1.inv' + m2.inv'), with
        d m.inv' / d h = m.inv' * d log(m.inv') / d h taken from the owning map's own log-Jacobian gradient by a
        4th-order central difference in h (exact maps, scalar hypers: a handful of O(N) evaluations)."""
        di1, _ = self.m1.grads(y, p)
        di2, _ = self.m2.grads(y, p)
        dinv = dict(di1)
        dinv.update(di2)
        tot = self.dinv_dy(y, p)
        dld = {}
        for m, di in ((self.m1, di1), (self.m2, di2)):
            for h in di:
                dld[h] = _fd_hyper(lambda pp: m.dinv_dy(y, pp), p, h, weight=1.0 / tot)
        return dinv, dld


def _fd_hyper(fun, p, h, weight):
    """sum(weight * d fun / d h) by 4th-order central differences on the natural-space value of hyper h (per component)."""
    base = np.atleast_1d(np.asarray(p(h), dtype=np.float64)).copy()
    out = np.zeros(base.size)
    for c in range(base.size):
        step = 1e-4 * max(1.0, abs(base[c]))

        def at(s):
            v = base.copy()
            v[c] += s * step
            val = v if not getattr(h, "scalar", base.size == 1) else float(v[0])
            return fun(lambda q: val if q is h else p(q))
        d = (-at(2.0) + 8.0 * at(1.0) - 8.0 * at(-1.0) + at(-2.0)) / (12.0 * step)
        out[c] = float(np.sum(weight * d))
    return out if out.size > 1 else float(out[0])


class Identity(Mapping):             # mappings.py:88-99
    def __init__(self, y=None, name=None):
        super().__init__(y, name)

    def __call__(self, z, p=None):
        return z

    def inv(self, y, p=None):
        return y

    def logdet_dinv(self, y, p=None):
        return 0.0

    def dinv_dy(self, y, p=None):
        return np.ones_like(y)

    def dlog_dinv_dy(self, y, p=None):
        return np.zeros_like(y)


class LinearMapping(Mapping):        # mappings.py:102-125
    def __init__(self, y=None, name=None, shift=None, scale=None):
        super().__init__(y, name)
        self.shift, self.scale = shift, scale

    def check_hypers(self, parent="", reg=None):
        self._reg(parent, reg, "shift", False)
        self._reg(parent, reg, "scale", True)

    def default_hypers(self, x=None, y=None):
        return {h: v for h, v in ((self.shift, 0.0), (self.scale, 1.0)) if isinstance(h, HyperVar)}

    def __call__(self, z, p):
        return _v(p, self.scale) * (z - _v(p, self.shift))

    def inv(self, y, p):
        return y / _v(p, self.scale) + _v(p, self.shift)

    def logdet_dinv(self, y, p):
        return -float(y.shape[0]) * np.log(_v(p, self.scale))

    def dinv_dy(self, y, p):
        return np.ones_like(y) / _v(p, self.scale)

    def dlog_dinv_dy(self, y, p):
        return np.zeros_like(y)

    def grads(self, y, p):
        s = _v(p, self.scale)
        n = float(y.shape[0])
        return ({self.shift: np.ones_like(y), self.scale: -y / s ** 2}, {self.shift: 0.0, self.scale: -n / s})


class LogShifted(Mapping):           # mappings.py:128-149
    def __init__(self, y=None, name=None, shift=None):
        super().__init__(y, name)
        self.shift = shift

    def check_hypers(self, parent="", reg=None):
        self._reg(parent, reg, "shift", False)

    def default_hypers(self, x=None, y=None):
        return {self.shift: float(np.min(y)) - 1.0} if isinstance(self.shift, HyperVar) else {}

    def __call__(self, z, p):
        return np.exp(z) + _v(p, self.shift)

    def inv(self, y, p):
        return np.log(np.maximum(y - _v(p, self.shift), _F32_1EM32))

    def logdet_dinv(self, y, p):
        with np.errstate(invalid="ignore", divide="ignore"):
            return -np.sum(np.log(y - _v(p, self.shift)))

    def dinv_dy(self, y, p):
        return 1.0 / (y - _v(p, self.shift))

    def dlog_dinv_dy(self, y, p):
        return -1.0 / (y - _v(p, self.shift))

    def grads(self, y, p):
        sh = y - _v(p, self.shift)
        live = sh > _F32_1EM32
        with np.errstate(invalid="ignore", divide="ignore"):
            return {self.shift: np.where(live, -1.0 / sh, 0.0)}, {self.shift: float(np.sum(1.0 / sh))}


class _BoxCox(Mapping):
    def _params(self, p):
        raise NotImplementedError

    def inv(self, y, p):
        shift, scale, power, thr = self._params(p)
        sh = scale * (y + shift)
        with np.errstate(invalid="ignore", divide="ignore"):
            if power < thr:
                return np.log(sh)
            return (np.sign(sh) * np.abs(sh) ** power - 1.0) / power

    def dinv_dy(self, y, p):
        shift, scale, power, thr = self._params(p)
        sh = scale * (y + shift)
        return np.abs(sh) ** (power - 1.0) * scale

    def dlog_dinv_dy(self, y, p):
        shift, scale, power, thr = self._params(p)
        return (power - 1.0) / (y + shift)

    def _grads(self, y, p):
        shift, scale, power, thr = self._params(p)
        n = float(y.shape[0])
        sh = scale * (y + shift)
        a = np.abs(sh)
        with np.errstate(invalid="ignore", divide="ignore"):
            sp = np.sign(sh) * a ** power
            d_shift = a ** (power - 1.0) * scale
            d_scale = a ** (power - 1.0) * (y + shift)
            d_power = (sp * np.log(a) * power - (sp - 1.0)) / power ** 2
            ld_shift = (power - 1.0) * float(np.sum(1.0 / (y + shift)))
            ld_power = float(np.sum(np.log(a)))
        return d_shift, d_scale, d_power, ld_shift, power * n / scale, ld_power


def _boxcox_batch(y, shift, scale, power, thr, want_grad):
    """The Box-Cox family for (B, 1) columns of hypers: one log and one exp pass over the (B, n) block serve the inverse, the
    log-Jacobian and every derivative (|sh|^(power-1) = exp((power-1) log|sh|)).  None if any row takes the log branch."""
    if np.any(power < thr):
        return None
    ys = y[None, :] + shift
    sh = scale * ys
    a = np.abs(sh)
    with np.errstate(invalid="ignore", divide="ignore"):
        la = np.log(a)
        ap1 = np.exp((power - 1.0) * la)
        sp = np.sign(sh) * (ap1 * a)
        inv = (sp - 1.0) / power
        if not want_grad:
            return inv, la, None
        d_shift = ap1 * scale
        d_scale = ap1 * ys
        d_power = (sp * la * power - (sp - 1.0)) / (power * power)
        ld_shift = (power[:, 0] - 1.0) * np.sum(1.0 / ys, axis=1)
        ld_power = np.sum(la, axis=1)
    return inv, la, (d_shift, d_scale, d_power, ld_shift, ld_power)


class BoxCoxShifted(_BoxCox):        # mappings.py:152-179
    def __init__(self, y=None, name="BoxShift", shift=None, power=None):
        super().__init__(y, name)
        self.shift, self.power = shift, power

    def check_hypers(self, parent="", reg=None):
        self._reg(parent, reg, "shift", False)
        self._reg(parent, reg, "power", True)

    def default_hypers(self, x=None, y=None):
        return {h: 1.0 for h in (self.shift, self.power) if isinstance(h, HyperVar)}

    def _params(self, p):
        return _v(p, self.shift), 1.0, _v(p, self.power), 1e-5

    def __call__(self, z, p):
        shift, _, power, _ = self._params(p)
        sc = power * z + 1.0
        return _signed_root(sc, power) - shift

    def logdet_dinv(self, y, p):
        shift, _, power, _ = self._params(p)
        with np.errstate(invalid="ignore", divide="ignore"):
            return (power - 1.0) * np.sum(np.log(np.abs(y + shift)))

    def grads(self, y, p):
        d_shift, _, d_power, ld_shift, _, ld_power = self._grads(y, p)
        return {self.shift: d_shift, self.power: d_power}, {self.shift: ld_shift, self.power: ld_power}

    def batch_terms(self, y, col, want_grad):
        if not (isinstance(self.shift, HyperVar) and isinstance(self.power, HyperVar)):
            return None
        shift, power = col(self.shift)[:, None], col(self.power)[:, None]
        out = _boxcox_batch(np.asarray(y, dtype=np.float64), shift, 1.0, power, 1e-5, want_grad)
        if out is None:
            return None
        inv, la, d = out
        logdet = (power[:, 0] - 1.0) * np.sum(la, axis=1)          # scale = 1: log|y + shift| = la
        if d is None:
            return inv, logdet, {}, {}
        d_shift, _, d_power, ld_shift, ld_power = d
        return inv, logdet, {self.shift: d_shift, self.power: d_power}, {self.shift: ld_shift, self.power: ld_power}


class BoxCoxLinear(_BoxCox):         # mappings.py:182-215
    def __init__(self, y=None, name=None, shift=None, scale=None, power=None):
        super().__init__(y, name)
        self.shift, self.scale, self.power = shift, scale, power

    def check_hypers(self, parent="", reg=None):
        self._reg(parent, reg, "shift", False)
        self._reg(parent, reg, "scale", True)
        self._reg(parent, reg, "power", True)

    def default_hypers(self, x=None, y=None):
        return {h: 1.0 for h in (self.shift, self.scale, self.power) if isinstance(h, HyperVar)}

    def _params(self, p):
        return _v(p, self.shift), _v(p, self.scale), _v(p, self.power), _F32_1EM5

    def __call__(self, z, p):
        shift, scale, power, _ = self._params(p)
        sc = power * z + 1.0
        return _signed_root(sc, power) / scale - shift

    def logdet_dinv(self, y, p):
        shift, scale, power, _ = self._params(p)
        with np.errstate(invalid="ignore", divide="ignore"):
            return (power - 1.0) * np.sum(np.log(np.abs(scale * (y + shift)))) + float(y.shape[0]) * np.log(scale)

    def grads(self, y, p):
        d_shift, d_scale, d_power, ld_shift, ld_scale, ld_power = self._grads(y, p)
        return ({self.shift: d_shift, self.scale: d_scale, self.power: d_power},
                {self.shift: ld_shift, self.scale: ld_scale, self.power: ld_power})


class BoxCoxLinear2(Mapping):        # mappings.py:218-251: shifted = scale * y + shift (BoxCoxLinear: scale * (y + shift))
    def __init__(self, y=None, name=None, shift=None, scale=None, power=None):
        super().__init__(y, name)
        self.shift, self.scale, self.power = shift, scale, power

    def check_hypers(self, parent="", reg=None):
        self._reg(parent, reg, "shift", False)
        self._reg(parent, reg, "scale", True)
        self._reg(parent, reg, "power", True)

    def default_hypers(self, x=None, y=None):
        return {h: 1.0 for h in (self.shift, self.scale, self.power) if isinstance(h, HyperVar)}

    def _params(self, p):
        return _v(p, self.shift), _v(p, self.scale), _v(p, self.power)

    def __call__(self, z, p):
        shift, scale, power = self._params(p)
        sc = power * z + 1.0
        return (_signed_root(sc, power) - shift) / scale

    def inv(self, y, p):
        shift, scale, power = self._params(p)
        sh = scale * y + shift
        with np.errstate(invalid="ignore", divide="ignore"):
            if power < _F32_1EM5:
                return np.log(sh)
            return (np.sign(sh) * np.abs(sh) ** power - 1.0) / power

    def logdet_dinv(self, y, p):
        shift, scale, power = self._params(p)
        e = -1.0 if power < _F32_1EM5 else power - 1.0
        with np.errstate(invalid="ignore", divide="ignore"):
            return e * np.sum(np.log(np.abs(scale * y + shift))) + float(y.shape[0]) * np.log(scale)

    def dinv_dy(self, y, p):
        shift, scale, power = self._params(p)
        return np.abs(scale * y + shift) ** (power - 1.0) * scale

    def dlog_dinv_dy(self, y, p):
        shift, scale, power = self._params(p)
        return (power - 1.0) * scale / (scale * y + shift)

    def grads(self, y, p):
        shift, scale, power = self._params(p)
        n = float(y.shape[0])
        sh = scale * y + shift
        a = np.abs(sh)
        with np.errstate(invalid="ignore", divide="ignore"):
            sp = np.sign(sh) * a ** power
            d_shift = a ** (power - 1.0)
            d_scale = a ** (power - 1.0) * y
            d_power = (sp * np.log(a) * power - (sp - 1.0)) / power ** 2
            ld_shift = (power - 1.0) * float(np.sum(1.0 / sh))
            ld_scale = (power - 1.0) * float(np.sum(y / sh)) + n / scale
            ld_power = float(np.sum(np.log(a)))
        return ({self.shift: d_shift, self.scale: d_scale, self.power: d_power},
                {self.shift: ld_shift, self.scale: ld_scale, self.power: ld_power})


class ArcsinhLinear(Mapping):        # mappings.py:309-333
    def __init__(self, y=None, name=None, shift=None, scale=None):
        super().__init__(y, name)
        self.shift, self.scale = shift, scale

    def check_hypers(self, parent="", reg=None):
        self._reg(parent, reg, "shift", False)
        self._reg(parent, reg, "scale", True)

    def default_hypers(self, x=None, y=None):
        return {h: v for h, v in ((self.shift, float(np.mean(y))), (self.scale, float(np.std(y)))) if isinstance(h, HyperVar)}

    def __call__(self, z, p):
        return np.sinh((z - _v(p, self.shift)) / _v(p, self.scale))

    def inv(self, y, p):
        return np.arcsinh(y) * _v(p, self.scale) + _v(p, self.shift)

    def logdet_dinv(self, y, p):
        return float(y.shape[0]) * np.log(_v(p, self.scale)) - 0.5 * np.sum(np.log1p(y ** 2))

    def dinv_dy(self, y, p):
        return _v(p, self.scale) / np.sqrt(1.0 + y ** 2)

    def dlog_dinv_dy(self, y, p):
        return -y / (1.0 + y ** 2)

    def grads(self, y, p):
        n = float(y.shape[0])
        return ({self.shift: np.ones_like(y), self.scale: np.arcsinh(y)}, {self.shift: 0.0, self.scale: n / _v(p, self.scale)})


class SinhArcsinh(Mapping):          # mappings.py:336-358
    def __init__(self, y=None, name=None, shift=None, scale=None):
        super().__init__(y, name)
        self.shift, self.scale = shift, scale

    def check_hypers(self, parent="", reg=None):
        self._reg(parent, reg, "shift", False)
        self._reg(parent, reg, "scale", True)

    def default_hypers(self, x=None, y=None):
        return {h: v for h, v in ((self.shift, 0.0), (self.scale, 1.0)) if isinstance(h, HyperVar)}

    def __call__(self, z, p):
        return np.sinh((np.arcsinh(z) - _v(p, self.shift)) / _v(p, self.scale))

    def inv(self, y, p):
        return np.sinh(_v(p, self.shift) + _v(p, self.scale) * np.arcsinh(y))

    def logdet_dinv(self, y, p):
        w = _v(p, self.shift) + _v(p, self.scale) * np.arcsinh(y)
        return (np.sum(np.log(np.cosh(w))) + float(y.shape[0]) * np.log(_v(p, self.scale)) - 0.5 * np.sum(np.log1p(y ** 2)))

    def dinv_dy(self, y, p):
        w = _v(p, self.shift) + _v(p, self.scale) * np.arcsinh(y)
        return np.cosh(w) * _v(p, self.scale) / np.sqrt(1.0 + y ** 2)

    def dlog_dinv_dy(self, y, p):
        w = _v(p, self.shift) + _v(p, self.scale) * np.arcsinh(y)
        return np.tanh(w) * _v(p, self.scale) / np.sqrt(1.0 + y ** 2) - y / (1.0 + y ** 2)

    def grads(self, y, p):
        s = _v(p, self.scale)
        ash = np.arcsinh(y)
        w = _v(p, self.shift) + s * ash
        n = float(y.shape[0])
        return ({self.shift: np.cosh(w), self.scale: np.cosh(w) * ash},
                {self.shift: float(np.sum(np.tanh(w))), self.scale: float(np.sum(np.tanh(w) * ash)) + n / s})


class Logistic(Mapping):             # mappings.py:363-397
    def __init__(self, y=None, name=None, lower=None, high=None, location=None, scale=None):
        super().__init__(y, name)
        self.lower, self.high, self.location, self.scale = lower, high, location, scale

    def check_hypers(self, parent="", reg=None):
        self._reg(parent, reg, "lower", False)
        self._reg(parent, reg, "high", True)
        self._reg(parent, reg, "location", False)
        self._reg(parent, reg, "scale", True)

    def default_hypers(self, x=None, y=None):
        y = np.asarray(y, dtype=np.float64)
        d = {self.lower: 1.5 * np.min(y) - 0.5 * np.max(y), self.high: 2.0 * (np.max(y) - np.min(y)),
             self.location: np.mean(y), self.scale: np.std(y)}
        return {h: float(v) for h, v in d.items() if isinstance(h, HyperVar)}

    def _p(self, y, p):
        lo, hi = _v(p, self.lower), _v(p, self.high)
        return np.where(y < lo, 0.0, np.where(y > lo + hi, 1.0, (y - lo) / hi))

    def __call__(self, z, p):
        return _v(p, self.lower) + _v(p, self.high) * (0.5 + 0.5 * np.tanh((z - _v(p, self.location)) / (2 * _v(p, self.scale))))

    def inv(self, y, p):
        with np.errstate(all="ignore"):
            q = self._p(y, p)
            return _v(p, self.location) + _v(p, self.scale) * _tt_to_num(np.log(q / (1 - q)))

    def logdet_dinv(self, y, p):
        with np.errstate(all="ignore"):
            q = self._p(y, p)
            return float(np.sum(_tt_to_num(np.log(_v(p, self.scale) / (_v(p, self.high) * q * (1 - q))))))

    def dinv_dy(self, y, p):
        q = self._p(y, p)
        with np.errstate(all="ignore"):
            return _tt_to_num(_v(p, self.scale) / (_v(p, self.high) * q * (1 - q)))

    def dlog_dinv_dy(self, y, p):
        q = self._p(y, p)
        with np.errstate(all="ignore"):
            return np.where((q > 0.0) & (q < 1.0), (-1.0 / q + 1.0 / (1.0 - q)) / _v(p, self.high), 0.0)

    def grads(self, y, p):
        hi, sc = _v(p, self.high), _v(p, self.scale)
        q = self._p(y, p)
        live = (q > 0.0) & (q < 1.0)                  # saturated points are constants of the switch (mappings.py:392)
        with np.errstate(all="ignore"):
            dq = np.where(live, 1.0 / (q * (1 - q)), 0.0)
            dl = np.where(live, -1.0 / q + 1.0 / (1 - q), 0.0)
            lq = np.where(live, np.log(q / (1 - q)), 0.0)
        n_live = float(np.sum(live))
        dinv = {self.lower: sc * dq * (-1.0 / hi), self.high: sc * dq * (-q / hi), self.location: np.ones_like(y),
                self.scale: lq}
        dld = {self.lower: float(np.sum(dl * (-1.0 / hi))), self.high: float(-n_live / hi + np.sum(dl * (-q / hi))),
               self.location: 0.0, self.scale: n_live / sc}
        return ({h: v for h, v in dinv.items() if isinstance(h, HyperVar)},
                {h: v for h, v in dld.items() if isinstance(h, HyperVar)})


class _NewtonWarping(Mapping):
    elementwise_forward = False

    """Warpings given by their inverse only: the forward map is `inverse_function(self.inv, z)` (mappings.py:11-12,
    libs/tensors.py:134-145) -- a damped Newton iteration from 0 (step 0.1, slopes below 1 replaced by their sign)
    stopped when max|inv(x) - z| < 1e-3 over the whole vector.  Reproduced literally, so T(z) carries the reference's
    ~1e-3 error; `logdet_dinv` is the reference's `sum(log(diag(jacobian(inv))))` in closed form."""
    NAMES = ()

    def __init__(self, y=None, n=1, name=None):
        super().__init__(y, name)
        self.n = int(n)

    def _regn(self, parent, reg, attr, positive):
        h = getattr(self, attr)
        if h is None:
            h = (reg.FlatExp if positive else reg.Flat)(parent + self.name + "_" + attr, shape=self.n)
            setattr(self, attr, h)
        if isinstance(h, HyperVar) and h not in self.hypers:
            self.hypers += [h]

    def logdet_dinv(self, y, p):
        with np.errstate(all="ignore"):
            return float(np.sum(np.log(self.dinv_dy(y, p))))

    def __call__(self, z, p, tol=1e-3, n_steps=1024, alpha=0.1):
        z = np.asarray(z, dtype=np.float64)
        x = 0.0 * z
        for _ in range(n_steps):
            diff = self.inv(x, p) - z
            d = self.dinv_dy(x, p)
            d = np.where(np.abs(d) < 1.0, np.sign(d), d)
            x = x - alpha * diff / d
            if np.max(np.abs(diff)) < tol:
                break
        return x

    def grads(self, y, p):
        dinv, dD = self._partials(y, p)
        Dy = self.dinv_dy(y, p)
        hs = [getattr(self, a) for a in self.NAMES]
        return ({h: v for h, v in zip(hs, dinv) if isinstance(h, HyperVar)},
                {h: np.sum(v / Dy, axis=1) for h, v in zip(hs, dD) if isinstance(h, HyperVar)})


class WarpingTanh(_NewtonWarping):   # mappings.py:253-278: inv(y) = y + sum_j a_j tanh(b_j (y + c_j))
    NAMES = ("a", "b", "c")

    def __init__(self, y=None, n=1, name=None, a=None, b=None, c=None):
        super().__init__(y, n, name)
        self.a, self.b, self.c = a, b, c

    def check_hypers(self, parent="", reg=None):
        self._regn(parent, reg, "a", True)
        self._regn(parent, reg, "b", True)
        self._regn(parent, reg, "c", False)

    def default_hypers(self, x=None, y=None):
        y = np.asarray(y, dtype=np.float64)
        one = np.ones(self.n)
        d = {self.a: 0.1 * one * np.abs(y).max() / self.n, self.b: 0.1 * one / np.abs(y).max(), self.c: one * np.mean(y)}
        return {h: v for h, v in d.items() if isinstance(h, HyperVar)}

    def _abc(self, p):
        return _vec(p, self.a, self.n), _vec(p, self.b, self.n), _vec(p, self.c, self.n)

    def inv(self, y, p):
        a, b, c = self._abc(p)
        return y + np.dot(np.tanh(b * (y[:, None] + c)), a)

    def dinv_dy(self, y, p):
        a, b, c = self._abc(p)
        return 1.0 + np.dot(1.0 / np.cosh(b * (y[:, None] + c)) ** 2, a * b)

    def dlog_dinv_dy(self, y, p):
        a, b, c = self._abc(p)
        u = b * (y[:, None] + c)
        return np.dot(-2.0 / np.cosh(u) ** 2 * np.tanh(u), a * b * b) / self.dinv_dy(y, p)

    def _partials(self, y, p):
        a, b, c = self._abc(p)
        yc = y[:, None] + c
        u = b * yc
        t, s2 = np.tanh(u), 1.0 / np.cosh(u) ** 2
        dinv = (t.T, (a * s2 * yc).T, (a * b * s2).T)
        dD = ((b * s2).T, (a * s2 - 2.0 * a * b * s2 * t * yc).T, (-2.0 * a * b * b * s2 * t).T)
        return dinv, dD


class WarpingBoxCox(_NewtonWarping):  # mappings.py:281-306: inv(y) = sum_j w_j (sgn(s_j)|s_j|^p_j - 1)/p_j, s_j = y + shift_j
    NAMES = ("shift", "power", "w")

    def __init__(self, y=None, n=1, name=None, shift=None, power=None, w=None):
        super().__init__(y, n, name)
        self.shift, self.power, self.w = shift, power, w

    def check_hypers(self, parent="", reg=None):
        self._regn(parent, reg, "shift", True)
        self._regn(parent, reg, "power", True)
        self._regn(parent, reg, "w", True)

    def default_hypers(self, x=None, y=None):
        one = np.ones(self.n)
        d = {self.w: one / self.n, self.shift: one, self.power: one}
        return {h: v for h, v in d.items() if isinstance(h, HyperVar)}

    def _spw(self, p):
        return _vec(p, self.shift, self.n), _vec(p, self.power, self.n), _vec(p, self.w, self.n)

    def inv(self, y, p):
        shift, power, w = self._spw(p)
        sh = y[:, None] + shift
        with np.errstate(all="ignore"):
            return np.dot((np.sign(sh) * np.abs(sh) ** power - 1.0) / power, w)

    def dinv_dy(self, y, p):
        shift, power, w = self._spw(p)
        with np.errstate(all="ignore"):
            return np.dot(np.abs(y[:, None] + shift) ** (power - 1.0), w)

    def dlog_dinv_dy(self, y, p):
        shift, power, w = self._spw(p)
        sh = y[:, None] + shift
        with np.errstate(all="ignore"):
            return np.dot((power - 1.0) * np.abs(sh) ** (power - 2.0) * np.sign(sh), w) / self.dinv_dy(y, p)

    def _partials(self, y, p):
        shift, power, w = self._spw(p)
        sh = y[:, None] + shift
        with np.errstate(all="ignore"):
            ab = np.abs(sh)
            sp = np.sign(sh) * ab ** power
            dinv = ((w * ab ** (power - 1.0)).T, (w * (sp * np.log(ab) * power - (sp - 1.0)) / power ** 2).T,
                    ((sp - 1.0) / power).T)
            dD = ((w * (power - 1.0) * ab ** (power - 2.0) * np.sign(sh)).T, (w * ab ** (power - 1.0) * np.log(ab)).T,
                  (ab ** (power - 1.0)).T)
        return dinv, dD

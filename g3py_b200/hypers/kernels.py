"""Kernel algebra (mirror of g3py/processes/hypers/kernels.py for the hot-path kernels).

The reference's `Kernel.cov` builds a Theano expression; here each kernel *compiles* itself into the
post-order `g3_kernel_desc` tree that the CUDA Gram kernels interpret (include/g3b.h), so that
`k1 + k2`, `k1 * k2`, `c * k`, `c + k` (kernels.py:51-75) keep working unchanged.
"""
import numpy as np

from . import Hypers, HyperVar
from .. import _cabi as cabi

__all__ = ["Kernel", "KernelStationary", "KernelSum", "KernelProd", "KernelScale", "KernelShift", "KernelNoise", "WN",
           "SE", "OU", "MAT32", "MAT52", "RQ", "SIN", "COS", "SINC", "SM", "KernelPeriodic", "DescBuilder",
           "KernelDot", "LIN", "POL", "BW", "VAR", "KernelMax", "NN", "NIL", "KernelEquals", "KernelEquals2", "kernel_leaves"]


def kernel_leaves(k):
    """Leaves of a kernel expression, left to right."""
    if hasattr(k, "k1"):
        return kernel_leaves(k.k1) + kernel_leaves(k.k2)
    if hasattr(k, "k") and isinstance(getattr(k, "k"), Kernel):
        return kernel_leaves(k.k)
    return [k]


class DescBuilder:
    """Collects nodes and the kernel-theta layout (free hypers first come, constants in extra slots)."""

    def __init__(self, D):
        self.D = D
        self.nodes = []
        self.slots = []          # (HyperVar | None, size, constant ndarray | None)
        self.n_theta = 0

    def slot(self, h, size):
        """Index of hyper `h` (HyperVar, or a fixed number / array) in the kernel theta vector."""
        if isinstance(h, HyperVar):
            for hv, off, sz, _ in self.slots:
                if hv is h:
                    return off
            off = self.n_theta
            self.slots.append((h, off, h.size, None))
            self.n_theta += h.size
            return off
        val = np.broadcast_to(np.asarray(h, dtype=np.float64), (size,)).copy()
        off = self.n_theta
        self.slots.append((None, off, size, val))
        self.n_theta += size
        return off

    def node(self, op, dim0=0, dim1=0, var_idx=-1, p0=-1, p1=-1, flags=0, value=0.0):
        if len(self.nodes) >= cabi.G3_MAX_NODES:
            raise ValueError("kernel expression has more than %d nodes" % cabi.G3_MAX_NODES)
        self.nodes.append((op, dim0, dim1, var_idx, p0, p1, flags, float(value)))
        return len(self.nodes) - 1

    def finish(self):
        if self.n_theta > cabi.G3_MAX_THETA:
            raise ValueError("kernel has more than %d hyper slots" % cabi.G3_MAX_THETA)
        d = cabi.KernelDesc()
        d.n_nodes = len(self.nodes)
        d.n_theta = self.n_theta
        for i, (op, d0, d1, vi, p0, p1, fl, val) in enumerate(self.nodes):
            n = d.nodes[i]
            n.op, n.dim0, n.dim1, n.var_idx, n.p0_idx, n.p1_idx, n.flags, n.value = op, d0, d1, vi, p0, p1, fl, val
        return d


class Kernel(Hypers):
    """kernels.py:13-79."""

    def __init__(self, x=None, name=None, var=None):
        super().__init__(x, name)
        self.var = var

    def check_hypers(self, parent="", reg=None):
        if self.var is None:
            self.var = reg.FlatExp(parent + self.name + "_var")          # kernels.py:22-24
        if isinstance(self.var, HyperVar) and self.var not in self.hypers:
            self.hypers += [self.var]

    def default_hypers(self, x=None, y=None):
        return {self.var: float(np.var(y))} if isinstance(self.var, HyperVar) else {}   # kernels.py:33-40

    def compile(self, b, process_noise=False):
        raise NotImplementedError

    def nan_quirk_hypers(self):
        """Hypers whose gradient the reference returns as 0: Theano's autodiff gives NaN for them on the diagonal
        (d = 0) and `th_dlogp` scrubs NaN -> 0 (stochastic.py:308-309).  Confirmed by executing the reference
        (tests/golden/reference_g3py.json).  Empty for most kernels."""
        return []

    def cov(self, x1, x2=None, **hypers):
        """Numeric Kernel.cov through the device (kernels.py:106-110); hypers by bare name, natural space."""
        from ..processes import kernel_cov
        return kernel_cov(self, x1, x2, hypers)

    def __mul__(self, other):
        return KernelProd(self, other) if isinstance(other, Kernel) else KernelScale(self, other)
    __imul__ = __mul__

    def __rmul__(self, other):
        return KernelProd(other, self) if isinstance(other, Kernel) else KernelScale(self, other)

    def __add__(self, other):
        return KernelSum(self, other) if isinstance(other, Kernel) else KernelShift(self, other)
    __iadd__ = __add__

    def __radd__(self, other):
        return KernelSum(other, self) if isinstance(other, Kernel) else KernelShift(self, other)


class KernelOperation(Kernel):
    """kernels.py:112-139 — scalar (*) or (+) kernel."""
    OP = None

    def __init__(self, _k, _element):
        self.k = _k
        self.element = float(_element)
        self.hypers = []
        self.potential = None

    @property
    def name(self):
        return str(self.element) + " " + self.op + " " + self.k.name

    def check_hypers(self, parent="", reg=None):
        self.k.check_hypers(parent=parent, reg=reg)
        self.hypers = self.k.hypers

    def check_dims(self, x=None):
        self.k.check_dims(x)

    def default_hypers_dims(self, x=None, y=None):
        return self.k.default_hypers_dims(x, y)

    def compile(self, b, process_noise=False):
        c = self.k.compile(b)
        return b.node(self.OP, dim0=c, value=self.element)

    def nan_quirk_hypers(self):
        return self.k.nan_quirk_hypers()

    def check_potential(self, reg=None):                                 # kernels.py:131-133
        Hypers.check_potential(self, reg)
        self.k.check_potential(reg)

    def potential_hypers(self):
        return self.k.potential_hypers()


class KernelScale(KernelOperation):
    OP = cabi.K_SCALE
    op = "*"


class KernelShift(KernelOperation):
    OP = cabi.K_SHIFT
    op = "+"


class KernelComposition(Kernel):
    """kernels.py:142-189."""
    OP = None

    def __init__(self, _k1, _k2):
        self.k1 = _k1
        self.k2 = _k2
        self.hypers = []
        self.potential = None

    @property
    def name(self):
        # kernels.py:169-186: "k1(start,stop) op k2(start,stop)" with the reference's cascade of fallbacks (the name
        # ends up in hyper names, e.g. TKernel's 'Noise' + kernel.name, transports.py:205-206)
        def full(k):
            return k.name + "(" + str(k.dims.start) + "," + str(k.dims.stop) + ")"

        def plain(k):
            return k.name + "(" + str(k.dims) + ")"

        for a, b in ((full, full), (plain, plain), (None, plain), (plain, None)):
            try:
                return ((a(self.k1) if a else self.k1.name) + " " + self.op + " "
                        + (b(self.k2) if b else self.k2.name))
            except Exception:
                continue
        return self.k1.name + " " + self.op + " " + self.k2.name

    def check_hypers(self, parent="", reg=None):
        self.k1.check_hypers(parent=parent, reg=reg)
        self.k2.check_hypers(parent=parent, reg=reg)
        self.hypers = self.k1.hypers + self.k2.hypers

    def check_dims(self, x=None):
        self.k1.check_dims(x)
        self.k2.check_dims(x)

    def default_hypers_dims(self, x=None, y=None):
        return {**self.k1.default_hypers_dims(x, y), **self.k2.default_hypers_dims(x, y)}

    def compile(self, b, process_noise=False):
        l = self.k1.compile(b)
        r = self.k2.compile(b, process_noise=process_noise)
        return b.node(self.OP, dim0=l, dim1=r)

    def nan_quirk_hypers(self):
        return self.k1.nan_quirk_hypers() + self.k2.nan_quirk_hypers()

    def check_potential(self, reg=None):                                 # kernels.py:163-166
        Hypers.check_potential(self, reg)
        self.k1.check_potential(reg)
        self.k2.check_potential(reg)

    def potential_hypers(self):
        return self.k1.potential_hypers() + self.k2.potential_hypers()


class KernelProd(KernelComposition):
    OP = cabi.K_PROD
    op = "*"

    def __init__(self, _k1, _k2):
        super().__init__(_k1, _k2)
        # kernels.py:215-219: two leaf kernels that both have var=None -> the second var is fixed to 1.0
        if hasattr(self.k1, "var") and hasattr(self.k2, "var"):
            if self.k1.var is None and self.k2.var is None:
                self.k2.var = 1.0


class KernelSum(KernelComposition):
    OP = cabi.K_SUM
    op = "+"


class KernelMax(KernelComposition):  # kernels.py:247-257
    OP = cabi.K_MAX
    op = "max"


class KernelStationary(Kernel):
    """kernels.py:96-110: cov = var * k(metric.gram(x1, x2)); metric hypers follow `var`."""
    OPCODE = None
    RATE_DEFAULT = "l2"

    def __init__(self, x=None, name=None, var=None, rate=None):
        super().__init__(x, name, var)
        self.rate = rate

    def check_hypers(self, parent="", reg=None):
        super().check_hypers(parent, reg)
        if self.rate is None:                                            # metrics.py:79-83  (ARD.check_hypers)
            self.rate = reg.FlatExp(parent + self.name + "_rate", shape=self.shape)
        if isinstance(self.rate, HyperVar) and self.rate not in self.hypers:
            self.hypers += [self.rate]

    def default_hypers(self, x=None, y=None):
        d = super().default_hypers(x, y)
        if isinstance(self.rate, HyperVar):
            try:
                m = np.abs(x[1:] - x[:-1]).mean(axis=0)
                d[self.rate] = (0.5 if self.RATE_DEFAULT == "l2" else 1.0) / m   # metrics.py:93-94,104-108
            except Exception:
                pass
        return d

    def potential_hypers(self):
        # the ARD `rate` (and ARD_DotBias `bias`) live on the metric object in the reference (metrics.py:79-83), whose
        # check_potential is never called: only var (+ alpha / freq / periodic rate) can be regularised
        skip = [self.rate] + [getattr(self, "bias", None)]
        return [h for h in self.hypers if isinstance(h, HyperVar) and not any(h is q for q in skip)]

    def _common(self, b):
        d0, d1 = self.dim_range(b.D)
        nd = d1 - d0
        if isinstance(self.var, HyperVar):
            vi, val = b.slot(self.var, 1), 0.0
        else:
            vi, val = -1, float(self.var)
        return d0, d1, nd, vi, val

    def compile(self, b, process_noise=False):
        d0, d1, nd, vi, val = self._common(b)
        p0 = b.slot(self.rate, nd)
        return b.node(self.OPCODE, d0, d1, vi, p0, -1, 0, val)


class SE(KernelStationary):          # kernels.py:434-436
    OPCODE = cabi.K_SE


class OU(KernelStationary):          # kernels.py:429-431 (ARD_L1 metric)
    OPCODE = cabi.K_OU
    RATE_DEFAULT = "l1"


class MAT32(KernelStationary):       # kernels.py:406-412
    OPCODE = cabi.K_MAT32

    def nan_quirk_hypers(self):          # grad(sqrt(3 d)) at d = 0 is 0/0
        return [self.rate] if isinstance(self.rate, HyperVar) else []


class MAT52(KernelStationary):       # kernels.py:415-421
    OPCODE = cabi.K_MAT52

    def nan_quirk_hypers(self):          # grad(sqrt(5 d)) at d = 0 is 0/0
        return [self.rate] if isinstance(self.rate, HyperVar) else []


class RQ(KernelStationary):          # kernels.py:388-403
    OPCODE = cabi.K_RQ

    def __init__(self, x=None, name=None, var=None, rate=None, alpha=None):
        super().__init__(x, name, var, rate)
        self.alpha = alpha

    def check_hypers(self, parent="", reg=None):
        super().check_hypers(parent, reg)
        if self.alpha is None:
            self.alpha = reg.FlatExp(parent + self.name + "_alpha")
        if isinstance(self.alpha, HyperVar) and self.alpha not in self.hypers:
            self.hypers += [self.alpha]

    def default_hypers(self, x=None, y=None):
        d = super().default_hypers(x, y)
        if isinstance(self.alpha, HyperVar):
            d[self.alpha] = 1.0
        return d

    def compile(self, b, process_noise=False):
        d0, d1, nd, vi, val = self._common(b)
        p0 = b.slot(self.rate, nd)
        p1 = b.slot(self.alpha, 1)
        return b.node(self.OPCODE, d0, d1, vi, p0, p1, 0, val)


class KernelPeriodic(KernelStationary):
    """kernels.py:439-459: Difference metric (no hypers); freq is created before rate."""

    def __init__(self, x=None, name=None, var=None, freq=None, rate=None):
        super().__init__(x, name, var, rate)
        self.freq = freq

    def check_hypers(self, parent="", reg=None):
        Kernel.check_hypers(self, parent, reg)
        if self.freq is None:
            self.freq = reg.FlatExp(parent + self.name + "_freq", shape=self.shape)
        if self.rate is None:
            self.rate = reg.FlatExp(parent + self.name + "_rate", shape=self.shape)
        for h in (self.rate, self.freq):
            if isinstance(h, HyperVar) and h not in self.hypers:
                self.hypers += [h]

    def potential_hypers(self):                                          # kernels.py:446-454: rate and freq are the kernel's own
        return [h for h in self.hypers if isinstance(h, HyperVar)]

    def default_hypers(self, x=None, y=None):
        d = Kernel.default_hypers(self, x, y)
        if isinstance(self.freq, HyperVar):
            d[self.freq] = 1.0 / (x.max(axis=0) - x.min(axis=0))
        if isinstance(self.rate, HyperVar):
            d[self.rate] = 1.0 / np.abs(x[1:] - x[:-1]).mean(axis=0)
        return d


class SIN(KernelPeriodic):           # kernels.py:470-472
    OPCODE = cabi.K_SIN

    def compile(self, b, process_noise=False):
        d0, d1, nd, vi, val = self._common(b)
        p1 = b.slot(self.freq, nd)
        p0 = b.slot(self.rate, nd)
        return b.node(self.OPCODE, d0, d1, vi, p0, p1, 0, val)


class COS(KernelPeriodic):           # kernels.py:462-467: rate is the constant 1.0, only freq is a hyper
    OPCODE = cabi.K_COS

    def __init__(self, x=None, name=None, var=None, freq=None):
        super().__init__(x, name, var, freq, rate=1.0)

    def compile(self, b, process_noise=False):
        d0, d1, nd, vi, val = self._common(b)
        p1 = b.slot(self.freq, nd)
        return b.node(self.OPCODE, d0, d1, vi, -1, p1, 0, val)


class SINC(COS):                     # kernels.py:475-482
    OPCODE = cabi.K_SINC

    def nan_quirk_hypers(self):          # the unselected switch branch sin(0)/0 has a NaN gradient (kernels.py:479-480)
        return [self.freq] if isinstance(self.freq, HyperVar) else []


class SM(KernelPeriodic):            # kernels.py:485-487 (spectral mixture component)
    OPCODE = cabi.K_SM

    def compile(self, b, process_noise=False):
        d0, d1, nd, vi, val = self._common(b)
        p1 = b.slot(self.freq, nd)
        p0 = b.slot(self.rate, nd)
        return b.node(self.OPCODE, d0, d1, vi, p0, p1, 0, val)


class KernelDot(KernelStationary):
    """kernels.py:82-96 with the ARD_Dot metric (metrics.py:110-117): var * sum_k rate_k^2 x_ik x_jk.  Subclasses with
    BIAS use ARD_DotBias (metrics.py:120-137): bias + sum_k ..., hypers created rate first, then bias."""
    OPCODE = cabi.K_DOT
    BIAS = False
    POWER = 1
    FLAGS = 0

    def __init__(self, x=None, name=None, var=None, rate=None, bias=None):
        super().__init__(x, name, var, rate)
        self.bias = bias

    def check_hypers(self, parent="", reg=None):
        super().check_hypers(parent, reg)
        if self.BIAS:
            if self.bias is None:
                self.bias = reg.FlatExp(parent + self.name + "_bias")
            if isinstance(self.bias, HyperVar) and self.bias not in self.hypers:
                self.hypers += [self.bias]

    def default_hypers(self, x=None, y=None):
        d = Kernel.default_hypers(self, x, y)
        x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        if self.BIAS:                                                    # metrics.py:135-137
            if isinstance(self.bias, HyperVar):
                d[self.bias] = np.abs(y).mean() / np.abs(x).mean()
            if isinstance(self.rate, HyperVar):
                d[self.rate] = np.sqrt(np.abs(y)).mean(axis=0) / np.abs(x).mean(axis=0)
        elif isinstance(self.rate, HyperVar):                           # metrics.py:114-115
            d[self.rate] = 1.0 / (np.sqrt(np.abs(x)).mean(axis=0) / np.abs(y).mean(axis=0))
        return d

    def compile(self, b, process_noise=False):
        d0, d1, nd, vi, val = self._common(b)
        p0 = b.slot(self.rate, nd)
        p1 = b.slot(self.bias, 1) if self.BIAS else -1
        return b.node(self.OPCODE, d0, d1, vi, p0, p1, ((int(self.POWER) & 0xff) << 8) | self.FLAGS, val)


class LIN(KernelDot):                # kernels.py:319-321: var fixed to 1
    BIAS = True

    def __init__(self, x=None, name=None, var=1, rate=None, bias=None):
        super().__init__(x, name, var, rate, bias)


class POL(KernelDot):                # kernels.py:324-336: var * (bias + sum_k rate_k^2 x_ik x_jk) ** p
    BIAS = True

    def __init__(self, x=None, p=2, name=None, var=1, rate=None, bias=None):
        super().__init__(x, name, var, rate, bias)
        if int(p) != p or not 1 <= int(p) <= 255:
            raise ValueError("POL: p must be an integer in [1, 255]")
        self.p = self.POWER = int(p)


class NN(KernelDot):                 # kernels.py:339-351: var * arcsin(2m / (1 + 2m)^2), m = bias + sum_k rate_k^2 x_ik x_jk
    """The neural-network kernel as its single-argument `cov(x1)` is written in the reference: elementwise in the Gram
    entry m_ij (not normalised by m_ii, m_jj).  The two-argument form multiplies an N1xN1 by an N2xN2 matrix
    (kernels.py:351) and only broadcasts when N1 == N2, so the posterior methods raise here as they do there."""
    BIAS = True
    FLAGS = cabi.KF_NN
    TRAINING_GRAM_ONLY = True


class BW(Kernel):                    # kernels.py:291-293 with the Minimum metric (metrics.py:49-51): Brownian motion
    OPCODE = cabi.K_BW

    def compile(self, b, process_noise=False):
        d0, d1 = self.dim_range(b.D)
        if isinstance(self.var, HyperVar):
            vi, val = b.slot(self.var, 1), 0.0
        else:
            vi, val = -1, float(self.var)
        return b.node(self.OPCODE, d0, d1, vi, -1, -1, 0, val)


class VAR(BW):                       # kernels.py:296-306: constant kernel var * ones
    OPCODE = cabi.K_VAR


class NIL(Kernel):                   # kernels.py:309-320: zeros, no hypers (var=1 is never used)
    def __init__(self, x=None, name=None, var=1):
        super().__init__(x, name, 1)

    def compile(self, b, process_noise=False):
        d0, d1 = self.dim_range(b.D)
        return b.node(cabi.K_VAR, d0, d1, -1, -1, -1, 0, 0.0)


class KernelEquals(Kernel):          # kernels.py:262-274 over DeltaEq (metrics.py:38-43): sum_k [x_ik == eq][x_jk == eq]
    def __init__(self, x=None, name=None, eq=0):
        super().__init__(x, name, 1)
        self.eq = float(eq)

    def compile(self, b, process_noise=False):
        d0, d1 = self.dim_range(b.D)
        return b.node(cabi.K_EQ, d0, d1, -1, -1, -1, 0, self.eq)


class KernelEquals2(Kernel):         # kernels.py:277-288 over DeltaEq2 (metrics.py:46-51)
    def __init__(self, x=None, name=None, eq1=0, eq2=0):
        super().__init__(x, name, 1)
        self.eq1, self.eq2 = float(eq1), float(eq2)

    def compile(self, b, process_noise=False):
        import struct
        d0, d1 = self.dim_range(b.D)
        lo, hi = struct.unpack("<ii", struct.pack("<d", self.eq2))       # the second constant travels in the two index words
        return b.node(cabi.K_EQ, d0, d1, -1, lo, hi, cabi.KF_EQ2, self.eq1)


class KernelNoise(Kernel):           # kernels.py:360-371
    OPCODE = cabi.K_NOISE

    def compile(self, b, process_noise=False):
        if isinstance(self.var, HyperVar):
            vi, val = b.slot(self.var, 1), 0.0
        else:
            vi, val = -1, float(self.var)
        return b.node(self.OPCODE, 0, 0, vi, -1, -1, cabi.KF_PROCESS_NOISE if process_noise else 0, val)


class WN(Kernel):                    # kernels.py:374-385
    OPCODE = cabi.K_WN

    def compile(self, b, process_noise=False):
        d0, d1 = self.dim_range(b.D)
        if isinstance(self.var, HyperVar):
            vi, val = b.slot(self.var, 1), 0.0
        else:
            vi, val = -1, float(self.var)
        return b.node(self.OPCODE, d0, d1, vi, -1, -1, 0, val)

"""Transports — mirror of g3py/processes/hypers/transports.py.

A transport pushes white noise to the observations: y = t_1(t_2(... t_n(eps))), eps ~ N(0, I), composed with `@`
(`transports.py:59-64,93-119`).  The element-wise pieces (`TLocation`, `TMapping`, `ID`) are O(N) host work; the
linear piece `TKernel` is the Cholesky factor of a Gram matrix and is where the device path enters
(`transports.py:200-257`).  `TransportGaussianProcess` (g3py_b200/transport.py) evaluates chains of the form

    [ID | TMapping]* @ [TScale]? @ [ID | TLocation]* @ TKernel

i.e. any number of element-wise transports outside exactly one kernel transport, which is the structure whose
density is the warped-GP density (`processes/transport.py:214-246`).  `TScale` (`transports.py:165-181`) multiplies by a
parametric function of the INPUTS; between the warpings and the locations it acts as one more (input-dependent) linear
warping, y = T(s(x) (m(x) + L eps)).  `TTriangular` (`transports.py:260-263`, unfinished in the reference) is not built.
"""
from . import Hypers
from .kernels import Kernel, KernelSum, KernelNoise
from .mappings import Mapping
from .means import Mean

__all__ = ["Transport", "TransportComposed", "ID", "TLocation", "TMapping", "TKernel", "TScale"]


class Transport(Hypers):
    """transports.py:10-67."""

    def __init__(self, x=None, name=None):
        super().__init__(x, name)
        self.parametrics = []

    def chain(self):
        """Outermost-first list of the primitive transports."""
        return [self]

    def check_hypers(self, parent="", reg=None):
        for p in self.parametrics:
            p.check_hypers(parent, reg)
            self.hypers += [h for h in p.hypers if h not in self.hypers]

    def check_dims(self, x=None):
        for p in self.parametrics:
            p.check_dims(x)

    def default_hypers_dims(self, x=None, y=None):
        r = {}
        for p in self.parametrics:
            r.update(p.default_hypers_dims(x, y))
        return r

    def __matmul__(self, other):
        return TransportComposed(self, other)


class TransportComposed(Transport):
    """transports.py:93-119: (t1 @ t2)(eps) = t1(t2(eps)); hypers are created t1 first."""

    def __init__(self, t1, t2):
        self.t1, self.t2 = t1, t2
        self.hypers = []
        self.parametrics = []
        self.name = t1.name + " " + t2.name
        self.dims = None
        self.shape = None
        self.potential = None

    def chain(self):
        return self.t1.chain() + self.t2.chain()

    def check_hypers(self, parent="", reg=None):
        self.t1.check_hypers(parent=parent, reg=reg)
        self.t2.check_hypers(parent=parent, reg=reg)
        self.hypers = self.t1.hypers + self.t2.hypers

    def check_dims(self, x=None):
        self.t1.check_dims(x)
        self.t2.check_dims(x)

    def default_hypers_dims(self, x=None, y=None):
        return {**self.t1.default_hypers_dims(x, y), **self.t2.default_hypers_dims(x, y)}


class ID(Transport):
    """transports.py:122-130.  Its `logdet_dinv` is `tt.ones(())` = 1 in the reference (not 0): every ID in a chain
    adds 1 to logp; reproduced."""


class TLocation(Transport):
    """transports.py:146-162: outputs + location(inputs)."""

    def __init__(self, location=None, x=None, name=None):
        super().__init__(x, name)
        if not isinstance(location, Mean):
            raise TypeError("TLocation needs a Mean")
        self.location = location
        self.parametrics.append(location)


class TScale(Transport):
    """transports.py:165-181: outputs * scale(inputs); inverse outputs / scale(inputs); log|d inv| = -sum log scale(inputs)."""

    def __init__(self, scale=None, x=None, name=None):
        super().__init__(x, name)
        if not isinstance(scale, Mean):
            raise TypeError("TScale needs a Mean (a parametric function of the inputs)")
        self.scale = scale
        self.parametrics.append(scale)


class TMapping(Transport):
    """transports.py:184-197: mapping(outputs)."""

    def __init__(self, mapping=None, x=None, name=None):
        super().__init__(x, name)
        if not isinstance(mapping, Mapping):
            raise TypeError("TMapping needs a Mapping")
        self.mapping = mapping
        self.parametrics.append(mapping)


class TKernel(Transport):
    """transports.py:200-257: chol(kernel.cov(inputs)) @ outputs; `noisy=True` adds KernelNoise named
    'Noise' + kernel.name (`:205-206`)."""

    def __init__(self, kernel, noisy=False, x=None, name=None):
        super().__init__(x, name)
        if not isinstance(kernel, Kernel):
            raise TypeError("TKernel needs a Kernel")
        self.kernel = kernel
        self.is_noisy = bool(noisy)
        self.noisy = KernelSum(kernel, KernelNoise(name="Noise" + kernel.name)) if noisy else kernel
        self.parametrics.append(self.noisy)

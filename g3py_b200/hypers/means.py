"""Mean (location) algebra — mirror of g3py/processes/hypers/means.py.

O(N*D) work: it stays on the host and enters the device path as
`delta = T^-1(y) - m(X)`.  `eval`/`jac` take the natural-space hyper values through `p(h)`.
"""
import numpy as np

from . import Hypers, HyperVar

__all__ = ["Mean", "Location", "Zero", "Bias", "Linear", "Power", "BlackBox", "MeanSum", "MeanProd", "MeanScale", "MeanShift"]


def _val(p, h):
    return p(h) if isinstance(h, HyperVar) else np.asarray(h, dtype=np.float64)


class Mean(Hypers):
    def __mul__(self, other):
        return MeanProd(self, other) if isinstance(other, Mean) else MeanScale(self, other)
    __imul__ = __mul__
    __rmul__ = __mul__

    def __add__(self, other):
        return MeanSum(self, other) if isinstance(other, Mean) else MeanShift(self, other)
    __iadd__ = __add__
    __radd__ = __add__

    def eval(self, x, p):
        raise NotImplementedError

    def jac(self, x, p):
        """{HyperVar: d mean / d hyper, shape (size, n)} in natural space."""
        return {}

    def __call__(self, x, p):
        x = np.asarray(x)
        return self.eval(x[:, self.dims] if self.dims is not None else x, p)      # means.py:26-27

    def jacobian(self, x, p):
        x = np.asarray(x)
        return self.jac(x[:, self.dims] if self.dims is not None else x, p)


Location = Mean


class _MeanOp(Mean):
    def __init__(self, _m, _element):
        self.m = _m
        self.element = float(_element)
        self.hypers = []
        self.name = "op"
        self.dims = None

    def check_hypers(self, parent="", reg=None):
        self.m.check_hypers(parent=parent, reg=reg)
        self.hypers = self.m.hypers

    def check_dims(self, x=None):
        self.m.check_dims(x)

    def default_hypers_dims(self, x=None, y=None):
        return self.m.default_hypers_dims(x, y)


class MeanScale(_MeanOp):            # means.py:63-69
    def __call__(self, x, p):
        return self.element * self.m(x, p)

    def jacobian(self, x, p):
        return {h: self.element * j for h, j in self.m.jacobian(x, p).items()}


class MeanShift(_MeanOp):            # means.py:72-78
    def __call__(self, x, p):
        return self.element + self.m(x, p)

    def jacobian(self, x, p):
        return self.m.jacobian(x, p)


class _MeanComp(Mean):
    def __init__(self, _m1, _m2):
        self.m1 = _m1
        self.m2 = _m2
        self.hypers = []
        self.name = "comp"
        self.dims = None

    def check_hypers(self, parent="", reg=None):
        self.m1.check_hypers(parent=parent, reg=reg)
        self.m2.check_hypers(parent=parent, reg=reg)
        self.hypers = self.m1.hypers + self.m2.hypers

    def check_dims(self, x=None):
        self.m1.check_dims(x)
        self.m2.check_dims(x)

    def default_hypers_dims(self, x=None, y=None):
        return {**self.m1.default_hypers_dims(x, y), **self.m2.default_hypers_dims(x, y)}


class MeanProd(_MeanComp):           # means.py:99-104
    def __call__(self, x, p):
        return self.m1(x, p) * self.m2(x, p)

    def jacobian(self, x, p):
        a, b = self.m1(x, p), self.m2(x, p)
        out = {h: j * b for h, j in self.m1.jacobian(x, p).items()}
        for h, j in self.m2.jacobian(x, p).items():
            out[h] = out.get(h, 0.0) + a * j
        return out


class MeanSum(_MeanComp):            # means.py:107-114
    def __call__(self, x, p):
        return self.m1(x, p) + self.m2(x, p)

    def jacobian(self, x, p):
        out = dict(self.m1.jacobian(x, p))
        for h, j in self.m2.jacobian(x, p).items():
            out[h] = out.get(h, 0.0) + j
        return out


class Zero(Mean):                    # means.py:117-119
    def eval(self, x, p):
        return np.zeros(x.shape[0])


class Bias(Mean):                    # means.py:122-137
    def __init__(self, x=None, name=None, bias=None):
        super().__init__(x, name)
        self.bias = bias

    def check_hypers(self, parent="", reg=None):
        if self.bias is None:
            self.bias = reg.Flat(parent + self.name + "_Bias")
        if isinstance(self.bias, HyperVar) and self.bias not in self.hypers:
            self.hypers += [self.bias]

    def default_hypers(self, x=None, y=None):
        return {self.bias: float(np.mean(y))} if isinstance(self.bias, HyperVar) else {}

    def eval(self, x, p):
        return float(_val(p, self.bias)) * np.ones(x.shape[0])

    def jac(self, x, p):
        return {self.bias: np.ones((1, x.shape[0]))} if isinstance(self.bias, HyperVar) else {}


class Linear(Mean):                  # means.py:140-159
    def __init__(self, x=None, name=None, constant=None, coeff=None):
        super().__init__(x, name)
        self.constant = constant
        self.coeff = coeff

    def check_hypers(self, parent="", reg=None):
        if self.constant is None:
            self.constant = reg.Flat(parent + self.name + "_Constant")
        if self.coeff is None:
            self.coeff = reg.Flat(parent + self.name + "_Coeff", shape=self.shape)
        for h in (self.constant, self.coeff):
            if isinstance(h, HyperVar) and h not in self.hypers:
                self.hypers += [h]

    def default_hypers(self, x=None, y=None):
        d = {}
        if isinstance(self.constant, HyperVar):
            d[self.constant] = float(np.mean(y))
        if isinstance(self.coeff, HyperVar):
            d[self.coeff] = np.mean(y) / x.mean(axis=0)
        return d

    def eval(self, x, p):
        return float(_val(p, self.constant)) + x.dot(np.atleast_1d(_val(p, self.coeff)))

    def jac(self, x, p):
        out = {}
        if isinstance(self.constant, HyperVar):
            out[self.constant] = np.ones((1, x.shape[0]))
        if isinstance(self.coeff, HyperVar):
            out[self.coeff] = x.T.copy()
        return out


class Power(Mean):                   # means.py:162-182: constant + dot(x**n, coeff)
    def __init__(self, x=None, name=None, constant=None, coeff=None, n=2):
        super().__init__(x, name)
        self.constant = constant
        self.coeff = coeff
        self.n = n

    def check_hypers(self, parent="", reg=None):
        if self.constant is None:
            self.constant = reg.Flat(parent + self.name + "_Constant")
        if self.coeff is None:
            self.coeff = reg.Flat(parent + self.name + "_Coeff", shape=self.shape)
        for h in (self.constant, self.coeff):
            if isinstance(h, HyperVar) and h not in self.hypers:
                self.hypers += [h]

    def default_hypers(self, x=None, y=None):
        d = {}
        if isinstance(self.constant, HyperVar):
            d[self.constant] = float(np.mean(y))
        if isinstance(self.coeff, HyperVar):
            d[self.coeff] = np.mean(y) / (x ** self.n).mean(axis=0)
        return d

    def eval(self, x, p):
        return float(_val(p, self.constant)) + (x ** self.n).dot(np.atleast_1d(_val(p, self.coeff)))

    def jac(self, x, p):
        out = {}
        if isinstance(self.constant, HyperVar):
            out[self.constant] = np.ones((1, x.shape[0]))
        if isinstance(self.coeff, HyperVar):
            out[self.coeff] = (x ** self.n).T.copy()
        return out


class BlackBox(Mean):                # means.py:32-41: a fixed vector, element[:len(x)]; no hypers
    def __init__(self, element, x=None, name=None):
        super().__init__(x, name)
        self.element = np.asarray(element, dtype=np.float64).reshape(-1)

    def eval(self, x, p):
        return self.element[:x.shape[0]].copy()

    def __call__(self, x, p):
        return self.element[:np.asarray(x).shape[0]].copy()

    def jacobian(self, x, p):
        return {}

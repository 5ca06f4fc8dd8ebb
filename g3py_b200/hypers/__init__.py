"""Hyper-parameter plumbing of the operator algebra (mirror of g3py/processes/hypers/__init__.py).

The reference creates one PyMC3 free RV per hyper inside `check_hypers` (Flat prior; positive
hypers use `FlatExp` = Flat + `NonTransformLog`, i.e. they live in log space with a zero Jacobian
and a -inf barrier at exp(theta) <= 1e-6, hypers/__init__.py:116-126,190-202).  Here a `HyperVar`
records the same facts (name, shape, positivity, creation order); the flat theta vector follows the
creation order, like PyMC3's `ArrayOrdering` over `model.cont_vars` (bayesian/models.py:143-145).
"""
import numpy as np


class HyperVar:
    """Stand-in for a PyMC3 free RV created by Hypers.Flat / Hypers.FlatExp."""

    def __init__(self, name, size=1, positive=False, scalar=True):
        self.name = name
        self.size = int(size)
        self.positive = bool(positive)
        self.scalar = scalar          # shape () vs shape (size,)
        self.offset = None            # position in the process theta vector (set by the process)

    @property
    def tname(self):
        """Name of the transformed variable as PyMC3 >= 3.1 spells it (bayesian/average.py:104-112)."""
        return self.name + "_log__" if self.positive else self.name

    def __repr__(self):
        return "HyperVar(%s, %d%s)" % (self.name, self.size, ", +" if self.positive else "")


class Registry:
    """Creation-ordered list of hypers and potentials (the role `pm.Model` plays for g3py)."""

    def __init__(self):
        self.vars = []
        self.potentials = []     # (name, reg, c, [HyperVar]) -- pm.Potential terms (hypers/__init__.py:94-109)

    def Flat(self, name, shape=()):
        return self._add(name, shape, False)

    def FlatExp(self, name, shape=()):
        return self._add(name, shape, True)

    def _add(self, name, shape, positive):
        for v in self.vars:
            if v.name == name:
                raise ValueError("hyper name %r used twice: give the components distinct `name`s" % name)
        scalar = shape == () or shape is None
        size = 1 if scalar else int(shape if np.isscalar(shape) else shape[0])
        v = HyperVar(name, size, positive, scalar)
        self.vars.append(v)
        return v


class Hypers:
    """hypers/__init__.py:35-109 — name, `dims` column slice, hyper list, defaults."""

    def __init__(self, x=None, name=None):
        self.name = self.__class__.__name__ if name is None else name
        self.hypers = []
        self.shape = None
        self.dims = None
        self.potential = None
        if x is not None:
            self.check_dims(x)

    def check_dims(self, x=None):
        # hypers/__init__.py:55-83
        if self.shape is not None:
            return
        if x is None:
            self.shape = None
            self.dims = slice(None)
            return
        if isinstance(x, list):
            d = np.array(x)
            if d.size and np.all(np.diff(d) == 1):
                self.dims = slice(int(d[0]), int(d[-1]) + 1)
            else:
                raise NotImplementedError("only contiguous column lists are supported as dims")
            self.shape = int(d.size)
        elif isinstance(x, tuple):
            domain, dims = x
            self.dims = dims
            domain = np.asarray(domain)
            full = domain.shape[1] if domain.ndim > 1 else 1
            if isinstance(dims, slice):
                self.shape = len(range(*dims.indices(full)))
            else:
                self.shape = full
        else:
            x = np.asarray(x)
            self.shape = x.shape[1] if x.ndim > 1 else 1
            self.dims = slice(0, self.shape)

    def dim_range(self, D):
        """Resolved [start, stop) of `dims` for inputs with D columns."""
        if self.dims is None:
            return 0, D
        start, stop, step = self.dims.indices(D)
        if step != 1:
            raise NotImplementedError("strided dims are not supported")
        return start, stop

    def check_hypers(self, parent="", reg=None):
        pass

    def set_potential(self, hypers="", reg="L1", c=1):
        """hypers/__init__.py:91-92: regulariser -c * sum |h| (L1) or -c * sum h^2 (L2) over this component's hypers
        whose name contains `hypers` (not at position 0), added to logp as a `pm.Potential`."""
        self.potential = (hypers, reg, c)

    def potential_hypers(self):
        """The hypers the reference keeps in `self.hypers` of this component (what `check_potential` iterates over)."""
        return [h for h in self.hypers if isinstance(h, HyperVar)]

    def check_potential(self, reg=None):
        # hypers/__init__.py:94-109
        if getattr(self, "potential", None) is None:
            return
        hypers, kind, c = self.potential
        sel = [h for h in self.potential_hypers() if h.name.find(hypers) > 0]
        reg.potentials.append((self.name + "_" + hypers + "_" + kind, kind, float(c), sel))

    def default_hypers(self, x=None, y=None):
        return {}

    def default_hypers_dims(self, x=None, y=None):
        x = np.asarray(x)
        return dict(self.default_hypers(x[:, self.dims] if self.dims is not None else x, y))

    def __str__(self):
        if len(self.hypers) == 0:
            return str(self.__class__.__name__)
        return str(self.__class__.__name__) + "[h=" + str(self.hypers) + "]"
    __repr__ = __str__


class Freedom(Hypers):
    """hypers/__init__.py:144-160 — nu = bound + degree, degree positive (log space)."""

    def __init__(self, x=None, name=None, degree=None, bound=2.0):
        super().__init__(x, name)
        self.degree = degree
        self.bound = float(bound)

    def check_hypers(self, parent="", reg=None):
        if self.degree is None:
            self.degree = reg.FlatExp(parent + self.name + "_degree")
        if isinstance(self.degree, HyperVar):
            self.hypers += [self.degree]

    def default_hypers(self, x=None, y=None):
        return {self.degree: float(np.asarray(y).shape[0])} if isinstance(self.degree, HyperVar) else {}


from .kernels import *   # noqa: E402,F401,F403
from .means import *     # noqa: E402,F401,F403
from .mappings import *  # noqa: E402,F401,F403

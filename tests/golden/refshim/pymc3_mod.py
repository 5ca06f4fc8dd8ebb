"""Stand-in for the slice of PyMC3 (3.1+) the reference uses -- TEST INFRASTRUCTURE (see __init__.py).

Model context stack, Flat / Continuous distributions with optional element-wise transforms (a transformed
variable `x` becomes the free variable `x_<transform>__` plus the deterministic `x = backward(free)`, with
logp = dist.logp(backward(free)) + jacobian_det(free)), observed variables, potentials, `inputvars` /
`cont_inputs` graph traversal and the dict<->array bijection in free-variable creation order.
"""
import sys
import threading
import types

import numpy as np


def install(lazy, tt, config):
    TensorVariable = lazy.TensorVariable

    class Context:
        contexts = threading.local()

        def __enter__(self):
            type(self).get_contexts().append(self)
            return self

        def __exit__(self, *a):
            type(self).get_contexts().pop()

        @classmethod
        def get_contexts(cls):
            if not hasattr(cls.contexts, 'stack'):
                cls.contexts.stack = []
            return cls.contexts.stack

        @classmethod
        def get_context(cls):
            try:
                return cls.get_contexts()[-1]
            except IndexError:
                raise TypeError('No context on context stack')

    class FreeRV(TensorVariable):
        pass

    class ObservedRV(TensorVariable):
        pass

    class TransformedRV(TensorVariable):
        pass

    class Model(Context):
        def __init__(self, name='', model=None):
            self.name = name
            self.named_vars = {}
            self.free_RVs = []
            self.observed_RVs = []
            self.deterministics = []
            self.potentials = []
            self.missing_values = []

        def Var(self, name, dist, data=None, total_size=None):
            if data is None:
                if getattr(dist, 'transform', None) is None:
                    var = FreeRV('input', name=name, dtype=dist.dtype)
                    var.distribution = dist
                    var.model = self
                    var.tag.test_value = np.asarray(dist.default(), dtype=dist.dtype)
                    var.dshape = tuple(var.tag.test_value.shape)
                    var.dsize = int(np.prod(var.dshape))
                    var.logp_elemwiset = dist.logp(var)
                    var.logpt = tt.sum(var.logp_elemwiset)
                    self.free_RVs.append(var)
                else:
                    tr = dist.transform
                    free = self.Var('%s_%s__' % (name, tr.name), tr.apply(dist))
                    normal = tr.backward(free)
                    var = TransformedRV('op', fn=lambda t: t, args=(normal,), name=name, dtype=dist.dtype)
                    var.transformed = free
                    var.distribution = dist
                    var.model = self
                    self.deterministics.append(var)
            else:
                data = lazy.as_tensor_variable(data)
                var = ObservedRV('op', fn=lambda t: t, args=(data,), name=name, dtype=dist.dtype)
                var.distribution = dist
                var.model = self
                var.logp_elemwiset = dist.logp(data)
                var.logpt = tt.sum(var.logp_elemwiset)
                self.observed_RVs.append(var)
            self.named_vars[name] = var
            return var

        def __getitem__(self, key):
            return self.named_vars[key]

        @property
        def vars(self):
            return self.free_RVs

        @property
        def cont_vars(self):
            return [v for v in self.free_RVs if v.dtype.startswith('float')]

        @property
        def basic_RVs(self):
            return self.free_RVs + self.observed_RVs

        @property
        def unobserved_RVs(self):
            return self.free_RVs + self.deterministics

        @property
        def test_point(self):
            return {v.name: np.array(v.tag.test_value) for v in self.free_RVs}

        @property
        def ndim(self):
            return sum(v.dsize for v in self.free_RVs)

        @property
        def bijection(self):
            return DictToArrayBijection(ArrayOrdering(self.free_RVs), self.test_point)

        @property
        def logpt(self):
            return tt.add(*[tt.sum(v.logpt) for v in self.basic_RVs] + [tt.sum(p) for p in self.potentials])

    def modelcontext(model=None):
        return Model.get_context() if model is None else model

    # ---- distributions
    class Distribution:
        def __new__(cls, name, *args, **kwargs):
            if isinstance(name, str):
                model = Model.get_context()
                data = kwargs.pop('observed', None)
                kwargs.pop('total_size', None)
                dist = cls.dist(*args, **kwargs)
                return model.Var(name, dist, data)
            raise TypeError('name needs to be a string')

        def __getnewargs__(self):
            return ('_',)

        @classmethod
        def dist(cls, *args, **kwargs):
            d = object.__new__(cls)
            d.__init__(*args, **kwargs)
            return d

        def __init__(self, shape=(), dtype=None, testval=None, defaults=(), transform=None, broadcastable=None,
                     **kw):
            self.shape = shape
            self.dtype = str(dtype or config.floatX)
            self.testval = testval
            self.defaults = defaults
            self.transform = transform

        def default(self):
            tv = self.testval
            if isinstance(tv, lazy.Node):
                tv = tv.eval() if tv.kind != 'shared' else tv.get_value()
            if tv is None:
                tv = np.zeros(self.shape if isinstance(self.shape, tuple) else (self.shape,))
            return np.asarray(tv, dtype=self.dtype)

    class Continuous(Distribution):
        def __init__(self, shape=(), dtype=None, defaults=('median', 'mean', 'mode'), *args, **kwargs):
            super().__init__(shape, dtype, defaults=defaults, *args, **kwargs)

    class NoDistribution(Distribution):
        def logp(self, value):
            return 0

    class Flat(Continuous):
        def logp(self, value):
            return tt.zeros_like(value)

    class _Unused(Continuous):
        def logp(self, value):
            raise NotImplementedError(type(self).__name__ + ': outside the shim')

    class TransformedDistribution(Distribution):
        def __init__(self, dist, transform, *args, **kwargs):
            testval = transform.forward(dist.default()).eval()
            super().__init__(shape=dist.shape, dtype=dist.dtype, testval=testval)
            self.dist = dist
            self.transform_used = transform

        def logp(self, x):
            return self.dist.logp(self.transform_used.backward(x)) + self.transform_used.jacobian_det(x)

    class Transform:
        name = ''

        def apply(self, dist):
            return TransformedDistribution.dist(dist, self)

        def jacobian_det(self, x):
            raise NotImplementedError

    class ElemwiseTransform(Transform):
        pass

    class Log(ElemwiseTransform):
        name = 'log'

        def backward(self, x):
            return tt.exp(x)

        def forward(self, x):
            return tt.log(x)

        def jacobian_det(self, x):
            return x

    def Potential(name, var, model=None):
        model = modelcontext(model)
        var = lazy.as_tensor_variable(var)
        var.name = name
        model.potentials.append(var)
        model.named_vars[name] = var
        return var

    # ---- graph helpers / bijection
    def inputvars(a):
        if not isinstance(a, (list, tuple)):
            a = [a]
        return [v for v in lazy.graph_inputs(a) if isinstance(v, TensorVariable)]

    def cont_inputs(f):
        return [v for v in inputvars(f) if v.dtype.startswith('float')]

    class VarMap:
        def __init__(self, var, slc, shp, dtyp):
            self.var, self.slc, self.shp, self.dtyp = var, slc, shp, dtyp

    class ArrayOrdering:
        def __init__(self, vars):
            self.vmap = []
            dim = 0
            for var in vars:
                shp = tuple(np.shape(var.tag.test_value))
                size = int(np.prod(shp))
                self.vmap.append(VarMap(str(var), slice(dim, dim + size), shp, var.dtype))
                dim += size
            self.dimensions = dim

    class DictToArrayBijection:
        def __init__(self, ordering, dpoint):
            self.ordering = ordering
            self.dpt = dpoint

        def map(self, dpt):
            apt = np.empty(self.ordering.dimensions)
            for m in self.ordering.vmap:
                apt[m.slc] = np.asarray(dpt[m.var]).ravel()
            return apt

        def rmap(self, apt):
            dpt = dict(self.dpt)
            for m in self.ordering.vmap:
                dpt[m.var] = np.asarray(apt)[m.slc].reshape(m.shp).astype(m.dtyp)
            return dpt

    def _mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    def _outside(*a, **k):
        raise NotImplementedError('outside the shim')

    transforms = _mod('pymc3.distributions.transforms', ElemwiseTransform=ElemwiseTransform, Transform=Transform,
                      log=Log())
    distributions = _mod('pymc3.distributions', transforms=transforms, Continuous=Continuous,
                         Distribution=Distribution)
    model_mod = _mod('pymc3.model', Model=Model, FreeRV=FreeRV, ObservedRV=ObservedRV,
                     TransformedRV=TransformedRV, modelcontext=modelcontext)
    plots = _mod('pymc3.plots', utils=None, artists=None)
    tracetab = _mod('pymc3.backends.tracetab', create_flat_names=_outside)
    backends = _mod('pymc3.backends', tracetab=tracetab)

    class _Log:
        def warning(self, *a, **k):
            pass

    _mod('pymc3', Model=Model, modelcontext=modelcontext, Continuous=Continuous, Flat=Flat,
         NoDistribution=NoDistribution, Uniform=_Unused, Exponential=_Unused, Normal=_Unused, StudentT=_Unused,
         Potential=Potential, inputvars=inputvars, cont_inputs=cont_inputs, gradient=_outside, jacobian=_outside,
         DictToArrayBijection=DictToArrayBijection, ArrayOrdering=ArrayOrdering, distributions=distributions,
         model=model_mod, plots=plots, backends=backends, _log=_Log(), __version__='shim-3.x')

"""Stand-ins for `theano`, `pymc3` and the plotting/sampling imports of the reference -- TEST INFRASTRUCTURE.

`install()` registers fake modules in `sys.modules` so that `import g3py` (the unmodified reference under
/root/reference) works in this container.  Only the API surface the reference touches is provided (see
`lazy.py` for the evaluation semantics).  Used solely by `tests/golden/make_reference_goldens.py`; the
generated fixtures are committed, this shim never runs on the GPU box and is never imported by the product.
"""
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np
import scipy.special
import torch

from . import lazy
from .lazy import op, _T

STUB_ROOTS = ('matplotlib', 'seaborn', 'emcee', 'ipywidgets', 'IPython', 'statsmodels', 'mpl_toolkits')


class _Dummy:
    """Attribute/call sink for plotting and sampler libraries that are imported but never exercised."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Dummy()


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Dummy()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# ---------------------------------------------------------------------------------------------- tt
def _elem(fn):
    return lambda x: op(lambda t: fn(_T(t)), x)


def _switch(c, a, b):
    def f(c, a, b):
        a, b = _T(a), _T(b)
        dt = torch.promote_types(a.dtype, b.dtype)
        return torch.where(_T(c).bool(), a.to(dt), b.to(dt))
    return op(f, c, a, b)


def _cmp(fn):
    return lambda a, b: op(lambda x, y: fn(_T(x), _T(y) if not isinstance(y, torch.Tensor) else y), a, b)


def _concatenate(tensors, axis=0):
    return op(lambda ts: torch.cat([_T(t).reshape(-1) if _T(t).dim() == 0 else _T(t) for t in ts], dim=axis),
              list(tensors))


def _diag(x):
    return op(lambda t: torch.diag(t), x)


def _alloc(fill):
    def f(shape, dtype=None):
        return op(lambda s: torch.full(lazy._shape_arg(s), fill, dtype=getattr(torch, str(dtype or config.floatX))),
                  shape)
    return f


def _eye(n, m=None, k=0, dtype=None):
    return op(lambda n, m: torch.eye(int(n), int(m) if m is not None else int(n),
                                     dtype=getattr(torch, str(dtype or config.floatX))), n, m)


def _tt_sum(x, axis=None, dtype=None, keepdims=False):
    return op(lazy._sum, x, axis)


def _tt_prod(x, axis=None, dtype=None, keepdims=False, **kw):
    return op(lazy._prod, x, axis)


def _tt_add(*xs):
    r = xs[0]
    for x in xs[1:]:
        r = r + x
    return r


def _var(x, axis=None):
    return op(lambda t: t.var(unbiased=False) if axis is None else t.var(dim=axis, unbiased=False), x)


def _gammaln(x):
    return op(lambda t: torch.lgamma(_T(t)), x)


def _solve(A, b):
    return op(lambda A, b: torch.linalg.solve(A, b), A, b)


def _solve_tri(lower):
    def f(A, b):
        def g(A, b):
            if b.dim() == 1:
                return torch.linalg.solve_triangular(A, b[:, None], upper=not lower)[:, 0]
            return torch.linalg.solve_triangular(A, b, upper=not lower)
        return op(g, A, b)
    return f


class _Solve:
    def __init__(self, A_structure='general', lower=False, **kw):
        self.A_structure = A_structure
        self.lower = lower

    def __call__(self, A, b):
        if self.A_structure == 'lower_triangular':
            return _solve_tri(True)(A, b)
        if self.A_structure == 'upper_triangular':
            return _solve_tri(False)(A, b)
        return _solve(A, b)


class _Config:
    floatX = 'float64'
    mode = 'FAST_RUN'
    on_unused_input = 'ignore'
    warn_float64 = 'ignore'
    cast_policy = 'custom'
    int_division = 'int'

    class lib:
        amdlibm = False


config = _Config()


def _print_op(name=''):
    return lambda x: x


def uninstall():
    """Remove the stand-in modules again (unit tests call this so that nothing leaks into other test modules)."""
    for name in [n for n in sys.modules if n.split('.')[0] in ('theano', 'pymc3') + STUB_ROOTS]:
        mod = sys.modules[name]
        if name.split('.')[0] in STUB_ROOTS and not isinstance(mod, _StubModule):
            continue
        if name.split('.')[0] == 'theano' and not getattr(sys.modules.get('theano'), '_g3b_shim', False):
            continue
        del sys.modules[name]
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, _StubFinder)]


def install():
    if 'theano' in sys.modules and getattr(sys.modules['theano'], '_g3b_shim', False):
        return
    sys.meta_path.insert(0, _StubFinder())
    if not hasattr(np, 'float'):
        np.float = float        # the reference predates NumPy 1.24 (mappings.py:129)

    def typed(ndim_unused):
        return lambda name=None, dtype=None: lazy.input_var(name, dtype or config.floatX)

    tt = _mod(
        'theano.tensor',
        TensorVariable=lazy.TensorVariable, as_tensor_variable=lazy.as_tensor_variable,
        scalar=typed(0), vector=typed(1), matrix=typed(2),
        log=_elem(torch.log), exp=_elem(torch.exp), sqrt=_elem(torch.sqrt), abs_=_elem(torch.abs),
        sgn=_elem(torch.sign), sin=_elem(torch.sin), cos=_elem(torch.cos), sinh=_elem(torch.sinh),
        cosh=_elem(torch.cosh), tanh=_elem(torch.tanh), arcsinh=_elem(torch.asinh), arcsin=_elem(torch.asin),
        log1p=_elem(torch.log1p), gammaln=_gammaln, isnan=_elem(torch.isnan), isinf=_elem(torch.isinf),
        isnan_=_elem(torch.isnan), isinf_=_elem(torch.isinf), zeros_like=_elem(torch.zeros_like),
        pow=lambda a, b: op(lazy._pow, a, b),
        maximum=_cmp(torch.maximum), minimum=_cmp(torch.minimum),
        eq=_cmp(torch.eq), neq=_cmp(torch.ne), le=_cmp(torch.le), lt=_cmp(torch.lt), ge=_cmp(torch.ge),
        gt=_cmp(torch.gt), or_=_cmp(torch.logical_or), and_=_cmp(torch.logical_and),
        any=lambda x, axis=None: op(lambda t: _T(t).bool().any(), x),
        all=lambda x, axis=None: op(lambda t: _T(t).bool().all(), x),
        switch=_switch, sum=_tt_sum, prod=_tt_prod, add=_tt_add,
        mean=lambda x, axis=None: op(lazy._mean, x, axis), var=_var,
        min=lambda x, axis=None: op(lazy._min, x, axis), max=lambda x, axis=None: op(lazy._max, x, axis),
        dot=lambda a, b: op(lazy._dot, a, b), concatenate=_concatenate,
        diag=_diag, diagonal=lambda x: op(lambda t: torch.diagonal(t), x),
        eye=_eye, zeros=_alloc(0.0), ones=_alloc(1.0),
        tril=lambda x, k=0: op(lambda t: torch.tril(t, k), x), triu=lambda x, k=0: op(lambda t: torch.triu(t, k), x),
        flatten=lambda x, ndim=1: op(lambda t: _T(t).reshape(-1), x),
        grad=lazy.grad,
        jacobian=lazy.jacobian,
    )
    tsl = _mod('theano.tensor.slinalg', solve=_solve, solve_lower_triangular=_solve_tri(True),
               solve_upper_triangular=_solve_tri(False), Solve=_Solve)
    tnl = _mod('theano.tensor.nlinalg', diag=_diag, extract_diag=lambda x: op(lambda t: torch.diagonal(t), x),
               alloc_diag=lambda x: op(lambda t: torch.diag(t), x))
    tt.slinalg, tt.nlinalg = tsl, tnl
    ife = _mod('theano.ifelse', ifelse=lazy.ifelse)
    gof = _mod('theano.gof', Op=lazy.Op, Apply=lazy.Apply)
    printing = _mod('theano.printing', Print=_print_op, pydotprint=_Dummy(), debugprint=_Dummy())
    scan_module = _mod('theano.scan_module', until=lazy.Until)
    sandbox_linalg = _mod('theano.sandbox.linalg', det=_Dummy())
    sandbox = _mod('theano.sandbox', linalg=sandbox_linalg)

    gradient = _mod('theano.gradient', DisconnectedType=lazy.DisconnectedType, grad_undefined=lazy.grad_undefined)
    th = _mod('theano', tensor=tt, ifelse=ife, gof=gof, printing=printing, scan_module=scan_module, gradient=gradient,
              sandbox=sandbox, config=config, shared=lazy.shared, function=lazy.Function, scan=lazy.scan,
              In=_Dummy, __version__='shim-1.0 (torch %s)' % torch.__version__, _g3b_shim=True)
    th.__path__ = []
    tt.__path__ = []
    sandbox.__path__ = []

    from . import pymc3_mod
    pymc3_mod.install(lazy, tt, config)

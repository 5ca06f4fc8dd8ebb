"""Lazy expression graph evaluated with torch (CPU, dtype-preserving) -- TEST INFRASTRUCTURE ONLY.

This is the core of a *minimal stand-in for the Theano API* whose only purpose is to let the UNMODIFIED
reference (`/root/reference/g3py`, Theano + PyMC3 code) execute inside this container, where neither Theano
nor PyMC3 can be installed, so that golden input/output vectors can be generated from the reference's own
source (see `tests/golden/make_reference_goldens.py`).  Nothing in the product (`g3py_b200/`) imports it.

Semantics kept on purpose:
  * symbolic graph first, numbers later: `theano.function(inputs, outputs, givens=...)` evaluates the graph
    for the bound inputs; shared variables can be replaced through `givens`;
  * dtypes are preserved: `np.float32` literals stay float32 constants, so an expression made only of float32
    constants (e.g. `tt.log(np.float32(2*pi))`) is computed in float32 exactly as Theano's constant folding
    does, while float32-constant (op) float64-tensor promotes to float64;
  * `tt.grad` is reverse-mode autodiff of the *un-optimised* graph (torch.autograd), custom `Op`s run their own
    `perform` and their own symbolic `grad`;
  * `switch` evaluates both branches (NaN can leak through the gradient of the unselected branch exactly as in
    Theano), `ifelse` evaluates one branch.
"""
import numpy as np
import torch

torch.set_grad_enabled(True)


class _Tag:
    pass


def _has_node(a):
    if isinstance(a, Node):
        return True
    if isinstance(a, (list, tuple)):
        return any(_has_node(x) for x in a)
    if isinstance(a, slice):
        return _has_node((a.start, a.stop, a.step))
    return False


def to_tensor(a):
    """numpy / python value -> torch value, dtype preserved (python scalars stay weakly typed)."""
    if isinstance(a, torch.Tensor):
        return a
    if isinstance(a, np.ndarray):
        return torch.from_numpy(np.array(a, copy=True))
    if isinstance(a, np.generic):
        return torch.from_numpy(np.array(a))
    return a


def _deep_eval(a, env):
    if isinstance(a, Node):
        return evaluate(a, env)
    if isinstance(a, list):
        return [_deep_eval(x, env) for x in a]
    if isinstance(a, tuple):
        return tuple(_deep_eval(x, env) for x in a)
    if isinstance(a, slice):
        return slice(_deep_eval(a.start, env), _deep_eval(a.stop, env), _deep_eval(a.step, env))
    return to_tensor(a)


class MissingInputError(Exception):
    pass


def evaluate(node, env):
    """env: {'v': {id(node): tensor}, 'givens': {id(shared): replacement}, 'test': bool}"""
    if not isinstance(node, Node):
        return to_tensor(node)
    memo = env['v']
    key = id(node)
    if key in memo:
        return memo[key]
    giv = env.get('givens')
    if giv and key in giv:
        t = evaluate(giv[key], env)
    elif node.kind == 'const':
        t = node.value
    elif node.kind == 'shared':
        t = torch.from_numpy(np.array(node.value, copy=True))
        if t.is_floating_point():
            t.requires_grad_(True)
    elif node.kind == 'input':
        if env.get('test') and hasattr(node.tag, 'test_value'):
            t = torch.from_numpy(np.array(node.tag.test_value, copy=True))
            if str(t.dtype).replace('torch.', '') != node.dtype and t.is_floating_point():
                t = t.to(getattr(torch, node.dtype))
            if t.is_floating_point():
                t.requires_grad_(True)
        else:
            raise MissingInputError('no value for input %r' % (node.name,))
    elif node.kind == 'op':
        args = [_deep_eval(a, env) for a in node.args]
        kwargs = {k: _deep_eval(v, env) for k, v in node.kwargs.items()}
        t = node.fn(*args, **kwargs)
    elif node.kind == 'ifelse':
        c = evaluate(node.args[0], env)
        t = evaluate(node.args[1] if bool(c) else node.args[2], env)
        t = to_tensor(t)
        if not isinstance(t, torch.Tensor):
            t = torch.tensor(t)
    elif node.kind == 'grad':
        f = evaluate(node.args[0], env)
        w = evaluate(node.args[1], env)
        if not (isinstance(f, torch.Tensor) and f.requires_grad):
            t = torch.zeros_like(w)
        else:
            (g,) = torch.autograd.grad(f, w, retain_graph=True, allow_unused=True)
            t = torch.zeros_like(w) if g is None else g
    elif node.kind == 'jacobian':
        f = evaluate(node.args[0], env)
        w = evaluate(node.args[1], env)
        rows = []
        for i in range(f.numel()):
            (g,) = torch.autograd.grad(f.reshape(-1)[i], w, retain_graph=True, create_graph=True, allow_unused=True)
            rows.append(torch.zeros_like(w) if g is None else g)
        t = torch.stack(rows)
    elif node.kind == 'scan':
        t = _run_scan(node, env)
    elif node.kind == 'opout':
        akey = ('apply', id(node.apply))
        if akey not in memo:                          # all outputs of one Apply come from a single perform()
            xs = [evaluate(a, env) for a in node.apply.inputs]
            memo[akey] = _run_op(node.apply, xs)
        t = memo[akey][node.out_index]
    else:
        raise RuntimeError(node.kind)
    memo[key] = t
    return t


def _new_env(givens=None, test=False):
    return {'v': {}, 'givens': givens or {}, 'test': test}


def _to_numpy(t):
    if isinstance(t, torch.Tensor):
        return t.detach().numpy().copy()
    return np.asarray(t)


# --------------------------------------------------------------------------------------------------
class TensorType:
    def __init__(self, dtype, ndim_of=None):
        self.dtype = dtype
        self._ndim_of = ndim_of

    def __call__(self, name=None):
        v = TensorVariable('pending', name=name, dtype=self.dtype)
        return v

    make_variable = __call__


class Node:
    __array_ufunc__ = None      # numpy scalars defer to our reflected operators
    __array_priority__ = 1000

    def __init__(self, kind, fn=None, args=(), kwargs=None, name=None, dtype='float64', value=None):
        self.kind = kind
        self.fn = fn
        self.args = tuple(args)
        self.kwargs = kwargs or {}
        self.name = name
        self.dtype = dtype
        self.value = value
        self.tag = _Tag()
        self.apply = None

    # -- graph children (for input discovery)
    def children(self):
        out = []

        def walk(a):
            if isinstance(a, Node):
                out.append(a)
            elif isinstance(a, (list, tuple)):
                for x in a:
                    walk(x)
            elif isinstance(a, slice):
                walk((a.start, a.stop, a.step))
        if self.kind == 'opout':
            walk(list(self.apply.inputs))
        else:
            walk(list(self.args))
            walk(list(self.kwargs.values()))
        return out

    @property
    def owner(self):
        return None if self.kind in ('input', 'shared', 'const') else self

    # -- evaluation helpers
    def eval(self, inputs_to_values=None):
        env = _new_env()
        if inputs_to_values:
            for k, v in inputs_to_values.items():
                env['v'][id(k)] = _bind(k, v)
        return _to_numpy(evaluate(self, env))

    @property
    def test_value(self):
        return _to_numpy(evaluate(self, _new_env(test=True)))

    @property
    def ndim(self):
        return self.test_value.ndim

    @property
    def type(self):
        return TensorType(self.dtype)

    @property
    def shape(self):
        return op(lambda t: torch.tensor(list(t.shape), dtype=torch.int64), self)

    @property
    def size(self):
        return op(lambda t: torch.tensor(t.numel(), dtype=torch.int64), self)

    @property
    def T(self):
        return op(lambda t: t.T if t.dim() == 2 else t, self)

    def __repr__(self):
        return self.name if self.name else '<%s>' % self.kind
    __str__ = __repr__

    def __hash__(self):
        return id(self)

    def __bool__(self):
        raise TypeError('symbolic variables have no truth value')

    def __len__(self):
        raise TypeError('symbolic variables have no length')

    def __iter__(self):
        raise TypeError('symbolic variables are not iterable')

    # -- arithmetic
    def __add__(self, o): return op(_add, self, o)
    def __radd__(self, o): return op(_add, o, self)
    def __sub__(self, o): return op(_sub, self, o)
    def __rsub__(self, o): return op(_sub, o, self)
    def __mul__(self, o): return op(_mul, self, o)
    def __rmul__(self, o): return op(_mul, o, self)
    def __truediv__(self, o): return op(_div, self, o)
    def __rtruediv__(self, o): return op(_div, o, self)
    def __pow__(self, o): return op(_pow, self, o)
    def __rpow__(self, o): return op(_pow, o, self)
    def __neg__(self): return op(lambda a: -a, self)
    def __abs__(self): return op(torch.abs, self)
    def __lt__(self, o): return op(lambda a, b: _T(a) < b, self, o)
    def __le__(self, o): return op(lambda a, b: _T(a) <= b, self, o)
    def __gt__(self, o): return op(lambda a, b: _T(a) > b, self, o)
    def __ge__(self, o): return op(lambda a, b: _T(a) >= b, self, o)

    def __getitem__(self, idx):
        return op(_getitem, self, idx)

    # -- methods used by the reference
    def dot(self, o): return op(_dot, self, o)
    def astype(self, dtype): return op(_astype, self, str(dtype))
    def flatten(self, ndim=1): return op(lambda t: t.reshape(-1), self)
    def reshape(self, shape, ndim=None): return op(_reshape, self, shape)
    def sum(self, axis=None, dtype=None, keepdims=False): return op(_sum, self, axis)
    def mean(self, axis=None): return op(_mean, self, axis)
    def prod(self, axis=None, dtype=None): return op(_prod, self, axis)
    def max(self, axis=None): return op(_max, self, axis)
    def min(self, axis=None): return op(_min, self, axis)

    def dimshuffle(self, *pattern):
        if len(pattern) == 1 and isinstance(pattern[0], (list, tuple)):
            pattern = tuple(pattern[0])
        return op(_dimshuffle, self, tuple(pattern))


class TensorVariable(Node):
    """Symbolic variable (what `isinstance(v, tt.TensorVariable)` is true for)."""


class TensorConstant(Node):
    pass


class SharedVariable(Node):
    def get_value(self, borrow=False):
        return np.array(self.value, copy=True)

    def set_value(self, value, borrow=False):
        self.value = np.array(value, dtype=self.dtype, copy=True)


def _T(a):
    return a if isinstance(a, torch.Tensor) else torch.tensor(a)


def _add(a, b): return a + b
def _sub(a, b): return a - b
def _mul(a, b): return a * b
def _div(a, b):
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return torch.true_divide(_T(a), b) if not isinstance(b, torch.Tensor) or not isinstance(a, torch.Tensor) \
            else torch.true_divide(a, b)
    return a / b


def _pow(a, b):
    if not isinstance(a, torch.Tensor) and not isinstance(b, torch.Tensor):
        return a ** b
    return torch.pow(a, b) if isinstance(a, torch.Tensor) else torch.pow(torch.tensor(float(a), dtype=b.dtype), b)


def _getitem(t, idx):
    def conv(i):
        if isinstance(i, torch.Tensor):
            return int(i) if i.dim() == 0 and not i.is_floating_point() else i.long()
        if isinstance(i, slice):
            return slice(conv(i.start), conv(i.stop), conv(i.step))
        return i
    if isinstance(idx, tuple):
        return t[tuple(conv(i) for i in idx)]
    return t[conv(idx)]


def _dot(a, b):
    a, b = _T(a), _T(b)
    if a.dtype != b.dtype:
        dt = torch.promote_types(a.dtype, b.dtype)
        a, b = a.to(dt), b.to(dt)
    if a.dim() == 0 or b.dim() == 0:
        return a * b
    return torch.matmul(a, b)


def _astype(t, dtype):
    return _T(t).to(getattr(torch, dtype))


def _shape_arg(shape):
    if isinstance(shape, torch.Tensor):
        return [int(s) for s in shape.reshape(-1)]
    if isinstance(shape, (list, tuple)):
        return [int(s) for s in shape]
    return [int(shape)]


def _reshape(t, shape):
    return t.reshape(_shape_arg(shape))


def _sum(t, axis=None):
    if isinstance(t, (list, tuple)):
        t = torch.stack([_T(x) for x in t]) if len(t) else torch.zeros(0, dtype=torch.float64)
    t = _T(t)
    if t.dtype == torch.bool:
        t = t.to(torch.int64)
    return t.sum() if axis is None else t.sum(dim=axis)


def _mean(t, axis=None):
    return _T(t).mean() if axis is None else _T(t).mean(dim=axis)


def _prod(t, axis=None):
    return _T(t).prod() if axis is None else _T(t).prod(dim=axis)


def _max(t, axis=None):
    return _T(t).max() if axis is None else _T(t).max(dim=axis).values


def _min(t, axis=None):
    return _T(t).min() if axis is None else _T(t).min(dim=axis).values


def _dimshuffle(t, pattern):
    perm = [p for p in pattern if p != 'x']
    dropped = [d for d in range(t.dim()) if d not in perm]
    for d in dropped:
        assert t.shape[d] == 1
    t = t.permute(*(perm + dropped)).reshape([t.shape[p] for p in perm])
    for pos, p in enumerate(pattern):
        if p == 'x':
            t = t.unsqueeze(pos)
    return t


def op(fn, *args, **kwargs):
    """Build an op node; an expression without symbolic inputs is folded to a constant right away."""
    if not _has_node(args) and not _has_node(tuple(kwargs.values())):
        val = fn(*[_deep_eval(a, None) for a in args], **{k: _deep_eval(v, None) for k, v in kwargs.items()})
        return constant(val)
    return TensorVariable('op', fn=fn, args=args, kwargs=kwargs)


def constant(val, name=None):
    t = to_tensor(val)
    if not isinstance(t, torch.Tensor):
        t = torch.tensor(t)
    t = t.detach()
    return TensorConstant('const', value=t, name=name, dtype=str(t.dtype).replace('torch.', ''))


def as_tensor_variable(x, name=None, ndim=None):
    if isinstance(x, Node):
        return x
    return constant(np.asarray(x) if not isinstance(x, (np.ndarray, np.generic)) else x, name=name)


def input_var(name=None, dtype='float64'):
    return TensorVariable('input', name=name, dtype=str(dtype))


def shared(value, name=None, borrow=False, allow_downcast=None, **kw):
    value = np.array(value, copy=True)
    return SharedVariable('shared', name=name, dtype=str(value.dtype), value=value)


def _bind(var, value):
    arr = np.asarray(value)
    if isinstance(var, Node) and var.dtype.startswith('float') and arr.dtype.kind in 'fiub':
        arr = arr.astype(var.dtype)        # allow_input_downcast / upcast to the variable's dtype
    t = torch.from_numpy(np.array(arr, copy=True))
    if t.is_floating_point():
        t.requires_grad_(True)
    return t


# --------------------------------------------------------------------------------------------------
class Function:
    """theano.function: positional inputs in order, keyword inputs by variable name."""

    def __init__(self, inputs, outputs, givens=None, **kw):
        self.inputs = list(inputs)
        self.outputs = outputs
        self.givens = {}
        if givens:
            items = givens.items() if isinstance(givens, dict) else givens
            for k, v in items:
                self.givens[id(k)] = v
        self._keep = givens

    def __call__(self, *args, **kwargs):
        env = _new_env(self.givens)
        names = {v.name: v for v in self.inputs}
        for var, val in zip(self.inputs, args):
            env['v'][id(var)] = _bind(var, val)
        for k, val in kwargs.items():
            if k in names:
                env['v'][id(names[k])] = _bind(names[k], val)
            else:
                raise TypeError('unknown input %r' % k)
        outs = self.outputs
        if isinstance(outs, (list, tuple)):
            return [_to_numpy(evaluate(o, env)) for o in outs]
        return _to_numpy(evaluate(outs, env))


def grad(cost, wrt, disconnected_inputs='raise', **kw):
    if isinstance(wrt, (list, tuple)):
        return [grad(cost, w) for w in wrt]
    return TensorVariable('grad', args=(cost, wrt))


def jacobian(expression, wrt, **kw):
    return TensorVariable('jacobian', args=(expression, wrt))


class Until:
    """theano.scan_module.until(condition)."""

    def __init__(self, condition):
        self.condition = condition


def scan(fn, sequences=None, outputs_info=None, non_sequences=None, n_steps=None, **kw):
    """The one form the reference uses (libs/tensors.py:134-145): a single recurrent output, one non-sequence,
    `fn` returning (next value, until(condition)).  Returns (all step values stacked, updates)."""
    assert sequences is None and n_steps is not None
    x_in = input_var('scan_x')
    z_in = input_var('scan_z')
    r, until = fn(x_in, z_in)
    node = TensorVariable('scan', args=(outputs_info, non_sequences, r, until.condition))
    node.scan_inputs = (x_in, z_in)
    node.n_steps = int(n_steps)
    return node, {}


def _run_scan(node, env):
    init, nonseq, r, cond = node.args
    x_in, z_in = node.scan_inputs
    x = evaluate(init, env).detach()
    z = evaluate(nonseq, env).detach()
    base = dict(env['v'])
    vals = []
    for _ in range(node.n_steps):
        e2 = {'v': dict(base), 'givens': env.get('givens'), 'test': env.get('test')}
        e2['v'][id(x_in)] = x.clone().requires_grad_(True)
        e2['v'][id(z_in)] = z
        x = evaluate(r, e2).detach()
        vals.append(x)
        if bool(evaluate(cond, e2)):
            break
    return torch.stack(vals)


def ifelse(cond, a, b, name=None):
    if not _has_node((cond,)):
        return a if bool(np.asarray(cond)) else b
    return TensorVariable('ifelse', args=(cond, a, b))


def graph_inputs(variables):
    """Leaves of the graph, depth-first, left-most operand first (theano.gof.graph.inputs order)."""
    seen, out = set(), []
    stack = list(reversed(list(variables)))
    while stack:
        v = stack.pop()
        if not isinstance(v, Node) or id(v) in seen:
            continue
        seen.add(id(v))
        if v.kind in ('input', 'shared', 'const'):
            out.append(v)
        else:
            stack.extend(reversed(v.children()))
    return out


# --------------------------------------------------------------------------------------------------
class Apply:
    def __init__(self, op, inputs, outputs):
        self.op = op
        self.inputs = list(inputs)
        self.outputs = list(outputs)


class DisconnectedType:
    """theano.gradient.DisconnectedType: `DisconnectedType()()` marks an input no gradient flows to."""

    def __call__(self):
        return self


class _Undefined:
    pass


def grad_undefined(op, x_pos, x, comment=''):
    return _Undefined()


class Op:
    """theano.gof.Op protocol: make_node / perform / grad; one or several outputs."""

    def __call__(self, *inputs):
        node = self.make_node(*inputs)
        for i, out in enumerate(node.outputs):
            out.kind = 'opout'
            out.apply = node
            out.out_index = i
        return node.outputs[0] if len(node.outputs) == 1 else list(node.outputs)


class _OpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, apply, *xs):
        storage = [[None] for _ in apply.outputs]
        apply.op.perform(apply, [x.detach().numpy() for x in xs], storage)
        ctx.g3_node = apply
        ctx.save_for_backward(*xs)
        outs = tuple(torch.from_numpy(np.array(s[0], copy=True)) for s in storage)
        return outs

    @staticmethod
    def backward(ctx, *gs):
        xs = ctx.saved_tensors
        with torch.enable_grad():
            xin = [constant(x.detach()) for x in xs]
            gin = [constant(torch.zeros(()) if g is None else g.detach()) for g in gs]
            grads = ctx.g3_node.op.grad(xin, gin)
            vals = []
            for x, v in zip(xs, grads):
                if v is None or isinstance(v, (DisconnectedType, _Undefined)):
                    vals.append(None)
                else:
                    t = evaluate(v, _new_env())
                    t = t if isinstance(t, torch.Tensor) else torch.tensor(t)
                    vals.append(t.detach().to(x.dtype).reshape(x.shape) if x.is_floating_point() else None)
        return (None, *vals)


def _run_op(apply, xs):
    return _OpFunction.apply(apply, *xs)              # tuple, one tensor per output

"""Minimal stand-in for the parts of Theano's Op protocol that g3py's own Ops use
(g3py/libs/tensors.py:111-263): gof.Op (__props__ equality, __call__ -> make_node), gof.Apply,
tensor.as_tensor_variable, gradient.grad_undefined / DisconnectedType.  Enough to drive
make_node / perform / grad of g3py_b200.theano_ops without Theano (absent from this image)."""
import types

import numpy as np


class Var:
    def __init__(self, value=None, owner=None, index=0, ndim=None, dtype="float64"):
        self.value = None if value is None else np.asarray(value, dtype=dtype)
        self.owner, self.index = owner, index
        self.ndim = self.value.ndim if self.value is not None else ndim
        self.dtype = dtype

    def type(self):
        return Var(ndim=self.ndim, dtype=self.dtype)

    def __mul__(self, other):
        return Mul(self, other)
    __rmul__ = __mul__

    def eval(self):
        if self.value is not None:
            return self.value
        node = self.owner
        ins = [v.eval() for v in node.inputs]
        storage = [[None] for _ in node.outputs]
        node.op.perform(node, ins, storage)
        for o, s in zip(node.outputs, storage):
            o.value = np.asarray(s[0])
        return self.value


class Mul(Var):
    def __init__(self, a, b):
        super().__init__(ndim=max(getattr(a, "ndim", 0) or 0, getattr(b, "ndim", 0) or 0))
        self.a, self.b = a, b

    def eval(self):
        f = lambda v: v.eval() if isinstance(v, Var) else np.asarray(v)
        return f(self.a) * f(self.b)


class Apply:
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)
        for i, o in enumerate(self.outputs):
            o.owner, o.index = self, i


class Op:
    __props__ = ()

    def _key(self):
        return (type(self).__name__,) + tuple(getattr(self, p) for p in self.__props__)

    def __eq__(self, other):
        return type(self) is type(other) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __call__(self, *inputs):
        node = self.make_node(*inputs)
        return node.outputs[0] if len(node.outputs) == 1 else node.outputs


class DisconnectedType:
    def __call__(self):
        return "disconnected"


def make_module():
    th = types.SimpleNamespace()
    th.gof = types.SimpleNamespace(Op=Op, Apply=Apply)
    th.tensor = types.SimpleNamespace(as_tensor_variable=lambda x: x if isinstance(x, Var) else Var(x))
    th.gradient = types.SimpleNamespace(grad_undefined=lambda op, i, x: "undefined", DisconnectedType=DisconnectedType)
    return th

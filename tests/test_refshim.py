"""Unit tests of the Theano/PyMC3 API stand-in (tests/golden/refshim) that produced tests/golden/reference_g3py.json.
They need neither the reference nor a GPU: they pin the evaluation semantics the goldens rely on (float32 constant
folding, NaN leaking through the unselected `switch` branch, `ifelse` laziness, the `Op` protocol with a symbolic
`grad`, `givens`, the transformed-variable bookkeeping and the dict<->array bijection)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def mods():
    import torch
    prev = torch.get_default_dtype()
    import refshim
    refshim.install()
    torch.set_default_dtype(torch.float64)
    import theano as th
    import theano.tensor as tt
    import pymc3 as pm
    th.config.floatX = "float64"
    yield th, tt, pm
    torch.set_default_dtype(prev)
    refshim.uninstall()


def test_float32_constants_fold_in_float32(mods):
    th, tt, pm = mods
    c = tt.log(np.float32(2.0 * np.pi))                  # gaussian.py:218
    v = c.eval()
    assert v.dtype == np.float32 and float(v) == 1.8378770351409912
    x = tt.vector("x")
    y = np.float32(-0.5) * x * c                          # float32 constants times a float64 tensor -> float64
    out = th.function([x], y)(np.array([2.0, 4.0]))
    assert out.dtype == np.float64 and np.allclose(out, [-1.8378770351409912, -3.6757540702819824], rtol=0, atol=0)


def test_switch_leaks_nan_gradient_like_theano(mods):
    """d/dr of (1 + s) exp(-s), s = sqrt(3 * 0.5 r^2 d2), summed over d2 = [0, 1]: the d2 = 0 element is 0/0."""
    th, tt, pm = mods
    r = tt.scalar("r")
    d2 = np.array([0.0, 1.0])
    s = tt.sqrt(3 * (0.5 * r ** 2 * d2))
    k = tt.sum((1 + s) * tt.exp(-s))
    g = th.function([r], tt.grad(k, r))(0.7)
    assert np.isnan(g)                                    # the reference scrubs this to 0 (stochastic.py:308-309)
    safe = tt.switch(tt.isnan(tt.grad(k, r)), np.float32(0), tt.grad(k, r))
    assert th.function([r], safe)(0.7) == 0.0


def test_ifelse_is_lazy_and_blocks_the_gradient(mods):
    th, tt, pm = mods
    from theano.ifelse import ifelse
    x = tt.scalar("x")
    f = ifelse(x > 0, tt.log(x), np.float32(-1e30))
    fn = th.function([x], [f, tt.grad(f, x)])
    v, g = fn(2.0)
    assert v == pytest.approx(np.log(2.0)) and g == pytest.approx(0.5)
    v, g = fn(-1.0)
    assert float(v) == float(np.float32(-1e30)) and g == 0.0


def test_op_protocol_perform_and_symbolic_grad(mods):
    th, tt, pm = mods

    class Square(th.gof.Op):
        def make_node(self, x):
            x = tt.as_tensor_variable(x)
            return th.gof.Apply(self, [x], [x.type()])

        def perform(self, node, inputs, outputs):
            outputs[0][0] = inputs[0] ** 2

        def grad(self, inputs, gradients):
            return [gradients[0] * 2 * inputs[0]]

    x = tt.vector("x")
    y = tt.sum(Square()(x) * np.array([1.0, 10.0]))
    val, g = th.function([x], [y, tt.grad(y, x)])(np.array([3.0, 4.0]))
    assert val == 169.0 and np.array_equal(g, [6.0, 80.0])


def test_function_givens_replace_shared_variables(mods):
    th, tt, pm = mods
    s = th.shared(np.array([1.0, 2.0]), name="s")
    s_in = tt.vector("s_in")
    a = tt.scalar("a")
    f = th.function([s_in, a], tt.sum(s * a), givens=[(s, s_in)])
    assert f(np.array([10.0, 20.0]), a=2.0) == 60.0
    assert tt.sum(s).eval() == 3.0


def test_transformed_variable_and_bijection(mods):
    th, tt, pm = mods

    class NonTransformLog(pm.distributions.transforms.ElemwiseTransform):
        name = "log"

        def backward(self, x):
            return tt.exp(x)

        def forward(self, x):
            return tt.log(x)

        def jacobian_det(self, x):
            return tt.switch(tt.exp(x) > 1e-6, 0, -np.inf)

    with pm.Model() as model:
        b = pm.Flat("b", testval=np.zeros(()), dtype="float64")
        v = pm.Flat("v", transform=NonTransformLog(), shape=2, testval=np.array([2.0, 3.0]), dtype="float64")
    assert [x.name for x in model.vars] == ["b", "v_log__"] and type(v) is pm.model.TransformedRV
    assert np.allclose(model.test_point["v_log__"], np.log([2.0, 3.0]))
    bij = pm.DictToArrayBijection(pm.ArrayOrdering(pm.inputvars(model.cont_vars)), model.test_point)
    arr = bij.map({"b": np.array(0.5), "v_log__": np.array([1.0, -1.0])})
    assert np.array_equal(arr, [0.5, 1.0, -1.0])
    back = bij.rmap(arr)
    assert back["b"] == 0.5 and np.array_equal(back["v_log__"], [1.0, -1.0])
    # free-RV logp: Flat (0) + jacobian_det: 0, or -inf behind the 1e-6 barrier (hypers/__init__.py:200-201)
    f = th.function(model.vars, tt.sum(model.vars[1].logpt))
    assert f(b=0.0, v_log__=np.array([0.0, 0.0])) == 0.0
    assert f(b=0.0, v_log__=np.array([0.0, -20.0])) == -np.inf

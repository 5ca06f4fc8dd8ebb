"""Seeded synthetic inputs of the BASELINE.json configs.  Product-side generators
(bench.py must not import the oracle); tests check they equal the oracle's."""
import numpy as np


def c1_inputs():
    rng = np.random.default_rng(0)
    x = np.linspace(0, 10, 200)[:, None]
    y = np.sin(x[:, 0]) + 0.1 * rng.standard_normal(200)
    return x, y


def c2_inputs(N=4096, B=64, theta_seed=2):
    """GP SE+Matern-5/2 ARD on D=3, B hyper samples: theta = [Bias, SE var, SE rate[3], MAT52 var,
    MAT52 rate[3], Noise var] in log space for the positive ones."""
    rng = np.random.default_rng(1)
    X = rng.uniform(0, 10, size=(N, 3))
    f = np.sin(X[:, 0]) + np.cos(X[:, 1] / 2) + 0.1 * X[:, 2]
    y = f + 0.1 * rng.standard_normal(N)
    rng2 = np.random.default_rng(theta_seed)
    tbar = np.array([0.0, 1.0, 1.0, 1.0, 1.0, 0.5, 0.5, 0.5, 0.5, 0.05])
    Theta = np.tile(np.concatenate([[0.0], np.log(tbar[1:])]), (B, 1))
    Theta = Theta + 0.1 * rng2.standard_normal(Theta.shape)
    return X, y, Theta


def c3_inputs(N=2048, M=10000):
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0, 20, size=N))[:, None]
    f = np.sin(2 * np.pi * x[:, 0] / 5.0) * np.exp(-0.02 * x[:, 0]) + 0.05 * rng.standard_normal(N)
    y = np.exp(0.3 * f) + 0.5
    xs = np.linspace(0, 20, M)[:, None]
    return x, y, xs


def c4_inputs(N=16384, M=4096):
    rng = np.random.default_rng(4)
    X = rng.standard_normal((N, 5))
    f = np.sin(X[:, 0]) + 0.5 * X[:, 1] * X[:, 2] + np.cos(X[:, 3]) - 0.3 * X[:, 4]
    y = f + 0.1 * rng.standard_normal(N)
    Xs = rng.standard_normal((M, 5))
    return X, y, Xs


def c5_inputs(N=65536):
    rng = np.random.default_rng(5)
    X = rng.uniform(0, N ** (1.0 / 3.0), size=(N, 3))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
    return X, y

"""Golden vectors from the UNMODIFIED reference (`/root/reference/g3py`) -> tests/golden/reference_g3py.json.

The reference is Theano + PyMC3 code; neither can be installed in this container (no network, not in the
wheelhouse).  `tests/golden/refshim/` provides stand-ins for the slice of the Theano / PyMC3 API that g3py
touches (lazy expression graph evaluated with torch CPU, dtype-preserving; reverse-mode autodiff for
`tt.grad`; the reference's own `CholeskyRobust.perform` / `.grad` run as written), so that

    import g3py;  gp = g3py.GP(x, g3py.Bias(), g3py.SE(x));  gp.observed(x, y);  gp.logp(); gp.dlogp(); gp.predict()

execute the reference's own source files: kernels.py, metrics.py, means.py, mappings.py, tensors.py,
elliptical.py, gaussian.py, studentT.py, stochastic.py, models.py.  What this pins and what it cannot:
  + every formula, constant (incl. float32 literals), guard, hyper-parameter order and transform in those files;
  + the NaN->0 scrub of Matern rate gradients (Theano's sqrt gradient at d=0; torch has the same 0/0);
  - Theano's graph *optimiser* (rewrites may change rounding at the 1e-16 level and NaN propagation), its
    BLAS/LAPACK build, and `graph.inputs` ordering of `dlogp` components (stored per variable name here).

Run in this container only (needs /root/reference):   python tests/golden/make_reference_goldens.py
The fixture is committed; tests read the JSON, never the reference.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')


def _import_reference():
    import refshim
    refshim.install()
    import torch
    torch.set_default_dtype(torch.float64)
    sys.dont_write_bytecode = True          # never write __pycache__ into the (read-only by contract) reference tree
    sys.path.insert(0, '/root/reference')
    import g3py as g3
    import theano as th
    th.config.floatX = 'float64'           # g3py/config.py defaults to float32; this is the fp64 run
    return g3


# ------------------------------------------------------------------ neutral spec -> reference objects
def _x_arg(X, dims):
    # column subsets: the reference's `(domain, slice)` form sizes `rate` with the full D (hypers/__init__.py:63-68)
    # and then fails in `tt.dot`; the list-of-columns form (`:60-62`) is the one that works
    return X if dims is None else list(range(int(dims[0]), int(dims[1])))


def _pot(obj, spec):
    if 'potential' in spec:
        obj.set_potential(*spec['potential'])
    return obj


def ref_kernel(g3, spec, X):
    return _pot(_ref_kernel(g3, spec, X), spec)


def _ref_kernel(g3, spec, X):
    t = spec['type']
    if t == 'sum':
        return ref_kernel(g3, spec['k1'], X) + ref_kernel(g3, spec['k2'], X)
    if t == 'prod':
        return ref_kernel(g3, spec['k1'], X) * ref_kernel(g3, spec['k2'], X)
    if t == 'max':
        return g3.KernelMax(ref_kernel(g3, spec['k1'], X), ref_kernel(g3, spec['k2'], X))
    if t == 'scale':
        return spec['c'] * ref_kernel(g3, spec['k'], X)
    if t == 'shift':
        return spec['c'] + ref_kernel(g3, spec['k'], X)
    kw = {}
    if 'name' in spec:
        kw['name'] = spec['name']
    if spec.get('var') is not None:
        kw['var'] = spec['var']
    if t == 'POL' and 'p' in spec:
        kw['p'] = spec['p']
    for key in ('eq', 'eq1', 'eq2'):
        if key in spec:
            kw[key] = spec[key]
    cls = g3.KernelNoise if t == 'Noise' else getattr(g3, t)
    return cls(_x_arg(X, spec.get('dims')), **kw)


def ref_transport(g3, spec, X):
    parts = []
    for t in spec['chain']:
        if t['t'] == 'ID':
            parts.append(g3.ID())
        elif t['t'] == 'TMapping':
            parts.append(g3.TMapping(getattr(g3, t['mapping']['type'])()))
        elif t['t'] == 'TLocation':
            parts.append(g3.TLocation(getattr(g3, t['location']['type'])(_x_arg(X, t['location'].get('dims')))))
        elif t['t'] == 'TScale':
            sc = t['scale']
            parts.append(g3.TScale(getattr(g3, sc['type'])(_x_arg(X, sc.get('dims')), **({'name': sc['name']} if 'name' in sc else {}))))
        else:
            parts.append(g3.TKernel(ref_kernel(g3, t['kernel'], X), noisy=t.get('noisy', False)))
    tr = parts[0]
    for t in parts[1:]:
        tr = tr @ t
    return tr


def ref_mapping(g3, mp):
    if mp['type'] == 'composed':
        return ref_mapping(g3, mp['m1']) @ ref_mapping(g3, mp['m2'])
    if mp['type'] == 'invsum':
        from g3py.processes.hypers.mappings import MappingInvSum
        return MappingInvSum(ref_mapping(g3, mp['m1']), ref_mapping(g3, mp['m2']))
    if mp['type'] == 'BoxCoxLinear2':
        from g3py.processes.hypers.mappings import BoxCoxLinear2
        return _pot(BoxCoxLinear2(**({'name': mp['name']} if 'name' in mp else {})), mp)
    mkw = {'name': mp['name']} if 'name' in mp else {}
    if 'n' in mp:
        mkw['n'] = mp['n']
    return _pot(getattr(g3, mp['type'])(**mkw), mp)


def ref_process(g3, spec, X):
    kind = spec.get('kind', 'gauss')
    if kind == 'transport':
        return g3.TransportGaussianProcess(X, ref_transport(g3, spec, X))
    mp = spec.get('mapping', {'type': 'Identity'})
    warped = spec.get('warped', mp['type'] != 'Identity')
    cls = {('gauss', False): g3.GP, ('gauss', True): g3.WGP, ('student', False): g3.TP,
           ('student', True): g3.WTP}[(kind, warped)]
    loc = spec.get('location', {'type': 'Zero'})
    lkw = {'name': loc['name']} if 'name' in loc else {}
    if loc['type'] == 'Power':
        from g3py.processes.hypers.means import Power
        location = _pot(Power(_x_arg(X, loc.get('dims')), n=loc.get('n', 2), **lkw), loc)
    elif loc['type'] == 'BlackBox':
        from g3py.processes.hypers.means import BlackBox
        import theano.tensor as tt                      # the element is sliced with a symbolic length: it must be a tensor
        location = BlackBox(tt.as_tensor_variable(np.asarray(loc['element'], dtype=np.float64)), _x_arg(X, loc.get('dims')), **lkw)
    else:
        location = _pot(getattr(g3, loc['type'])(_x_arg(X, loc.get('dims')), **lkw), loc)
    mapping = ref_mapping(g3, mp)
    kw = {'name': spec['name']} if 'name' in spec else {}
    return cls(X, location, ref_kernel(g3, spec['kernel'], X), mapping, noisy=spec.get('noisy', True), **kw)


# ------------------------------------------------------------------ cases
def _data(seed, N, D, M, positive=False):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 6.0, size=(N, D))
    f = np.sin(X[:, 0]) + (0.5 * np.cos(0.7 * X[:, 1]) if D > 1 else 0.0) + (0.1 * X[:, -1] if D > 2 else 0.0)
    y = f + 0.1 * rng.standard_normal(N)
    if positive:
        y = np.exp(0.5 * y) + 0.2
    Xs = rng.uniform(0.0, 6.0, size=(M, D))
    return X, y, Xs


K = lambda t, **kw: dict(type=t, **kw)
CASES = {
    # BASELINE configs, down-sized
    'C1_gp_se':        dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('SE')), N=60, D=1, M=17, seed=10),
    'C2_gp_se_mat52':  dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('sum', k1=K('SE'), k2=K('MAT52'))),
                            N=64, D=3, M=13, seed=11),
    'C3_wgp_sinxse':   dict(spec=dict(kind='gauss', warped=True, location=K('Bias'),
                                      kernel=K('prod', k1=K('SIN'), k2=K('SE')), mapping=K('BoxCoxShifted')),
                            N=48, D=1, M=15, seed=12, positive=True),
    'C4_tp_se':        dict(spec=dict(kind='student', location=K('Bias'), kernel=K('SE')), N=56, D=5, M=11, seed=13),
    # multi-tile sizes (3 tiles of 128 with padding) straight from the reference
    'C2_mid_3tiles':   dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('sum', k1=K('SE'), k2=K('MAT52'))),
                            N=300, D=3, M=9, seed=14, slim=True),
    'C4_mid_3tiles':   dict(spec=dict(kind='student', location=K('Bias'), kernel=K('SE')), N=270, D=5, M=9, seed=15,
                            slim=True),
    # leaf zoo (SURVEY a2 + f4)
    'leaf_ou':         dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('OU')), N=24, D=2, M=7, seed=20),
    'leaf_mat32':      dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('MAT32')), N=24, D=2, M=7, seed=21),
    'leaf_rq':         dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('RQ')), N=24, D=2, M=7, seed=22),
    'leaf_sin':        dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('SIN')), N=24, D=1, M=7, seed=23,
                            sin_rate=0.001),
    'leaf_sin_indef':  dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('SIN')), N=24, D=1, M=7, seed=23,
                            sin_rate=0.1),    # indefinite K: logp through the jitter ladder, posterior through LU
    'leaf_cos':        dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('sum', k1=K('COS'), k2=K('SE'))),
                            N=24, D=2, M=7, seed=24),
    'leaf_sinc':       dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('sum', k1=K('SINC'), k2=K('SE'))),
                            N=24, D=1, M=7, seed=25),
    'leaf_sm':         dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('SM')), N=24, D=2, M=7, seed=26),
    'leaf_wn':         dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('sum', k1=K('SE'), k2=K('WN'))),
                            N=24, D=2, M=7, seed=27),
    # dot-product / Brownian / constant leaves and KernelMax (SURVEY f-4): non-stationary, diag depends on x
    'leaf_lin':        dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('sum', k1=K('LIN'), k2=K('SE'))),
                            N=24, D=2, M=7, seed=70),
    'leaf_pol3':       dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('POL', p=3)), N=24, D=2, M=7, seed=71),
    'leaf_dot_bw_var': dict(spec=dict(kind='gauss', location=K('Zero'),
                                      kernel=K('sum', k1=K('sum', k1=K('KernelDot'), k2=K('BW')), k2=K('VAR'))),
                            N=24, D=2, M=7, seed=72),
    'alg_max':         dict(spec=dict(kind='gauss', location=K('Zero'),
                                      kernel=K('max', k1=K('SE'), k2=K('scale', c=0.5, k=K('MAT32')))),
                            N=24, D=2, M=7, seed=73),
    # pm.Potential regularisers (hypers/__init__.py:94-109)
    'potentials':      dict(spec=dict(kind='gauss', warped=True, location=dict(type='Bias', potential=['Bias', 'L2', 0.3]),
                                      kernel=dict(type='sum', potential=['var', 'L1', 0.7],
                                                  k1=dict(type='SE'), k2=dict(type='RQ', potential=['alpha', 'L2', 0.2])),
                                      mapping=dict(type='BoxCoxShifted', potential=['power', 'L1', 0.4])),
                            N=24, D=2, M=7, seed=80, positive=True),
    # operator algebra (a4), column subsets, fixed variances
    'alg_scale_shift': dict(spec=dict(kind='gauss', location=K('Bias'),
                                      kernel=K('shift', c=0.3, k=K('scale', c=1.7, k=K('SE')))), N=24, D=2, M=7, seed=30),
    'alg_prod_dims':   dict(spec=dict(kind='gauss', location=K('Linear'),
                                      kernel=K('sum', k1=K('prod', k1=K('SE', dims=[0, 2]), k2=K('MAT32', dims=[2, 3])),
                                               k2=K('RQ', dims=[1, 3], name='RQb'))), N=32, D=3, M=9, seed=31),
    # mappings (a7 det_m, a12)
    'map_logshifted':  dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('SE'), mapping=K('LogShifted')),
                            N=24, D=1, M=7, seed=40, positive=True),
    'map_boxcoxlin':   dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('SE'), mapping=K('BoxCoxLinear')),
                            N=24, D=1, M=7, seed=41, positive=True),
    'map_arcsinh':     dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('SE'), mapping=K('ArcsinhLinear')),
                            N=24, D=1, M=7, seed=42),
    'map_sinharcsinh': dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('SE'), mapping=K('SinhArcsinh')),
                            N=24, D=1, M=7, seed=43),
    'map_linear':      dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('SE'), mapping=K('LinearMapping')),
                            N=24, D=1, M=7, seed=44),
    'map_logistic':    dict(spec=dict(kind='gauss', warped=True, location=K('Bias'), kernel=K('SE'), mapping=K('Logistic')),
                            N=24, D=1, M=7, seed=46, positive=True),
    # Newton-inverse warpings (SURVEY f-4): forward map = damped Newton with tol 1e-3 (libs/tensors.py:134-145)
    'map_warptanh':    dict(spec=dict(kind='gauss', warped=True, location=K('Bias'), kernel=K('SE'),
                                      mapping=K('WarpingTanh', n=2)), N=24, D=1, M=7, seed=47),
    'map_warpboxcox':  dict(spec=dict(kind='gauss', warped=True, location=K('Bias'), kernel=K('SE'),
                                      mapping=K('WarpingBoxCox', n=2)), N=24, D=1, M=7, seed=48, positive=True),
    # composed warpings m1 @ m2 (mappings.py:57-70): m1's hypers move the argument of m2's log-Jacobian
    'map_comp_log_lin': dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('SE'),
                                       mapping=K('composed', m1=K('LogShifted'), m2=K('LinearMapping'))),
                             N=24, D=1, M=7, seed=51, positive=True),
    'map_comp_boxcox_sas': dict(spec=dict(kind='student', warped=True, location=K('Bias'), kernel=K('MAT52'),
                                          mapping=K('composed', m1=K('BoxCoxShifted'), m2=K('SinhArcsinh'))),
                                N=28, D=2, M=8, seed=52, positive=True),
    'map_comp_three':  dict(spec=dict(kind='gauss', location=K('Bias'), kernel=K('SE'),
                                      mapping=K('composed', m1=K('composed', m1=K('LinearMapping'), m2=K('ArcsinhLinear')),
                                                m2=K('SinhArcsinh'))), N=24, D=1, M=7, seed=53),
    'map_comp_lin_warptanh': dict(spec=dict(kind='gauss', warped=True, location=K('Bias'), kernel=K('SE'),
                                            mapping=K('composed', m1=K('LinearMapping'), m2=K('WarpingTanh', n=2))),
                                  N=24, D=1, M=7, seed=54),
    # operator-surface leftovers: Power / BlackBox means, BoxCoxLinear2, MappingInvSum
    'mean_power':      dict(spec=dict(kind='gauss', location=K('Power', n=2), kernel=K('SE')), N=24, D=2, M=7, seed=90),
    'mean_blackbox':   dict(spec=dict(kind='gauss', location=K('BlackBox', element=[round(0.3 * np.sin(0.7 * i), 6) for i in range(32)]),
                                      kernel=K('MAT32')), N=24, D=2, M=7, seed=91),
    'map_boxcoxlin2':  dict(spec=dict(kind='gauss', warped=True, location=K('Bias'), kernel=K('SE'), mapping=K('BoxCoxLinear2')),
                            N=24, D=1, M=7, seed=92, positive=True),
    # NN (training Gram only: its two-argument cov does not broadcast, kernels.py:351), NIL, KernelEquals / KernelEquals2 on
    # inputs whose first column is rounded to integers so that the equality metrics are not identically zero
    'leaf_nn':         dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('NN')), N=24, D=2, M=7, seed=93, logp_only=True,
                            theta_shift={'NN_var': -1.5, 'Noise_var': 2.0}),
    'leaf_nil_equals': dict(spec=dict(kind='gauss', location=K('Bias'),
                                      kernel=K('sum', k1=K('sum', k1=K('SE'), k2=K('NIL')),
                                               k2=K('sum', k1=K('scale', c=0.3, k=K('KernelEquals', eq=1.0, dims=[0, 1])),
                                                    k2=K('scale', c=0.01, k=K('KernelEquals2', eq1=0.0, eq2=1.0, dims=[0, 1]))))),
                            N=24, D=2, M=7, seed=94, int_col=True, theta_shift={'Noise_var': 2.0}),
    # MappingInvSum cannot be pinned: its `__call__` is `pass`, so EllipticalProcess.th_define_process (elliptical.py:64:
    # tt_to_num(self.f_mapping(self.th_outputs))) raises before any method exists - dead code in the reference.
    'wtp_boxcox':      dict(spec=dict(kind='student', warped=True, location=K('Bias'), kernel=K('MAT52'),
                                      mapping=K('BoxCoxShifted')), N=32, D=2, M=9, seed=45, positive=True),
    # TransportGaussianProcess (SURVEY f-3): chains [ID | TMapping | TLocation]* @ TKernel
    'tgp_canonical':   dict(spec=dict(kind='transport', chain=[dict(t='TMapping', mapping=K('BoxCoxShifted')),
                                                               dict(t='TLocation', location=K('Bias')),
                                                               dict(t='TKernel', kernel=K('SE'), noisy=True)]),
                            N=24, D=1, M=7, seed=60, positive=True),
    'tgp_id_linear':   dict(spec=dict(kind='transport', chain=[dict(t='ID'), dict(t='TLocation', location=K('Linear')),
                                                               dict(t='TKernel', kernel=K('MAT32'), noisy=True)]),
                            N=24, D=2, M=7, seed=61),
    'tgp_bare_kernel': dict(spec=dict(kind='transport', chain=[dict(t='TKernel', kernel=K('SE'), noisy=True)]),
                            N=20, D=2, M=6, seed=62),
    'tgp_noise_free':  dict(spec=dict(kind='transport', chain=[dict(t='TLocation', location=K('Bias')),
                                                               dict(t='TKernel', kernel=K('sum', k1=K('SE'), k2=K('WN')),
                                                                    noisy=False)]),
                            N=20, D=1, M=6, seed=63),
    # TScale (transports.py:165-181) between the location and nothing: y = s (m + L eps), s a Bias mean named 'Scale'
    'tgp_scale':       dict(spec=dict(kind='transport', chain=[dict(t='TScale', scale=K('Bias', name='Scale')),
                                                               dict(t='TLocation', location=K('Bias')),
                                                               dict(t='TKernel', kernel=K('SE'), noisy=True)]),
                            N=24, D=1, M=7, seed=64, theta_shift={'Scale_Bias': 1.2}),
    # tt_to_cov (a5): a negative shift makes the Gram diagonal <= 0, so the reference adds (1e-6 - min diag) I.
    # logp only: Theano differentiates THROUGH the shift, and its max/min gradient gives every tied element the full
    # upstream gradient (all N diagonal entries of a stationary kernel tie), which torch (even split) does not mimic;
    # the oracle and the CUDA path treat the shift as a constant.
    'tt_to_cov_shift': dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('shift', c=-3.0, k=K('SE'))),
                            N=20, D=1, M=5, seed=51, logp_only=True, skip_dlogp=True),
    # robustness (a5, a6): noise-free kernel on duplicated inputs -> jitter ladder of CholeskyRobust
    'jitter_ladder':   dict(spec=dict(kind='gauss', location=K('Zero'), kernel=K('SE'), noisy=False), N=20, D=1, M=5,
                            seed=50, duplicate=True, logp_only=True),   # LU `tsl.solve` of the posterior is singular here
}


def theta_for(layout, y, rng, case):
    """Moderate hyper-parameters in the oracle's layout order (positive ones as logs)."""
    th = []
    for name, size, pos in layout:
        v = 0.15 * rng.standard_normal(size)
        if name.endswith('Noise_var'):
            v += np.log(0.05 * max(np.nanvar(y), 1e-3))
        elif name.endswith('_var'):
            v += np.log(max(np.nanvar(y), 1e-3))
        elif name.endswith('SIN_rate'):
            # SIN = exp(+2 r sin^2) >= its own diagonal (kernels.py:472): K is indefinite unless r is tiny
            v = np.full(size, np.log(case.get('sin_rate', 0.01)))
        elif name.endswith('_bias'):
            v += np.log(0.5)
        elif name.endswith('LIN_rate') or name.endswith('POL_rate') or name.endswith('KernelDot_rate'):
            v += np.log(0.3)
        elif name.endswith('_freq'):
            v += np.log(0.2)
        elif name.endswith('SM_rate'):
            v += np.log(0.15)
        elif name.endswith('Freedom_degree'):
            v = np.full(size, np.log(5.0))
        elif name.endswith('_power'):
            v = np.full(size, np.log(0.7))
        elif name.endswith('_Bias') or name.endswith('_Constant'):
            v += np.nanmean(y) if 'mapping' not in case['spec'] else 0.0
        elif name.endswith('_Coeff'):
            v *= 0.3
        elif name.endswith('Logistic_lower'):
            v = np.full(size, np.nanmin(y) - 0.4)
        elif name.endswith('Logistic_high'):
            v = np.full(size, np.log(np.nanmax(y) - np.nanmin(y) + 0.9))
        elif name.endswith('Logistic_location'):
            v = np.full(size, 0.1)
        elif name.endswith('WarpingTanh_a'):
            v += np.log(0.3)
        elif name.endswith('WarpingTanh_c'):
            v += -np.nanmean(y)
        elif name.endswith('WarpingBoxCox_w'):
            v += np.log(0.5)
        elif name.endswith('LogShifted_shift'):
            v = np.full(size, np.nanmin(y) - 0.5)
        elif name.endswith('_shift'):
            v = 0.05 * v
        for suffix, shift in case.get('theta_shift', {}).items():      # per-case nudges (log space for positive hypers)
            if name.endswith(suffix):
                v = v + shift
        th.append(v)
    return np.concatenate(th) if th else np.zeros(0)


def short(name, proc):
    s = name[len(proc) + 1:]
    for suf in ('_log__', '_log_'):
        if s.endswith(suf):
            s = s[:-len(suf)]
    return s


def shim_self_check(g3):
    """The stand-in against REAL Theano output: notebooks/07-Student-t-Process.ipynb (cell 4) prints r1, r2, r3, det_m of
    the two evaluations `_compile_methods` makes on the 2-point dummy data set (float32 Theano run).  The reference
    executed through the stand-in must reproduce their sums to float32 print precision."""
    x = np.linspace(1700, 2008, 309)[:, None]
    tp = g3.WarpedStudentTProcess(x, g3.Bias(), g3.SE(x), g3.ArcsinhLinear())
    X2, y2 = np.array([[0.0], [1.0]]), np.array([0.0, 1.0])
    rows = [(np.zeros(7), (-0.8902489542961121, -0.7392648458480835, -0.6449083089828491, -0.3465735912322998)),
            (np.array([0.5, np.log(0.25), np.log(0.5), np.log(0.25), 0.5, np.log(0.5), 0.0]),
             (-0.9840160012245178, -0.7392648458480835, 0.8014175891876221, -1.732867956161499))]
    out = []
    for th, terms in rows:
        got = float(tp.logp(th, inputs=X2, outputs=y2, array=True))
        assert abs(got - sum(terms)) < 5e-6, (got, sum(terms))
        out.append(dict(theta=th.tolist(), notebook_terms=list(terms), notebook_sum=sum(terms), shim_logp=got))
    with open(os.path.join(HERE, 'reference_shim_check.json'), 'w') as f:
        json.dump(dict(source='notebooks/07-Student-t-Process.ipynb cell 4 (real Theano, float32)', rows=out), f, indent=1)
    print('shim vs real-Theano notebook prints:', [(r['shim_logp'], r['notebook_sum']) for r in out])


def c1_find_map(g3):
    """BASELINE config 1 (the tutorial example) at full size: the reference's own default hypers and its own find_MAP
    (scipy BFGS through its logp / dlogp), -> tests/golden/reference_c1_find_map.json."""
    from g3py_b200 import workloads
    x, y = workloads.c1_inputs()
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    gp.observed(x, y)
    names = [m.var for m in gp.active.bijection.ordering.vmap]
    p0 = gp.params
    th0 = gp.active.dict_to_array(p0)
    pm_ = gp.find_MAP(start=None, points=1, display=False, plot=False, powell=False)
    thm = gp.active.dict_to_array(pm_)
    rec = dict(names=names, theta_default=th0.tolist(), logp_default=float(gp.logp(th0, array=True)),
               dlogp_default=np.asarray(gp.dlogp(th0, array=True)).tolist(),
               theta_map=thm.tolist(), logp_map=float(gp.logp(thm, array=True)),
               dlogp_map=np.asarray(gp.dlogp(thm, array=True)).tolist())
    with open(os.path.join(HERE, 'reference_c1_find_map.json'), 'w') as f:
        json.dump(rec, f, indent=1)
    print('C1 find_MAP (reference): logp %.6f -> %.6f, theta_MAP %s' % (rec['logp_default'], rec['logp_map'], thm))


def fullsize(g3):
    """BASELINE configs 2 and 3 at FULL size from the executed reference (inputs are the seeded generators of
    g3py_b200/workloads.py, so only theta and the outputs are stored) -> tests/golden/reference_fullsize.json."""
    import time
    import pymc3 as pm
    from g3py_b200 import workloads
    out = {}

    def grads_by_name(proc, th):
        wrt = pm.inputvars(pm.cont_inputs(proc.th_logp()))
        flat = np.asarray(proc.dlogp(th, array=True), dtype=np.float64)
        off, d = 0, {}
        for v in wrt:
            size = int(np.prod(np.shape(v.tag.test_value)))
            d[v.name] = flat[off:off + size].tolist()
            off += size
        return d

    # config 2: GP Bias + SE + MAT52 (ARD, D=3) + noise, N=4096, first two rows of the 64-sample theta batch
    t0 = time.time()
    X, y, Theta = workloads.c2_inputs(4096, 64)
    gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X))
    gp.observed(X, y)
    names = [m.var for m in gp.active.bijection.ordering.vmap]
    rows = []
    for b in (0, 1):
        rows.append(dict(b=b, theta=Theta[b].tolist(), logp=float(gp.logp(Theta[b], array=True)),
                         dlogp=grads_by_name(gp, Theta[b])))
    out['C2'] = dict(N=4096, names=names, rows=rows)
    print('C2 full size: logp', [r['logp'] for r in rows], '%.0f s' % (time.time() - t0))
    # config 3: warped GP, periodic x SE, BoxCoxShifted, N=2048; posterior on 64 of the 10k test points
    t0 = time.time()
    x, y, xs = workloads.c3_inputs(2048, 10000)
    wgp = g3.WGP(x, g3.Bias(), g3.SIN(x) * g3.SE(x), g3.BoxCoxShifted())
    wgp.observed(x, y)
    names = [m.var for m in wgp.active.bijection.ordering.vmap]
    th = wgp.active.dict_to_array(wgp.params)
    lay = {n: i for i, n in enumerate(names)}
    th[lay['WGP_SIN_rate_log__']] = np.log(0.01)
    th[lay['WGP_SIN_freq_log__']] = np.log(0.2)
    th[lay['WGP_BoxShift_power_log__']] = np.log(0.7)
    sel = np.arange(0, 10000, 157)[:64]
    kw = dict(params=wgp.active.array_to_dict(th), space=xs[sel], inputs=x, outputs=y)
    out['C3'] = dict(N=2048, names=names, theta=th.tolist(), space_index=sel.tolist(),
                     logp=float(wgp.logp(th, array=True)), dlogp=grads_by_name(wgp, th),
                     location=np.asarray(wgp.location(prior=False, noise=False, **kw)).tolist(),
                     kernel_diag=np.asarray(wgp.kernel_diag(prior=False, noise=False, **kw)).tolist(),
                     mean=np.asarray(wgp.mean(prior=False, noise=False, **kw)).tolist(),
                     variance=np.asarray(wgp.variance(prior=False, noise=False, **kw)).tolist())
    print('C3 full size: logp', out['C3']['logp'], '%.0f s' % (time.time() - t0))
    # config 4 (Student-t process, SE ARD, D=5) at N=4096 -- the reference's N x N x D metric tensor makes N=16384
    # (10.7 GB, plus autodiff copies) impractical here; same generator, first 4096 rows
    t0 = time.time()
    X4, y4, Xs4 = workloads.c4_inputs(16384, 4096)
    X4, y4, Xs4 = X4[:4096], y4[:4096], Xs4[:32]
    tp = g3.TP(X4, g3.Bias(), g3.SE(X4))
    tp.observed(X4, y4)
    names = [m.var for m in tp.active.bijection.ordering.vmap]
    th = tp.active.dict_to_array(tp.params)
    lay = {n: i for i, n in enumerate(names)}
    th[lay['TP_Freedom_degree_log__']] = np.log(5.0)
    th[lay['TP_Noise_var_log__']] = np.log(0.05)
    kw = dict(params=tp.active.array_to_dict(th), space=Xs4, inputs=X4, outputs=y4)
    out['C4_4096'] = dict(N=4096, names=names, theta=th.tolist(), logp=float(tp.logp(th, array=True)),
                          dlogp=grads_by_name(tp, th),
                          location=np.asarray(tp.location(prior=False, noise=True, **kw)).tolist(),
                          variance=np.asarray(tp.variance(prior=False, noise=True, **kw)).tolist(),
                          quantile_up=np.asarray(tp.quantiler(q=0.975, prior=False, noise=True, **kw)).tolist())
    print('C4 at N=4096: logp', out['C4_4096']['logp'], '%.0f s' % (time.time() - t0))
    with open(os.path.join(HERE, 'reference_fullsize.json'), 'w') as f:
        json.dump(out, f)


def main():
    g3 = _import_reference()
    shim_self_check(g3)
    if not any(a.startswith('--only=') for a in sys.argv):
        c1_find_map(g3)
        if '--no-fullsize' not in sys.argv:
            fullsize(g3)
    from oracle import g3_oracle as orc     # only for the neutral layout (names / order), not for any value
    out = {}
    only = [a.split('=', 1)[1].split(',') for a in sys.argv if a.startswith('--only=')]
    if only:                                 # regenerate the named cases only and merge them into the committed file
        with open(os.path.join(HERE, 'reference_g3py.json')) as f:
            out = json.load(f)
    for cname, case in CASES.items():
        if only and cname not in only[0]:
            continue
        X, y, Xs = _data(case['seed'], case['N'], case['D'], case['M'], case.get('positive', False))
        if case.get('duplicate'):
            X[1::2] = X[0::2]                       # exact duplicates: singular noise-free Gram
            y[1::2] = y[0::2]
        if case.get('int_col'):
            X[:, 0] = np.floor(X[:, 0] / 2.0)            # {0, 1, 2}
            Xs[:, 0] = np.floor(Xs[:, 0] / 2.0)
        spec = case['spec']
        proc = ref_process(g3, spec, X)
        proc.observed(X, y)
        vmap = proc.active.bijection.ordering.vmap
        names = [m.var for m in vmap]
        layout = orc.build_process(spec, X.shape[1]).layout()
        assert [short(n, proc.name) for n in names] == [l[0] for l in layout], (names, layout)
        assert [int(np.prod(m.shp)) for m in vmap] == [l[1] for l in layout]
        assert [n.endswith('__') for n in names] == [bool(l[2]) for l in layout]
        th = theta_for(layout, y, np.random.default_rng(1000 + case['seed']), case)
        params = proc.active.array_to_dict(th)
        rec = dict(spec=spec, X=X.tolist(), y=y.tolist(), Xs=Xs.tolist(), theta=th.tolist(),
                   ref_names=names, layout=[list(l) for l in layout],
                   default_params={k: np.asarray(v, dtype=np.float64).ravel().tolist()
                                   for k, v in proc.params_default.items()})
        transport = spec.get('kind') == 'transport'
        if transport:
            rec['min_eig_K'] = 1.0
        else:
            Kin = np.asarray(proc.kernel(params=params, space=X, inputs=X, outputs=y, prior=True, noise=True))
            rec['min_eig_K'] = float(np.linalg.eigvalsh(0.5 * (Kin + Kin.T)).min())
        rec['logp'] = float(proc.logp(th, array=True))
        rec['logp_prior'] = float(proc.logp(th, array=True, prior=True))
        rec['loglike'] = float(proc.loglike(th, array=True))
        # dlogp: reference order is graph-traversal order -> store per variable
        import pymc3 as pm
        wrt = pm.inputvars(pm.cont_inputs(proc.th_logp()))
        flat = np.asarray(proc.dlogp(th, array=True), dtype=np.float64)
        off, d = 0, {}
        for v in wrt:
            size = int(np.prod(np.shape(v.tag.test_value)))
            d[v.name] = flat[off:off + size].tolist()
            off += size
        assert off == flat.size
        rec['dlogp_order'] = [v.name for v in wrt]
        rec['dlogp'] = d
        if case.get('skip_dlogp'):
            rec['skip_dlogp'] = True
        kw = dict(params=params, space=Xs, inputs=X, outputs=y)
        if transport:
            vec = np.random.default_rng(7000 + case['seed']).standard_normal(len(Xs))
            rec['vector'] = vec.tolist()
            for sel in ('transport', 'transport_inv', 'transport_diag'):
                for prior in (False, True):
                    for noise in (False, True):
                        # the inverse needs a vector in the range of the warping: use the transported draw
                        v_in = vec
                        if sel == 'transport_inv' and prior:
                            v_in = np.asarray(proc.transport(vector=vec, prior=True, noise=noise, **kw))
                        out_v = getattr(proc, sel)(vector=v_in, prior=prior, noise=noise, **kw)
                        rec['%s_prior%d_noise%d' % (sel, prior, noise)] = np.asarray(out_v, dtype=np.float64).tolist()
            out[cname] = rec
            print('%-18s N=%3d P=%2d logp=% .12e  |dlogp|=%.3e  (transport)' % (cname, len(y), len(th), rec['logp'],
                  np.linalg.norm(flat)))
            continue
        for noise in (() if case.get('logp_only') else (False, True)):
            r = {}
            for key in ('location', 'kernel_diag', 'kernel_sd', 'mean', 'median', 'variance', 'std'):
                r[key] = np.asarray(getattr(proc, key)(prior=False, noise=noise, **kw), dtype=np.float64).tolist()
            r['kernel'] = np.asarray(proc.kernel(prior=False, noise=noise, **kw)).tolist()
            r['prior_kernel'] = np.asarray(proc.kernel(prior=True, noise=noise, **kw)).tolist()
            r['quantile_up'] = np.asarray(proc.quantiler(q=0.975, prior=False, noise=noise, **kw)).tolist()
            r['quantile_down'] = np.asarray(proc.quantiler(q=0.025, prior=False, noise=noise, **kw)).tolist()
            if hasattr(proc, 'covariance') and proc.th_covariance() is not None:
                r['covariance'] = np.asarray(proc.covariance(prior=False, noise=noise, **kw)).tolist()
            rec['post_noise%d' % noise] = r
        if spec.get('kind', 'gauss') == 'gauss' and not case.get('logp_only'):
            v = np.asarray(rec['post_noise1']['median']) * 1.01 + 0.01
            rec['logpredictive_at'] = v.tolist()
            rec['logpredictive'] = float(proc.logpredictive(vector=v, prior=False, noise=True, **kw))
        if spec.get('kind') == 'student':
            rec['freedom_post'] = float(proc.freedom(prior=False, **kw))
        out[cname] = rec
        print('%-18s N=%3d P=%2d logp=% .12e  |dlogp|=%.3e  min eig K=% .2e' % (cname, len(y), len(th), rec['logp'],
              np.linalg.norm(flat), rec['min_eig_K']))
    with open(os.path.join(HERE, 'reference_g3py.json'), 'w') as f:
        json.dump(out, f)
    print('wrote reference_g3py.json', os.path.getsize(os.path.join(HERE, 'reference_g3py.json')), 'bytes')


if __name__ == '__main__':
    main()

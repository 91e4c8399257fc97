"""Benchmark instances (data only): the frozen problem matrices under data/model_*.npz (package data) and the
controller built from them.  The matrices were produced once by oracle/models.py from the reference's
own model files (notebooks/cart_pole_with_walls/mld_dynamics.py, controller.py:8-27); loading them
needs neither the reference nor the oracle."""
import os
import numpy as np

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


def load_model(name):
    z = np.load(os.path.join(DATA, 'model_%s.npz' % name), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d['name'] = str(d['name']); d['nub'] = int(d['nub']); d['T'] = int(d['T'])
    return d


def controller_from_model(model, **kw):
    from .mld_system import MLDSystem
    from .controller import HybridModelPredictiveController
    mld = MLDSystem([model['A'], model['B']], [model['F'], model['G'], model['h']], int(model['nub']))
    return HybridModelPredictiveController(mld, int(model['T']), [model['Q'], model['R'], model['Q_T']],
                                           [model['F_T'], model['h_T']], **kw)


def load_initial_states(lo=0, hi=None):
    """Frozen set of 4096 feasible initial states of the two-wall cart-pole (BASELINE configs[2])."""
    x = np.load(os.path.join(DATA, 'cp20_instances.npy'))
    hi = len(x) if hi is None else hi
    return np.ascontiguousarray(x[np.arange(lo, hi) % len(x)])

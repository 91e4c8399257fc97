"""Benchmark instances (data only): the frozen problem matrices under tests/golden/model_*.npz and the
controller built from them.  The matrices were produced once by oracle/models.py from the reference's
own model files (notebooks/cart_pole_with_walls/mld_dynamics.py, controller.py:8-27); loading them
needs neither the reference nor the oracle."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def load_model(name):
    z = np.load(os.path.join(GOLDEN, 'model_%s.npz' % name), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d['name'] = str(d['name']); d['nub'] = int(d['nub']); d['T'] = int(d['T'])
    return d


def controller_from_model(model, **kw):
    from .mld_system import MLDSystem
    from .controller import HybridModelPredictiveController
    mld = MLDSystem([model['A'], model['B']], [model['F'], model['G'], model['h']], int(model['nub']))
    return HybridModelPredictiveController(mld, int(model['T']), [model['Q'], model['R'], model['Q_T']],
                                           [model['F_T'], model['h_T']], **kw)

"""ctypes binding of the C ABI declared in include/wshmpc.h (libwshmpc.so, built in-tree by
``__graft_entry__.build()``).  PyTorch is used only to own device memory and streams.

There is NO CPU fallback: a missing library or a missing GPU raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libwshmpc.so')
_lib = None


class _Problem(C.Structure):
    _fields_ = ([(k, C.c_int) for k in ('nx', 'nu', 'nub', 'T', 'nh', 'nh1', 'nq', 'nqT', 'nr', 'n', 'm', 'mc', 'nb')]
                + [(k, C.c_void_p) for k in ('A', 'B', 'F', 'G', 'h', 'F_Tm1', 'G_Tm1', 'h_Tm1', 'Q', 'R', 'Q_T',
                                             'M_mu', 'M_rho', 'Mh', 'nrm', 'vscale', 'Eh', 'hh', 'Rinv', 'Kx',
                                             'Zmap', 'bin_idx')]
                + [(k, C.c_double) for k in ('eps', 'tol_p', 'tol_d', 'tol_sing', 'tol_ray', 'prox_tol')]
                + [(k, C.c_int) for k in ('max_iter', 'max_prox')])


class Layout(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('primal', 'dual', 'off_lam', 'off_mu', 'off_nu_lb', 'off_nu_ub',
                                       'off_rho', 'off_sigma')]


def load_library():
    """Loads libwshmpc.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('CUDA library %s not built: run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU fallback for the hot path)' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.wshmpc_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(lib, name)           # AttributeError if the ABI is incomplete
        _lib = lib
    return _lib


EXPORTS = ('wshmpc_last_error', 'wshmpc_create', 'wshmpc_destroy', 'wshmpc_get_layout', 'wshmpc_solve_nodes')


def _check(rc):
    if rc != 0:
        raise RuntimeError('wshmpc error %d: %s' % (rc, load_library().wshmpc_last_error().decode()))


class Handle(object):
    """Owns one wshmpc_handle (one GPU, one stream, `n_slots` solver states)."""

    def __init__(self, pd, device=0, n_slots=1, stream=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('no CUDA device: the B&B hot path has no CPU fallback')
        self.lib = load_library()
        self.pd = pd
        self.device = device
        self.torch_device = torch.device('cuda', device)
        self.n_slots = n_slots
        p = _Problem()
        for k in ('nx', 'nu', 'nub', 'T', 'nh', 'nh1', 'nq', 'nqT', 'nr', 'n', 'm', 'mc', 'nb', 'max_iter', 'max_prox',
                  'eps', 'tol_p', 'tol_d', 'tol_sing', 'tol_ray', 'prox_tol'):
            setattr(p, k, getattr(pd, k))
        for k in ('A', 'B', 'F', 'G', 'h', 'F_Tm1', 'G_Tm1', 'h_Tm1', 'Q', 'R', 'Q_T', 'M_mu', 'M_rho', 'Mh', 'nrm',
                  'vscale', 'Eh', 'hh', 'Rinv', 'Kx', 'Zmap', 'bin_idx'):
            a = getattr(pd, k)
            assert a.flags['C_CONTIGUOUS']
            setattr(p, k, a.ctypes.data)
        self._h = C.c_void_p()
        self.stream = stream
        sptr = C.c_void_p(stream.cuda_stream if stream is not None else 0)
        _check(self.lib.wshmpc_create(C.byref(p), device, n_slots, sptr, C.byref(self._h)))
        self.layout = Layout()
        _check(self.lib.wshmpc_get_layout(self._h, C.byref(self.layout)))

    def close(self):
        if self._h:
            self.lib.wshmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- K1 -------------------------------------------------------------------------------------
    def solve_nodes(self, x0, lb, ub, slot=None, hot=None):
        """x0 [N, nx], lb/ub [N, nb] (torch CUDA fp64 or numpy).  Returns dict of torch CUDA tensors."""
        import torch
        dev = self.torch_device
        t = lambda a, dt: torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a, dtype=dt, device=dev).contiguous()
        x0, lb, ub = t(x0, torch.float64), t(lb, torch.float64), t(ub, torch.float64)
        N = x0.shape[0]
        assert x0.shape == (N, self.pd.nx) and lb.shape == (N, self.pd.nb) and ub.shape == (N, self.pd.nb)
        slot = torch.arange(N, dtype=torch.int32, device=dev) % self.n_slots if slot is None else t(slot, torch.int32)
        hot = torch.zeros(N, dtype=torch.int32, device=dev) if hot is None else t(hot, torch.int32)
        if int(slot.max()) >= self.n_slots:
            raise ValueError('slot index out of range')
        out = dict(status=torch.zeros(N, dtype=torch.int32, device=dev),
                   cost=torch.zeros(N, dtype=torch.float64, device=dev),
                   dobj=torch.zeros(N, dtype=torch.float64, device=dev),
                   iters=torch.zeros(N, dtype=torch.int32, device=dev),
                   primal=torch.zeros((N, self.layout.primal), dtype=torch.float64, device=dev),
                   dual=torch.zeros((N, self.layout.dual), dtype=torch.float64, device=dev))
        P = lambda a: C.c_void_p(a.data_ptr())
        _check(self.lib.wshmpc_solve_nodes(self._h, N, P(x0), P(lb), P(ub), P(slot), P(hot), P(out['status']),
                                           P(out['cost']), P(out['dobj']), P(out['iters']), P(out['primal']),
                                           P(out['dual'])))
        return out

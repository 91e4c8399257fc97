"""MLD system container -- host-side modelling object consumed by the controller.

Mirrors the constructor contract of the reference's ``MLDSystem`` (mld_system.py:9-66): same
attribute names (A, B, F, G, h, nx, nu, nub, nuc, V) and the same ValueErrors on inconsistent sizes.
The symbolic / PWA builders of the reference (mld_system.py:68-214) are host modelling helpers and
out of scope of the hot path; build the matrices with the reference or numpy.
"""
import numpy as np


class MLDSystem(object):
    """x(t+1) = A x(t) + B u(t),  F x(t) + G u(t) <= h,  last `nub` inputs binary."""

    def __init__(self, dynamics, constraints, nub):
        [self.A, self.B] = [np.asarray(M, dtype=float) for M in dynamics]
        [self.F, self.G, self.h] = [np.asarray(M, dtype=float) for M in constraints]
        self.nx = self.A.shape[1]
        self.nu = self.B.shape[1]
        self.nub = nub
        self.nuc = self.nu - nub
        self.V = np.hstack((np.zeros((nub, self.nuc)), np.eye(nub)))
        self._check_input_sizes()

    def _check_input_sizes(self):
        if self.A.shape[0] != self.A.shape[1]:
            raise ValueError('Nonsquare A matrix.')
        if self.B.shape[0] != self.nx:
            raise ValueError('A and B matrices have incompatible size.')
        if self.F.shape != (self.h.size, self.nx):
            raise ValueError('Matrix F has incompatible size.')
        if self.G.shape != (self.h.size, self.nu):
            raise ValueError('Matrix G has incompatible size.')

"""Restatement of ``filterpy.kalman.KalmanFilter`` (the parts SORT uses).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Parity unpinned: filterpy is
an unpinned third-party dependency of the reference
(``/root/reference/environment.yml:35``; imported at
``/root/reference/tracking/sort/sort.py:30``; used at ``sort.py:97,164,172``)
and is not installed here.  Restated from the published algorithm of
``filterpy/kalman/kalman_filter.py`` (1.4.5): linear predict
``x = Fx, P = a^2 F P F' + Q`` and Joseph-form update with ``numpy.linalg.inv``.
Only the attributes the reference touches are provided.
"""
import numpy as np


class KalmanFilter:
    def __init__(self, dim_x, dim_z, dim_u=0):
        self.dim_x = dim_x
        self.dim_z = dim_z
        self.x = np.zeros((dim_x, 1))
        self.P = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.F = np.eye(dim_x)
        self.H = np.zeros((dim_z, dim_x))
        self.R = np.eye(dim_z)
        self._alpha_sq = 1.0
        self._I = np.eye(dim_x)
        self.z = np.array([[None] * dim_z]).T
        self.K = np.zeros((dim_x, dim_z))
        self.y = np.zeros((dim_z, 1))
        self.S = np.zeros((dim_z, dim_z))
        self.SI = np.zeros((dim_z, dim_z))
        self.x_prior = self.x.copy()
        self.P_prior = self.P.copy()
        self.x_post = self.x.copy()
        self.P_post = self.P.copy()
        self.inv = np.linalg.inv

    def predict(self):
        F = self.F
        self.x = np.dot(F, self.x)
        self.P = self._alpha_sq * np.dot(np.dot(F, self.P), F.T) + self.Q
        self.x_prior = self.x.copy()
        self.P_prior = self.P.copy()

    def update(self, z):
        z = np.asarray(z).reshape(self.dim_z, 1)
        H, R = self.H, self.R
        self.y = z - np.dot(H, self.x)
        PHT = np.dot(self.P, H.T)
        self.S = np.dot(H, PHT) + R
        self.SI = self.inv(self.S)
        self.K = np.dot(PHT, self.SI)
        self.x = self.x + np.dot(self.K, self.y)
        I_KH = self._I - np.dot(self.K, H)
        self.P = np.dot(np.dot(I_KH, self.P), I_KH.T) + np.dot(np.dot(self.K, R), self.K.T)
        self.z = z.copy()
        self.x_post = self.x.copy()
        self.P_post = self.P.copy()

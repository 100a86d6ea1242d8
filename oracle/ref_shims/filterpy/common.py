"""Stand-in for filterpy.common (1.4.5): only reshape_z is executed by oc_sort/kalmanfilter.py:493."""
import numpy as np


def reshape_z(z, dim_z, ndim):
    z = np.atleast_2d(z)
    if z.shape[1] == dim_z:
        z = z.T
    if z.shape != (dim_z, 1):
        raise ValueError("z (shape {}) must be convertible to shape ({}, 1)".format(z.shape, dim_z))
    if ndim == 1:
        z = z[:, 0]
    if ndim == 0:
        z = z[0, 0]
    return z


def pretty_str(label, arr):
    return "{} = {}".format(label, arr)

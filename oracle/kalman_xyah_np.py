"""8-d (x, y, a, h, vx, vy, va, vh) constant-velocity Kalman filter oracles (test infrastructure).

Three noise models exist in the reference; they differ only in how the standard deviations are
built, so one restatement takes the model as a parameter:

  * ``"bytetrack"``  — /root/reference/plugins/track/byte_track/kalman_filter.py:23-269
        position/velocity std = w * h, aspect std constant (1e-2 / 1e-5; R: 1e-1)
  * ``"strongsort"`` — /root/reference/plugins/track/strong_sort/sort/kalman_filter.py:21-214
        std scaled by x, y, a, h individually; R std scaled by (1 - confidence)
  * ``"bpbreid"``    — /root/reference/plugins/track/bpbreid_strong_sort/sort/kalman_filter.py:21-227
        all std (incl. aspect) scaled by h

The NumPy calls (np.dot / multi_dot / cho_factor / cho_solve / solve_triangular) are the same as
the reference's so dtype promotion and BLAS rounding are reproduced bit for bit.
"""
import numpy as np
import scipy.linalg

W_POS = 1.0 / 20
W_VEL = 1.0 / 160

_F = np.eye(8, 8)
for _i in range(4):
    _F[_i, 4 + _i] = 1.0
_H = np.eye(4, 8)

CHI2INV95_4 = 9.4877  # strong_sort/sort/kalman_filter.py:9-18


def bt_initiate(z):
    """byte_track/kalman_filter.py:55-86 — z may be float32 (STrack._tlwh is float32)."""
    mean = np.r_[z, np.zeros_like(z)]
    std = [2 * W_POS * z[3], 2 * W_POS * z[3], 1e-2, 2 * W_POS * z[3],
           10 * W_VEL * z[3], 10 * W_VEL * z[3], 1e-5, 10 * W_VEL * z[3]]
    return mean, np.diag(np.square(std))


def bt_multi_predict(mean, cov):
    """byte_track/kalman_filter.py:155-192 (vectorised over tracks; mean may be a float32 array)."""
    h = mean[:, 3]
    std_pos = [W_POS * h, W_POS * h, 1e-2 * np.ones_like(h), W_POS * h]
    std_vel = [W_VEL * h, W_VEL * h, 1e-5 * np.ones_like(h), W_VEL * h]
    sqr = np.square(np.r_[std_pos, std_vel]).T
    q = np.asarray([np.diag(sqr[i]) for i in range(len(mean))])
    mean = np.dot(mean, _F.T)
    left = np.dot(_F, cov).transpose((1, 0, 2))
    cov = np.dot(left, _F.T) + q
    return mean, cov


def bt_project(mean, cov):
    """byte_track/kalman_filter.py:126-153."""
    std = [W_POS * mean[3], W_POS * mean[3], 1e-1, W_POS * mean[3]]
    r = np.diag(np.square(std))
    return np.dot(_H, mean), np.linalg.multi_dot((_H, cov, _H.T)) + r


def _correct(mean, cov, z, pm, pc):
    """Cholesky-based gain + correction shared by the three models (e.g. byte_track/kalman_filter.py:212-226)."""
    chol, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
    gain = scipy.linalg.cho_solve((chol, lower), np.dot(cov, _H.T).T, check_finite=False).T
    innovation = z - pm
    new_mean = mean + np.dot(innovation, gain.T)
    new_cov = cov - np.linalg.multi_dot((gain, pc, gain.T))
    return new_mean, new_cov


def bt_update(mean, cov, z):
    """byte_track/kalman_filter.py:194-226."""
    pm, pc = bt_project(mean, cov)
    return _correct(mean, cov, z, pm, pc)


def maha_sq(pm, pc, zs):
    """Squared Mahalanobis distance of measurements ``zs[M,4]`` (e.g. strong_sort/sort/kalman_filter.py:202-214)."""
    d = zs - pm
    chol = np.linalg.cholesky(pc)
    z = scipy.linalg.solve_triangular(chol, d.T, lower=True, check_finite=False, overwrite_b=True)
    return np.sum(z * z, axis=0)

"""GPU parity of the stateless cost-matrix kernels and the batched assignment solver vs the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _boxes(rng, B, N, W=1920, H=1080):
    xy = rng.uniform(0, [W - 200, H - 300], size=(B, N, 2))
    wh = rng.uniform([20, 40], [200, 300], size=(B, N, 2))
    return np.concatenate([xy, xy + wh], axis=2)


@pytest.mark.parametrize("variant", ["iou", "giou", "diou", "ciou", "ct_dist"])
@pytest.mark.parametrize("shape", [(1, 20, 20), (3, 40, 37), (2, 150, 150), (1, 1, 9)])
def test_iou_family_matches_oracle(variant, shape):
    from oracle.ocsort_np import ASSO
    from tracklab_b200 import kernels
    B, N, M = shape
    rng = np.random.default_rng(N * 7 + M)
    a, b = _boxes(rng, B, N), _boxes(rng, B, M)
    b[:, : min(N, M)] = a[:, : min(N, M)] + rng.normal(0, 3, size=(B, min(N, M), 4))   # overlapping pairs
    out = kernels.iou_matrix(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), variant).cpu().numpy()
    for p in range(B):
        ref = ASSO[variant](a[p], b[p])
        if variant in ("iou", "giou", "diou"):
            assert np.array_equal(out[p], ref) or np.abs(out[p] - ref).max() < 1e-15
        else:
            assert np.abs(out[p] - ref).max() < 1e-12   # atan differs by ulps


@pytest.mark.parametrize("shape", [(1, 40, 40), (4, 150, 150), (2, 7, 64)])
def test_iou_plus1_f32_bit_exact(shape):
    from oracle.boxes_np import iou_plus1_f32
    from tracklab_b200 import kernels
    B, N, M = shape
    rng = np.random.default_rng(N + M)
    a, b = _boxes(rng, B, N).astype(np.float32), _boxes(rng, B, M).astype(np.float32)
    b[:, : min(N, M)] = a[:, : min(N, M)] + rng.normal(0, 3, size=(B, min(N, M), 4)).astype(np.float32)
    out = kernels.iou_p1_dist(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    for p in range(B):
        assert np.array_equal(out[p], 1 - iou_plus1_f32(a[p], b[p]))


@pytest.mark.parametrize("shape", [(1, 40, 40, 512), (2, 150, 150, 256), (1, 40, 37, 2048), (3, 5, 9, 100)])
def test_cosine_distance_within_1e4(shape):
    """fp32 distances within 1e-4 of the reference's float32 NumPy formula (nn_matching.py:30-49)."""
    from tracklab_b200 import kernels
    B, N, M, E = shape
    rng = np.random.default_rng(E)
    a = rng.normal(0, 1, size=(B, N, E)).astype(np.float32)
    b = (a[:, :1].repeat(M, axis=1) * 0 + rng.normal(0, 1, size=(B, M, E))).astype(np.float32)
    b[:, : min(N, M)] = a[:, : min(N, M)] + 0.15 * rng.normal(0, 1, size=(B, min(N, M), E)).astype(np.float32)
    out = kernels.cosine_dist(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    for p in range(B):
        an = a[p] / np.linalg.norm(a[p], axis=1, keepdims=True)
        bn = b[p] / np.linalg.norm(b[p], axis=1, keepdims=True)
        ref = 1.0 - np.dot(an, bn.T)
        assert np.abs(out[p] - ref).max() < 1e-4


@pytest.mark.parametrize("shape,limit", [((8, 40, 40), 0.8), ((4, 150, 150), 0.8), ((6, 37, 52), 0.5), ((6, 52, 37), None),
                                         ((3, 150, 120), None), ((2, 1, 30), 0.7), ((64, 20, 20), 0.8)])
def test_lap_batched_matches_oracle(shape, limit):
    from oracle.assign_np import lapjv_extended
    from tracklab_b200 import kernels
    B, N, M = shape
    rng = np.random.default_rng(N * 3 + M)
    cost = rng.uniform(0.0, 1.0, size=(B, N, M))
    k = min(N, M)
    for p in range(B):   # plant a strong diagonal so many rows are uncontested, plus contested neighbours
        idx = rng.permutation(k)
        cost[p, np.arange(k), idx] = rng.uniform(0.0, 0.3, size=k)
    x, y, st = kernels.lap_batched(torch.from_numpy(cost).cuda(), limit)
    assert int(st.item()) == 0
    x, y = x.cpu().numpy(), y.cpu().numpy()
    for p in range(B):
        rx, ry = lapjv_extended(cost[p], np.inf if limit is None else limit)
        assert np.array_equal(x[p], rx) and np.array_equal(y[p], ry)


@pytest.mark.parametrize("shape", [(1, 40, 37, 6, 512), (2, 9, 150, 5, 64), (1, 1, 1, 3, 20)])
def test_part_distance_matches_oracle(shape):
    """Stateless part-based distance (BPBReID / KPR embeddings) vs oracle/bpbreid_np.part_distance, float32 within 1e-5."""
    from oracle.bpbreid_np import part_distance
    from tracklab_b200 import kernels
    B, N, M, K, E = shape
    rng = np.random.default_rng(11)
    a = rng.normal(size=(B, N, K, E)).astype(np.float32); b = rng.normal(size=(B, M, K, E)).astype(np.float32)
    va = (rng.uniform(size=(B, N, K)) < 0.8).astype(np.float32); vb = (rng.uniform(size=(B, M, K)) < 0.8).astype(np.float32)
    va[..., 0] = 1.0; vb[..., 0] = 1.0
    got = kernels.part_dist(*(torch.from_numpy(x).cuda() for x in (a, va, b, vb))).cpu().numpy()
    for p in range(B):
        for i in range(N):
            ref = part_distance(a[p, i], va[p, i], b[p], vb[p])
            assert np.abs(got[p, i] - ref).max() < 1e-5


@pytest.mark.parametrize("aspect_const", [True, False])
def test_kf_gate_matches_oracle(aspect_const):
    from oracle import bpbreid_np, strongsort_np
    from tracklab_b200 import kernels
    rng = np.random.default_rng(12)
    T, D = 37, 50
    mean = np.zeros((T, 8)); mean[:, 0] = rng.uniform(100, 1800, T); mean[:, 1] = rng.uniform(100, 900, T)
    mean[:, 2] = rng.uniform(0.3, 0.6, T); mean[:, 3] = rng.uniform(80, 300, T); mean[:, 4:] = rng.normal(0, 2, (T, 4))
    A = rng.normal(size=(T, 8, 8)); cov = A @ A.transpose(0, 2, 1) + 5.0 * np.eye(8)
    z = np.column_stack([rng.uniform(100, 1800, D), rng.uniform(100, 900, D), rng.uniform(0.3, 0.6, D), rng.uniform(80, 300, D)])
    got, st = kernels.kf_gate(torch.from_numpy(mean).cuda(), torch.from_numpy(cov).cuda(), torch.from_numpy(z).cuda(), aspect_const)
    got = got.cpu().numpy()
    gate = strongsort_np.kf_gating if aspect_const else bpbreid_np.kf_gating
    ref = np.stack([gate(mean[t], cov[t], z) for t in range(T)])
    assert int(st.item()) == 0 and np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()


def test_lsap_scipy_batched_reproduces_scipy_including_ties():
    """tk_lsap_scipy_batched (csrc/lsap_scipy.cuh, the solver of the StrongSORT / BPBReID whole-video kernels) against
    scipy.optimize.linear_sum_assignment itself on tie-heavy matrices, tall and wide: identical pairs, not just equal cost."""
    from scipy.optimize import linear_sum_assignment
    from tracklab_b200 import kernels
    rng = np.random.default_rng(1)
    for (N, M) in [(1, 1), (5, 9), (9, 5), (40, 44), (44, 40), (37, 300), (300, 37), (128, 128)]:
        B = 24
        cs = []
        for b in range(B):
            kind = b % 4
            if kind == 0:
                c = rng.integers(0, 3, size=(N, M)).astype(float)
            elif kind == 1:
                c = rng.uniform(0, 1, size=(N, M)); c[c > 0.3] = 0.3 + 1e-5
            elif kind == 2:
                c = rng.uniform(0, 1, size=(N, M)); c[rng.uniform(size=(N, M)) < 0.8] = 0.8 + 1e-5
            else:
                c = np.full((N, M), 0.5)
            cs.append(c)
        cost = np.stack(cs)
        x, y, st = kernels.lsap_scipy_batched(torch.from_numpy(cost).cuda())
        assert int(st.item()) == 0
        x, y = x.cpu().numpy(), y.cpu().numpy()
        for b in range(B):
            rows, cols = linear_sum_assignment(cost[b])
            want_x = -np.ones(N, dtype=np.int64); want_x[rows] = cols
            want_y = -np.ones(M, dtype=np.int64); want_y[cols] = rows
            assert np.array_equal(x[b], want_x) and np.array_equal(y[b], want_y), (N, M, b)

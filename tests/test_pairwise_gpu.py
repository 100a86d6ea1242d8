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


@pytest.mark.parametrize("model", ["bytetrack", "botsort"])
def test_stateless_kf_predict_update_match_oracle(model):
    """tk_kf_predict / tk_kf_update vs the NumPy restatements of the two filters (oracle/kalman_xyah_np.py, oracle/botsort_np.py):
    predict is sums of two entries with one rounding each (bit-equal); update goes through a 4x4 Cholesky (<= 1e-9 relative)."""
    import oracle.botsort_np as bo
    import oracle.kalman_xyah_np as kx
    from tracklab_b200 import kernels
    rng = np.random.default_rng(5)
    n = 37
    if model == "bytetrack":
        z0 = np.column_stack([rng.uniform(100, 1800, n), rng.uniform(100, 1000, n), rng.uniform(0.3, 0.6, n), rng.uniform(80, 400, n)])
        init, predict, update = kx.bt_initiate, kx.bt_multi_predict, kx.bt_update
    else:
        z0 = np.column_stack([rng.uniform(100, 1800, n), rng.uniform(100, 1000, n), rng.uniform(40, 200, n), rng.uniform(80, 400, n)])
        init, predict, update = bo.kf_initiate, bo.kf_multi_predict, bo.kf_update
    mc = [init(z) for z in z0]
    mean = np.stack([m for m, _ in mc]).astype(np.float64)
    cov = np.stack([c for _, c in mc]).astype(np.float64)
    mean[:, 4:] = rng.normal(0, 2, (n, 4))
    dm, dc = torch.from_numpy(mean.copy()).cuda(), torch.from_numpy(cov.copy()).cuda()
    for step in range(3):
        mean, cov = predict(mean, cov)
        kernels.kf_predict(dm, dc, model)
        assert np.array_equal(dm.cpu().numpy(), mean) and np.array_equal(dc.cpu().numpy(), cov)
        z = mean[:, :4] + rng.normal(0, 1.0, (n, 4)) * np.array([3.0, 3.0, 0.01 if model == "bytetrack" else 2.0, 3.0])
        upd = [update(mean[i], cov[i], z[i]) for i in range(n)]
        mean, cov = np.stack([u[0] for u in upd]), np.stack([u[1] for u in upd])
        _, _, st = kernels.kf_update(dm, dc, torch.from_numpy(z).cuda(), model)
        assert int(st.item()) == 0
        assert np.allclose(dm.cpu().numpy(), mean, rtol=1e-9, atol=1e-9) and np.allclose(dc.cpu().numpy(), cov, rtol=1e-9, atol=1e-9)
        dm.copy_(torch.from_numpy(mean)); dc.copy_(torch.from_numpy(cov))       # keep both sides on the same trajectory


def test_vdc_cost_matches_reference_formula():
    """tk_vdc_cost vs the angle term of oc_sort/association.py:246-266 as restated inside oracle/ocsort_np.associate."""
    from tracklab_b200 import kernels
    rng = np.random.default_rng(9)
    D, T = 23, 31
    dets = np.column_stack([rng.uniform(0, 1800, (D, 2)), rng.uniform(0, 1, D), rng.uniform(0, 3, D)])
    dets = np.column_stack([dets[:, 0], dets[:, 1], dets[:, 0] + rng.uniform(30, 200, D), dets[:, 1] + rng.uniform(60, 400, D), dets[:, 2], dets[:, 3]])
    prev = np.column_stack([rng.uniform(0, 1800, (T, 2)), np.zeros((T, 2)), rng.uniform(0.2, 1, T)])
    prev[:, 2] = prev[:, 0] + 80; prev[:, 3] = prev[:, 1] + 200
    prev[::5] = -1.0                                                   # placeholder observations
    vel = rng.normal(0, 1, (T, 2)); vel /= np.linalg.norm(vel, axis=1, keepdims=True)
    inertia = 0.3941737016672115
    pt = prev[..., np.newaxis]
    cx1, cy1 = (dets[:, 0] + dets[:, 2]) / 2.0, (dets[:, 1] + dets[:, 3]) / 2.0
    cx2, cy2 = (pt[:, 0] + pt[:, 2]) / 2.0, (pt[:, 1] + pt[:, 3]) / 2.0
    dx, dy = cx1 - cx2, cy1 - cy2
    norm = np.sqrt(dx ** 2 + dy ** 2) + 1e-6
    X, Y = dx / norm, dy / norm
    iy = np.repeat(vel[:, 0][:, np.newaxis], Y.shape[1], axis=1)
    ix = np.repeat(vel[:, 1][:, np.newaxis], X.shape[1], axis=1)
    ang = np.arccos(np.clip(ix * X + iy * Y, a_min=-1, a_max=1))
    ang = (np.pi / 2.0 - np.abs(ang)) / np.pi
    valid = np.ones(T); valid[np.where(prev[:, 4] < 0)] = 0
    valid = np.repeat(valid[:, np.newaxis], X.shape[1], axis=1)
    scores = np.repeat(dets[:, -1][:, np.newaxis], T, axis=1)
    ref = ((valid * ang) * inertia).T * scores
    out = kernels.vdc_cost(torch.from_numpy(dets).cuda(), torch.from_numpy(prev).cuda(), torch.from_numpy(vel).cuda(), inertia, 5).cpu().numpy()
    assert np.abs(out - ref).max() < 1e-12

"""Stand-in for the two torchreid (BPBReID fork, un-vendored git dependency — README.md:131) functions imported by
/root/reference/plugins/track/bpbreid_strong_sort/sort/nn_matching.py:4-8. PARITY UNPINNED: restated from the published
BPBReID code as documented in SURVEY.md §8c [3P-memory]: per-part Euclidean distance between (already L2-normalised) part
embeddings, combined as the visibility-weighted mean  sum_k vq_k*vg_k*d_k / sum_k vq_k*vg_k ; returns
(pairwise_dist [Q,G], body_part_pairwise_dist [K,Q,G]) like the original."""
import torch


def compute_distance_matrix(a, b, metric="euclidean"):
    if metric == "cosine":
        a = torch.nn.functional.normalize(a, dim=1)
        b = torch.nn.functional.normalize(b, dim=1)
        return 1 - a @ b.t()
    return torch.cdist(a, b, compute_mode="donot_use_mm_for_euclid_dist")


def compute_distance_matrix_using_bp_features(qf, gf, qf_parts_visibility=None, gf_parts_visibility=None,
                                              dist_combine_strat="mean", batch_size_pairwise_dist_matrix=5000,
                                              use_gpu=False, metric="euclidean"):
    q = qf.transpose(1, 0)   # [K, Q, E]
    g = gf.transpose(1, 0)   # [K, G, E]
    part = torch.cdist(q, g, compute_mode="donot_use_mm_for_euclid_dist")   # [K, Q, G]
    if qf_parts_visibility is None or gf_parts_visibility is None:
        return part.mean(0), part
    vq = qf_parts_visibility.t().to(part.dtype)   # [K, Q]
    vg = gf_parts_visibility.t().to(part.dtype)   # [K, G]
    w = vq.unsqueeze(2) * vg.unsqueeze(1)         # [K, Q, G]
    pair = (part * w).sum(0) / w.sum(0)
    return pair, part

from .distance import compute_distance_matrix, compute_distance_matrix_using_bp_features  # noqa: F401

"""PyTorch definitions of the detector / ReID backbones (tensor-core work stays in PyTorch/cuDNN, per
BASELINE.json north_star). Pure module definitions: device-agnostic, seeded random weights."""

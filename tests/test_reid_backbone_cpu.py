"""CPU (build container only): the ResNet-50 ReID restatement equals the reference's vendored definition."""
import os
import sys

import pytest
import torch

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_resnet50_reid_equals_vendored_reference():
    sys.path.insert(0, os.path.join(REF, "plugins", "track"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_resnet", os.path.join(REF, "plugins/track/strong_sort/deep/models/resnet.py"))
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    torch.manual_seed(0)
    ref = ref_mod.resnet50(num_classes=10, pretrained=False).eval()
    # give BatchNorm non-trivial statistics so the folding is really exercised
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
    from tracklab_b200.nets.resnet_reid import ResNet50ReID
    mine = ResNet50ReID().eval().from_reference_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 256, 128)
    with torch.no_grad():
        a, b = ref(x), mine(x)
    assert a.shape == b.shape == (2, 2048)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-4 * a.abs().max().item())


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("ibn", [False, True])
def test_osnet_reid_equals_vendored_reference(ibn):
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_osnet", os.path.join(REF, "plugins/track/strong_sort/deep/models/osnet.py"))
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    torch.manual_seed(1)
    ref = (ref_mod.osnet_ibn_x1_0 if ibn else ref_mod.osnet_x1_0)(num_classes=10, pretrained=False).eval()
    for m in ref.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
        elif isinstance(m, torch.nn.InstanceNorm2d):
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
        elif isinstance(m, torch.nn.Linear):
            m.weight.data.normal_(0, 0.05); m.bias.data.normal_(0, 0.1)
    from tracklab_b200.nets.osnet_reid import OSNetReID
    mine = OSNetReID(ibn=ibn).eval().from_reference_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 256, 128)
    with torch.no_grad():
        a, b = ref(x), mine(x)
    assert a.shape == b.shape == (2, 512)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-4 * a.abs().max().item())


def test_space_to_depth_stem_weights_reproduce_the_7x7_stride2_convolution():
    """Host logic of the fused ReID executor: the 4x4 / stride-1 weights on the padded 2x2 space-to-depth crop (the layout
    tk_crop_resize_norm writes, TK_CROP_LAYOUT_S2D16) give the 7x7 / stride-2 / pad-3 stem output exactly (fp64)."""
    import torch.nn.functional as F
    from tracklab_b200.nets.resnet_fused import stem_weight_s2d16
    torch.manual_seed(3)
    x = torch.randn(2, 3, 32, 16, dtype=torch.float64)
    w = torch.randn(8, 3, 7, 7, dtype=torch.float64)
    ref = F.conv2d(x, w, None, 2, 3)
    N, _, H, W = x.shape
    buf = torch.zeros(N, H // 2 + 3, W // 2 + 3, 16, dtype=torch.float64)                       # NHWC, zero border 2 before / 1 after
    for yy in range(H):
        for xx in range(W):
            c0 = ((yy & 1) * 2 + (xx & 1)) * 3
            buf[:, (yy >> 1) + 2, (xx >> 1) + 2, c0:c0 + 3] = x[:, :, yy, xx]
    got = F.conv2d(buf.permute(0, 3, 1, 2), stem_weight_s2d16(w), None, 1, 0)
    assert got.shape == ref.shape and torch.allclose(got, ref, rtol=0, atol=1e-12)


def test_streaming_drain_schedule_and_crop_buckets():
    from tracklab_b200.nets.resnet_fused import ResNet50Fused
    from tracklab_b200.video_pipeline import DetectTrackPipeline

    class _Det:
        tail_sizes = set()

    pipe = DetectTrackPipeline.__new__(DetectTrackPipeline)
    pipe.batch, pipe.det = 50, _Det()
    sched = pipe._schedule(500, True)
    assert sched[:9] == [(i, i + 50) for i in range(0, 450, 50)] and sched[9:] == [(450, 475), (475, 488), (488, 500)]
    assert sorted(_Det.tail_sizes) == [12, 13, 25]
    assert pipe._schedule(500, False) == [(i, i + 50) for i in range(0, 500, 50)]                 # HBM-resident frames: no drain
    assert pipe._schedule(480, True)[-1] == (450, 480)                                             # ragged tail: left as it is
    covered = [f for a, b in sched for f in range(a, b)]
    assert covered == list(range(500))
    assert [ResNet50Fused.bucket(n) for n in (1, 64, 65, 572, 760)] == [64, 64, 128, 576, 768]

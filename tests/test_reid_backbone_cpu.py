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

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100) device; run with -m gpu on a B200")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")


def pytest_collection_modifyitems(config, items):
    import torch

    have_gpu = torch.cuda.is_available()
    ngpu = torch.cuda.device_count() if have_gpu else 0
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "multigpu" in item.keywords and ngpu < 2:
            item.add_marker(pytest.mark.skip(reason="needs >= 2 GPUs"))


@pytest.fixture(scope="session")
def synth_root(tmp_path_factory):
    from distributed_vgg_f_b200.data.synthetic import make_synthetic_imagefolder

    root = tmp_path_factory.mktemp("synth")
    make_synthetic_imagefolder(str(root), train_per_class=8, val_per_class=4, size=128, seed=1)
    return str(root)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_fragments.npz"))


@pytest.fixture(autouse=True)
def _poison_recycled_device_memory(request):
    """GPU tests run on a DIRTY caching-allocator pool: before each one a block of 0xFF bytes is allocated and handed
    back, so scratch buffers are carved out of poisoned memory instead of freshly mapped (zero) pages.  A kernel that
    reads a slot nobody wrote then sees 0xFFFFFFFF ids / NaNs and fails loudly instead of passing by luck (this is how
    the speculative-capacity overflow bug of DESIGN.md section 5 surfaced).  GSR_TEST_POISON_MB=0 turns it off."""
    mb = int(os.environ.get("GSR_TEST_POISON_MB", "512"))
    if mb > 0 and request.node.get_closest_marker("gpu") is not None:
        import torch
        if torch.cuda.is_available():
            for dev in range(torch.cuda.device_count()):
                junk = torch.full((mb << 18,), -1, dtype=torch.int32, device=f"cuda:{dev}")
                del junk
    yield

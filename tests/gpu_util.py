import pytest
import torch


def need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device")
    cc = torch.cuda.get_device_capability(0)
    if cc[0] != 10:
        pytest.fail(f"sm_100 device required, got sm_{cc[0]}{cc[1]}")

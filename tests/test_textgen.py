"""The synthetic-text generator (tests/support): the numpy restatement used by bench.py's reference arm == the C loop of
libacgen.so, and (GPU) == the device kernel, for shards at arbitrary offsets."""
import numpy as np
import pytest

from helpers import random_patterns, textgen

CASES = [(100000, 0, 0, 4096), (70000, 12345, 1, 512), (5000, 999999, 0, 64), (33, 7, 1, 0), (200000, 4096 * 3 - 5, 0, 4096), (50000, 0, 0, 16)]


@pytest.mark.parametrize("nb,first,kind,period", CASES)
def test_numpy_generator_equals_c_loop(nb, first, kind, period):
    flat, offsets = random_patterns(500)
    a = textgen.host_text(nb, first, kind, plant_period=period, dict_flat=flat, dict_offsets=offsets)
    b = textgen.host_text_c(nb, first, kind, plant_period=period, dict_flat=flat, dict_offsets=offsets)
    assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("nb,first,kind,period", CASES)
def test_device_generator_equals_host(nb, first, kind, period):
    import torch

    flat, offsets = random_patterns(500)
    d = torch.empty(nb + 64, dtype=torch.uint8, device="cuda")
    textgen.generate_text(nb, first, kind, plant_period=period, dict_flat=flat, dict_offsets=offsets, device_ptr=d.data_ptr())
    want = textgen.host_text(nb, first, kind, plant_period=period, dict_flat=flat, dict_offsets=offsets)
    assert np.array_equal(d[:nb].cpu().numpy(), want)

"""The factorisation the CUDA kernels implement (tests/packed_model.py: 4 packed real spectra, Hermitian
half, two-for-one transforms) reproduces the reference outputs — CPU, float64 model."""
import numpy as np
import pytest

import packed_model as M
from conftest import assert_maps_close, load_golden


@pytest.mark.parametrize("name", ["n16_default", "n64_default", "n64_wind"])
def test_packed_factorisation_matches_reference(name):
    g, params = load_golden(name)
    h0 = g["h0"]
    for i, t in enumerate(g["t"]):
        a, disp, norm, mn, mx = M.compute_waves(params["tile_size"], params["tile_length"], params["lam"],
                                                h0[..., 0], h0[..., 1], h0[..., 4], float(t))
        assert_maps_close(disp, norm, g["disp"][i], g["norm"][i], f"{name} t={t}")
        assert abs(a - g["A"][i]) <= 1e-6 * g["A"][i]
        assert abs(mn - g["minh"][i]) <= 1e-6 * abs(g["minh"][i])


def test_slot_permutation_is_a_bijection():
    for n in (16, 64, 1024):
        col = M.column_of_slot(n)
        assert sorted(col.tolist()) == list(range(n))
        assert col[0] == 0 and col[n // 2] == n // 2 and col[n // 2 + 1] == n - 1

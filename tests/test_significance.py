"""CPU: nolzss_b200.genomics.significance against vectors produced by the reference's own module
(tests/golden/make_significance_golden.py imports /root/reference/src/noLZSS/genomics/significance.py), plus the
file path (binary factor files -> lengths) and the error behaviour of the reference's API."""
import importlib.util
import json
import os
import warnings

import numpy as np
import pytest

from nolzss_b200.genomics import significance as sig

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "significance_golden.json")))


def _gen():
    spec = importlib.util.spec_from_file_location("make_sig_gold", os.path.join(HERE, "golden", "make_significance_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _hex(xs):
    return [float(v).hex() for v in xs]


def test_clopper_pearson_matches_reference_vectors():
    for c in GOLD["clopper_pearson_upper"]:
        assert float(sig.clopper_pearson_upper(c["k"], c["n"], c["alpha"])).hex() == c["value"], c
    for bad in [(1, 0, 0.05), (-1, 5, 0.05), (6, 5, 0.05), (1, 5, 0.0), (1, 5, 1.0)]:
        with pytest.raises(ValueError):
            sig.clopper_pearson_upper(*bad)


def test_infer_length_significance_matches_reference_vectors():
    gen = _gen()
    for g in GOLD["infer_length_significance"]:
        c = g["case"]
        real, shuf = gen.lengths(*c["real"]), gen.lengths(*c["shuf"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = sig.infer_length_significance(real, shuf, tau_expected_fp=c["tau"], alpha_cp=c["alpha"])
        assert (r["N_real"], r["N_shuf"], r["L_star"]) == (g["N_real"], g["N_shuf"], g["L_star"])
        assert [int(v) for v in r["uniq_L"]] == g["uniq_L"]
        assert _hex(r["S0"]) == g["S0"]
        assert _hex(r["S0_upper"]) == g["S0_upper"]
        assert _hex(r["expected_fp_upper"]) == g["expected_fp_upper"]
        assert _hex(r["rarity_scores_real"][:200]) == g["rarity_scores_real"]
        assert float(np.sum(r["rarity_scores_real"])).hex() == g["rarity_sum"]
        assert [[L, float(r["p_any_ge"](L)).hex()] for L, _ in g["p_any_ge"]] == g["p_any_ge"]
        assert r["tau_expected_fp"] == c["tau"] and r["alpha_cp"] == c["alpha"]


def test_docstring_example_and_errors():
    r = sig.infer_length_significance(np.array([5, 10, 15, 20, 25]), np.array([2, 3, 4, 5, 6, 7, 8, 9, 10]), tau_expected_fp=0.5)
    assert r["L_star"] is None or isinstance(r["L_star"], int)
    assert np.allclose(r["rarity_scores_real"][:3], [6 / 9, 1 / 9, 0.0])   # S0(5) = 6 of 9 shuffled lengths >= 5
    with pytest.raises(ValueError, match="Shuffled genome must have at least one factor"):
        sig.infer_length_significance([1, 2], [])
    with pytest.warns(UserWarning, match="Real genome has no factors"):
        sig.infer_length_significance([], [3, 4, 5])
    with pytest.raises(ValueError, match="alpha must be in"):
        sig.infer_length_significance([1], [1, 2], alpha_cp=1.5)
    assert list(sig.extract_factor_lengths([(0, 5, 0), (5, 3, 2), (8, 10, 1)])) == [5, 3, 10]
    assert sig.extract_factor_lengths([]).dtype == np.int64
    with pytest.raises(ValueError, match="at least 2 elements"):
        sig.extract_factor_lengths([(1,)])
    with pytest.raises(ValueError, match="list of tuples or a file path"):
        sig.extract_factor_lengths(7)


def test_threshold_from_binary_factor_files(tmp_path):
    """configs[4] tail end: two noLZSSv2 files -> calculate_factor_length_threshold; the file path must give exactly
    what the in-memory path gives on the same lengths."""
    import ctypes

    from nolzss_b200 import _lib as L

    gen = _gen()
    real, shuf = gen.lengths("heavy", 30_000, 21), gen.lengths("geometric", 25_000, 22)
    lib = L.load()

    def write(path, lens):
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
        tr = np.stack([starts, lens.astype(np.uint64), starts], axis=1).astype(np.uint64).copy()
        L.check(lib.nlz_write_factor_file(os.fsencode(str(path)), tr.ctypes.data_as(ctypes.c_void_p), len(lens), None, 0, 0, 0,
                                          int(lens.sum())))

    write(tmp_path / "real.bin", real)
    write(tmp_path / "shuf.bin", shuf)
    assert np.array_equal(sig.extract_factor_lengths(tmp_path / "real.bin"), real)
    a = sig.calculate_factor_length_threshold(tmp_path / "real.bin", str(tmp_path / "shuf.bin"), tau_expected_fp=10.0)
    b = sig.infer_length_significance(real, shuf, tau_expected_fp=10.0)
    assert a["L_star"] == b["L_star"] and a["L_star"] is not None
    assert np.array_equal(a["S0_upper"], b["S0_upper"]) and np.array_equal(a["rarity_scores_real"], b["rarity_scores_real"])
    with pytest.raises(FileNotFoundError, match="Real factors file not found"):
        sig.calculate_factor_length_threshold(tmp_path / "missing.bin", tmp_path / "shuf.bin")
    with pytest.raises(FileNotFoundError, match="Shuffled factors file not found"):
        sig.calculate_factor_length_threshold(tmp_path / "real.bin", tmp_path / "missing.bin")


def test_scales_to_many_factors():
    """10^7 factors in seconds (the reference's O(U * z) loops and per-factor Python reads take minutes here)."""
    import time

    rng = np.random.default_rng(3)
    shuf = (8 + rng.geometric(0.3, 10_000_000)).astype(np.int64)
    real = shuf.copy()
    real[rng.integers(0, len(real), 100_000)] = rng.integers(30, 100_000, 100_000)
    t0 = time.perf_counter()
    r = sig.infer_length_significance(real, shuf, tau_expected_fp=100.0)
    assert time.perf_counter() - t0 < 30
    assert r["L_star"] is not None and r["N_real"] == 10_000_000

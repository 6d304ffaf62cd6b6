"""CPU, authoring container only: the reference's OWN pure-Python readers (src/noLZSS/utils.py:106-357, the format
oracle named in SURVEY.md section 8c) must read the files this repo writes, and this repo's vectorised readers must
return what they return.  /root/reference does not exist on the GPU box: skipped there (not a gpu test anyway)."""
import importlib.util
import os
import struct

import numpy as np
import pytest

from nolzss_b200 import _lib as L
from nolzss_b200 import utils

REF_UTILS = "/root/reference/src/noLZSS/utils.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF_UTILS), reason="reference tree not present")


def _ref():
    spec = importlib.util.spec_from_file_location("reference_utils", REF_UTILS)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _write(path, tr, meta=b"", nseq=0, nsent=0, total=0):
    tr = np.ascontiguousarray(tr, dtype=np.uint64)
    L.check(L.load().nlz_write_factor_file(os.fsencode(str(path)), tr.ctypes.data, len(tr), meta if meta else None, len(meta),
                                          nseq, nsent, total))


def _random_factors(rng, z):
    lens = rng.integers(1, 50, z).astype(np.uint64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    refs = (rng.integers(0, 1 << 40, z).astype(np.uint64)) | (rng.integers(0, 2, z).astype(np.uint64) << np.uint64(63))
    return np.stack([starts, lens, refs], axis=1)


def test_reference_reader_reads_our_files(tmp_path):
    ref = _ref()
    rng = np.random.default_rng(4)
    shapes = [
        dict(z=0, meta=b"", nseq=0, nsent=0),                                                # empty factor list (footer only)
        dict(z=1, meta=b"", nseq=0, nsent=0),                                                # V1 / V3 / V6: no metadata
        dict(z=500, meta=b"\x00", nseq=1, nsent=0),                                          # V2: a lone NUL name
        dict(z=500, meta=b"", nseq=2, nsent=1),                                              # V4 / V5: counts without metadata
        dict(z=2000, meta=b"chr1\x00chr2 \x00c\x00" + struct.pack("<QQ", 7, 1200), nseq=3, nsent=2),   # V7
        dict(z=64, meta=b"rec000017\x00", nseq=1, nsent=0),                                  # V8
    ]
    for k, sh in enumerate(shapes):
        tr = _random_factors(rng, sh["z"]) if sh["z"] else np.zeros((0, 3), dtype=np.uint64)
        path = tmp_path / f"f{k}.bin"
        _write(path, tr, sh["meta"], sh["nseq"], sh["nsent"], int(tr[:, 1].sum()) if sh["z"] else 0)
        want = [(int(a), int(b), int(c)) for a, b, c in tr]
        assert ref.read_factors_binary_file(path) == want, sh
        assert utils.read_factors_binary_file(path) == want, sh
        arr = utils.read_factors_array(path)
        assert arr.shape == (sh["z"], 3) and np.array_equal(arr, tr)
        consistent = len(sh["meta"]) == 0 and sh["nseq"] == 0 or len(sh["meta"]) > 0
        if consistent:
            # (V4 / V5 claim 2 sequences + 1 sentinel but write no metadata -- SURVEY App. B.10; the reference's own
            # metadata reader cannot parse those files either)
            assert ref.read_binary_file_metadata(path) == utils.read_binary_file_metadata(path), sh
            a, b = ref.read_factors_binary_file_with_metadata(path), utils.read_factors_binary_file_with_metadata(path)
            assert a == b, sh


def test_input_helpers_match_reference():
    ref = _ref()
    for data in [b"abracadabra", "ACGTACGT", b"\x01\x02\x03" * 5, "x"]:
        assert ref.validate_input(data) == utils.validate_input(data)
        assert ref.analyze_alphabet(data) == utils.analyze_alphabet(data)
    for bad in ["", b"", "héllo", b"a\x00b", 5]:
        e_ref = e_us = None
        try:
            ref.validate_input(bad)
        except Exception as e:  # noqa: BLE001
            e_ref = type(e).__name__
        try:
            utils.validate_input(bad)
        except Exception as e:  # noqa: BLE001
            e_us = type(e).__name__
        assert e_ref == e_us, (bad, e_ref, e_us)

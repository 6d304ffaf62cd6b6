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


def _ref_genomics():
    """The reference's pure-Python genomics modules (sequences.py, fasta.py) under a throw-away package name; their
    package-relative imports are satisfied by the reference's utils.py and a stub for core.factorize / the compiled
    extension (neither is needed by the functions compared here)."""
    import sys
    import types

    base = "/root/reference/src/noLZSS"
    pkg = types.ModuleType("_refpkg"); pkg.__path__ = [base]
    gen = types.ModuleType("_refpkg.genomics"); gen.__path__ = [base + "/genomics"]
    core = types.ModuleType("_refpkg.core")
    core.factorize = lambda data: (_ for _ in ()).throw(RuntimeError("stub"))
    sys.modules.update({"_refpkg": pkg, "_refpkg.genomics": gen, "_refpkg.core": core})

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    load("_refpkg.utils", base + "/utils.py")
    seq = load("_refpkg.genomics.sequences", base + "/genomics/sequences.py")
    fa = load("_refpkg.genomics.fasta", base + "/genomics/fasta.py")
    return seq, fa


def test_sequence_helpers_and_fasta_readers_match_reference(tmp_path):
    from nolzss_b200.genomics import fasta as our_fa
    from nolzss_b200.genomics import sequences as our_seq

    ref_seq, ref_fa = _ref_genomics()
    samples = ["ACGT", "acgtn", "ACGU", "MKVLAAGIV", "hello world", "", "ACGT\n", b"ACGT", b"\xff\xfe", b"MKV", "12345", "ACGTRYKM",
               "A" * 50 + "X", b"acgt"]
    for s in samples:
        assert ref_seq.is_dna_sequence(s) == our_seq.is_dna_sequence(s), s
        assert ref_seq.is_protein_sequence(s) == our_seq.is_protein_sequence(s), s
        assert ref_seq.detect_sequence_type(s) == our_seq.detect_sequence_type(s), s
    contents = [
        ">a desc\nACGT\nAC GT\n\n>b\nMKVLA\n>c|x y\n\nacgtNN\n",
        "no header line\nACGT\n>late\nGG\n",
        ">only_header\n",
        "",
        ">dup\nAC\n>dup\nGT\n",
        ">x\r\nACGT\r\n>y\r\nTT\r\n",
    ]
    for c in contents:
        try:
            want = ("ok", ref_fa._parse_fasta_content(c))
        except Exception as e:  # noqa: BLE001
            want = ("err", type(e).__name__)
        try:
            got = ("ok", our_fa._parse_fasta_content(c))
        except Exception as e:  # noqa: BLE001
            got = ("err", type(e).__name__)
        assert want == got, c
    prot = tmp_path / "p.fasta"
    prot.write_text(">p1 first\nMKVLAAGIVGLLLAQW\nPEFF\n>p2\nmkvw\n")
    assert ref_fa.read_protein_fasta(prot) == our_fa.read_protein_fasta(prot)
    assert ref_fa.read_fasta_auto(prot) == our_fa.read_fasta_auto(prot)
    bad = tmp_path / "bad.fasta"
    bad.write_text(">p1\nMKV1LA\n")
    for fn in ("read_protein_fasta",):
        with pytest.raises(Exception) as e1:
            getattr(ref_fa, fn)(bad)
        with pytest.raises(Exception) as e2:
            getattr(our_fa, fn)(bad)
        assert type(e1.value).__name__ == type(e2.value).__name__
    for fn in ("read_protein_fasta", "read_nucleotide_fasta", "read_fasta_auto"):
        with pytest.raises(FileNotFoundError):
            getattr(ref_fa, fn)(tmp_path / "missing.fasta")
        with pytest.raises(FileNotFoundError):
            getattr(our_fa, fn)(tmp_path / "missing.fasta")


class _FakeExt:
    """Recording stand-in for the compiled extension: every function records its call and returns a canned value."""

    def __init__(self, tmp):
        self.calls = []
        self.tmp = tmp

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def fn(*args, **kwargs):
            self.calls.append((name, tuple(a if not isinstance(a, str) else a.replace(str(self.tmp), "<tmp>") for a in args),
                               tuple(sorted(kwargs.items()))))
            if name.endswith("_to_file") or name.startswith("count") or name.startswith("write"):
                # the *_to_file functions are expected to leave a factor file behind (parallel_factorize reads it back)
                for a in args:
                    if isinstance(a, str) and a.endswith(".bin"):
                        tr = np.array([[0, 1, 0], [1, 2, 0]], dtype=np.uint64)
                        _write(a, tr, b"", 0, 0, 3)
                return 2
            return [(0, 1, 0), (1, 2, 0)]

        return fn


def _load_reference_module(name, fake):
    """A reference module (core.py / parallel.py) imported under a throw-away package whose _noLZSS is `fake`."""
    import sys
    import types

    base = "/root/reference/src/noLZSS"
    pkgname = f"_refpkg_{name}"
    pkg = types.ModuleType(pkgname); pkg.__path__ = [base]
    ext = types.ModuleType(pkgname + "._noLZSS")
    ext.__getattr__ = lambda attr: getattr(fake, attr)                       # PEP 562: `from ._noLZSS import x`
    sys.modules[pkgname] = pkg
    sys.modules[pkgname + "._noLZSS"] = ext
    for sub in ("utils", name):
        spec = importlib.util.spec_from_file_location(f"{pkgname}.{sub}", f"{base}/{sub}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"{pkgname}.{sub}"] = m
        spec.loader.exec_module(m)
    return sys.modules[f"{pkgname}.{name}"]


def _outcome(fn, *args, **kwargs):
    try:
        return ("ok", fn(*args, **kwargs))
    except Exception as e:  # noqa: BLE001
        return ("err", type(e).__name__, str(e))


def test_python_wrappers_drive_the_extension_like_the_reference(tmp_path, monkeypatch):
    """core.py and parallel.py of the reference against this repo's, both on top of the same recording fake
    extension: same results, same exceptions (type and message), same calls into the extension (names, argument
    order and conversions) -- what a drop-in replacement of `_noLZSS` relies on."""
    import nolzss_b200.core as our_core
    import nolzss_b200.parallel as our_par

    existing = tmp_path / "in.txt"
    existing.write_bytes(b"abracadabra")
    out = tmp_path / "sub" / "o.bin"
    cases = {
        "core": [
            ("factorize", ("abracadabra",), {}), ("factorize", (b"abc",), {"validate": False}), ("factorize", ("",), {}),
            ("factorize", (5,), {}), ("factorize", ("a\x00b",), {}), ("factorize_file", (existing,), {}),
            ("factorize_file", (str(existing), 7), {}), ("factorize_file", (tmp_path / "nope",), {}),
            ("count_factors", ("abc",), {}), ("count_factors", (b"",), {}), ("count_factors_file", (existing,), {}),
            ("count_factors_file", (tmp_path / "nope",), {}), ("write_factors_binary_file", ("abc", out), {}),
            ("write_factors_binary_file", ("", out), {}), ("factorize_with_info", ("abracadabra",), {}),
            ("factorize_w_reference", ("ACGT", "ACGA"), {}), ("factorize_w_reference", (b"ACGT", "ACGA"), {"validate": False}),
            ("factorize_w_reference", ("", "ACGA"), {}), ("factorize_w_reference_file", ("ACGT", b"ACGA", out), {}),
        ],
        "parallel": [
            ("parallel_factorize_to_file", ("abc", out), {}), ("parallel_factorize_to_file", (b"abc", str(out), 4, 1), {"validate": False}),
            ("parallel_factorize_to_file", ("", out), {}), ("parallel_factorize_file_to_file", (existing, out), {}),
            ("parallel_factorize_file_to_file", (tmp_path / "nope", out, 2), {}), ("parallel_factorize", ("abracadabra",), {}),
            ("parallel_factorize", ("abracadabra", 3, 2), {}), ("parallel_factorize", ("",), {}),
            ("parallel_factorize_dna_w_rc_to_file", ("ACGT", out, 8), {}), ("parallel_factorize_dna_w_rc_to_file", ("", out), {}),
            ("parallel_factorize_file_dna_w_rc_to_file", (existing, str(out)), {}),
            ("parallel_factorize_file_dna_w_rc_to_file", (tmp_path / "nope", out), {}),
        ],
    }
    for modname, ours in (("core", our_core), ("parallel", our_par)):
        fake_ref, fake_us = _FakeExt(tmp_path), _FakeExt(tmp_path)
        ref_mod = _load_reference_module(modname, fake_ref)
        monkeypatch.setattr(ours, "_ext", fake_us)
        for fname, args, kwargs in cases[modname]:
            want = _outcome(getattr(ref_mod, fname), *args, **kwargs)
            got = _outcome(getattr(ours, fname), *args, **kwargs)
            strip = lambda o: tuple(str(x).replace(str(tmp_path), "<tmp>") if isinstance(x, str) else x for x in o)   # noqa: E731
            assert strip(want) == strip(got), (modname, fname, args, kwargs)

        def norm(calls):        # temp-file names of parallel_factorize differ between the two runs
            return [(n, tuple("<tempfile>" if isinstance(a, str) and a.endswith(".bin") and "<tmp>" not in a else a for a in args), kw)
                    for n, args, kw in calls]

        assert len(fake_ref.calls) >= 7, (modname, fake_ref.calls)
        assert norm(fake_ref.calls) == norm(fake_us.calls), modname


def test_genomics_wrappers_drive_the_extension_like_the_reference(tmp_path, monkeypatch):
    """genomics/fasta.py and genomics/sequences.py of the reference against this repo's on the recording fake
    extension: FASTA readers that factorize, the auto-detecting reader, the reference+target wrappers."""
    import sys
    import types

    import nolzss_b200.core as our_core
    import nolzss_b200.genomics.fasta as our_fa
    import nolzss_b200.genomics.sequences as our_seq

    fake_ref, fake_us = _FakeExt(tmp_path), _FakeExt(tmp_path)
    base = "/root/reference/src/noLZSS"
    pkgname = "_refpkg_gen"
    pkg = types.ModuleType(pkgname); pkg.__path__ = [base]
    gen = types.ModuleType(pkgname + ".genomics"); gen.__path__ = [base + "/genomics"]
    ext = types.ModuleType(pkgname + "._noLZSS")
    ext.__getattr__ = lambda attr: getattr(fake_ref, attr)
    sys.modules.update({pkgname: pkg, pkgname + ".genomics": gen, pkgname + "._noLZSS": ext})
    mods = {}
    for sub, path in (("utils", "utils.py"), ("core", "core.py"), ("genomics.sequences", "genomics/sequences.py"),
                      ("genomics.fasta", "genomics/fasta.py")):
        spec = importlib.util.spec_from_file_location(f"{pkgname}.{sub}", f"{base}/{path}")
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"{pkgname}.{sub}"] = m
        spec.loader.exec_module(m)
        mods[sub] = m
    for mod in (our_core, our_fa, our_seq):
        monkeypatch.setattr(mod, "_ext", fake_us)

    dna = tmp_path / "dna.fasta"
    dna.write_text(">s1 d\nACGTAC\nGT\n>s2\nacgt\n")
    mixed = tmp_path / "mixed.fasta"
    mixed.write_text(">s1\nACGTNNAC\n")
    prot = tmp_path / "prot.fasta"
    prot.write_text(">p\nMKVLW\n")
    text = tmp_path / "text.fasta"
    text.write_text(">t\nhello world 123\n")
    out = tmp_path / "o" / "x.bin"
    calls = [
        ("genomics.fasta", "read_nucleotide_fasta", (dna,), {}), ("genomics.fasta", "read_nucleotide_fasta", (mixed,), {}),
        ("genomics.fasta", "read_nucleotide_fasta", (prot,), {}), ("genomics.fasta", "read_fasta_auto", (dna,), {}),
        ("genomics.fasta", "read_fasta_auto", (prot,), {}), ("genomics.fasta", "read_fasta_auto", (text,), {}),
        ("genomics.fasta", "read_fasta_auto", (mixed,), {}),
        ("genomics.fasta", "write_factors_dna_w_reference_fasta_files_to_binary", (dna, dna, out), {}),
        ("genomics.fasta", "write_factors_dna_w_reference_fasta_files_to_binary", (tmp_path / "nope.fa", dna, out), {}),
        ("genomics.sequences", "factorize_dna_w_reference_seq", ("ACGT", "ACGA"), {}),
        ("genomics.sequences", "factorize_dna_w_reference_seq", (b"ACGT", b"ACGA"), {"validate": False}),
        ("genomics.sequences", "factorize_dna_w_reference_seq", ("", "ACGA"), {}),
        ("genomics.sequences", "factorize_dna_w_reference_seq_file", ("ACGT", "ACGA", out), {}),
        ("genomics.sequences", "factorize_dna_w_reference_seq_file", ("ACGT", "ACGA", str(out)), {"validate": False}),
    ]
    ours = {"genomics.fasta": our_fa, "genomics.sequences": our_seq}
    strip = lambda o: tuple(str(x).replace(str(tmp_path), "<tmp>") if isinstance(x, str) else x for x in o)   # noqa: E731
    for sub, fname, args, kwargs in calls:
        want = _outcome(getattr(mods[sub], fname), *args, **kwargs)
        got = _outcome(getattr(ours[sub], fname), *args, **kwargs)
        assert strip(want) == strip(got), (sub, fname, args, kwargs)
    assert len(fake_ref.calls) >= 6 and fake_ref.calls == fake_us.calls


def test_public_names_of_the_reference_python_layer_exist_here():
    """Every public function / class the reference defines in core.py, utils.py, parallel.py, genomics/fasta.py,
    genomics/sequences.py and genomics/significance.py exists under the same name in the mirrored module, with the
    same parameter names (plotting is the documented exception)."""
    import ast
    import inspect

    import nolzss_b200
    from nolzss_b200 import core, genomics, parallel
    from nolzss_b200 import utils as our_utils
    from nolzss_b200.genomics import fasta, sequences, significance

    base = "/root/reference/src/noLZSS"
    pairs = [("core.py", core, nolzss_b200), ("utils.py", our_utils, nolzss_b200), ("parallel.py", parallel, None),
             ("genomics/fasta.py", fasta, genomics), ("genomics/sequences.py", sequences, genomics),
             ("genomics/significance.py", significance, genomics)]
    skip = {"plot_factor_lengths"}                       # utils.py:360-: matplotlib plot, outside the hot path
    checked = 0
    for rel, mod, pkg in pairs:
        tree = ast.parse(open(f"{base}/{rel}").read())
        for node in tree.body:
            if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and not node.name.startswith("_") and node.name not in skip:
                assert hasattr(mod, node.name), (rel, node.name)
                if pkg is not None:
                    assert hasattr(pkg, node.name), (rel, node.name, "not re-exported")
                if isinstance(node, ast.FunctionDef):
                    want = [a.arg for a in node.args.posonlyargs + node.args.args + node.args.kwonlyargs]
                    got = list(inspect.signature(getattr(mod, node.name)).parameters)
                    assert want == got, (rel, node.name, want, got)
                    defaults_ref = [ast.literal_eval(d) for d in node.args.defaults]
                    sig = inspect.signature(getattr(mod, node.name))
                    defaults_us = [p.default for p in sig.parameters.values() if p.default is not inspect.Parameter.empty]
                    assert defaults_ref == defaults_us, (rel, node.name, defaults_ref, defaults_us)
                checked += 1
    assert checked >= 30


def test_all_binding_signatures_match_bindings_cpp():
    """src/cpp/bindings.cpp is the drop-in boundary: every m.def(name, ..., py::arg(...) [= default])
    is parsed from the reference's source and compared with the ctypes shim's Python signature."""
    import inspect
    import re

    from nolzss_b200 import _noLZSS as ext

    src = open("/root/reference/src/cpp/bindings.cpp").read()
    starts = [m.start() for m in re.finditer(r'\bm\.def\("', src)]
    assert len(starts) == 42
    for a, b in zip(starts, starts[1:] + [len(src)]):
        block = src[a:b]
        name = re.match(r'm\.def\("(\w+)"', block).group(1)
        doc_at = block.find('R"doc(')
        head = block if doc_at < 0 else block[:doc_at]
        args = re.findall(r'py::arg\("(\w+)"\)\s*(?:=\s*([^,\n]+?))?\s*(?:,|\))', head)
        want_names = [n for n, _ in args]
        want_defaults = []
        for _, d in args:
            d = d.strip()
            if not d:
                continue
            if d.startswith("std::string("):
                d = d[len("std::string("):-1]
            want_defaults.append(d.strip('"') if d.startswith('"') else int(d))
        sig = inspect.signature(getattr(ext, name))
        assert list(sig.parameters) == want_names, (name, want_names, list(sig.parameters))
        got_defaults = [p.default for p in sig.parameters.values() if p.default is not inspect.Parameter.empty]
        assert got_defaults == want_defaults, (name, want_defaults, got_defaults)

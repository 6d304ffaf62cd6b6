"""CPU: host side of the path (prepare, FASTA parser, factor files, readers, API surface)."""
import inspect
import json
import os
import random
import struct

import numpy as np
import pytest

import treewalk_model as tm
import oracle_py as orc
from kats import GENERAL_KATS, RC_KATS, plain_tuples, rc_tuples
from nolzss_b200 import _lib as L
from nolzss_b200 import _noLZSS as ext
from nolzss_b200 import utils

GOLD = os.path.join(os.path.dirname(__file__), "golden")

REFERENCE_BINDINGS = """factorize factorize_file count_factors count_factors_file write_factors_binary_file
factorize_dna_w_rc factorize_file_dna_w_rc count_factors_dna_w_rc count_factors_file_dna_w_rc
write_factors_binary_file_dna_w_rc factorize_multiple_dna_w_rc factorize_file_multiple_dna_w_rc
count_factors_multiple_dna_w_rc count_factors_file_multiple_dna_w_rc write_factors_binary_file_multiple_dna_w_rc
factorize_fasta_multiple_dna_w_rc factorize_dna_rc_w_ref_fasta_files factorize_fasta_multiple_dna_no_rc
write_factors_binary_file_fasta_multiple_dna_w_rc write_factors_binary_file_fasta_multiple_dna_no_rc
prepare_multiple_dna_sequences_w_rc prepare_multiple_dna_sequences_no_rc factorize_dna_w_reference_seq
factorize_dna_w_reference_seq_file factorize_w_reference factorize_w_reference_file
write_factors_dna_w_reference_fasta_files_to_binary parallel_factorize_to_file parallel_factorize_file_to_file
parallel_factorize_dna_w_rc_to_file parallel_factorize_file_dna_w_rc_to_file
parallel_write_factors_binary_file_fasta_multiple_dna_w_rc parallel_write_factors_binary_file_fasta_multiple_dna_no_rc
parallel_write_factors_dna_w_reference_fasta_files_to_binary factorize_fasta_dna_w_rc_per_sequence
factorize_fasta_dna_no_rc_per_sequence write_factors_binary_file_fasta_dna_w_rc_per_sequence
write_factors_binary_file_fasta_dna_no_rc_per_sequence count_factors_fasta_dna_w_rc_per_sequence
count_factors_fasta_dna_no_rc_per_sequence parallel_write_factors_binary_file_fasta_dna_w_rc_per_sequence
parallel_write_factors_binary_file_fasta_dna_no_rc_per_sequence""".split()


def test_all_42_binding_names_exist_with_reference_argument_names():
    assert len(REFERENCE_BINDINGS) == 42
    for name in REFERENCE_BINDINGS:
        assert callable(getattr(ext, name)), name
    for cls in ("Factor", "FastaFactorizationResult", "FastaPerSequenceFactorizationResult"):
        assert hasattr(ext, cls)
    sig = lambda f: list(inspect.signature(getattr(ext, f)).parameters)
    assert sig("factorize") == ["data"]
    assert sig("factorize_file") == ["path", "reserve_hint"]
    assert sig("write_factors_binary_file") == ["in_path", "out_path"]
    assert sig("parallel_factorize_to_file") == ["text", "output_path", "num_threads", "start_pos"]
    assert sig("parallel_write_factors_binary_file_fasta_dna_w_rc_per_sequence") == ["fasta_path", "out_dir", "num_threads", "sanitize_mode"]
    assert sig("factorize_dna_rc_w_ref_fasta_files") == ["reference_fasta_path", "target_fasta_path", "sanitize_mode"]
    f = ext.Factor(3, 4, (1 << 63) | 9)
    assert (f.start, f.length, f.ref, f.is_rc) == (3, 4, 9, True)


def test_prepare_matches_reference_restatement():
    rnd = random.Random(4)
    for _ in range(60):
        seqs = [bytes(rnd.choice(b"ACGTacgt") for _ in range(rnd.randint(0, 12))) for _ in range(rnd.randint(1, 6))]
        if all(len(s) == 0 for s in seqs):
            continue
        for fn, ref in ((ext.prepare_multiple_dna_sequences_w_rc, tm.prepare_multiple_dna_sequences_w_rc),
                        (ext.prepare_multiple_dna_sequences_no_rc, tm.prepare_multiple_dna_sequences_no_rc)):
            s, ol, sent = fn([q.decode() for q in seqs])
            es, eol, esent = ref(seqs)
            assert (s.encode("latin-1"), ol, sent) == (es, eol, esent)
    # sentinel bytes beyond 127 survive the str round trip the reference's tests use
    s, ol, sent = ext.prepare_multiple_dna_sequences_w_rc(["A"] * 125)
    raw = s.encode("latin-1")
    assert len(raw) == 500 and len(set(raw[1::2])) == 250 and not (set(raw[1::2]) & set(b"ACGT\x00"))
    assert ext.prepare_multiple_dna_sequences_w_rc([]) == ("", 0, [])


def test_prepare_errors():
    with pytest.raises(ValueError, match="Too many sequences"):
        ext.prepare_multiple_dna_sequences_w_rc(["A"] * 126)
    with pytest.raises(ValueError, match="Too many sequences"):
        ext.prepare_multiple_dna_sequences_no_rc(["A"] * 251)
    assert len(ext.prepare_multiple_dna_sequences_no_rc(["A"] * 250)[2]) == 249
    with pytest.raises(RuntimeError, match="Invalid nucleotide"):
        ext.prepare_multiple_dna_sequences_w_rc(["ACGT", "ACNT"])
    with pytest.raises(RuntimeError, match="All sequences are empty"):
        ext.prepare_multiple_dna_sequences_no_rc(["", ""])
    s, ol, sent = ext.prepare_multiple_dna_sequences_w_rc(["AC", "", "GT"])     # empty records are skipped
    assert s.encode("latin-1") == b"AC\x01GT\x02AC\x03GT\x04"


def _parse(path, mode=0):
    import ctypes

    lib = L.load()
    h = ctypes.c_void_p()
    L.check(lib.nlz_fasta_parse(os.fsencode(path), mode, ctypes.byref(h)))
    out = []
    for i in range(lib.nlz_fasta_num_sequences(h)):
        n = ctypes.c_uint64(0)
        p = lib.nlz_fasta_sequence(h, i, ctypes.byref(n))
        out.append((lib.nlz_fasta_id(h, i).decode(), ctypes.string_at(p, n.value)))
    lib.nlz_fasta_free(h)
    return out


def test_fasta_parser(tmp_path):
    assert _parse(os.path.join(GOLD, "two_records.fasta")) == [("seq1", b"ATCGATCGATTAGC"), ("seq2", b"GCTAGCTAGGCATCGATCGAT")]
    recs = _parse(os.path.join(GOLD, "messy.fasta"))
    assert recs == [("chrA", b"ACGTACGTACGTACGGGTTTAAACCC"), ("chrB/with:odd*chars", b"TTAGGGTTAGGGTTAGGGTTAGGG"),
                    ("chrC", b"ACGTTGCA")]
    with pytest.raises(RuntimeError, match="Invalid nucleotide 'N' found in sequence with ID: chrA"):
        _parse(os.path.join(GOLD, "messy.fasta"), 1)
    with pytest.raises(RuntimeError, match="Cannot open FASTA file"):
        _parse(str(tmp_path / "missing.fa"))
    p = tmp_path / "empty.fa"
    p.write_text(">only_header\n\n")
    with pytest.raises(RuntimeError, match="No valid sequences found"):
        _parse(str(p))
    with pytest.raises(ValueError, match="sanitize_mode"):
        ext.factorize_fasta_multiple_dna_w_rc(str(p), "bogus")


def test_factor_file_writer_and_readers_round_trip(tmp_path):
    import ctypes

    tr = np.array([[0, 1, 0], [1, 1, 1], [2, 5, (1 << 63) | 0], [7, 1, 7]], dtype=np.uint64)
    meta = b"seqA\x00seqB\x00" + struct.pack("<Q", 3)
    path = str(tmp_path / "f.bin")
    L.check(L.load().nlz_write_factor_file(os.fsencode(path), tr.ctypes.data, len(tr), meta, len(meta), 2, 1, 8))
    raw = open(path, "rb").read()
    assert len(raw) == 4 * 24 + len(meta) + 48 and raw[-48:-40] == b"noLZSSv2"
    assert struct.unpack("<QQQQQ", raw[-40:]) == (4, 2, 1, 48 + len(meta), 8)
    assert utils.read_factors_binary_file(path) == plain_tuples(tr)
    md = utils.read_binary_file_metadata(path)
    assert md == {"sentinel_factor_indices": [3], "sequence_names": ["seqA", "seqB"], "num_sequences": 2,
                  "num_sentinels": 1, "num_factors": 4, "total_length": 8}
    full = utils.read_factors_binary_file_with_metadata(path)
    assert full["factors"] == [(0, 1, 0, False), (1, 1, 1, False), (2, 5, 0, True), (7, 1, 7, False)]
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"x" * 60)
    with pytest.raises(utils.NoLZSSError, match="magic"):
        utils.read_factors_binary_file(bad)
    with pytest.raises(utils.NoLZSSError, match="File not found"):
        utils.read_factors_binary_file(tmp_path / "nope.bin")


def test_sentinel_factor_identification():
    import ctypes

    tr = np.array([[0, 3, 0], [3, 1, 3], [4, 2, 0], [6, 1, 6], [7, 2, 1]], dtype=np.uint64)
    pos = np.array([3, 6, 20], dtype=np.uint64)
    out, n = ctypes.POINTER(ctypes.c_uint64)(), ctypes.c_uint64(0)
    L.check(L.load().nlz_identify_sentinel_factors(tr.ctypes.data, len(tr), pos.ctypes.data, len(pos), ctypes.byref(out), ctypes.byref(n)))
    assert [out[i] for i in range(n.value)] == [1, 3]
    L.load().nlz_free(out)
    bad = np.array([[0, 3, 0], [3, 2, 3]], dtype=np.uint64)
    with pytest.raises(RuntimeError, match="unexpected length"):
        L.check(L.load().nlz_identify_sentinel_factors(bad.ctypes.data, 2, pos.ctypes.data, 1, ctypes.byref(out), ctypes.byref(n)))


def test_python_layer_validation():
    import nolzss_b200 as nz

    with pytest.raises(utils.InvalidInputError):
        nz.factorize(b"")
    with pytest.raises(utils.InvalidInputError):
        nz.factorize("héllo")
    with pytest.raises(TypeError):
        nz.factorize(123)
    with pytest.raises(FileNotFoundError):
        nz.factorize_file("/nonexistent/file.txt")
    a = nz.analyze_alphabet(b"AACG")
    assert a["size"] == 3 and a["total_length"] == 4 and abs(a["entropy"] - 1.5) < 1e-12
    from nolzss_b200.genomics import detect_sequence_type, is_dna_sequence

    assert is_dna_sequence("acgt") and not is_dna_sequence("acgu")
    assert detect_sequence_type("MKVLAAGIVGL") == "protein"


def test_golden_vectors_agree_with_oracle():
    cases = json.load(open(os.path.join(GOLD, "golden_factors.json")))["cases"]
    assert len(cases) >= 20
    for c in cases:
        t = c["text"].encode("latin-1")
        if c["mode"] == "general":
            got = orc.factorize(t, c["start_pos"])
        elif c["mode"] == "dna_rc":
            got = orc.factorize_multiple_dna_w_rc(tm.prepare_multiple_dna_sequences_w_rc([t])[0])
        else:
            got = orc.factorize_multiple_dna_w_rc(t, c["start_pos"])
        assert [[int(x) for x in r] for r in got] == c["factors"]
    for text, exp in GENERAL_KATS.items():
        assert any(c["text"].encode("latin-1") == text and [tuple(r) for r in c["factors"]] == exp for c in cases)


def test_shuffle_control_matches_reference_algorithm(tmp_path):
    """shuffle_fasta_sequences(method="reference") = batch_factorize.py:209-270: random.seed(seed), one
    random.shuffle per record in file order, 80-character lines, headers kept."""
    import random

    from nolzss_b200.genomics import shuffle_fasta_sequences

    recs = {"chr1 some description": "ACGT" * 50 + "GATTACA", "chr2": "TTTTGGGGCCCCAAAA" * 11, "chr3": "A"}
    src = tmp_path / "in.fasta"
    src.write_text("".join(f">{k}\n{v[:70]}\n{v[70:]}\n" for k, v in recs.items()))
    assert shuffle_fasta_sequences(src, tmp_path / "o" / "ref.fasta", seed=5)
    random.seed(5)
    exp = ""
    for k, v in recs.items():
        lst = list(v)
        random.shuffle(lst)
        sh = "".join(lst)
        exp += f">{k.split()[0]}\n" + "".join(sh[i:i + 80] + "\n" for i in range(0, len(sh), 80))
    assert (tmp_path / "o" / "ref.fasta").read_text() == exp
    assert shuffle_fasta_sequences(src, tmp_path / "np.fasta", seed=5, method="numpy")
    from nolzss_b200.genomics.fasta import _parse_fasta_content
    got = _parse_fasta_content((tmp_path / "np.fasta").read_text())
    assert list(got) == [k.split()[0] for k in recs] and all(sorted(got[k.split()[0]]) == sorted(recs[k]) for k in recs)
    assert not shuffle_fasta_sequences(tmp_path / "missing.fasta", tmp_path / "x.fasta")


def _py_fasta_model(data: bytes, strict: bool):
    """Serial restatement of fasta_processor.cpp:28-128 (the semantics the threaded parser must keep)."""
    recs, cur_id, cur = [], None, bytearray()

    def flush():
        nonlocal cur
        if cur_id is None:
            return
        if cur:
            recs.append((cur_id, bytes(cur)))
        cur = bytearray()

    for line in data.split(b"\n"):
        line = line.rstrip(b" \t\r\n\v\f")
        if not line:
            continue
        if line[:1] == b">":
            flush()
            toks = line[1:].split()
            if not toks:
                raise RuntimeError("Empty sequence header in FASTA file")
            cur_id = toks[0].decode()
        else:
            for c in line:
                ch = bytes([c])
                if ch in b"ACGTacgt":
                    cur += ch.upper()
                elif ch.isspace():
                    continue
                elif strict:
                    raise RuntimeError(f"Invalid nucleotide '{ch.decode()}' found in sequence with ID: {cur_id}")
    flush()
    return recs


def test_threaded_fasta_parser_equals_serial_model(tmp_path):
    """Files large enough to be cut into several ranges (>= 4 MiB per host thread): records, ids, dropped
    characters, empty records and the first strict-mode error must be those of a serial scan."""
    import random as _r
    rnd = _r.Random(9)
    rng = np.random.default_rng(9)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    chunks = []
    for k in range(1500):
        n = rnd.choice([0, 1, 50, 5_000, 40_000])
        seq = acgt[rng.integers(0, 4, n)].tobytes()
        if k % 7 == 0 and n:
            seq = seq.lower()
        if k % 11 == 0 and n > 10:
            seq = seq[:7] + b"NN-" + seq[7:]
        width = rnd.choice([60, 80, 10_000_000])
        body = b"".join(seq[i:i + width] + rnd.choice([b"\n", b"\r\n", b" \n"]) for i in range(0, len(seq), width))
        chunks.append(b">rec%d some description > with a bracket\n" % k + body + (b"\n" if k % 5 == 0 else b""))
    data = b"".join(chunks)
    assert len(data) > 12 * (1 << 20)
    path = tmp_path / "big.fasta"
    path.write_bytes(data)
    assert _parse(str(path), 0) == _py_fasta_model(data, False)
    with pytest.raises(RuntimeError) as e1:
        _py_fasta_model(data, True)
    with pytest.raises(RuntimeError) as e2:
        _parse(str(path), 1)
    assert str(e1.value) in str(e2.value)
    # an error that only a late range sees
    clean = b"".join(b">c%d\n" % k + acgt[rng.integers(0, 4, 30_000)].tobytes() + b"\n" for k in range(500))
    late = clean + b">bad\nACGTXACGT\n" + b">after\nACGT\n"
    path.write_bytes(late)
    with pytest.raises(RuntimeError, match="Invalid nucleotide 'X' found in sequence with ID: bad"):
        _parse(str(path), 1)
    got = _parse(str(path), 0)
    assert got[-2] == ("bad", b"ACGTACGT") and got[-1] == ("after", b"ACGT") and len(got) == 502

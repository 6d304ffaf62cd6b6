"""GPU: the reference-facing Python API (the 42 `_noLZSS` names and the package wrappers) end to end,
following the cross-API consistency strategy of the reference's own tests
(tests/test_cpp_bindings.py, test_parallel_fasta.py, test_per_sequence_fasta.py, test_reference_seq.py)."""
import json
import os
import struct

import numpy as np
import pytest

import oracle_py as orc
import treewalk_model as tm
from kats import GENERAL_KATS, RC_KATS
from nolzss_b200 import _noLZSS as ext
from nolzss_b200 import utils
from nolzss_b200 import workloads as wl

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RC_MASK = 1 << 63


def _footer(path):
    raw = open(path, "rb").read()
    assert raw[-48:-40] == b"noLZSSv2"
    return raw, struct.unpack("<QQQQQ", raw[-40:])


def _rc4(arr):
    return [(int(s), int(l), int(r) & ~RC_MASK, bool(int(r) & RC_MASK)) for s, l, r in arr]


def test_golden_vectors():
    for c in json.load(open(os.path.join(GOLD, "golden_factors.json")))["cases"]:
        t = c["text"].encode("latin-1")
        exp = [tuple(r) for r in c["factors"]]
        if c["mode"] == "general" and c["start_pos"] == 0:
            assert ext.factorize(t) == exp
            assert ext.count_factors(t) == len(exp)
        elif c["mode"] == "dna_rc":
            assert ext.factorize_dna_w_rc(t) == [(s, l, r & ~RC_MASK, bool(r & RC_MASK)) for s, l, r in exp]
            assert ext.count_factors_dna_w_rc(t) == len(exp)
        elif c["mode"] == "rc_prepared" and c["start_pos"] == 0:
            assert ext.factorize_multiple_dna_w_rc(t) == [(s, l, r & ~RC_MASK, bool(r & RC_MASK)) for s, l, r in exp]
            assert ext.count_factors_multiple_dna_w_rc(t) == len(exp)


def test_kats_and_invariants():
    for text, exp in GENERAL_KATS.items():
        assert ext.factorize(text) == exp
    for text, exp in RC_KATS.items():
        assert ext.factorize_dna_w_rc(text) == exp
    f = ext.factorize(wl.planted_dna(30_000, 3, scale=0.05).tobytes())
    assert f[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(f, f[1:])) and f[-1][0] + f[-1][1] == 30_000
    for bad in (np.zeros(4, dtype=np.uint16), np.zeros((2, 2), dtype=np.uint8)):
        with pytest.raises(ValueError):
            ext.factorize(bad)
    assert ext.factorize(bytearray(b"abracadabra")) == ext.factorize(memoryview(b"abracadabra"))


def test_memory_file_count_binary_consistency(tmp_path):
    t = wl.planted_dna(40_000, 8, scale=0.05).tobytes()
    src = tmp_path / "in.txt"
    src.write_bytes(t)
    # general
    mem = ext.factorize(t)
    assert ext.factorize_file(str(src)) == mem and ext.factorize_file(str(src), 100) == mem
    assert ext.count_factors(t) == ext.count_factors_file(str(src)) == len(mem)
    out = str(tmp_path / "g.bin")
    assert ext.write_factors_binary_file(str(src), out) == len(mem)
    raw, ft = _footer(out)
    assert ft == (len(mem), 0, 0, 48, len(t)) and len(raw) == 24 * len(mem) + 48          # footer V1
    assert utils.read_factors_binary_file(out) == mem
    # DNA with RC
    mem = ext.factorize_dna_w_rc(t)
    assert ext.factorize_file_dna_w_rc(str(src)) == mem
    assert ext.count_factors_dna_w_rc(t) == ext.count_factors_file_dna_w_rc(str(src)) == len(mem)
    out = str(tmp_path / "d.bin")
    assert ext.write_factors_binary_file_dna_w_rc(str(src), out) == len(mem)
    raw, ft = _footer(out)
    assert ft == (len(mem), 1, 0, 49, len(t)) and len(raw) == 24 * len(mem) + 49          # footer V2
    assert utils.read_factors_binary_file_with_metadata(out)["factors"] == mem
    # prepared multi-sequence text
    S, ol, sent = ext.prepare_multiple_dna_sequences_w_rc([t[:15_000].decode(), t[15_000:22_000].decode()])
    Sb = S.encode("latin-1")
    psrc = tmp_path / "prep.txt"
    psrc.write_bytes(Sb)
    mem = ext.factorize_multiple_dna_w_rc(Sb)
    assert _rc4(orc.factorize_multiple_dna_w_rc(Sb)) == mem
    assert ext.factorize_file_multiple_dna_w_rc(str(psrc)) == mem
    assert ext.count_factors_multiple_dna_w_rc(Sb) == ext.count_factors_file_multiple_dna_w_rc(str(psrc)) == len(mem)
    out = str(tmp_path / "m.bin")
    assert ext.write_factors_binary_file_multiple_dna_w_rc(str(psrc), out) == len(mem)
    assert _footer(out)[1] == (len(mem), 0, 0, 48, len(Sb))                                  # footer V3
    with pytest.raises(RuntimeError, match="Cannot open"):
        ext.factorize_file(str(tmp_path / "missing"))
    with pytest.raises(RuntimeError, match="Invalid nucleotide"):
        (tmp_path / "nl.txt").write_bytes(b"ACGT\n")
        ext.factorize_file_dna_w_rc(str(tmp_path / "nl.txt"))


def test_fasta_concatenated_and_writers(tmp_path):
    fa = os.path.join(GOLD, "messy.fasta")
    factors, sidx, ids = ext.factorize_fasta_multiple_dna_w_rc(fa)
    assert ids == ["chrA", "chrB/with:odd*chars", "chrC"]
    seqs = [b"ACGTACGTACGTACGGGTTTAAACCC", b"TTAGGGTTAGGGTTAGGGTTAGGG", b"ACGTTGCA"]
    S, ol, sent = tm.prepare_multiple_dna_sequences_w_rc(seqs)
    assert factors == _rc4(orc.factorize_multiple_dna_w_rc(S))
    assert [factors[i][0] for i in sidx] == sent[:2] and all(factors[i][1] == 1 for i in sidx)
    out = str(tmp_path / "w.bin")
    for fn in (ext.write_factors_binary_file_fasta_multiple_dna_w_rc, ext.parallel_write_factors_binary_file_fasta_multiple_dna_w_rc):
        assert fn(fa, out) == len(factors)
        md = utils.read_factors_binary_file_with_metadata(out)
        assert md["factors"] == factors and md["sequence_names"] == ids and md["sentinel_factor_indices"] == sidx
        raw, ft = _footer(out)
        names = sum(len(i) + 1 for i in ids)
        assert ft == (len(factors), 3, len(sidx), 48 + names + 8 * len(sidx), sum(f[1] for f in factors))   # footer V7
    # no RC: general factorization of T1 s0 T2 s1 T3 (sentinels between records only)
    f2, s2, ids2 = ext.factorize_fasta_multiple_dna_no_rc(fa)
    S2, _, sent2 = tm.prepare_multiple_dna_sequences_no_rc(seqs)
    assert f2 == _rc4(orc.factorize(S2)) and [f2[i][0] for i in s2] == sent2 and ids2 == ids
    assert ext.write_factors_binary_file_fasta_multiple_dna_no_rc(fa, out) == len(f2)
    assert utils.read_factors_binary_file_with_metadata(out)["factors"] == f2
    with pytest.raises(RuntimeError, match="Invalid nucleotide"):
        ext.factorize_fasta_multiple_dna_w_rc(fa, "strict")
    with pytest.raises(RuntimeError, match="Cannot open FASTA file"):
        ext.factorize_fasta_multiple_dna_w_rc(str(tmp_path / "missing.fa"))


def test_per_sequence(tmp_path):
    fa = os.path.join(GOLD, "messy.fasta")
    seqs = [b"ACGTACGTACGTACGGGTTTAAACCC", b"TTAGGGTTAGGGTTAGGGTTAGGG", b"ACGTTGCA"]
    per, ids = ext.factorize_fasta_dna_w_rc_per_sequence(fa)
    assert ids == ["chrA", "chrB/with:odd*chars", "chrC"]
    for got, s in zip(per, seqs):
        assert got == _rc4(orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s)))
    counts, ids2, total = ext.count_factors_fasta_dna_w_rc_per_sequence(fa)
    assert counts == [len(p) for p in per] and total == sum(counts) and ids2 == ids
    # no-RC per-sequence drops the last base of every record (fasta_processor.cpp:469-471)
    per_n, _ = ext.factorize_fasta_dna_no_rc_per_sequence(fa)
    for got, s in zip(per_n, seqs):
        assert got == _rc4(orc.factorize(s[:-1]))
    assert ext.count_factors_fasta_dna_no_rc_per_sequence(fa)[0] == [len(p) for p in per_n]
    d = str(tmp_path / "out" / "nested")
    for fn in (ext.write_factors_binary_file_fasta_dna_w_rc_per_sequence,
               lambda a, b: ext.parallel_write_factors_binary_file_fasta_dna_w_rc_per_sequence(a, b, 4)):
        assert fn(fa, d) == total
        files = sorted(os.listdir(d))
        assert files == ["chrA.bin", "chrB_with_odd_chars.bin", "chrC.bin"]
        md = utils.read_factors_binary_file_with_metadata(os.path.join(d, "chrB_with_odd_chars.bin"))
        assert md["factors"] == per[1] and md["sequence_names"] == ["chrB/with:odd*chars"] and md["num_sentinels"] == 0
        _, ft = _footer(os.path.join(d, "chrC.bin"))
        assert ft == (len(per[2]), 1, 0, 48 + 5, 8)                                          # footer V8
    assert ext.write_factors_binary_file_fasta_dna_no_rc_per_sequence(fa, str(tmp_path / "o2")) == sum(len(p) for p in per_n)


def test_reference_plus_target(tmp_path):
    ref, tgt = "ACGTACGTTTGACCA", "TGGTCAAACGTACGT"
    f = ext.factorize_dna_w_reference_seq(ref, tgt)
    S, _, _ = tm.prepare_multiple_dna_sequences_w_rc([ref.encode(), tgt.encode()])
    assert f == _rc4(orc.factorize_multiple_dna_w_rc(S, len(ref) + 1)) and f[0][0] == len(ref) + 1
    out = str(tmp_path / "r.bin")
    assert ext.factorize_dna_w_reference_seq_file(ref, tgt, out) == len(f)
    raw, ft = _footer(out)
    assert ft == (len(f), 2, 1, 48, len(tgt)) and len(raw) == 24 * len(f) + 48               # footer V4 (no names)
    g = ext.factorize_w_reference("abracadabra", "cadabra_abra")
    assert g == [tuple(int(x) for x in r) for r in orc.factorize(b"abracadabra\x01cadabra_abra", 12)]
    assert ext.factorize_w_reference_file("abracadabra", "cadabra_abra", out) == len(g)
    assert _footer(out)[1] == (len(g), 2, 1, 48, 12)                                         # footer V5
    # FASTA reference + target
    two = os.path.join(GOLD, "two_records.fasta")
    messy = os.path.join(GOLD, "messy.fasta")
    ff, sidx, ids = ext.factorize_dna_rc_w_ref_fasta_files(two, messy)
    seqs = [b"ATCGATCGATTAGC", b"GCTAGCTAGGCATCGATCGAT", b"ACGTACGTACGTACGGGTTTAAACCC", b"TTAGGGTTAGGGTTAGGGTTAGGG", b"ACGTTGCA"]
    S, _, sent = tm.prepare_multiple_dna_sequences_w_rc(seqs)
    start = len(seqs[0]) + 1 + len(seqs[1]) + 1
    assert ff == _rc4(orc.factorize_multiple_dna_w_rc(S, start)) and ids[:2] == ["seq1", "seq2"] and len(ids) == 5
    for fn in (ext.write_factors_dna_w_reference_fasta_files_to_binary,
               ext.parallel_write_factors_dna_w_reference_fasta_files_to_binary):
        assert fn(two, messy, out) == len(ff)
        md = utils.read_factors_binary_file_with_metadata(out)
        assert md["factors"] == ff and md["sequence_names"] == ids and md["sentinel_factor_indices"] == sidx


def test_parallel_entry_points(tmp_path):
    t = wl.planted_dna(20_000, 5, scale=0.05).tobytes()
    out = str(tmp_path / "p.bin")
    for sp in (0, 4321):
        n = ext.parallel_factorize_to_file(t, out, 3, sp)
        exp = [tuple(int(x) for x in r) for r in orc.factorize(t, sp)]
        assert n == len(exp) and utils.read_factors_binary_file(out) == exp
        assert _footer(out)[1] == (n, 0, 0, 48, sum(e[1] for e in exp))                      # footer V6
    src = tmp_path / "in.txt"
    src.write_bytes(t)
    assert ext.parallel_factorize_file_to_file(str(src), out) == len(ext.factorize(t))
    n = ext.parallel_factorize_dna_w_rc_to_file(t, out, 8)
    assert utils.read_factors_binary_file_with_metadata(out)["factors"] == ext.factorize_dna_w_rc(t) and n > 0
    assert ext.parallel_factorize_file_dna_w_rc_to_file(str(src), out) == n
    with pytest.raises(ValueError, match="start_pos"):
        ext.parallel_factorize_to_file(t, out, 0, len(t))
    assert ext.parallel_factorize_to_file(b"", str(tmp_path / "never.bin")) == 0 and not (tmp_path / "never.bin").exists()
    from nolzss_b200 import parallel as par

    assert par.parallel_factorize(b"CGACACGTA", num_threads=2) == ext.factorize(b"CGACACGTA")


def test_package_level_api(tmp_path):
    import nolzss_b200 as nz
    from nolzss_b200 import genomics

    assert nz.factorize("abracadabra") == GENERAL_KATS[b"abracadabra"]
    info = nz.factorize_with_info(b"abracadabra")
    assert info["num_factors"] == 8 and info["input_size"] == 11 and info["alphabet_info"]["size"] == 5
    src = tmp_path / "t.txt"
    src.write_bytes(b"abracadabra")
    assert nz.factorize_file(src) == nz.factorize(b"abracadabra") and nz.count_factors_file(src) == 8
    nz.write_factors_binary_file(str(src), tmp_path / "sub" / "o.bin")      # `data` is the input path (core.py:132)
    assert nz.read_factors_binary_file(tmp_path / "sub" / "o.bin") == nz.factorize(b"abracadabra")
    res = genomics.read_nucleotide_fasta(os.path.join(GOLD, "two_records.fasta"))
    assert [r[0] for r in res] == ["seq1", "seq2"] and res[0][1] == nz.factorize(b"ATCGATCGATTAGC")
    with pytest.raises(genomics.FASTAError):
        genomics.read_nucleotide_fasta(os.path.join(GOLD, "messy.fasta"))
    assert genomics.factorize_dna_w_reference_seq("ACGTACGT", "ACGTACGT")[0] == (9, 8, 0, False)


def test_multi_record_fasta_config3_style(tmp_path):
    """configs[2] in miniature: many records, per-sequence RC factorization, counts, per-record files, and the
    same records sharded over 2 'ranks' by nolzss_b200.sharding.assign_records."""
    from nolzss_b200.sharding import assign_records

    recs = wl.c3_records(40, 3000, seed=3)
    fa = tmp_path / "many.fasta"
    with open(fa, "w") as f:
        for rid, s in recs:
            f.write(f">{rid} synthetic\n")
            for k in range(0, len(s), 70):
                f.write(s[k:k + 70].decode() + "\n")
    per, ids = ext.factorize_fasta_dna_w_rc_per_sequence(str(fa))
    assert ids == [r for r, _ in recs]
    for got, (_, s) in zip(per, recs):
        assert got == _rc4(orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s)))
    counts, _, total = ext.count_factors_fasta_dna_w_rc_per_sequence(str(fa))
    assert counts == [len(p) for p in per] and total == sum(counts)
    out = tmp_path / "bins"
    assert ext.parallel_write_factors_binary_file_fasta_dna_w_rc_per_sequence(str(fa), str(out), 8) == total
    assert sorted(os.listdir(out)) == sorted(f"{r}.bin" for r, _ in recs)
    shares = assign_records([len(s) for _, s in recs], 2)
    merged = {}
    for share in shares:                      # what each rank would compute, gathered by record index
        for i in share:
            merged[i] = ext.factorize_dna_w_rc(recs[i][1])
    assert [merged[i] for i in range(len(recs))] == per


def test_shuffled_control_on_the_gpu_and_two_file_threshold(tmp_path):
    """configs[4] tail: per-record permutation on the GPU (same multiset of bases per record, headers
    kept, deterministic for a seed), then real + shuffled factor files (footer V7) -> calculate_factor_length_threshold."""
    from collections import Counter

    from nolzss_b200 import genomics
    from nolzss_b200.genomics import shuffle_control

    fa = tmp_path / "g.fasta"
    recs = [wl.planted_dna(120_000, 50 + r, scale=0.2).tobytes() for r in range(3)]
    with open(fa, "wb") as f:
        for r, s in enumerate(recs):
            f.write(b">chr%d desc\n" % r)
            for i in range(0, len(s), 80):
                f.write(s[i:i + 80] + b"\n")
    a, b = tmp_path / "a.fasta", tmp_path / "b.fasta"
    assert shuffle_control.shuffle_fasta_sequences(fa, a, seed=7, method="gpu")
    assert shuffle_control.shuffle_fasta_sequences(fa, b, seed=7, method="gpu")
    assert a.read_bytes() == b.read_bytes()
    from nolzss_b200.genomics.fasta import _parse_fasta_content
    sh = _parse_fasta_content(a.read_text())
    assert list(sh.keys()) == ["chr0", "chr1", "chr2"]
    for r, s in enumerate(recs):
        assert Counter(sh[f"chr{r}"].encode()) == Counter(s) and sh[f"chr{r}"].encode() != s
    real_bin, shuf_bin, n_real, n_shuf = shuffle_control.factorize_with_shuffled_control(fa, tmp_path / "out", seed=6, method="gpu")
    res = genomics.calculate_factor_length_threshold(real_bin, shuf_bin, tau_expected_fp=10.0)
    assert res["N_real"] == n_real and res["N_shuf"] == n_shuf and n_shuf > n_real
    assert res["L_star"] is not None and res["L_star"] <= 40

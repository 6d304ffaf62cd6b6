"""GPU: the reference's real-sequence fixtures (tests/golden/ref_fixtures, copied from the reference's
tests/resources) through the FASTA entry points -- bit-exact against the CPU oracle, plus the property checks of the
reference's own tests/test_factorization_validation.py:92-211 (every factor is a true match or reverse-complement
match; the factors cover the text without gaps; basic invariants)."""
import os

import numpy as np
import pytest

import oracle_py as orc
import treewalk_model as tm
from nolzss_b200 import _noLZSS as ext
from nolzss_b200 import workloads as wl

pytestmark = pytest.mark.gpu
FIX = os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures")
RC_MASK = 1 << 63
SMALL = ["T3.fasta", "T7.fasta", "short_dna1.fasta", "short_dna2.fasta", "test_viral_dna.fna", "test_bacterial_dna.fna"]


def _rc4(arr):
    return [(int(s), int(l), int(r) & ~RC_MASK, bool(int(r) & RC_MASK)) for s, l, r in arr]


def _parse(path):
    """Serial model of parse_fasta_sequences_and_ids in its default mode (fasta_processor.cpp:28-128): id = first
    token after '>', non-ACGT characters dropped, upper-cased, records without bases dropped."""
    ids, seqs, cur_id, cur = [], [], None, []
    keep = set(b"ACGTacgt")
    for line in open(path, "rb"):
        line = line.strip()
        if not line:
            continue
        if line.startswith(b">"):
            if cur_id is not None and b"".join(cur):
                ids.append(cur_id); seqs.append(b"".join(cur))
            tok = line[1:].split()
            cur_id, cur = (tok[0].decode() if tok else ""), []
        else:
            cur.append(bytes(ch for ch in line if ch in keep).upper())
    if cur_id is not None and b"".join(cur):
        ids.append(cur_id); seqs.append(b"".join(cur))
    return ids, seqs


def _vibrio():
    for p in (os.path.join(FIX, "Vibrio_cholerae.fna"), "/root/reference/tests/resources/Vibrio_cholerae.fna",
              os.path.join(os.path.dirname(FIX), "..", "..", "scratch_ab", "Vibrio_cholerae.fna")):
        if os.path.exists(p):
            return p
    return None


def _check_properties(factors, S, original_length):
    """tests/test_factorization_validation.py:92-211 of the reference, vectorised per factor."""
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    end = 0
    for start, length, ref, is_rc in factors:
        assert length > 0 and start == end, (start, end)                     # coverage without gaps (:172-199)
        end = start + length
        if start >= original_length:
            continue
        assert start + length <= original_length and ref + length <= len(S)
        sub, src = S[start:start + length], S[ref:ref + length]
        if is_rc:
            assert sub == src.translate(comp)[::-1], (start, length, ref)    # :139-150
        elif not (length == 1 and ref == start):
            assert sub == src and ref < start, (start, length, ref)         # :152-158
    assert end >= original_length - 1


@pytest.mark.parametrize("name", SMALL)
def test_fasta_concatenated_rc_vs_oracle(name):
    path = os.path.join(FIX, name)
    ids, seqs = _parse(path)
    S, orig_len, sent = tm.prepare_multiple_dna_sequences_w_rc(seqs)
    f, sidx, got_ids = ext.factorize_fasta_multiple_dna_w_rc(path)
    assert got_ids == ids
    assert f == _rc4(orc.factorize_multiple_dna_w_rc(S))
    assert [f[i][0] for i in sidx] == sent[: len(seqs) - 1]
    _check_properties(f, S, orig_len)
    # the binding-level prepare function gives the same string (tests/test_factorization_validation.py:109)
    S2, ol2, sp2 = ext.prepare_multiple_dna_sequences_w_rc([s.decode() for s in seqs])
    assert S2.encode("latin-1") == S and ol2 == orig_len and list(sp2) == list(sent)


@pytest.mark.parametrize("name", SMALL)
def test_fasta_no_rc_and_per_sequence_vs_oracle(name):
    path = os.path.join(FIX, name)
    ids, seqs = _parse(path)
    S, _, sent = tm.prepare_multiple_dna_sequences_no_rc(seqs)
    f, sidx, got_ids = ext.factorize_fasta_multiple_dna_no_rc(path)
    assert got_ids == ids and f == _rc4(orc.factorize(S))
    per, ids2 = ext.factorize_fasta_dna_w_rc_per_sequence(path)
    assert ids2 == ids
    for got, s in zip(per, seqs):
        assert got == _rc4(orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s)))
    counts, _, total = ext.count_factors_fasta_dna_w_rc_per_sequence(path)
    assert counts == [len(p) for p in per] and total == sum(counts)
    per_n, _ = ext.factorize_fasta_dna_no_rc_per_sequence(path)
    for got, s in zip(per_n, seqs):
        assert got == _rc4(orc.factorize(s[:-1]))          # the reference's last-base quirk (fasta_processor.cpp:469-471)


def test_reference_plus_target_t7_with_t3_reference(tmp_path):
    """tests/test_reference_seq.py of the reference factorizes T7 against T3; its .bin golden is stale (SURVEY 8c trap
    2: a genuine 13-base forward match at factor 35), so the oracle is the authority."""
    t3, t7 = os.path.join(FIX, "T3.fasta"), os.path.join(FIX, "T7.fasta")
    ids3, s3 = _parse(t3)
    ids7, s7 = _parse(t7)
    seqs = s3 + s7
    S, _, _ = tm.prepare_multiple_dna_sequences_w_rc(seqs)
    start = sum(len(s) + 1 for s in s3)
    f, sidx, ids = ext.factorize_dna_rc_w_ref_fasta_files(t3, t7)
    exp = _rc4(orc.factorize_multiple_dna_w_rc(S, start))
    assert f == exp and ids == ids3 + ids7 and f[0][0] == start
    out = str(tmp_path / "t7_w_t3.bin")
    assert ext.write_factors_dna_w_reference_fasta_files_to_binary(t3, t7, out) == len(f)
    f2 = ext.factorize_dna_w_reference_seq(s3[0].decode(), s7[0].decode())
    assert f2 == exp


def test_vibrio_cholerae_when_available():
    path = _vibrio()
    if path is None:
        pytest.skip("Vibrio_cholerae.fna (4.1 Mbp, not committed) not found")
    ids, seqs = _parse(path)
    S, orig_len, sent = tm.prepare_multiple_dna_sequences_w_rc(seqs)
    f, sidx, got_ids = ext.factorize_fasta_multiple_dna_w_rc(path)
    exp = orc.factorize_multiple_dna_w_rc(S)
    assert got_ids == ids and len(f) == len(exp)
    assert np.array_equal(np.array([(s, l, r | (RC_MASK if rc else 0)) for s, l, r, rc in f], dtype=np.uint64), exp)
    _check_properties(f, S, orig_len)

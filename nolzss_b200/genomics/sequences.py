"""Sequence-type detection and reference+target DNA factorization
(reference: /root/reference/src/noLZSS/genomics/sequences.py:12-221)."""
from __future__ import annotations

import re
from pathlib import Path
from typing import Union

from .. import _noLZSS as _ext
from ..utils import validate_input

# sequences.py:12-59 of the reference: case-insensitive; the protein alphabet is the 20 standard amino acids plus the
# extension codes B, J, O, U, X, Z; `$` also matches before a trailing newline (re.match semantics kept).
_DNA = re.compile(r"^[ATGC]+$", re.IGNORECASE)
_PROTEIN = re.compile(r"^[ACDEFGHIKLMNPQRSTVWYBJOUXZ]+$", re.IGNORECASE)
_AMINO_ONLY = frozenset("EFHIKLMNPQRSVWY")        # amino-acid codes that are not nucleotides (:95)
_AMINO = frozenset("ACDEFGHIKLMNPQRSTVWY")        # (:103; the detector uses the 20 standard codes only)
_NUCLEOTIDES = frozenset("ACGT")


def _ascii_str(data):
    """bytes -> ASCII str (None if not ASCII); other types unchanged."""
    if isinstance(data, bytes):
        try:
            return data.decode("ascii")
        except UnicodeDecodeError:
            return None
    return data


def is_dna_sequence(data: Union[str, bytes]) -> bool:
    text = _ascii_str(data)
    return isinstance(text, str) and bool(_DNA.match(text))


def is_protein_sequence(data: Union[str, bytes]) -> bool:
    text = _ascii_str(data)
    return isinstance(text, str) and bool(_PROTEIN.match(text))


def detect_sequence_type(data: Union[str, bytes]) -> str:
    """'dna', 'protein', 'text' or 'binary' (reference: sequences.py:62-117): non-ASCII bytes and non-strings are
    binary; anything with a non-alphabetic character, and the empty string, is text; then amino-acid-specific
    letters decide between protein and DNA."""
    text = _ascii_str(data)
    if not isinstance(text, str):
        return "binary"
    up = text.upper()
    if not up or not all(c.isalpha() for c in up):
        return "text"
    letters = set(up)
    has_amino_specific = bool(letters & _AMINO_ONLY)
    all_amino = letters <= _AMINO
    if has_amino_specific and all_amino:
        return "protein"
    if letters <= _NUCLEOTIDES and not has_amino_specific:
        return "dna"
    return "protein" if all_amino else "text"


def _as_str(x):
    return x.decode("ascii") if isinstance(x, bytes) else x


def factorize_dna_w_reference_seq(reference_seq, target_seq, validate: bool = True):
    if validate:
        reference_seq, target_seq = validate_input(reference_seq), validate_input(target_seq)
    return _ext.factorize_dna_w_reference_seq(_as_str(reference_seq), _as_str(target_seq))


def factorize_dna_w_reference_seq_file(reference_seq, target_seq, output_path: Union[str, Path], validate: bool = True) -> int:
    if validate:
        reference_seq, target_seq = validate_input(reference_seq), validate_input(target_seq)
    return _ext.factorize_dna_w_reference_seq_file(_as_str(reference_seq), _as_str(target_seq), str(output_path))

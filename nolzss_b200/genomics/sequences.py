"""Sequence-type detection and reference+target DNA factorization
(reference: /root/reference/src/noLZSS/genomics/sequences.py:12-221)."""
from __future__ import annotations

import re
from pathlib import Path
from typing import Union

from .. import _noLZSS as _ext
from ..utils import validate_input

_DNA = re.compile(r"^[ACGT]+$")
_PROTEIN = re.compile(r"^[ACDEFGHIKLMNPQRSTVWY]+$")


def _to_str(data: Union[str, bytes]) -> str:
    if isinstance(data, bytes):
        try:
            return data.decode("ascii")
        except UnicodeDecodeError:
            return ""
    return data


def is_dna_sequence(data: Union[str, bytes]) -> bool:
    return bool(_DNA.match(_to_str(data).upper()))


def is_protein_sequence(data: Union[str, bytes]) -> bool:
    return bool(_PROTEIN.match(_to_str(data).upper()))


def detect_sequence_type(data: Union[str, bytes]) -> str:
    if isinstance(data, bytes):
        try:
            text = data.decode("ascii")
        except UnicodeDecodeError:
            return "binary"
    else:
        text = data
    if not text:
        return "text"
    up = text.upper()
    if _DNA.match(up):
        return "dna"
    if _PROTEIN.match(up):
        return "protein"
    if all(32 <= ord(c) <= 126 or c in "\t\n\r" for c in text):
        return "text"
    return "binary"


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

"""Pure-Python FASTA readers (reference: /root/reference/src/noLZSS/genomics/fasta.py:23-290).
`read_nucleotide_fasta` factorizes every record separately in general (no-RC) mode on the GPU."""
from __future__ import annotations

import re
from pathlib import Path
from typing import Dict, List, Tuple, Union

from .. import _noLZSS as _ext
from ..core import factorize
from ..utils import NoLZSSError
from .sequences import detect_sequence_type


class FASTAError(NoLZSSError):
    """Raised when FASTA file parsing or validation fails."""


def _parse_fasta_content(content: str) -> Dict[str, str]:
    sequences: Dict[str, str] = {}
    cur_id, cur = None, []
    for line_num, line in enumerate(content.splitlines(), 1):
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            if cur_id is not None:
                sequences[cur_id] = "".join(cur)
            header = line[1:].strip()
            if not header:
                raise FASTAError(f"Empty sequence header at line {line_num}")
            cur_id, cur = header.split()[0], []
        else:
            if cur_id is None:
                raise FASTAError(f"Sequence data before header at line {line_num}")
            cur.append(re.sub(r"\s", "", line.upper()))
    if cur_id is not None:
        sequences[cur_id] = "".join(cur)
    if not sequences:
        raise FASTAError("No valid sequences found in FASTA file")
    return sequences


def _read(filepath) -> Dict[str, str]:
    filepath = Path(filepath)
    if not filepath.exists():
        raise FileNotFoundError(f"FASTA file not found: {filepath}")
    try:
        content = filepath.read_text(encoding="utf-8")
    except UnicodeDecodeError as e:
        raise FASTAError(f"File encoding error: {e}")
    return _parse_fasta_content(content)


def read_nucleotide_fasta(filepath: Union[str, Path]) -> List[Tuple[str, List[Tuple[int, int, int]]]]:
    results = []
    for seq_id, seq in _read(filepath).items():
        seq = seq.upper()
        if not re.match(r"^[ACGT]+$", seq):
            raise FASTAError(f"Sequence '{seq_id}' contains invalid nucleotides: {set(seq) - set('ACGT')}")
        try:
            results.append((seq_id, factorize(seq.encode("ascii"))))
        except Exception as e:   # noqa: BLE001 - mirrors the reference's wrapping
            raise FASTAError(f"Failed to factorize sequence '{seq_id}': {e}")
    return results


def read_protein_fasta(filepath: Union[str, Path]) -> List[Tuple[str, str]]:
    valid = set("ACDEFGHIKLMNPQRSTVWY")
    results = []
    for seq_id, seq in _read(filepath).items():
        seq = seq.upper()
        if not seq or set(seq) - valid:
            raise FASTAError(f"Sequence '{seq_id}' contains invalid amino acids: {set(seq) - valid}")
        results.append((seq_id, seq))
    return results


def read_fasta_auto(filepath: Union[str, Path]):
    """Nucleotide FASTA -> [(id, factors)], protein FASTA -> [(id, sequence)], decided on the first record
    (reference: fasta.py:177-228)."""
    seqs = _read(filepath)
    if not seqs:
        raise FASTAError("No sequences found in FASTA file")
    kind = detect_sequence_type(next(iter(seqs.values())))
    if kind == "dna":
        return read_nucleotide_fasta(filepath)
    if kind == "protein":
        return read_protein_fasta(filepath)
    raise FASTAError(f"Cannot determine sequence type. Detected: {kind}. "
                     f"Expected DNA (A,C,T,G) or protein (amino acids) sequences.")


def write_factors_dna_w_reference_fasta_files_to_binary(reference_fasta_path: Union[str, Path],
                                                        target_fasta_path: Union[str, Path],
                                                        output_path: Union[str, Path],
                                                        sanitize_mode: str = "remove_ambiguous") -> int:
    """Reference + target FASTA files -> one binary factor file of the target (reference: fasta.py:231-292; path
    checks, output directory creation and the sanitize_mode check happen here, the work in the extension)."""
    reference_path = Path(reference_fasta_path)
    target_path = Path(target_fasta_path)
    if not reference_path.exists():
        raise FileNotFoundError(f"Reference FASTA file not found: {reference_path}")
    if not target_path.exists():
        raise FileNotFoundError(f"Target FASTA file not found: {target_path}")
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    if sanitize_mode not in {"remove_ambiguous", "strict"}:
        raise ValueError("sanitize_mode must be 'remove_ambiguous' or 'strict'")
    return _ext.write_factors_dna_w_reference_fasta_files_to_binary(str(reference_path), str(target_path), str(output_path),
                                                                    sanitize_mode)

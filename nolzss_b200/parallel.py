"""Wrappers over the `parallel_*` entry points (reference: /root/reference/src/noLZSS/parallel.py:23-226).
`num_threads` is accepted for compatibility; the GPU pipeline has no thread count to choose."""
from __future__ import annotations

import os
import tempfile
from pathlib import Path
from collections import namedtuple
from typing import List, Union

from . import _noLZSS as _ext
from .utils import read_factors_binary_file, validate_input

# the reference reads the temporary file back into namedtuples with .start / .length / .ref (parallel.py:147-153),
# which its own tests/test_parallel.py:58 relies on
Factor = namedtuple("Factor", ["start", "length", "ref"])


def parallel_factorize_to_file(text, output_path, num_threads: int = 0, start_pos: int = 0, validate: bool = True) -> int:
    if validate:
        text = validate_input(text)
    return _ext.parallel_factorize_to_file(text, str(output_path), num_threads, start_pos)


def parallel_factorize_file_to_file(input_path, output_path, num_threads: int = 0, start_pos: int = 0) -> int:
    input_path = Path(input_path)
    if not input_path.exists():
        raise FileNotFoundError(f"Input file not found: {input_path}")
    return _ext.parallel_factorize_file_to_file(str(input_path), str(output_path), num_threads, start_pos)


def parallel_factorize(text, num_threads: int = 0, start_pos: int = 0, validate: bool = True) -> List[Factor]:
    if validate:
        text = validate_input(text)
    fd, tmp = tempfile.mkstemp(suffix=".bin")
    os.close(fd)
    try:
        _ext.parallel_factorize_to_file(text, tmp, num_threads, start_pos)
        return [Factor(*f) for f in read_factors_binary_file(tmp)]
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)


def parallel_factorize_dna_w_rc_to_file(text, output_path, num_threads: int = 0, validate: bool = True) -> int:
    if validate:
        text = validate_input(text)
    return _ext.parallel_factorize_dna_w_rc_to_file(text, str(output_path), num_threads)


def parallel_factorize_file_dna_w_rc_to_file(input_path, output_path, num_threads: int = 0) -> int:
    input_path = Path(input_path)
    if not input_path.exists():
        raise FileNotFoundError(f"Input file not found: {input_path}")
    return _ext.parallel_factorize_file_dna_w_rc_to_file(str(input_path), str(output_path), num_threads)

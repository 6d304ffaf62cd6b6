"""Python-level API of the general (byte text) mode; same names, arguments and errors as
/root/reference/src/noLZSS/core.py:25-257, on top of the CUDA-backed `_noLZSS` shim."""
from __future__ import annotations

from pathlib import Path
from typing import List, Tuple, Union

from . import _noLZSS as _ext
from .utils import analyze_alphabet, validate_input


def factorize(data: Union[str, bytes], validate: bool = True) -> List[Tuple[int, int, int]]:
    if validate:
        data = validate_input(data)
    return _ext.factorize(data)


def factorize_file(filepath: Union[str, Path], reserve_hint: int = 0) -> List[Tuple[int, int, int]]:
    filepath = Path(filepath)
    if not filepath.exists():
        raise FileNotFoundError(f"File not found: {filepath}")
    return _ext.factorize_file(str(filepath), reserve_hint)


def count_factors(data: Union[str, bytes], validate: bool = True) -> int:
    if validate:
        data = validate_input(data)
    return _ext.count_factors(data)


def count_factors_file(filepath: Union[str, Path], validate: bool = True) -> int:
    filepath = Path(filepath)
    if not filepath.exists():
        raise FileNotFoundError(f"File not found: {filepath}")
    return _ext.count_factors_file(str(filepath))


def write_factors_binary_file(data: Union[str, bytes], output_filepath: Union[str, Path]) -> None:
    # Like the reference (core.py:110-132), the validated `data` is forwarded as the INPUT PATH of the
    # extension's write_factors_binary_file(in_path, out_path).
    data = validate_input(data)
    output_filepath = Path(output_filepath)
    output_filepath.parent.mkdir(parents=True, exist_ok=True)
    _ext.write_factors_binary_file(data, str(output_filepath))


def factorize_with_info(data: Union[str, bytes], validate: bool = True) -> dict:
    if validate:
        data = validate_input(data)
    factors = _ext.factorize(data)
    return {"factors": factors, "alphabet_info": analyze_alphabet(data), "input_size": len(data),
            "num_factors": len(factors)}


def _as_str(x):
    return x.decode("ascii") if isinstance(x, bytes) else x


def factorize_w_reference(reference_seq, target_seq, validate: bool = True):
    if validate:
        reference_seq, target_seq = validate_input(reference_seq), validate_input(target_seq)
    return _ext.factorize_w_reference(_as_str(reference_seq), _as_str(target_seq))


def factorize_w_reference_file(reference_seq, target_seq, output_path, validate: bool = True) -> int:
    if validate:
        reference_seq, target_seq = validate_input(reference_seq), validate_input(target_seq)
    return _ext.factorize_w_reference_file(_as_str(reference_seq), _as_str(target_seq), str(output_path))

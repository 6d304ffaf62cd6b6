"""Input validation, alphabet analysis and noLZSSv2 binary factor-file readers.

Mirrors the behaviour of /root/reference/src/noLZSS/utils.py (validate_input :26-58,
analyze_alphabet :61-103, read_factors_binary_file :106-155, read_binary_file_metadata :158-247,
read_factors_binary_file_with_metadata :250-357).  The readers parse the factor block with one numpy
view instead of a per-factor struct.unpack loop (same return values).  Plotting is out of scope.
"""
from __future__ import annotations

import math
import struct
from collections import Counter
from pathlib import Path
from typing import Any, Dict, List, Tuple, Union

import numpy as np

RC_MASK = 1 << 63
_FOOTER = struct.Struct("<8sQQQQQ")   # factorizer.hpp:64-77


class NoLZSSError(Exception):
    """Base exception for noLZSS-related errors."""


class InvalidInputError(NoLZSSError):
    """Raised when input data is invalid for factorization."""


def validate_input(data: Union[str, bytes]) -> bytes:
    if isinstance(data, str):
        try:
            data = data.encode("ascii")
        except UnicodeEncodeError as e:
            raise InvalidInputError(f"Input string must contain only ASCII characters (1 byte each): {e}")
    elif not isinstance(data, bytes):
        raise TypeError(f"Input must be str or bytes, got {type(data)}")
    if len(data) == 0:
        raise InvalidInputError("Input data cannot be empty")
    if b"\x00" in data[:-1]:
        raise InvalidInputError("Input data contains null bytes")
    return data


def analyze_alphabet(data: Union[str, bytes]) -> Dict[str, Any]:
    if isinstance(data, bytes):
        chars = data.decode("ascii")
    elif isinstance(data, str):
        chars = data
    else:
        raise TypeError(f"Input must be str or bytes, got {type(data)}")
    dist = Counter(chars)
    total = len(chars)
    entropy = -sum((c / total) * math.log2(c / total) for c in dist.values()) if total else 0.0
    return {
        "size": len(dist),
        "characters": set(dist),
        "distribution": dist,
        "entropy": entropy,
        "most_common": dist.most_common(10),
        "total_length": total,
    }


def _read_footer(f, path) -> Tuple[int, int, int, int, int]:
    try:
        f.seek(-48, 2)
    except OSError:
        raise NoLZSSError("File too small to contain valid footer")
    raw = f.read(48)
    if len(raw) != 48:
        raise NoLZSSError("File too small to contain valid footer")
    magic, nf, nseq, nsent, fsize, total = _FOOTER.unpack(raw)
    if magic != b"noLZSSv2":
        raise NoLZSSError("Invalid file format: missing noLZSS magic footer (expected v2 format)")
    return nf, nseq, nsent, fsize, total


def _read_factor_block(f, num_factors: int) -> np.ndarray:
    f.seek(0)
    arr = np.fromfile(f, dtype="<u8", count=num_factors * 3)
    if arr.size != num_factors * 3:
        raise NoLZSSError(f"Insufficient data for factor {arr.size // 3}")
    return arr.reshape(num_factors, 3)


def read_factors_array(filepath: Union[str, Path]) -> np.ndarray:
    """(num_factors, 3) uint64 view of a factor file (vectorised sibling of read_factors_binary_file)."""
    filepath = Path(filepath)
    if not filepath.exists():
        raise NoLZSSError(f"File not found: {filepath}")
    try:
        with open(filepath, "rb") as f:
            nf, *_ = _read_footer(f, filepath)
            return _read_factor_block(f, nf)
    except IOError as e:
        raise NoLZSSError(f"Error reading file {filepath}: {e}")


def read_factors_binary_file(filepath: Union[str, Path]) -> List[Tuple[int, int, int]]:
    return list(map(tuple, read_factors_array(filepath).tolist()))


def _read_metadata(f, path):
    nf, nseq, nsent, fsize, total = _read_footer(f, path)
    f.seek(-fsize, 2)
    full = f.read(fsize)
    if len(full) != fsize:
        raise NoLZSSError(f"Could not read full footer: expected {fsize}, got {len(full)}")
    meta = full[: fsize - 48]
    names, off = [], 0
    for _ in range(nseq):
        end = meta.find(b"\x00", off)
        if end < 0:
            raise NoLZSSError("Invalid sequence name format")
        names.append(meta[off:end].decode("utf-8"))
        off = end + 1
    if len(meta) - off < 8 * nsent:
        raise NoLZSSError("Insufficient data for sentinel indices")
    sent = list(struct.unpack(f"<{nsent}Q", meta[off:off + 8 * nsent])) if nsent else []
    return nf, nseq, nsent, total, names, sent


def read_binary_file_metadata(filepath: Union[str, Path]) -> Dict[str, Any]:
    filepath = Path(filepath)
    if not filepath.exists():
        raise NoLZSSError(f"File not found: {filepath}")
    try:
        with open(filepath, "rb") as f:
            nf, nseq, nsent, total, names, sent = _read_metadata(f, filepath)
    except IOError as e:
        raise NoLZSSError(f"Error reading file {filepath}: {e}")
    except struct.error as e:
        raise NoLZSSError(f"Error unpacking binary data: {e}")
    return {"sentinel_factor_indices": sent, "sequence_names": names, "num_sequences": nseq,
            "num_sentinels": nsent, "num_factors": nf, "total_length": total}


def read_factors_binary_file_with_metadata(filepath: Union[str, Path]) -> Dict[str, Any]:
    filepath = Path(filepath)
    if not filepath.exists():
        raise NoLZSSError(f"File not found: {filepath}")
    try:
        with open(filepath, "rb") as f:
            nf, nseq, nsent, total, names, sent = _read_metadata(f, filepath)
            arr = _read_factor_block(f, nf)
    except IOError as e:
        raise NoLZSSError(f"Error reading file {filepath}: {e}")
    except struct.error as e:
        raise NoLZSSError(f"Error unpacking binary data: {e}")
    ref = arr[:, 2] if nf else np.zeros(0, dtype=np.uint64)
    factors = list(zip(arr[:, 0].tolist(), arr[:, 1].tolist(), (ref & np.uint64(RC_MASK - 1)).tolist(),
                       (ref >> np.uint64(63)).astype(bool).tolist())) if nf else []
    return {"factors": factors, "sentinel_factor_indices": sent, "sequence_names": names, "num_sequences": nseq,
            "num_sentinels": nsent, "total_length": total}


def plot_factor_lengths(factors_or_file, save_path=None, show_plot: bool = True) -> None:
    """Cumulative factor length against factor index (reference: utils.py:360-430).  Plotting is outside the hot
    path: the input handling and error types of the reference are kept so that callers and the
    reference's tests/test_utils.py behave the same; the figure itself needs matplotlib (a warning without it)."""
    import warnings

    if isinstance(factors_or_file, (str, Path)):
        lengths = read_factors_array(factors_or_file)[:, 1].astype(np.int64)
    elif isinstance(factors_or_file, list):
        lengths = np.array([f[1] for f in factors_or_file], dtype=np.int64)
    else:
        raise TypeError("factors_or_file must be a list of tuples or a path to a binary factors file")
    if lengths.size == 0:
        raise ValueError("No factors to plot")
    try:
        import matplotlib
        if not show_plot:
            matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        warnings.warn("matplotlib is required for plotting. Install with: pip install matplotlib", UserWarning)
        return
    fig, ax = plt.subplots(figsize=(10, 6))
    ax.scatter(np.cumsum(lengths), np.arange(1, lengths.size + 1), s=4, alpha=0.7)
    ax.set_xlabel("Cumulative factor length")
    ax.set_ylabel("Factor index")
    ax.set_title("Factor length accumulation")
    if save_path is not None:
        fig.savefig(str(save_path), dpi=150, bbox_inches="tight")
    if show_plot:
        plt.show()
    plt.close(fig)

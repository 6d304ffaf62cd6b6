"""Shuffled control for the factor-length significance test.

Reference: `shuffle_fasta_sequences` in /root/reference/src/noLZSS/genomics/batch_factorize.py:209-270 -- every
record is shuffled on its own, headers are kept, lines are 80 characters.  `method="reference"` reproduces the
reference's output byte for byte for a given seed (same `random.seed` / `random.shuffle` calls in record order);
`method="numpy"` permutes with numpy (vectorised: the reference needs about a microsecond per base in Python,
i.e. the better part of an hour for a 3.1 Gbp genome) and is NOT seed-compatible with the reference;
`method="gpu"` permutes every record on the GPU (`torch.randperm` on the device the factorizer runs on, seeded
generator; deterministic for a given seed and device type, not seed-compatible with the reference either).

`factorize_with_shuffled_control` is the two-job driver: real and shuffled FASTA -> two noLZSSv2 factor files
(the inputs of `calculate_factor_length_threshold`).
"""
from __future__ import annotations

import logging
import random
from pathlib import Path
from typing import Optional, Union

import numpy as np

from .fasta import _parse_fasta_content


def shuffle_fasta_sequences(input_path: Union[str, Path], output_path: Union[str, Path], seed: Optional[int] = None,
                            logger: Optional[logging.Logger] = None, method: str = "reference") -> bool:
    logger = logger or logging.getLogger(__name__)
    input_path, output_path = Path(input_path), Path(output_path)
    if method not in ("reference", "numpy", "gpu"):
        raise ValueError("method must be 'reference', 'numpy' or 'gpu'")
    try:
        logger.info(f"Creating shuffled version of {input_path}")
        with open(input_path, "r", encoding="utf-8") as f:
            content = f.read()
        sequences = _parse_fasta_content(content)
        if not sequences:
            logger.error(f"No sequences found in {input_path}")
            return False
        rng = None
        if method == "reference":
            if seed is not None:
                random.seed(seed)                              # batch_factorize.py:244-246 (module-level generator)
        elif method == "numpy":
            rng = np.random.default_rng(seed)
        else:
            import torch

            if not torch.cuda.is_available():
                raise RuntimeError("method='gpu' needs a CUDA device")
            gen = torch.Generator(device="cuda")
            gen.manual_seed(0 if seed is None else int(seed))
        output_path.parent.mkdir(parents=True, exist_ok=True)
        with open(output_path, "w", encoding="utf-8") as f:
            for seq_id, sequence in sequences.items():
                if method == "reference":
                    seq_list = list(sequence)
                    random.shuffle(seq_list)                   # :255-257
                    shuffled = "".join(seq_list)
                elif method == "numpy":
                    arr = np.frombuffer(sequence.encode("ascii"), dtype=np.uint8)
                    shuffled = rng.permutation(arr).tobytes().decode("ascii")
                else:
                    arr = torch.frombuffer(bytearray(sequence.encode("ascii")), dtype=torch.uint8).cuda()
                    perm = torch.randperm(arr.numel(), device="cuda", generator=gen)
                    shuffled = arr[perm].cpu().numpy().tobytes().decode("ascii")
                f.write(f">{seq_id}\n")
                for i in range(0, len(shuffled), 80):          # :261-263
                    f.write(shuffled[i:i + 80] + "\n")
        logger.info(f"Successfully created shuffled FASTA at {output_path}")
        return True
    except Exception as e:                                      # the reference logs and returns False (:268-270)
        logger.error(f"Failed to shuffle {input_path}: {e}")
        return False


def factorize_with_shuffled_control(fasta_path: Union[str, Path], out_dir: Union[str, Path], seed: Optional[int] = None,
                                    with_rc: bool = True, method: str = "numpy"):
    """Writes <out_dir>/<stem>.bin and <out_dir>/<stem>.shuffled.bin (concatenated multi-record DNA factorization,
    footer variant V7: names + sentinel factor indices) and returns (real_bin, shuffled_bin, n_real, n_shuffled)."""
    from .. import _noLZSS as ext

    fasta_path, out_dir = Path(fasta_path), Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    shuffled_fa = out_dir / (fasta_path.stem + ".shuffled.fasta")
    if not shuffle_fasta_sequences(fasta_path, shuffled_fa, seed, method=method):
        raise RuntimeError(f"could not create the shuffled control of {fasta_path}")
    write = ext.write_factors_binary_file_fasta_multiple_dna_w_rc if with_rc else ext.write_factors_binary_file_fasta_multiple_dna_no_rc
    real_bin, shuf_bin = out_dir / (fasta_path.stem + ".bin"), out_dir / (fasta_path.stem + ".shuffled.bin")
    n_real = write(str(fasta_path), str(real_bin))
    n_shuf = write(str(shuffled_fa), str(shuf_bin))
    return real_bin, shuf_bin, n_real, n_shuf

"""Generates tests/golden/significance_golden.json from the REFERENCE's own Python module
(/root/reference/src/noLZSS/genomics/significance.py), imported in the authoring container.  The reference cannot
travel to the GPU box, the vectors can.  Run:  python tests/golden/make_significance_golden.py

The module's only package-relative import (`from ..utils import read_factors_binary_file, NoLZSSError`) needs the
compiled extension, so the source is executed with that line replaced by stubs; everything that is recorded here
(`clopper_pearson_upper`, `infer_length_significance`) is pure numpy / scipy code of the reference, unmodified.
Floats are stored with float.hex() so that the comparison is exact."""
import json
import os
import sys

import numpy as np

REF = "/root/reference/src/noLZSS/genomics/significance.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    src = open(REF).read()
    stub = "class NoLZSSError(Exception):\n    pass\n\n\ndef read_factors_binary_file(path):\n    raise NotImplementedError\n"
    needle = "from ..utils import read_factors_binary_file, NoLZSSError"
    assert needle in src
    ns = {"__name__": "reference_significance"}
    exec(compile(src.replace(needle, stub), REF, "exec"), ns)
    return ns


def lengths(kind: str, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "geometric":                      # shuffled-genome-like: short factors, geometric tail
        return (8 + rng.geometric(0.35, n)).astype(np.int64)
    if kind == "heavy":                          # real-genome-like: the same bulk plus a heavy tail of repeats
        x = (8 + rng.geometric(0.35, n)).astype(np.int64)
        if n == 0:
            return x
        k = max(1, n // 50)
        x[rng.integers(0, n, k)] = rng.integers(20, 5000, k)
        return x
    if kind == "tiny":
        return rng.integers(1, 6, n).astype(np.int64)
    raise ValueError(kind)


CASES = [
    dict(real=("heavy", 5000, 1), shuf=("geometric", 5000, 2), tau=1.0, alpha=0.05),
    dict(real=("heavy", 20000, 3), shuf=("geometric", 15000, 4), tau=0.5, alpha=0.01),
    dict(real=("heavy", 300, 5), shuf=("geometric", 40000, 6), tau=5.0, alpha=0.1),
    dict(real=("tiny", 50, 7), shuf=("tiny", 60, 8), tau=1.0, alpha=0.05),
    dict(real=("geometric", 1000, 9), shuf=("geometric", 7, 10), tau=1.0, alpha=0.05),      # no L* at all
    dict(real=("heavy", 0, 11), shuf=("geometric", 100, 12), tau=1.0, alpha=0.05),           # empty real set
]
CP_CASES = [(5, 100, 0.05), (0, 100, 0.05), (100, 100, 0.05), (1, 2, 0.5), (37, 123456, 0.001), (99999, 100000, 0.2)]


def main():
    import warnings

    ref = load_reference()
    out = {"source": REF, "clopper_pearson_upper": [], "infer_length_significance": []}
    for k, n, a in CP_CASES:
        out["clopper_pearson_upper"].append({"k": k, "n": n, "alpha": a, "value": float(ref["clopper_pearson_upper"](k, n, a)).hex()})
    for c in CASES:
        real, shuf = lengths(*c["real"]), lengths(*c["shuf"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = ref["infer_length_significance"](real, shuf, tau_expected_fp=c["tau"], alpha_cp=c["alpha"])
        probes = [0.0, 1.0, 9.0, 10.5, 12.0, 25.0, 1e6]
        out["infer_length_significance"].append({
            "case": c, "N_real": int(r["N_real"]), "N_shuf": int(r["N_shuf"]), "L_star": r["L_star"],
            "uniq_L": [int(v) for v in r["uniq_L"]],
            "S0": [float(v).hex() for v in r["S0"]],
            "S0_upper": [float(v).hex() for v in r["S0_upper"]],
            "expected_fp_upper": [float(v).hex() for v in r["expected_fp_upper"]],
            "rarity_scores_real": [float(v).hex() for v in r["rarity_scores_real"][:200]],
            "rarity_sum": float(np.sum(r["rarity_scores_real"])).hex(),
            "p_any_ge": [[L, float(r["p_any_ge"](L)).hex()] for L in probes],
        })
    with open(os.path.join(HERE, "significance_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["infer_length_significance"]), "cases")


if __name__ == "__main__":
    sys.exit(main())

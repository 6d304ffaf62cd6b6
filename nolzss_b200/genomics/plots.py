"""Plotting is OUT OF SCOPE of this build (only the factorization hot path and the data formats
either side of it are rebuilt).  This module exists so that `from noLZSS.genomics.plots import PlotError, ...` -- which
the reference's own tests/test_genomics.py does at import time -- resolves when the package is used as a drop-in;
every plotting entry point raises PlotError."""
from __future__ import annotations


class PlotError(Exception):
    """Raised for plotting errors (reference: genomics/plots.py)."""


def _out_of_scope(name):
    def fn(*args, **kwargs):
        raise PlotError(f"{name}: plotting is not part of the B200 hot-path build; use the reference's "
                        "noLZSS.genomics.plots on the factor files this package writes (same noLZSSv2 format)")
    fn.__name__ = name
    return fn


plot_single_seq_accum_factors_from_file = _out_of_scope("plot_single_seq_accum_factors_from_file")
plot_multiple_seq_self_lz_factor_plot_from_file = _out_of_scope("plot_multiple_seq_self_lz_factor_plot_from_file")
plot_reference_seq_lz_factor_plot_simple = _out_of_scope("plot_reference_seq_lz_factor_plot_simple")
plot_reference_seq_lz_factor_plot = _out_of_scope("plot_reference_seq_lz_factor_plot")
plot_strand_bias_heatmap = _out_of_scope("plot_strand_bias_heatmap")

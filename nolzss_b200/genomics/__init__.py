"""Genomics entry points (reference: /root/reference/src/noLZSS/genomics/__init__.py:8-34): the RC-aware
bindings re-exported, the FASTA readers and the sequence helpers.  Plotting, batch drivers and the
significance analysis are outside the hot path."""
from .._noLZSS import (  # noqa: F401
    count_factors_dna_w_rc,
    count_factors_file_dna_w_rc,
    count_factors_file_multiple_dna_w_rc,
    count_factors_multiple_dna_w_rc,
    factorize_dna_w_rc,
    factorize_fasta_multiple_dna_w_rc,
    factorize_file_dna_w_rc,
    factorize_file_multiple_dna_w_rc,
    factorize_multiple_dna_w_rc,
    prepare_multiple_dna_sequences_w_rc,
    write_factors_binary_file_dna_w_rc,
    write_factors_binary_file_multiple_dna_w_rc,
)
from .fasta import *  # noqa: F401,F403
from .fasta import FASTAError, read_fasta_auto, read_nucleotide_fasta, read_protein_fasta  # noqa: F401
from .sequences import *  # noqa: F401,F403
from .sequences import (  # noqa: F401
    detect_sequence_type,
    factorize_dna_w_reference_seq,
    factorize_dna_w_reference_seq_file,
    is_dna_sequence,
    is_protein_sequence,
)
from .shuffle_control import factorize_with_shuffled_control, shuffle_fasta_sequences  # noqa: F401,E402

"""nolzss_b200: B200-native (sm_100a CUDA) non-overlapping LZSS factorization behind the Python API
of OmerKerner/noLZSS (reference: /root/reference/src/noLZSS/__init__.py).

    import nolzss_b200 as noLZSS
    noLZSS.factorize(b"abracadabra")

The package mirrors the reference's modules: `core`, `utils`, `parallel`, `genomics` and the
extension module `_noLZSS` (here a ctypes shim over libnolzss_b200.so).  There is no CPU fallback.
"""
from ._noLZSS import __version__  # noqa: F401
from .core import (  # noqa: F401
    count_factors,
    count_factors_file,
    factorize,
    factorize_file,
    factorize_w_reference,
    factorize_w_reference_file,
    factorize_with_info,
    write_factors_binary_file,
)
from .utils import (  # noqa: F401
    InvalidInputError,
    NoLZSSError,
    analyze_alphabet,
    read_binary_file_metadata,
    read_factors_array,
    read_factors_binary_file,
    read_factors_binary_file_with_metadata,
    validate_input,
)

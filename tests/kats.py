"""Known-answer vectors held by the reference's own tests / README for the hot path."""

# /root/reference/README.md:48-50 (general mode)
GENERAL_KATS = {
    b"abracadabra": [(0, 1, 0), (1, 1, 1), (2, 1, 2), (3, 1, 0), (4, 1, 4), (5, 1, 0), (6, 1, 6), (7, 4, 0)],
}

# /root/reference/tests/test_cpp_bindings.py:715-747 (factorize_dna_w_rc: (start, len, ref, is_rc))
RC_KATS = {
    b"AC": [(0, 1, 0, False), (1, 1, 1, False)],
    b"ACTGA": [(0, 1, 0, False), (1, 1, 1, False), (2, 1, 2, False), (3, 1, 3, False), (4, 1, 0, False)],
    b"ATGAT": [(0, 1, 0, False), (1, 1, 1, False), (2, 1, 2, False), (3, 2, 0, False)],
    b"ATGCAT": [(0, 1, 0, False), (1, 1, 1, False), (2, 1, 2, False), (3, 3, 0, True)],
    b"ATGATCTCA": [(0, 1, 0, False), (1, 1, 1, False), (2, 1, 2, False), (3, 2, 0, False), (5, 1, 5, False),
                   (6, 3, 1, True)],
    b"TATACATAG": [(0, 1, 0, False), (1, 1, 1, False), (2, 2, 0, False), (4, 1, 4, False), (5, 3, 1, False),
                   (8, 1, 8, False)],
}

# /root/reference/tests/test_genomics.py:277-285: third factor of "ATAT" in RC mode
ATAT_THIRD = (2, 2, 0, False)

RC_MASK = 1 << 63


def rc_tuples(arr):
    """(z,3) triples with RC_MASK -> [(start, len, ref&~mask, is_rc)] (bindings.cpp:226)."""
    return [(int(s), int(l), int(r) & ~RC_MASK, bool(int(r) & RC_MASK)) for s, l, r in arr]


def plain_tuples(arr):
    return [(int(s), int(l), int(r)) for s, l, r in arr]

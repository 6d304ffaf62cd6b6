"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Literal, brute-force restatement of the reference's two suffix-tree walks for SMALL inputs
(quadratic/cubic time; use for strings up to a few hundred characters).  It exists to pin the C
oracle (`oracle/nolzss_oracle.c`) and, through it, the CUDA path.

The reference walks an SDSL compressed suffix tree (sdsl-lite v3.0.3, not vendored under
/root/reference -- see DESIGN.md).  Every SDSL query used on the path has a unique mathematical
definition over the suffix array of `text + '\\0'`; this model evaluates those definitions
directly:

  cst.csa[k]                 -> sa[k]                     (suffix array of text·$, $ smallest)
  cst.csa.isa[i]             -> isa[i]
  path nodes of leaf r       -> the nested LCP-intervals around rank r (root ... leaf)
  bp_support.level_anc(l,D-d)-> path[d]                   (node at node-depth d on that path)
  cst.lb(v), cst.rb(v)       -> interval bounds
  cst.depth(v)               -> string depth of the interval (leaf: |text·$| - sa[r])
  rmq(lb, rb)                -> argmin over the interval
  lcp(cst, i, j)             -> factorizer_helpers.hpp:20-24 (LCA string depth)
  next_leaf(cst, l, len)     -> leaf of suffix sn + len   (factorizer_helpers.hpp:38-44)

Followed line by line:
  general mode : /root/reference/src/cpp/factorizer_core.hpp:51-119   (detail::nolzss)
  RC mode      : /root/reference/src/cpp/factorizer_core.hpp:177-383  (detail::nolzss_multiple_dna_w_rc)
  prepare w/ RC: /root/reference/src/cpp/factorizer.cpp:54-172
  prepare no RC: /root/reference/src/cpp/factorizer.cpp:194-294
"""
from __future__ import annotations

RC_MASK = 1 << 63
INF = (2**64 - 1) // 2


class _Cst:
    """Explicit suffix array + LCP-interval view of the suffix tree of data·$."""

    def __init__(self, data: bytes):
        # SDSL appends a 0 terminator that is smaller than every text byte (construct_im(..., 1)).
        self.x = bytes(data) + b"\x00"
        n1 = len(self.x)
        self.n1 = n1
        self.sa = sorted(range(n1), key=lambda i: self.x[i:])
        self.isa = [0] * n1
        for k, s in enumerate(self.sa):
            self.isa[s] = k
        self.lcp = [0] * (n1 + 1)
        for k in range(1, n1):
            a, b = self.sa[k - 1], self.sa[k]
            l = 0
            while a + l < n1 and b + l < n1 and self.x[a + l] == self.x[b + l]:
                l += 1
            self.lcp[k] = l

    def size(self):
        return self.n1

    def path(self, r):
        """Nodes on the root->leaf path of leaf rank r as (lb, rb, depth); path[0] is the root."""
        nodes = [(r, r, self.n1 - self.sa[r])]
        lo = hi = r
        while True:
            left = self.lcp[lo] if lo > 0 else -1
            right = self.lcp[hi + 1] if hi + 1 < self.n1 else -1
            d = max(left, right)
            if d <= 0:
                break
            while lo > 0 and self.lcp[lo] >= d:
                lo -= 1
            while hi + 1 < self.n1 and self.lcp[hi + 1] >= d:
                hi += 1
            nodes.append((lo, hi, d))
        if nodes[-1] != (0, self.n1 - 1, 0):
            nodes.append((0, self.n1 - 1, 0))
        nodes.reverse()
        return nodes

    def lcp_of(self, i, j):
        # factorizer_helpers.hpp:20-24
        if i == j:
            return self.n1 - i
        l = 0
        while self.x[i + l] == self.x[j + l]:
            l += 1
        return l


def nolzss(data: bytes, start_pos: int = 0):
    """factorizer_core.hpp:51-119.  Returns [(start, length, ref)]."""
    cst = _Cst(data)
    str_len = cst.size() - 1                                     # :54
    out = []
    lam = start_pos                                               # :56-58 (leaf of suffix start_pos)
    while lam < str_len:                                          # :66
        path = cst.path(cst.isa[lam])
        d = 1                                                     # :68
        u_min = 0
        while True:                                               # :70
            lb, rb, l = path[d]                                   # :71, :73
            v_min = min(cst.sa[lb:rb + 1])                        # :72
            if v_min + l - 1 < lam:                               # :75
                u_min = v_min                                     # :76
                d += 1
                continue
            u_lb, u_rb, u_depth = path[d - 1]                     # :79-80
            if v_min == lam:                                      # :82
                if d - 1 == 0:                                    # :83  u == root
                    l = 1
                    out.append((lam, l, lam))                     # :85
                else:
                    l = u_depth                                   # :90
                    out.append((lam, l, u_min))                   # :91
                break
            l = min(cst.lcp_of(lam, v_min), lam - v_min)          # :96-97
            if l <= u_depth:                                      # :98
                l = u_depth
                out.append((lam, l, u_min))                       # :100
            else:
                out.append((lam, l, v_min))                       # :105
            break
        lam = lam + l                                             # :113-115
    return out


def nolzss_multiple_dna_w_rc(S: bytes, start_pos: int = 0):
    """factorizer_core.hpp:177-383.  Returns [(start, length, ref_with_RC_MASK)]."""
    if len(S) == 0:                                               # :180
        return []
    if len(S) < 4:                                                # :189
        return []
    N = len(S) // 2 - 1                                           # :195
    if N == 0:                                                    # :196
        return []
    if start_pos >= N:                                            # :203
        raise ValueError("start_pos must be less than the original sequence length")
    cst = _Cst(S)                                                 # :208
    size = cst.size()
    fwd_starts = [INF] * size                                     # :212-213
    rc_ends = [INF] * size
    T_end, R_beg, R_end = N, N + 1, len(S) - 1                    # :215-217
    for k in range(size):                                         # :219-230
        posS = cst.sa[k]
        if posS < T_end:
            fwd_starts[k] = posS
        elif R_beg <= posS < R_end:
            rc_ends[k] = N - (posS - R_beg) - 1
    out = []
    i = start_pos
    while i < N:                                                  # :241
        path = cst.path(cst.isa[i])
        node_depth = len(path) - 1
        have_fwd = have_rc = False
        best_fwd_start = best_fwd_depth = 0
        best_rc_end = best_rc_posS = best_rc_depth = 0
        for step in range(1, node_depth + 1):                     # :256
            lb, rb, ell = path[step]                              # :257-258
            if ell == 0:                                          # :259
                break
            kF = min(range(lb, rb + 1), key=lambda k: fwd_starts[k])   # :264
            jF = fwd_starts[kF]
            okF = jF != INF and jF + ell - 1 < i                  # :266
            kR = min(range(lb, rb + 1), key=lambda k: rc_ends[k])      # :269
            endRC = rc_ends[kR]
            okR = endRC != INF and endRC < i                      # :271
            if not okF and not okR:                               # :273
                break
            if okF:                                               # :280-287
                if ell > best_fwd_depth or (
                    ell == best_fwd_depth and (jF + ell - 1) < (best_fwd_start + best_fwd_depth - 1)
                ):
                    best_fwd_depth, best_fwd_start, have_fwd = ell, jF, True
            if okR:                                               # :290-299
                posS_R = cst.sa[kR]
                if ell > best_rc_depth or (ell == best_rc_depth and endRC < best_rc_end):
                    best_rc_depth, best_rc_end, best_rc_posS, have_rc = ell, endRC, posS_R, True
        if not have_fwd and not have_rc:                          # :305-316
            out.append((i, 1, i))
            i += 1
            continue
        fwd_true_len = rc_true_len = 0
        if have_fwd:                                              # :322-326
            fwd_true_len = min(cst.lcp_of(i, best_fwd_start), i - best_fwd_start)
        if have_rc:                                               # :328-330
            rc_true_len = cst.lcp_of(i, best_rc_posS)
        use_fwd = use_literal = False                             # :335-352
        if have_fwd and fwd_true_len >= 1:
            use_fwd = not (have_rc and rc_true_len > fwd_true_len)
        else:
            if have_rc and rc_true_len > 1:
                use_fwd = False
            else:
                use_literal = True
        if use_literal:                                           # :354-365
            emit_len, emit_ref = 1, i
        elif use_fwd:
            emit_len, emit_ref = fwd_true_len, best_fwd_start
        else:
            emit_len = rc_true_len
            emit_ref = RC_MASK | (best_rc_end - emit_len + 1)
        assert emit_len > 0                                       # :368
        out.append((i, emit_len, emit_ref))
        i += emit_len                                             # :377-379
    return out


_COMP = {ord("A"): ord("T"), ord("C"): ord("G"), ord("G"): ord("C"), ord("T"): ord("A")}


def sentinel_byte(index: int) -> int:
    """factorizer.cpp:110-125 -- index-th byte of 1..255 (wrapping) that is not A/C/G/T."""
    s, count = 1, 0
    while True:
        if s not in (0, 65, 67, 71, 84):
            if count == index:
                return s
            count += 1
        s = (s + 1) & 0xFF
        if s == 0:
            s = 1


def _validate_and_upper(sequences, limit, what):
    if not sequences:
        return None
    non_empty = [s for s in sequences if len(s)]
    if not non_empty:
        raise RuntimeError("All sequences are empty - cannot prepare for factorization")
    if len(non_empty) > limit:
        raise ValueError(
            f"Too many sequences: maximum {limit} sequences supported (due to sentinel character limitations)"
        )
    for idx, s in enumerate(sequences):
        for c in s:
            if c not in b"ACGTacgt":
                raise RuntimeError(f"Invalid nucleotide '{chr(c)}' found in sequence {idx}")
    return [bytes(s).upper() for s in non_empty]


def prepare_multiple_dna_sequences_w_rc(sequences):
    """factorizer.cpp:54-172 -> (prepared bytes, original_length, sentinel_positions)."""
    seqs = _validate_and_upper(sequences, 125, "w_rc")
    if seqs is None:
        return b"", 0, []
    out = bytearray()
    sent = []
    idx = 0
    for s in seqs:
        out += s
        sent.append(len(out))
        out.append(sentinel_byte(idx))
        idx += 1
    original_length = len(out)
    for s in reversed(seqs):
        out += bytes(_COMP[c] for c in reversed(s))
        sent.append(len(out))
        out.append(sentinel_byte(idx))
        idx += 1
    return bytes(out), original_length, sent


def prepare_multiple_dna_sequences_no_rc(sequences):
    """factorizer.cpp:194-294."""
    seqs = _validate_and_upper(sequences, 250, "no_rc")
    if seqs is None:
        return b"", 0, []
    out = bytearray()
    sent = []
    for k, s in enumerate(seqs):
        out += s
        if k + 1 < len(seqs):
            sent.append(len(out))
            out.append(sentinel_byte(k))
    return bytes(out), len(out), sent


def factorize(data: bytes, start_pos: int = 0):
    return nolzss(bytes(data), start_pos)


def factorize_dna_w_rc(data: bytes):
    """factorizer_core.hpp:140-151 + bindings.cpp:226 tuple shape (start, len, ref&~MASK, is_rc)."""
    if len(data) == 0:
        return []
    S, _, _ = prepare_multiple_dna_sequences_w_rc([bytes(data)])
    return [(s, l, r & ~RC_MASK, bool(r & RC_MASK)) for (s, l, r) in nolzss_multiple_dna_w_rc(S)]

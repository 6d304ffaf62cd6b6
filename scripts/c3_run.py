"""configs[2]: 10 000 records x 10 kbp multi-record FASTA, per-sequence RC factorization (count + files)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nolzss_b200 import _lib as L, _noLZSS as ext, workloads as wl

nrec = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
recs = wl.c3_records(nrec, 10_000, seed=3)
d = tempfile.mkdtemp()
fa = os.path.join(d, "c3.fasta")
with open(fa, "wb") as f:
    for rid, s in recs:
        f.write(b">" + rid.encode() + b"\n" + s + b"\n")
nb = sum(len(s) for _, s in recs)
ext.count_factors_fasta_dna_w_rc_per_sequence(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "two_records.fasta"))
for threads in (1, 4, 8, 16):
    t0 = time.perf_counter()
    total = ext.parallel_write_factors_binary_file_fasta_dna_w_rc_per_sequence(fa, os.path.join(d, f"out{threads}"), threads)
    dt = time.perf_counter() - t0
    print(f"threads={threads}: {nrec} records, {nb/1e6:.0f} Mbp, {total} factors, {dt:.2f} s -> {nb/dt/1e6:.1f} Mbases/s (incl. FASTA parse + {nrec} files)", flush=True)
t0 = time.perf_counter()
counts, ids, total2 = ext.count_factors_fasta_dna_w_rc_per_sequence(fa)
print(f"count only (8 threads): {time.perf_counter()-t0:.2f} s, total {total2}")
assert total2 == total
st = L.stats()
print("last batch device stages:", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
# device-only view of the same batch (records already parsed)
t0 = time.perf_counter()
_, cnts = L.factorize_batch([s for _, s in recs], True, want_factors=False)
dt = time.perf_counter() - t0
st = L.stats()
print(f"nlz_factorize_batch count-only: wall {dt*1e3:.0f} ms, device {st['ms_total']:.1f} ms -> {nb/st['ms_total']/1e3:.1f} Mbases/s on the device; "
      f"rounds={st['doubling_rounds']} key_syms={st['key_syms']} stages: prep {st['ms_prepare']:.1f} keys {st['ms_keys']:.1f} sort0 {st['ms_sort0']:.1f} "
      f"doubling {st['ms_doubling']:.1f} lcp {st['ms_lcp']:.1f} lpnf {st['ms_lpnf']:.1f} chain {st['ms_chain']:.1f}")
assert int(cnts.sum()) == total

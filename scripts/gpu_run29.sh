cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=60
python - <<'PY'
import numpy as np, os
from nolzss_b200 import _lib as L, workloads as wl
lib=L.load(); L.context(0)
p="/dev/shm/nlz_reg_test.npy"
mm=np.lib.format.open_memmap(p, mode="w+", dtype=np.uint8, shape=(50_000_000,)); mm[:]=wl.uniform_dna(50_000_000, 3); mm.flush(); del mm
for mode in ("r", "r+"):
    t=np.load(p, mmap_mode=mode)
    base=t.ctypes.data; a0=base//4096*4096; a1=(base+len(t)+4095)//4096*4096
    rc=lib.nlz_host_register(a0, a1-a0)
    print(mode, "register rc", rc, lib.nlz_last_error() if rc else b"")
    if rc==0:
        import time; t0=time.perf_counter(); z=L.count(L.MODE_DNA_RC, t); dt=time.perf_counter()-t0
        print("count", z, "ms_prepare", L.stats()["ms_prepare"], "wall", round(dt,3))
        print("unregister", lib.nlz_host_unregister(a0))
    del t
os.unlink(p)
PY
python -m pytest tests/test_gpu_dist.py -x -q -k "device_resident" 2>&1 | tail -2

"""Runs the reference's OWN test files, unmodified, against this repo: `noLZSS` (and every submodule) is aliased to
`nolzss_b200` by an import hook, so `import noLZSS._noLZSS as cpp` binds the B200 shim.

    python scripts/run_reference_tests.py --stage     # in the authoring container: copy /root/reference/tests into
                                                      # scratch_ab/ref_suite (git-ignored; it travels with gpurun)
    python scripts/run_reference_tests.py             # on the GPU box: run every file, one pytest process each;
                                                      # results -> gpurun_out/ref_suite_results.json + .md

The reference's test sources are never committed (they stay under the git-ignored scratch_ab/)."""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, "scratch_ab", "ref_suite")
REF_TESTS = "/root/reference/tests"

CONFTEST = '''
import importlib, importlib.abc, importlib.util, os, sys
ROOT = %r
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

class _Alias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """noLZSS[.x] -> nolzss_b200[.x]; the same module object under both names."""
    def find_spec(self, name, path=None, target=None):
        if name != "noLZSS" and not name.startswith("noLZSS."):
            return None
        real = "nolzss_b200" + name[len("noLZSS"):]
        try:
            importlib.import_module(real)
        except ImportError:
            return None
        return importlib.util.spec_from_loader(name, self, is_package=hasattr(sys.modules[real], "__path__"))
    def create_module(self, spec):
        return sys.modules["nolzss_b200" + spec.name[len("noLZSS"):]]
    def exec_module(self, module):
        pass

sys.meta_path.insert(0, _Alias())
'''


def stage():
    if os.path.isdir(SUITE):
        shutil.rmtree(SUITE)
    os.makedirs(SUITE)
    shutil.copytree(REF_TESTS, os.path.join(SUITE, "tests"))
    with open(os.path.join(SUITE, "conftest.py"), "w") as f:
        f.write(CONFTEST % ROOT)
    print("staged", SUITE)


def run():
    if not os.path.isdir(os.path.join(SUITE, "tests")):
        print("nothing staged under", SUITE)
        return 1
    with open(os.path.join(SUITE, "conftest.py"), "w") as f:      # ROOT differs on the GPU box
        f.write(CONFTEST % ROOT)
    files = sorted(f for f in os.listdir(os.path.join(SUITE, "tests")) if f.startswith("test_") and f.endswith(".py"))
    results = {}
    for fn in files:
        p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--tb=line",
                            os.path.join("tests", fn)], cwd=SUITE, capture_output=True, text=True, timeout=1800)
        tail = (p.stdout or "").strip().splitlines()
        summary = tail[-1] if tail else ""
        counts = {k: int(v) for v, k in re.findall(r"(\\d+) (passed|failed|skipped|error|errors|xfailed|xpassed)", summary)}
        fails = [ln for ln in tail if ln.startswith(("FAILED", "ERROR")) or " Error" in ln or "Error:" in ln][:12]
        results[fn] = {"rc": p.returncode, "summary": summary, "counts": counts, "failures": fails}
        print(fn, "->", summary, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_suite_results.json"), "w") as f:
        json.dump(results, f, indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "ref_suite_results.md"), "w") as f:
        f.write("| reference test file | result | notes |\n|---|---|---|\n")
        for fn, r in results.items():
            f.write(f"| `tests/{fn}` | {r['summary']} | {'; '.join(r['failures'][:3])} |\n")
    return 0


if __name__ == "__main__":
    sys.exit(stage() if "--stage" in sys.argv else run())

"""One-off soak: the randomised differential test of tests/test_gpu_fuzz.py on more seeds (needs a GPU)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("fz", os.path.join(ROOT, "tests", "test_gpu_fuzz.py"))
fz = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fz)
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    try:
        fz.test_random_pass_against_oracle(seed)
    except Exception as e:      # noqa: BLE001
        bad += 1
        print("seed", seed, "FAILED:", type(e).__name__, str(e)[:200])
print("seeds %d..%d: %d failures" % (lo, hi - 1, bad))

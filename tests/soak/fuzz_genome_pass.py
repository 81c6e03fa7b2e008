"""One-off soak: the streaming K4 against the general kernel (tests/test_gpu_genome_pass.py::test_split_k4_matches_the_direct_kernel)
and exact mode against speculative mode on more seeds (needs a GPU).

    python tests/soak/fuzz_genome_pass.py LO HI
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pytest

spec = importlib.util.spec_from_file_location("gpt", os.path.join(ROOT, "tests", "test_gpu_genome_pass.py"))
gpt = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gpt)


def exact_vs_speculative(seed):
    """test_exact_mode_equals_speculative_mode, except that a fit whose guard trips by itself (spline above 1/16: both runs
    are exact-mode runs) counts as a skipped case and not as a failure."""
    import torch
    from blueberry_b200.distributed import GenomePass
    dev = torch.device("cuda", 0)
    eng, shards = gpt._random_shards(seed, dev)
    gp = GenomePass(eng, group=False, q_values=True)
    gp.attach(shards)
    gp.run()
    if gp.last_score.exact:
        pytest.skip("guard tripped")
    p_spec, q_spec = gp.p.clone(), gp.q.clone()
    gp.p.fill_(7.0); gp.q.fill_(7.0)
    gp.force_exact = True
    gp.run()
    assert gp.last_score.exact == 1
    assert gpt._same(gp.p.cpu().numpy()[:gp.rows], p_spec.cpu().numpy()[:gp.rows])
    assert gpt._same(gp.q.cpu().numpy()[:gp.rows], q_spec.cpu().numpy()[:gp.rows])


lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = skipped = 0
for seed in range(lo, hi):
    for what, fn in (("split", lambda: gpt.test_split_k4_matches_the_direct_kernel(seed, False)),
                     ("exact", lambda: exact_vs_speculative(seed))):
        try:
            fn()
        except pytest.skip.Exception:
            skipped += 1
        except (ZeroDivisionError, ValueError) as e:       # degenerate random case: the fit raises, as the reference would
            skipped += 1
        except Exception as e:      # noqa: BLE001
            bad += 1
            print("seed", seed, what, "FAILED:", type(e).__name__, str(e)[:200])
print("seeds %d..%d: %d failures, %d degenerate cases skipped" % (lo, hi - 1, bad, skipped))

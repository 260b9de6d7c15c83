"""Randomised differential run: device sampler (both RNG modes) against the oracle over random dataset layouts, sampler
settings, classes and batch sizes (tests/fuzz_util.py).  usage: python scratch/fuzz.py [n_cases] [seed]
FUZZ_ORACLE_ONLY=1 exercises the generator and the oracle alone (no GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.fuzz_util import run_case

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
master = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
oracle_only = bool(os.environ.get('FUZZ_ORACLE_ONLY'))
failures, t0 = 0, time.time()
for case in range(n_cases):
    seed = int(master.integers(0, 2**31))
    try:
        run_case(np.random.default_rng(seed), oracle_only)
    except Exception as exc:
        failures += 1
        print(f'FAIL case {case} (generator seed {seed}): {type(exc).__name__}: {str(exc)[:1500]}')
print(f'{n_cases} cases, {failures} failures, {time.time() - t0:.1f} s')
sys.exit(1 if failures else 0)

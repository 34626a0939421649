"""`-m gpu`: BASELINE.json's configurations at their FULL size (1e9 rows on one GPU), checked against the oracle over ALL rows —
the raw output buffer bit for bit, floating-point SUM within 1e-6 (baseline-hash results as rows sorted by key).  The rows are never
materialised on the host: the oracle regenerates the counter-based columns slab by slab (oracle_execute_generated), the device
generates the same values in HBM (b2q_gen_column) — the machinery `bench.py` uses for the `parity_check` of its timed result.
The 32-bit low-word shared-memory accumulators, the split (lo | hi) L2 accumulators and the radix passes are exactly the code that
can be right at 3e6 rows and wrong at 1e9 (carries, 4-byte COUNT slots, region capacities)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", ["c2", "c2all", "c3", "c4", "c4s"])
def test_full_size_configuration_matches_the_oracle(cfg):
    import torch

    import bench
    free_b, _ = torch.cuda.mem_get_info()
    rows = 1_000_000_000
    if free_b < 60 * 2**30:       # a smaller device: the check still runs, on what fits
        rows = 250_000_000
    r = bench.Runner(cfg, rows, 0, 1, None, torch)
    try:
        rs = r.step()
        out = bench.parity_check(cfg, r.unit, rs, r.all_frags(), r.guess)
        assert out["ok"], out
        assert out["rows"] == rows
        del rs
    finally:
        r.free(torch)

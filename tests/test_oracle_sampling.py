"""Pin the index-stream oracle: (1) against CPython's own ``random`` (the reference's
dependency), (2) against what the real reference ReplayBuffer returned (golden/sampling.json)."""
import json
import os
import random

import numpy as np
import pytest

from oracle.mt_sample import MT19937, logical_to_slot


@pytest.mark.parametrize("seed", [0, 1, 12345, 2**40 + 7])
def test_mt_matches_stdlib(seed):
    for n, k in [(800, 256), (5000, 256), (5000, 1024), (1_000_000, 256), (1045, 256), (1046, 256), (300, 256), (7, 3), (5, 5), (9, 0)]:
        random.seed(seed)
        m = MT19937(seed)
        for _ in range(2):
            assert random.sample(range(n), k) == m.sample_indices(n, k)


def test_sample_rejects_oversize():
    with pytest.raises(ValueError):
        MT19937(0).sample_indices(5, 6)


def test_reference_buffer_golden(golden_dir):
    with open(os.path.join(golden_dir, "sampling.json")) as f:
        g = json.load(f)
    assert g["underfilled_raises"] == "ValueError"
    for c in g["cases"]:
        m = MT19937(c["seed"])
        n = min(c["pushes"], c["capacity"])
        oldest = max(c["pushes"] - c["capacity"], 0)
        for ids in c["push_ids"]:
            logical = m.sample_indices(n, c["k"])
            # deque position j holds push number oldest + j  (F3)
            assert [oldest + j for j in logical] == ids
            slots = logical_to_slot(np.array(logical), c["pushes"], c["capacity"])
            assert np.array_equal(slots, np.array(ids) % c["capacity"])

"""A short run of the option-space fuzz (tools/option_fuzz.py): random option sets on seeded reads of every shape, the
oracle against the unmodified reference binary.  Skipped where oracle/_ref/TideHunter does not exist (the GPU box has
it, the repository history does not)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [3, 4])
def test_oracle_matches_reference_on_random_option_sets(seed):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "TideHunter")):
        pytest.skip("oracle/_ref/TideHunter not built (needs /root/reference)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "option_fuzz.py"), "14", str(seed)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    assert r.returncode == 0, r.stdout.decode()[-3000:]

"""Device leg of the differential fuzz (tests/golden/fuzz_scenarios.py: 478 aggregate + 100 select / search queries the
real reference answered). Its CPU legs are green — the oracle (test_oracle_golden.py), the planner's programs row by row
(test_planner_fuzz.py), the host post-aggregation in Python and in C++ (test_host_fuzz.py, test_adapter_mock.py).

This leg was written after the round's GPU budget was spent: it had NEVER run on a B200 when it was committed. It is
therefore marked xfail(strict=False) — XPASS in the report means the device reproduces the reference on the whole fuzz
corpus, XFAIL means a kernel / table-layout finding (the JSON line of the runner names the records) — and each part runs
in a process of its own (tests/gpu_fuzz_runner.py), last in the suite, so that nothing it does can touch the other tests."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.xfail(strict=False, reason="first run of the fuzz corpus on a B200 (committed without GPU minutes): XPASS == green")
@pytest.mark.parametrize("which", ["aggregate", "aggregate_forced_hash", "select_search"])
def test_device_reproduces_the_reference_on_the_fuzz_corpus(which, built_lib):
    p = subprocess.run([sys.executable, os.path.join(HERE, "gpu_fuzz_runner.py"), which], capture_output=True, text=True, timeout=240)   # bounded: ~75 s expected per part
    print(p.stdout[-3000:])
    print(p.stderr[-1500:], file=sys.stderr)
    assert p.returncode == 0, (p.returncode, p.stdout[-1500:], p.stderr[-800:])

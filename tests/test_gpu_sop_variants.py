"""GPU suite: the build-time / run-time variants of the sum-of-products kernel that are not the default, so that an A/B
switch can never ship an unchecked kernel.  Each variant re-runs tests/test_gpu_sop.py in a child process (the knobs are
read once per process)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env", [{"ZK_B200_SOP_WIDE": "1"}, {"ZK_B200_SOP_GROUP": "0", "ZK_B200_SOP_SCHED": "static"}, {"ZK_B200_SOP_FOLD_PIPE": "f64"},
                                  {"ZK_B200_SOP_TOOM": "0"}],
                         ids=["deferred-reduction", "no-grouping-static-split", "fp64-folds", "points-0-to-3"])
def test_sum_of_products_kernel_variant(env):
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_gpu_sop.py")],
                       capture_output=True, text=True, env=dict(os.environ, **env), timeout=900, cwd=ROOT)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail, r.stdout[-2500:] + r.stderr[-1500:]

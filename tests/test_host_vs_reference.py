"""CPU, authoring container only: the device-independent host helpers of this package
(zdataset, renormalize, the rewriter's crop / paste geometry, zca_from_cov, nethook
subsequence / InstrumentedModel, FixedSubsetSampler) against the LIVE reference on seeded random
inputs.  Runs tests/_ref_host_compare.py in a subprocess because the reference shim
monkey-patches torch; skipped where /root/reference does not exist (the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not os.path.isdir('/root/reference/rewrite'),
                    reason='needs a checkout of the reference (authoring container only)')
def test_host_helpers_equal_the_live_reference():
    r = subprocess.run([sys.executable, '-W', 'ignore', os.path.join(HERE, '_ref_host_compare.py')],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')][-1]
    res = json.loads(line[len('RESULT '):])
    flags = {'renorm_url_roundtrip', 'subsequence_names_equal', 'subsequence_shares_weights', 'sampler'}
    for name, val in res.items():
        if name in flags:
            assert val == 1.0, name
        else:
            assert val == 0.0, (name, val)             # bit-identical on every check
    assert len(res) >= 19

"""The opt-in forms of the fused train step (measured-negative experiments kept behind ADER_B200_* switches) produce
the SAME BITS as the default: every switch is read once per process, so each form runs in its own interpreter
(tests/variant_probe.py) and the digests of losses, parameters, Adam slots, gradient and step counter are compared."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
VARIANTS = {
    "split_table_adam": {"ADER_B200_SPLIT_ADAM": "1"},          # k_adam_untouched beside the scatter + k_adam_touched behind it
    "standalone_drep_reduction": {"ADER_B200_FUSE_DREP": "0"},   # k_reduce_drep + k_lnf_bwd instead of k_lnf_bwd_drep
    "shared_memory_scatter": {"ADER_B200_SCATTER": "2"},         # k_scatter_apply2
    "single_cta_packing": {"ADER_B200_PACK": "1"},               # k_pack_small
}
SWITCHES = sorted({k for v in VARIANTS.values() for k in v})


def _digest(extra):
    env = {k: v for k, v in os.environ.items() if k not in SWITCHES}
    env.update(extra)
    r = subprocess.run([sys.executable, os.path.join(HERE, "variant_probe.py")], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("DIGEST ")]
    assert len(lines) == 1, r.stdout[-2000:]
    return lines[0].split()[1]


@pytest.fixture(scope="module")
def default_digest():
    a, b = _digest({}), _digest({})
    assert a == b, "the default step is not run-to-run bit-identical"
    return a


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_opt_in_form_is_bit_identical_to_the_default(name, default_digest):
    assert _digest(VARIANTS[name]) == default_digest

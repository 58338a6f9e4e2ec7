"""The reference's OWN model tests (tests/models/*_test.py, unmodified) run on the GPU against this package:
`deepctr_torch` is aliased to `deepctr_torch_b200` by tests/ref_alias_plugin.py.  They build every model with
the reference's fixtures (VarLen features, dropout 0.5, all branch toggles), call compile / fit with callbacks /
save + load of weights and of the pickled module.  Skipped when baseline/_ref/reftests is absent
(scripts/install_reference.sh creates it in the build container; it travels to the GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFTESTS = os.path.join(REPO, "baseline", "_ref", "reftests")
FILES = ["DeepFM", "xDeepFM", "FiBiNET", "DCN", "DCNMix", "WDL", "NFM", "AFM", "IFM", "DIFM"]


@pytest.mark.parametrize("model", FILES)
def test_reference_model_test_file_passes(model, tmp_path):
    path = os.path.join(REFTESTS, "tests", "models", model + "_test.py")
    if not os.path.exists(path):
        pytest.skip("reference tests not installed (scripts/install_reference.sh)")
    env = dict(os.environ, TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD="1", PYTHONDONTWRITEBYTECODE="1",
               PYTHONPATH=os.pathsep.join([os.path.join(REPO, "tests"), REPO, REFTESTS]))
    cmd = [sys.executable, "-m", "pytest", "-p", "ref_alias_plugin", "-p", "no:cacheprovider", "-q", "-x",
           "--rootdir", REFTESTS, path]
    r = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + "\n" + r.stderr)[-4000:]

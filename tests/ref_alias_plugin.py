"""pytest plugin (-p ref_alias_plugin): makes `import deepctr_torch` resolve to THIS repository's package, so that the
reference's own test files (baseline/_ref/reftests, copied unmodified by scripts/install_reference.sh) exercise the
B200 implementation through the reference's public API (SURVEY §4 / VERDICT r1 item 9)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import deepctr_torch_b200  # noqa: E402
from deepctr_torch_b200 import callbacks, inputs, layers, models  # noqa: E402

sys.modules["deepctr_torch"] = deepctr_torch_b200
for name, mod in (("inputs", inputs), ("models", models), ("layers", layers), ("callbacks", callbacks)):
    sys.modules["deepctr_torch." + name] = mod

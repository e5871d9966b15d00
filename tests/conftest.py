import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "vosk-api_b200"), os.path.join(ROOT, "vosk-api_b200", "tools"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def model_root(tmp_path_factory):
    """Directory holding generated synthetic model dirs, keyed by architecture (generated on demand)."""
    import vbmodel
    base = os.environ.get("VB_MODEL_CACHE") or str(tmp_path_factory.mktemp("models"))
    made = {}

    def get(arch="tiny", seed=0, overrides=None, tag=""):
        key = f"{arch}_{seed}{tag}"
        if key not in made:
            root = os.path.join(base, key)
            if not os.path.exists(os.path.join(root, "model", "graph", "HCLG.fst")):
                vbmodel.write_model_dir(root, arch, seed, overrides=overrides)
            made[key] = os.path.join(root, "model")
        return made[key]

    return get


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.lib()
    return oracle

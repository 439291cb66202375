import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def wfx():
    # a fresh checkout has no libwavefx.so (built artefacts are not tracked): compile it first --
    # nvcc cross-compiles sm_100a without a GPU, same as __graft_entry__.build()
    lib = os.environ.get("WFX_LIB") or os.path.join(ROOT, "wave-fenics_b200", "libwavefx.so")
    if not os.path.exists(lib):
        import importlib.util
        spec = importlib.util.spec_from_file_location("wfx_build", os.path.join(ROOT, "wave-fenics_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    import wave_fenics_b200
    return wave_fenics_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    return oracle

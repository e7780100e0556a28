import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The tests go through the built artefacts (libwrfb200.so, oracle/liboracle.so).  In a fresh checkout they
    # do not exist yet (built files are git-ignored): build them once, exactly as the driver's build check does.
    lib = os.path.join(ROOT, "wrf_model_cuda_sample_b200", "libwrfb200.so")
    ora = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(ora)):
        import __graft_entry__
        __graft_entry__.build()

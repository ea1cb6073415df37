import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# a protocol bug in a halo mode must fail a test within seconds, not hold the GPU box for the library's production limit of two minutes
# per device-side wait (TMQ_OPT_HALO_TIMEOUT_MS)
import os
os.environ.setdefault("TMQ_HALO_TIMEOUT_MS", "10000")

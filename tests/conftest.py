import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container)")


def pytest_sessionstart(session):
    """The tests need the in-tree library (host builders and probes on CPU, the engine on the
    GPU): build it if it is missing or older than its sources (no-op otherwise; concurrent
    sessions serialise on the build lock)."""
    import __graft_entry__ as ge

    ge.build()

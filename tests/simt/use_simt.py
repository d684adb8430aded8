"""TEST INFRASTRUCTURE ONLY — makes axiomr_b200.api talk to the SIMT-interpreter build of the CUDA sources (tests/simt).

The product loader (axiomr_b200.api.load_library) refuses that build unconditionally; the test harness therefore replaces the
loader, in the test process only, with one that binds whatever library api.LIB_PATH (or AXR_B200_LIB) names. Called from
tests/conftest.py when AXR_SIMT_TESTS_ONLY=1 and from the fuzz drivers in this directory; nothing under axiomr_b200/, bench.py or
__graft_entry__.py imports this module.
"""
import ctypes as C
import os


def install():
    from axiomr_b200 import api

    def load_library():
        if api._lib is None:
            api._lib = api._bind(C.CDLL(api.LIB_PATH))
        return api._lib

    if os.environ.get("AXR_B200_LIB"):
        api.LIB_PATH = os.environ["AXR_B200_LIB"]
    api.load_library = load_library
    api._lib = None

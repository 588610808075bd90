"""The C-ABI library loads and exports every symbol include/flybody_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as ge
from flybody_b200 import stepper

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'flybody_b200.h')).read()
    return sorted(set(re.findall(r'\b(fb_[a-z_]+)\s*\(', hdr)))


def test_build_and_exports():
    ge.build()
    lib = ctypes.CDLL(stepper.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in the header but not exported'
    assert sorted(stepper.EXPORTS) == syms
    assert b'sm_100a' in ctypes.c_char_p(ctypes.cast(lib.fb_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()).value


def test_product_path_has_no_cpu_fallback(tmp_path):
    with pytest.raises(stepper.StepperError):
        stepper.load_library(str(tmp_path / 'missing.so'))
    # the product package never references the oracle or the emulation build
    pkg = os.path.join(ROOT, 'flybody_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.h')) and 'compiler' not in dp:
                src = open(os.path.join(dp, f)).read()
                assert 'fly_oracle' not in src and 'libflyoracle' not in src, f
                if f.endswith('.py'):
                    assert 'libfb_emu' not in src, f

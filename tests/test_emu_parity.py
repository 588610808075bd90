"""CPU checks of the *kernel source* (host-emulation build, tests/_emu) against the oracle.
These are host-logic tests; the parity tests proper are tests/test_gpu_parity.py on the B200."""
import os

import numpy as np
import pytest

import __graft_entry__ as ge
from flybody_b200 import stepper as st
from flybody_b200.flymodel import load_model
from oracle import fly_oracle as fo
from conftest import walk_reset_qpos
from parity_common import (compare_stage_fields, teacher_forced_errors, summarize_tf, STAGE_TOL)


@pytest.fixture(scope='module')
def emu():
    ge.build()
    return ge.EMU


def test_stage_parity_walk(emu):
    m = load_model('walk')
    compare_stage_fields(m, st.BatchedStepper(m, 2, lib_path=emu), seed=0)


def test_stage_parity_flight(emu):
    m = load_model('flight')
    compare_stage_fields(m, st.BatchedStepper(m, 2, lib_path=emu), seed=1, vel_scale=20.0)


def test_teacher_forced_control_steps_walk(emu):
    m = load_model('walk')
    r = summarize_tf(*teacher_forced_errors(m, st.BatchedStepper(m, 1, lib_path=emu), n_steps=8, n_sub=10))
    assert r['p90_q'] < 2e-6 and r['p90_v'] < 5e-3 and r['events'] <= 1, r


@pytest.mark.parametrize('kb', ['0', '24'])
def test_solver_global_memory_fallback_paths(emu, kb, monkeypatch):
    """The solver keeps its vectors / Delassus matrix / Hessian factor in shared memory when they fit and
    falls back to the global arrays otherwise; both placements must give the same answer."""
    monkeypatch.setenv('FB_SOLVE_SMEM_KB', kb)
    m = load_model('walk')
    compare_stage_fields(m, st.BatchedStepper(m, 2, lib_path=emu), seed=0)

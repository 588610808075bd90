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


@pytest.mark.parametrize('ncap,seed', [('0', 0), ('32', 4)])
def test_solver_shared_and_global_memory_paths(emu, ncap, seed, monkeypatch):
    """The solver keeps an env's working set in its warp's shared-memory slice when nefc <= 32 and runs the
    same code on the global arrays otherwise; both placements must agree with the oracle.
    (seed 0 has nefc = 53 -> global path even with the default cap; seed 4 with pos_scale 0 has nefc = 18.)"""
    monkeypatch.setenv('FB_SOLVE_NCAP', ncap)
    m = load_model('walk')
    compare_stage_fields(m, st.BatchedStepper(m, 2, lib_path=emu), seed=seed, pos_scale=0.1 if seed == 0 else 0.0)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box)')


def walk_reset_qpos(m):
    """Config-1 reset state: qpos0 with wings at their springrefs."""
    q0 = m.qpos0.copy()
    for side in ('left', 'right'):
        for dof, val in (('yaw', 1.5), ('roll', 0.7), ('pitch', -1.0)):
            q0[m.jnt_qposadr_of(f'walker/wing_{dof}_{side}')] = val
    return q0

"""The PyTorch ports of the reference's actor networks (flybody_b200/policy_torch.py): shapes and the structural facts the reference
states about them (agents/network_factory.py:66-109, agents/network_factory_vis.py:141-293).  CPU; random weights."""
import numpy as np
import torch

from flybody_b200 import policy_torch as pt


def test_dmpo_policy_head():
    p = pt.DMPOPolicy(741, 59).eval()
    obs = torch.randn(5, 741)
    mean, scale = p.distribution(obs)
    assert mean.shape == scale.shape == (5, 59) and bool((scale > 0).all())
    assert torch.equal(p(obs, deterministic=True), mean)
    # MultivariateNormalDiagHead: a zero pre-activation gives scale = init_scale (softplus(0) * 0.7 / log 2)
    with torch.no_grad():
        p.scale.weight.zero_(); p.scale.bias.zero_()
    assert torch.allclose(p.distribution(obs)[1], torch.full((5, 59), 0.7), atol=1e-5)
    n = sum(x.numel() for x in p.parameters())
    assert n == 741 * 256 + 256 + 2 * 256 + 2 * (256 * 256 + 256) + 2 * (256 * 59 + 59)


def test_visnet_geometry_and_layout():
    v = pt.VisNet(vis_output_dim=8).eval()
    left = torch.randint(0, 255, (3, 32, 32, 3), dtype=torch.uint8); right = torch.randint(0, 255, (3, 32, 32, 3), dtype=torch.uint8)
    task, others = torch.randn(3, 2), torch.randn(3, 62)
    out = v(left, right, task, others)
    assert out.shape == (3, 2 + 8 + 62)
    assert torch.equal(out[:, :2], task) and torch.equal(out[:, 10:], others)          # [task_input, vis, rest] (network_factory_vis.py:207-211)
    assert v.head.in_features == 16 * 6 * 6                                             # 32 -> 30 -> 28 -> 13 -> 6 (VALID, strides 1 1 2 2)
    gray = left.float().mean(-1)                                                        # RGB -> one channel, then (x - 77) / 56
    assert torch.allclose(v(gray, right.float().mean(-1), task, others), out, atol=1e-5)


def test_two_level_controller_splices_the_steering_command():
    n_others, n_act = 62, 12
    c = pt.TwoLevelController(n_others, n_act, steering_idx=n_others).eval()
    assert c.steering_dim == 42 and c.offset == 10
    x = torch.randn(4, 10 + n_others)
    a = c(x, deterministic=True)
    assert a.shape == (4, n_act)
    # on initialisation the high-level output is close to the no-op steering command (zero displacement, identity quaternion)
    steer = c.hl(x) + c.ballpark
    assert float((steer - c.ballpark).detach().abs().max()) < 0.5
    assert all(not p.requires_grad for p in c.ll.parameters()) and any(p.requires_grad for p in c.hl.parameters())
    # the low-level controller sees [others[:idx], steering, others[idx:]]
    c2 = pt.TwoLevelController(n_others, n_act, steering_idx=20).eval()
    seen = {}
    c2.ll.torso.first.register_forward_hook(lambda m, inp, out: seen.update(x=inp[0]))
    c2(x, deterministic=True)
    assert torch.equal(seen['x'][:, :20], x[:, 10:30]) and torch.equal(seen['x'][:, 62:], x[:, 30:])

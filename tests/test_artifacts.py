"""Run artefacts in the reference's formats (checkpoint layout, action traces, nMAE) -- CPU tests;
the trace replay itself needs the GPU."""
import os

import numpy as np
import pytest
import torch
from torch import nn

from f16_model_b200 import artifacts


class _Pol:
    def __init__(self):
        mk = lambda: nn.Sequential(nn.Linear(5, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, 1))  # noqa: E731
        self.actor_mean, self.critic = mk(), mk()
        self.actor_logstd = nn.Parameter(torch.full((1,), -0.3))


def test_checkpoint_layout_matches_agent_save_and_roundtrips(tmp_path):
    a, b = _Pol(), _Pol()
    path = artifacts.save_agent(a, str(tmp_path), "run1")
    assert sorted(os.listdir(path)) == ["actor_logstd", "actor_mean", "critic", "gsde_logstd", "gsde_mean"]
    # what the reference's Agent.load does (control/ppo/ppo_model_gsde.py:117-132)
    actor = torch.load(os.path.join(path, "actor_mean"), map_location="cpu", weights_only=False)
    assert isinstance(actor, nn.Sequential) and actor[0].in_features == 5 and actor[4].out_features == 1
    assert torch.load(os.path.join(path, "actor_logstd"), weights_only=False).shape == (1, 1)
    artifacts.load_agent(b, path)
    x = torch.randn(7, 5)
    assert torch.equal(a.actor_mean(x), b.actor_mean(x)) and torch.equal(a.critic(x), b.critic(x))
    assert torch.equal(a.actor_logstd, b.actor_logstd)


def test_action_trace_format(tmp_path):
    tr = artifacts.ActionTrace("runX", log_dir=str(tmp_path))
    init = np.arange(9, dtype=float)
    tr.start(init)
    for a in (0.1, -0.25, 1.0):
        tr.action(np.array([a]))
    lines = open(tr.path).read().splitlines()
    assert lines[0] == str([float(v) for v in init]) and lines[1] == "[0.1]" and len(lines) == 4   # state_logger's format (numpy 1.26)
    open(tr.path, "a").write("[np.float64(0.5)]\n")                                      # ... and as numpy >= 2 prints it
    assert artifacts.read_trace(tr.path)[1][-1] == 0.5
    open(tr.path, "w").write("\n".join(lines) + "\n")
    i2, acts = artifacts.read_trace(tr.path)
    assert np.array_equal(i2, init) and np.array_equal(acts, [0.1, -0.25, 1.0])


def test_nmae_matches_reference_formula():
    y, p = np.array([1.0, -2.0, 0.5]), np.array([0.5, -1.0, 0.5])
    assert abs(artifacts.nMAE(y, p) - (np.mean(np.abs(y - p)) / np.mean(np.abs(y) + 1e-9))) < 1e-15
    obs = torch.zeros(4, 2, 5)
    w150 = np.radians(150)
    obs[:, :, 1] = 0.5                      # wz = 0
    obs[:, 0, 3], obs[:, 1, 3] = 0.1, 0.2   # constant references
    out = artifacts.episode_nmae(obs)
    assert torch.allclose(out, torch.ones(2), atol=1e-5)      # |ref - 0| / |ref| = 1
    assert abs((0.5 * 2 * w150) - w150) < 1e-12


@pytest.mark.gpu
def test_trace_replay_reproduces_logged_run(tmp_path, golden):
    """Write the golden config-1 episode as a reference-format trace, replay it on the GPU."""
    tr = artifacts.ActionTrace("cfg1", log_dir=str(tmp_path))
    tr.start(golden["traj_x0"][0])
    for a in golden["traj_actions"][:, 0]:
        tr.action(np.array([a]))
    cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "pure_step"}
    states, rewards, done_at = artifacts.replay_trace(tr.path, cfg)
    assert done_at == 1000
    err = np.abs(states - golden["traj_states"][:, 0]) / np.maximum(np.abs(golden["traj_states"][:, 0]), 1e-2)
    assert err.max() < 1e-10 and abs(rewards.sum() - golden["traj_return"][0]) < 1e-8

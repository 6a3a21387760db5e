"""Run artefacts in the reference's formats, so that its evaluation scripts and
notebooks can consume what this stack produces:

* agent checkpoints -- ``Agent.save`` / ``Agent.load`` layout (control/ppo/ppo_model_gsde.py:101-132):
  whole-module pickles ``actor_mean``, ``critic``, ``gsde_mean`` and tensors ``actor_logstd``,
  ``gsde_logstd`` under ``<save_dir>/<run_name>/``;
* action traces -- ``state_logger`` (control/ppo/utils.py:176-182): ``logs/<run>.txt`` with the initial state
  list on line 0 and one action list per following line, and the replay of such a file
  (test/env_test.py:66-103);
* metrics -- ``nMAE`` / ``mae`` (F16model/utils/control_metrics.py:16-24) and the per-episode nMAE of
  ``calculate_episode_nmae`` (control/ppo/utils.py:184-194), batched on the device;
* TensorBoard scalars with the reference's tags (``charts/*``, ``losses/*``; ppo_model_gsde.py:377-391)."""
import ast
import copy
import os

import numpy as np


# ---------------------------------------------------------------- checkpoints
def save_agent(policy, save_dir, run_name):
    """Write a checkpoint the reference's ``Agent.load(path)`` can read."""
    import torch

    path = os.path.join(save_dir, run_name)
    os.makedirs(path, exist_ok=True)
    actor = copy.deepcopy(policy.actor_mean).cpu()
    critic = copy.deepcopy(policy.critic).cpu()
    logstd = policy.actor_logstd.detach().cpu().reshape(1, 1).clone()
    torch.save(logstd, os.path.join(path, "actor_logstd"))
    torch.save(actor, os.path.join(path, "actor_mean"))
    torch.save(critic, os.path.join(path, "critic"))
    # the reference also stores its (unused) gSDE head; write same-shaped placeholders so Agent.load succeeds
    torch.save(torch.zeros(1, 1), os.path.join(path, "gsde_logstd"))
    gsde = copy.deepcopy(actor)
    with torch.no_grad():
        for p in gsde.parameters():
            p.zero_()
    torch.save(gsde, os.path.join(path, "gsde_mean"))
    return path


def load_agent(policy, path_name):
    """Load a checkpoint written by ``save_agent`` or by the reference's ``Agent.save`` into ``policy``."""
    import torch

    dev = policy.actor_logstd.device
    actor = torch.load(os.path.join(path_name, "actor_mean"), map_location=dev, weights_only=False)
    critic = torch.load(os.path.join(path_name, "critic"), map_location=dev, weights_only=False)
    logstd = torch.load(os.path.join(path_name, "actor_logstd"), map_location=dev, weights_only=False)
    with torch.no_grad():
        policy.actor_mean.load_state_dict(actor.state_dict())
        policy.critic.load_state_dict(critic.state_dict())
        policy.actor_logstd.copy_(torch.as_tensor(logstd).reshape(-1)[:1])
    if hasattr(policy, "sync_weights"):
        policy.sync_weights()
    return policy


# ---------------------------------------------------------------- action traces
class ActionTrace:
    """``state_logger``-compatible trace writer: line 0 = initial state, then one action per line."""

    def __init__(self, run_name, log_dir="logs"):
        os.makedirs(log_dir, exist_ok=True)
        self.path = os.path.join(log_dir, run_name + ".txt")

    def start(self, init_state):
        with open(self.path, "w") as f:
            f.write(str([float(v) for v in np.asarray(init_state, dtype=float)]) + "\n")

    def action(self, action):
        with open(self.path, "a") as f:
            f.write(str([float(v) for v in np.atleast_1d(np.asarray(action, dtype=float))]) + "\n")


def read_trace(path):
    """(init_state[9], actions[T]) from a ``logs/<run>.txt`` trace."""
    import re

    with open(path) as f:   # numpy >= 2 prints list items as np.float64(x); the reference's numpy 1.26 prints x
        lines = [re.sub(r"np\.float(?:32|64)\(([^)]*)\)", r"\1", ln.strip()) for ln in f if ln.strip()]
    init = np.array(ast.literal_eval(lines[0]), dtype=np.float64)
    acts = np.array([ast.literal_eval(ln)[0] for ln in lines[1:]], dtype=np.float64)
    return init, acts


def replay_trace(path, config, device="cuda"):
    """Re-fly a logged run (test/env_test.py:66-103): returns states [T, 9], rewards [T], done step."""
    import torch

    from .env import F16VecEnv

    init, acts = read_trace(path)
    T = len(acts)
    env = F16VecEnv(config, 1, device=device, dtype=torch.float64, autoreset=False)
    env.set_state(torch.tensor(init.reshape(9, 1)))
    states, rewards = np.zeros((T, 9)), np.zeros(T)
    done_at = None
    for k in range(T):
        _, r, d, _, _ = env.step(torch.tensor([acts[k]], dtype=torch.float64, device=env.device))
        states[k], rewards[k] = env.state[:, 0].cpu().numpy(), float(r[0])
        if bool(d[0]) and done_at is None:
            done_at = k + 1
    return states, rewards, done_at


# ---------------------------------------------------------------- metrics
def mae(y_true, predictions):
    y_true, predictions = np.array(y_true), np.array(predictions)
    return np.mean(np.abs(y_true - predictions))


def nMAE(y_true, predictions):
    return mae(y_true, predictions) / np.mean(np.abs(y_true) + 1e-9)


def episode_nmae(obs, upto=None):
    """Per-env nMAE between the tracked reference and the pitch rate over a rollout buffer of
    observations [T, N, 5] (``calculate_episode_nmae``: denormalised obs[1] vs obs[3]).  Torch, on device."""
    import torch

    from . import plane

    T = obs.shape[0] if upto is None else upto
    lo, hi = plane.state_bound["wz"]
    wz = obs[:T, :, 1] * (hi - lo) + lo
    ref = obs[:T, :, 3]
    # nMAE(y_true=theta_ref_obs, predictions=theta_obs) with the reference's argument order (utils.py:191)
    return (ref - wz).abs().mean(dim=0) / (ref.abs() + 1e-9).mean(dim=0)


# ---------------------------------------------------------------- TensorBoard
class TensorboardLogger:
    """Callable for ``PPOTrainer.train(log=...)`` writing the reference's scalar tags."""

    TAGS = {"value_loss": "losses/value_loss", "policy_loss": "losses/policy_loss", "entropy": "losses/entropy",
            "old_approx_kl": "losses/old_approx_kl", "approx_kl": "losses/approx_kl", "clipfrac": "losses/clipfrac",
            "explained_variance": "losses/explained_variance", "learning_rate": "charts/learning_rate",
            "rollout_sps": "charts/SPS", "mean_return_per_finished_episode": "charts/episodic_return"}

    def __init__(self, save_dir, run_name, hyperparameters=None):
        from torch.utils.tensorboard import SummaryWriter

        self.writer = SummaryWriter(os.path.join(save_dir, run_name))
        if hyperparameters:
            rows = "\n".join(f"|{k}|{v}|" for k, v in hyperparameters.items())
            self.writer.add_text("hyperparameters", "|param|value|\n|-|-|\n" + rows)

    def __call__(self, stats):
        step = stats.get("global_step", 0)
        for key, tag in self.TAGS.items():
            if key in stats and stats[key] is not None:
                self.writer.add_scalar(tag, stats[key], step)
        self.writer.add_scalar("train/step", step, step)

    def close(self):
        self.writer.close()

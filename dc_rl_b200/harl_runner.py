"""Driving the UNMODIFIED HARL runners with the CUDA env (BASELINE config 5).

`harl.runners.on_policy_base_runner` imports `make_train_env` / `make_eval_env` by name from `harl.utils.envs_tools` and
looks its logger up in `harl.envs.LOGGER_REGISTRY` (reference on_policy_base_runner.py:16-23); `install()` rebinds those
three names to this package's drop-ins -- the runner, algorithm, model and buffer code stays byte-identical.

Two ways to roll out:
  * the runner's own loop (`runner.run()`): per step the actors read numpy buffers, `envs.step` returns numpy arrays,
    `insert` copies them into the numpy buffers -- everything the reference does, with one CUDA launch instead of N pipes;
  * `DeviceRollout`: the same collect -> step -> insert sequence (on_policy_base_runner.py:241-282, 329-503) with the
    observations, actions, values, rewards and masks of the whole episode kept in torch CUDA tensors (`step_torch`), and ONE
    hand-over into the runner's numpy buffers per episode; `compute()` / `train()` then run unmodified.

    python -m dc_rl_b200.harl_runner --harl-root baseline/_ref --n-envs 4096 --episode-length 64 --episodes 3 [--device-rollout]
prints one JSON line with the env-steps/s delivered into the rollout buffers.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install(harl_root, use_vector_logger=True):
    """Makes `harl` importable from `harl_root` (an unmodified checkout) and points the names its runners import at the
    CUDA drop-ins.  Returns the `harl.runners.RUNNER_REGISTRY`."""
    harl_root = os.path.abspath(harl_root)
    for p in (os.path.join(REPO, "baseline", "shims"), harl_root):       # the tensorboardX stand-in only if the real one is absent
        if p not in sys.path:
            sys.path.append(p) if p.endswith("shims") else sys.path.insert(0, p)
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                   # numpy >= 2 dropped the alias the runner uses (on_policy_base_runner.py:52)
    from . import harl_env
    from .logger import SustainDCLogger
    import harl.envs
    import harl.utils.envs_tools as envs_tools
    envs_tools.make_train_env = harl_env.make_train_env
    envs_tools.make_eval_env = harl_env.make_eval_env
    if use_vector_logger:
        harl.envs.LOGGER_REGISTRY["sustaindc"] = SustainDCLogger
    import harl.runners.on_policy_base_runner as base
    base.make_train_env = harl_env.make_train_env
    base.make_eval_env = harl_env.make_eval_env
    try:
        import harl.runners.off_policy_base_runner as off
        off.make_train_env = harl_env.make_train_env
        off.make_eval_env = harl_env.make_eval_env
    except Exception:                     # noqa: BLE001 -- off-policy runners are optional here
        pass
    from harl.runners import RUNNER_REGISTRY
    return RUNNER_REGISTRY


def load_args(harl_root, algo="happo"):
    """(algo_args, env_args) from the reference's YAML files (harl/configs/, configs_tools.py:9-26)."""
    import yaml
    with open(os.path.join(harl_root, "harl", "configs", "algos_cfgs", algo + ".yaml")) as f:
        algo_args = yaml.safe_load(f)
    with open(os.path.join(harl_root, "harl", "configs", "envs_cfgs", "sustaindc.yaml")) as f:
        env_args = yaml.safe_load(f)
    return algo_args, env_args


class DeviceRollout:
    """One on-policy episode collected on the device for an OnPolicy*Runner whose `envs` is a CudaShareVecEnv."""

    def __init__(self, runner):
        import torch
        self.torch = torch
        r = self.r = runner
        self.envs = r.envs
        self.N = r.algo_args["train"]["n_rollout_threads"]
        self.T = r.algo_args["train"]["episode_length"]
        if r.state_type != "EP":
            raise NotImplementedError("DeviceRollout implements the EP state type (the shipped sustaindc configuration)")
        dev = self.dev = torch.device("cuda", self.envs.engine.device)
        N, T, A = self.N, self.T, r.num_agents
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)      # noqa: E731
        self.obs, self.share = z(T + 1, N, A, 26), z(T + 1, N, self.envs.share_observation_space[0].shape[0])
        self.actions, self.logp = z(T, N, A, 1), z(T, N, A, 1)
        self.values, self.rewards, self.masks = z(T, N, 1), z(T, N, 1), torch.ones(T + 1, N, 1, device=dev)
        self.rnn = z(N, r.recurrent_n, r.rnn_hidden_size)
        self.avail = torch.ones(N, 3, device=dev)
        self.ep_reward = z(N)
        self.done_rewards = []
        self.started = False

    def _share_row(self, obs, share):
        return share if self.envs.nonoverlapping else obs.reshape(self.N, -1)

    @property
    def _no_grad(self):
        return self.torch.no_grad()

    def collect(self):
        """collect -> envs.step -> insert for episode_length steps, all on the device (on_policy_base_runner.py:241-282)."""
        torch, r, T = self.torch, self.r, self.T
        with torch.no_grad():
            if not self.started:                     # warmup(): on_policy_base_runner.py:313-327
                obs, share = self.envs.reset_torch()
                self.obs[0].copy_(obs); self.share[0].copy_(self._share_row(obs, share))
                self.started = True
            else:                                    # after_update(): the last step becomes the first
                self.obs[0].copy_(self.obs[T]); self.share[0].copy_(self.share[T]); self.masks[0].copy_(self.masks[T])
            for t in range(T):
                for a in range(r.num_agents):
                    act, logp, _ = r.actor[a].get_actions(self.obs[t, :, a], self.rnn, self.masks[t], self.avail)
                    self.actions[t, :, a].copy_(act); self.logp[t, :, a].copy_(logp)
                value, _ = r.critic.get_values(self.share[t], self.rnn, self.masks[t])
                self.values[t].copy_(value)
                obs, share, rew, done = self.envs.step_torch(self.actions[t, :, :, 0])
                self.obs[t + 1].copy_(obs); self.share[t + 1].copy_(self._share_row(obs, share))
                self.rewards[t].copy_(rew[:, 0:1])                          # EP state: the critic learns agent 0's reward (:495-499)
                d = done.to(torch.float32)
                self.masks[t + 1].copy_((1.0 - d)[:, None])
                self.ep_reward += rew.mean(dim=1)                           # base_logger.py:57-64
                if bool(done.any()):
                    self.done_rewards.extend(self.ep_reward[done.bool()].tolist())
                    self.ep_reward *= (1.0 - d)
        return self.N * T

    def hand_over(self):
        """The episode's tensors -> the runner's numpy buffers, one copy per array (what T calls of insert() leave behind)."""
        r, A = self.r, self.r.num_agents
        obs, acts, logp, masks = self.obs.cpu().numpy(), self.actions.cpu().numpy(), self.logp.cpu().numpy(), self.masks.cpu().numpy()
        for a in range(A):
            b = r.actor_buffer[a]
            b.obs[:] = obs[:, :, a, :b.obs.shape[-1]]
            b.actions[:] = acts[:, :, a]; b.action_log_probs[:] = logp[:, :, a]
            b.masks[:] = masks; b.active_masks[:] = 1.0; b.rnn_states[:] = 0.0
            b.step = 0
        c = r.critic_buffer
        c.share_obs[:] = self.share.cpu().numpy()
        c.value_preds[:-1] = self.values.cpu().numpy(); c.rewards[:] = self.rewards.cpu().numpy()
        c.masks[:] = masks; c.bad_masks[:] = 1.0; c.rnn_states_critic[:] = 0.0
        c.step = 0
        lg = r.logger
        lg.done_episodes_rewards.extend(self.done_rewards)
        self.done_rewards = []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--harl-root", default=os.path.join(REPO, "baseline", "_ref"))
    ap.add_argument("--n-envs", type=int, default=4096)
    ap.add_argument("--episode-length", type=int, default=64)
    ap.add_argument("--episodes", type=int, default=3)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--device-rollout", action="store_true")
    ap.add_argument("--reference-logger", action="store_true", help="keep the reference's per-env Python logger (InfoBatch drop-in check)")
    ap.add_argument("--traces", default="data", choices=["data", "synthetic"])
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "harl_results"))
    a = ap.parse_args()
    import torch
    registry = install(a.harl_root, use_vector_logger=not a.reference_logger)
    algo_args, env_args = load_args(a.harl_root)
    N, T = a.n_envs, a.episode_length
    algo_args["train"].update(n_rollout_threads=N, episode_length=T, num_env_steps=N * T * a.episodes, log_interval=1, eval_interval=10 ** 9)
    algo_args["eval"]["use_eval"] = False
    algo_args["device"].update(cuda=torch.cuda.is_available(), torch_threads=4)
    algo_args["logger"]["log_dir"] = a.out
    algo_args["algo"].update(actor_num_mini_batch=1, critic_num_mini_batch=1, ppo_epoch=1, critic_epoch=1)
    env_args.update(location="ny", days_per_episode=7, device=a.device, output_views=True)
    env_args.pop("month", None)                       # months by rank, harl/utils/envs_tools.py:56-62
    if a.traces == "synthetic" or not os.path.isdir(os.path.join(a.harl_root, "data")):
        env_args["traces"] = "synthetic"
    else:
        env_args["data_root"] = os.path.join(a.harl_root, "data")
    if torch.cuda.is_available():
        torch.cuda.set_device(a.device)
    runner = registry["happo"]({"algo": "happo", "env": "sustaindc", "exp_name": "b200"}, algo_args, env_args)
    stats = {"insert_calls": 0, "env_steps": 0, "t_first": None, "t_last": None, "t_env": 0.0}
    if a.device_rollout:
        roll = DeviceRollout(runner)
        runner.logger.init(a.episodes)
        if hasattr(runner.logger, "attach"):
            runner.logger.attach(runner.envs)
        t_roll = 0.0
        for ep in range(1, a.episodes + 1):
            runner.logger.episode_init(ep)
            runner.prep_rollout()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            steps = roll.collect()
            torch.cuda.synchronize(); t_roll += time.perf_counter() - t0
            stats["env_steps"] += steps
            roll.hand_over()
            runner.compute(); runner.prep_training()
            actor_infos, critic_info = runner.train()
            runner.logger.episode_log(actor_infos, critic_info, runner.actor_buffer, runner.critic_buffer)
            runner.after_update()
        rate = stats["env_steps"] / t_roll
        mode = "DeviceRollout (torch CUDA buffers, one hand-over per episode)"
    else:
        if hasattr(runner.logger, "attach"):
            runner.logger.attach(runner.envs)
        insert, step, collect = runner.insert, runner.envs.step, runner.collect
        stats["t_rollout"] = 0.0

        def timed_collect(i):
            stats["t_c0"] = time.perf_counter()
            return collect(i)

        def timed_step(actions):
            t0 = time.perf_counter()
            out = step(actions)
            stats["t_env"] += time.perf_counter() - t0
            return out

        def counted_insert(data):
            now = time.perf_counter()
            stats["t_first"] = stats["t_first"] or now
            insert(data)
            stats["insert_calls"] += 1
            stats["env_steps"] += N
            stats["t_last"] = time.perf_counter()
            stats["t_rollout"] += stats["t_last"] - stats["t_c0"]          # collect -> envs.step -> logger.per_step -> insert
        runner.insert, runner.envs.step, runner.collect = counted_insert, timed_step, timed_collect
        t0 = time.perf_counter()
        runner.run()
        stats["wall_with_training"] = time.perf_counter() - t0            # what the reference's own FPS line counts (base_logger.py:86)
        rate = stats["env_steps"] / stats["t_rollout"]
        mode = "unmodified OnPolicyHARunner.run()"
    runner.close() if hasattr(runner, "close") else None
    print(json.dumps({"config": "configs[4]: HAPPO rollout through harl.runners", "mode": mode, "n_envs": N, "episode_length": T,
                      "episodes": a.episodes, "env_steps": stats["env_steps"], "env_steps_per_s_into_buffers": rate,
                      "env_step_call_s": stats["t_env"], "env_steps_per_s_through_envs_step": (stats["env_steps"] / stats["t_env"]) if stats["t_env"] else None,
                      "wall_with_training_s": stats.get("wall_with_training"), "logger": type(runner.logger).__module__}))


if __name__ == "__main__":
    main()

"""Vectorised `SustainDCLogger`: drop-in for the reference's per-step Python loop over `infos[i][0]`
(reference harl/envs/sustaindc/sustaindc_logger.py:5-272 over harl/common/base_logger.py:8-210).

Same constructor, methods, TensorBoard tags and printed lines; three ways to feed it, fastest first:
  * attached (`logger.attach(envs)` with a CudaShareVecEnv): `per_step` touches no info at all -- the sums come from the
    device-side accumulators the step kernel maintains (`sdc_metrics`) and the HVAC statistics from the device histogram
    (`sdc_hvac_histogram`; mean / max / 90th percentile to one bin = 0.02 % of the power range), read once per log;
  * `infos` is an `InfoBatch`: one vectorised column read per key (ten D2H copies of N floats, not N dict look-ups);
  * anything else (lists of dicts, e.g. the reference's own vec-env): the reference's loop.
At N = 65 536 envs the reference's loop is ~0.1 s of Python per step, two orders of magnitude above the step kernel
(SURVEY.md 8f-3).  Register it where the runners look it up:  harl.envs.LOGGER_REGISTRY["sustaindc"] = SustainDCLogger.
"""
import os
import time

import numpy as np

_SUM_KEYS = (            # (metric, info key)  sustaindc_logger.py:87-96
    ("net_energy_sum", "bat_total_energy_with_battery_KWh"), ("CO2_footprint_sum", "bat_CO2_footprint"),
    ("water_usage", "dc_water_usage"), ("load_left", "ls_unasigned_day_load_left"), ("ls_tasks_in_queue", "ls_tasks_in_queue"),
    ("ls_tasks_dropped", "ls_tasks_dropped"), ("ite_power_sum", "dc_ITE_total_power_kW"), ("ct_power_sum", "dc_CT_total_power_kW"),
    ("chiller_power_sum", "dc_Compressor_total_power_kW"), ("hvac_power_sum", "dc_HVAC_total_power_kW"))
_DEVICE_METRIC = {"net_energy_sum": "bat_total_energy_with_battery_KWh", "CO2_footprint_sum": "bat_CO2_footprint",
                  "water_usage": "dc_water_usage", "ls_tasks_in_queue": "ls_tasks_in_queue", "ls_tasks_dropped": "ls_tasks_dropped",
                  "ite_power_sum": "dc_ITE_total_power_kW", "ct_power_sum": "dc_CT_total_power_kW",
                  "chiller_power_sum": "dc_Compressor_total_power_kW", "hvac_power_sum": "dc_HVAC_total_power_kW",
                  "step_count": "env_steps"}
HVAC_SAMPLE_CAP = 1 << 25        # positive-HVAC samples kept verbatim per log interval (128 MB of fp32) before binning


def _fresh_metrics():
    return {"net_energy_sum": 0, "ite_power_sum": 0, "ct_power_sum": 0, "chiller_power_sum": 0, "hvac_power_sum": 0,
            "CO2_footprint_sum": 0, "water_usage": 0, "step_count": 0, "load_left": 0, "ls_tasks_in_queue": 0,
            "ls_tasks_dropped": 0, "instantaneous_net_energy": [], "hvac_power_on_used": [], "PUE": 0}


class SustainDCLogger:
    def __init__(self, args, algo_args, env_args, num_agents, writter, run_dir):
        self.args, self.algo_args, self.env_args = args, algo_args, env_args
        self.task_name = self.get_task_name()
        self.num_agents = num_agents
        self.writter = writter
        self.run_dir = run_dir
        self.log_file = open(os.path.join(run_dir, "progress.txt"), "w", encoding="utf-8")
        self.avg_eval_episode_reward = 0.0
        self._envs = None
        self.metrics = _fresh_metrics()
        self.is_off_policy = False

    # ---- additive: device-side accumulators ---------------------------------------------------------------------
    def attach(self, envs):
        """Take the training metrics from `envs` (CudaShareVecEnv) device accumulators instead of reading infos."""
        self._envs = envs
        envs.metrics(clear=True)
        envs.engine.hvac_histogram(clear=True)
        return self

    # ---- reference surface ----------------------------------------------------------------------------------------
    def get_task_name(self):
        loc = self.env_args["location"]
        return f"{loc if isinstance(loc, str) else 'mixed'}-discrete"

    def init(self, episodes):
        self.start = time.time()
        self.episodes = episodes
        self.train_episode_rewards = np.zeros(self.algo_args["train"]["n_rollout_threads"])
        self.done_episodes_rewards = []

    def episode_init(self, episode):
        self.episode = episode
        self.metrics = _fresh_metrics()
        self.is_off_policy = False

    def _accumulate(self, metrics, infos):
        if hasattr(infos, "column"):                                   # InfoBatch: one vectorised read per key
            for name, key in _SUM_KEYS:
                metrics[name] += float(np.sum(infos.column(key), dtype=np.float64))
            hvac = np.asarray(infos.column("dc_HVAC_total_power_kW"))
            pos = hvac[hvac > 0]
            if pos.size:
                metrics["hvac_power_on_used"].append(pos.astype(np.float32))
                self._maybe_bin(metrics)
            metrics["step_count"] += len(infos)
            return
        for i in range(len(infos)):                                    # the reference's loop (sustaindc_logger.py:86-101)
            row = infos[i][0]
            for name, key in _SUM_KEYS:
                metrics[name] += row.get(key, 0)
            if row.get("dc_HVAC_total_power_kW", 0) > 0:
                metrics["hvac_power_on_used"].append(row.get("dc_HVAC_total_power_kW", 0))
            metrics["step_count"] += 1

    @staticmethod
    def _maybe_bin(metrics):
        chunks = metrics["hvac_power_on_used"]
        if sum(np.size(c) for c in chunks) <= HVAC_SAMPLE_CAP or isinstance(chunks[0], dict):
            return
        allv = np.concatenate([np.ravel(c) for c in chunks])
        hi = float(allv.max()) * 2.0
        counts, _ = np.histogram(allv, bins=1 << 16, range=(0.0, hi))
        metrics["hvac_power_on_used"] = [{"counts": counts.astype(np.int64), "hi": hi, "sum": float(allv.sum(dtype=np.float64)), "max": float(allv.max())}]

    @staticmethod
    def _hvac_stats(samples):
        """(mean, max, p90) of the positive HVAC power samples, or None."""
        if len(samples) == 0:
            return None
        if isinstance(samples[0], dict):                               # binned (beyond HVAC_SAMPLE_CAP samples)
            h = samples[0]
            for c in samples[1:]:
                c = np.ravel(c)
                h["counts"] += np.histogram(c, bins=len(h["counts"]), range=(0.0, h["hi"]))[0]
                h["sum"] += float(c.sum(dtype=np.float64)); h["max"] = max(h["max"], float(c.max()))
            total = h["counts"].sum()
            cum = np.cumsum(h["counts"])
            b = int(np.searchsorted(cum, 0.9 * (total - 1) + 1))
            return h["sum"] / total, h["max"], (b + 0.5) * h["hi"] / len(h["counts"])
        allv = np.concatenate([np.ravel(np.asarray(c, np.float64)) for c in samples])
        return float(np.mean(allv)), float(np.max(allv)), float(np.percentile(allv, 90))

    def per_step(self, data):
        obs, share_obs, rewards, dones, infos = data[:5]
        dones_env = np.all(dones, axis=1)
        self.train_episode_rewards += np.mean(rewards, axis=1).flatten()       # base_logger.py:57-59
        idx = np.nonzero(dones_env)[0]
        if len(idx):
            self.done_episodes_rewards.extend(self.train_episode_rewards[idx].tolist())
            self.train_episode_rewards[idx] = 0
        if self._envs is None:
            self._accumulate(self.metrics, infos)

    def eval_per_step(self, eval_data):
        eval_rewards, eval_infos = eval_data[2], eval_data[4]
        for eval_i in range(self.algo_args["eval"]["n_eval_rollout_threads"]):
            self.one_episode_rewards[eval_i].append(eval_rewards[eval_i])
        self.eval_infos = eval_infos
        self._accumulate(self.eval_metrics, eval_infos)

    def _write_metrics(self, prefix, m):
        """The scalars of episode_log / eval_log (sustaindc_logger.py:126-171, 203-240).  Returns (avg energy, avg CO2, water)."""
        n = m["step_count"]
        avg = {k: (m[k] / n if n > 0 else 0) for k in ("net_energy_sum", "ite_power_sum", "ct_power_sum", "chiller_power_sum",
                                                      "hvac_power_sum", "CO2_footprint_sum")}
        stats = m.get("hvac_stats") or self._hvac_stats(m["hvac_power_on_used"])
        w, t = self.writter, self.total_num_steps
        if stats is not None:
            w.add_scalar(prefix + "/Average HVAC Power on use", stats[0], t)
            w.add_scalar(prefix + "/Max HVAC Power on use", stats[1], t)
            w.add_scalar(prefix + "/Percentile 90% HVAC Power on use", stats[2], t)
        w.add_scalar(prefix + "/Average Net Energy", avg["net_energy_sum"], t)
        w.add_scalar(prefix + "/Average ITE Power", avg["ite_power_sum"], t)
        w.add_scalar(prefix + "/Average CT Power", avg["ct_power_sum"], t)
        w.add_scalar(prefix + "/Average Chiller Power", avg["chiller_power_sum"], t)
        w.add_scalar(prefix + "/Average HVAC Power", avg["hvac_power_sum"], t)
        w.add_scalar(prefix + "/Average PUE", 1 + avg["hvac_power_sum"] / avg["ite_power_sum"] if avg["ite_power_sum"] else float("nan"), t)
        w.add_scalar(prefix + "/Average CO2 Footprint", avg["CO2_footprint_sum"], t)
        w.add_scalar(prefix + "/Total Water Usage", m["water_usage"] if n > 0 else 0, t)
        w.add_scalar(prefix + "/Total Tasks in Queue", m["ls_tasks_in_queue"] if n > 0 else 0, t)
        w.add_scalar(prefix + "/Total Tasks Dropped", m["ls_tasks_dropped"] if n > 0 else 0, t)
        return avg["net_energy_sum"], avg["CO2_footprint_sum"], m["water_usage"] if n > 0 else 0, m["ls_tasks_in_queue"], m["ls_tasks_dropped"]

    def episode_log(self, actor_train_infos, critic_train_info, actor_buffer, critic_buffer):
        train = self.algo_args["train"]
        self.total_num_steps = self.episode * train["episode_length"] * train["n_rollout_threads"]
        self.end = time.time()
        print("Env {} Task {} Algo {} Exp {} updates {}/{} episodes, total num timesteps {}/{}, FPS {}.".format(
            self.args["env"], self.task_name, self.args["algo"], self.args["exp_name"], self.episode, self.episodes,
            self.total_num_steps, train["num_env_steps"], int(self.total_num_steps / (self.end - self.start))))
        critic_train_info["average_step_rewards"] = critic_buffer.get_mean_rewards()
        self.log_train(actor_train_infos, critic_train_info)
        print("Average step reward is {}.".format(critic_train_info["average_step_rewards"]))
        if len(self.done_episodes_rewards) > 0:
            aver_episode_rewards = np.mean(self.done_episodes_rewards)
            print("Some episodes done, average episode reward is {}.\n".format(aver_episode_rewards))
            self.writter.add_scalar("train/average_step_rewards", aver_episode_rewards, self.total_num_steps)
            self.done_episodes_rewards = []
        if self._envs is not None:                                     # device accumulators: one read per log interval
            dm = self._envs.metrics(clear=True)
            for name, key in _DEVICE_METRIC.items():
                self.metrics[name] = dm[key]
            st = self._envs.hvac_power_stats(clear=True)
            self.metrics["hvac_stats"] = (st["mean"], st["max"], st["p90"]) if st["samples"] else None
            self.metrics["hvac_power_on_used"] = []
        e, c, wtr, q, d = self._write_metrics("metrics", self.metrics)
        print(f"Episode {self.episode}: Avg Net Energy={e:.3f}, Avg CO2={c:.3f}, Water Usage={wtr:.3f}")
        print(f"Tasks in Queue={q:.3f}, Tasks Dropped={d:.3f}")
        self.metrics = _fresh_metrics()

    def eval_init(self):
        train = self.algo_args["train"]
        self.total_num_steps = self.episode * train["episode_length"] * train["n_rollout_threads"]
        self._eval_reset()
        self.is_off_policy = False

    def eval_init_off_policy(self, total_num_steps):
        self.total_num_steps = total_num_steps
        self._eval_reset()
        self.is_off_policy = True

    def _eval_reset(self):
        n = self.algo_args["eval"]["n_eval_rollout_threads"]
        self.eval_episode_rewards = [[] for _ in range(n)]
        self.one_episode_rewards = [[] for _ in range(n)]
        self.eval_metrics = _fresh_metrics()

    def eval_thread_done(self, tid):
        self.eval_episode_rewards[tid].append(np.sum(self.one_episode_rewards[tid], axis=0))
        self.one_episode_rewards[tid] = []

    def eval_log(self, eval_episode):
        self.eval_episode_rewards = np.concatenate([rewards for rewards in self.eval_episode_rewards if rewards])
        self.log_env({"eval_average_episode_rewards": self.eval_episode_rewards,
                      "eval_max_episode_rewards": [np.max(self.eval_episode_rewards)]})
        eval_avg_rew = np.mean(self.eval_episode_rewards)
        print("Evaluation average episode reward is {}.\n".format(eval_avg_rew))
        self.log_file.write(",".join(map(str, [self.total_num_steps, eval_avg_rew])) + "\n")
        self.log_file.flush()
        e, c, wtr, q, d = self._write_metrics("eval_metrics", self.eval_metrics)
        if self.is_off_policy:
            train = self.algo_args["train"]
            self.episode = int(self.total_num_steps) // train["episode_length"] // train["n_rollout_threads"]
        print(f"Episode {self.episode}: Avg Net Energy={e:.3f}, Avg CO2={c:.3f}, Water Usage={wtr:.3f}")
        print(f"Tasks in Queue={q:.3f}, Tasks Dropped={d:.3f}")
        self.eval_metrics = _fresh_metrics()
        self.avg_eval_episode_reward = np.mean(self.eval_episode_rewards)

    def log_train(self, actor_train_infos, critic_train_info):
        for agent_id in range(self.num_agents):
            for k, v in actor_train_infos[agent_id].items():
                self.writter.add_scalar(f"train/agent{agent_id}/{k}", v, self.total_num_steps)
        for k, v in critic_train_info.items():
            self.writter.add_scalar(f"train/critic/{k}", v, self.total_num_steps)

    def log_env(self, env_infos):
        for k, v in env_infos.items():
            if len(v) > 0:
                self.writter.add_scalar(f"metrics/{k}", np.mean(v), self.total_num_steps)

    def save_weights_log(self):
        msg = "model weights at episode {} with average episode reward {}\n".format(self.episode, self.avg_eval_episode_reward)
        self.log_file.write("Saving " + msg)
        self.log_file.flush()
        print("We are saving " + msg)

    def close(self):
        self.log_file.close()

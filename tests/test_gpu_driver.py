"""The training loop of run_experiment.py:349-472 + save / resume (:692-714, :224-233, :489-509) on the GPU
(`dreamer_v3_b200/training/driver.py`): environment steps through `DreamerModel.forward_inference`, the HBM replay
buffer, graph-replayed updates, checkpoints written in the reference's file layout - and a resumed run that continues
EXACTLY where the uninterrupted one would have been (weights, optimizer slots, buffer, counters, noise stream)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class ToyEnvs:
    """`n` parallel 1-D point-mass environments (numpy): obs = (x, v, 1), action in {left, stay, right}, reward = -|x|,
    terminated when |x| > 3; episodes are cut into chunks per `collect` call like EnvRunnerV2.sample (env_runner_v2.py:221-330)."""

    def __init__(self, n, seed):
        self.n, self.rng = n, np.random.default_rng(seed)
        self.x = self.rng.uniform(-1, 1, n).astype(np.float32)
        self.v = np.zeros(n, np.float32)
        self.first = np.ones(n, np.float32)
        self.ids = [f"e{seed}-{i}-0" for i in range(n)]
        self.count = n

    def obs(self):
        return np.stack([self.x, self.v, np.ones_like(self.x)], -1).astype(np.float32)

    def collect(self, dm, steps):
        from dreamer_v3_b200.utils.episode import Episode
        states = getattr(self, "states", None)
        if states is None:
            init = dm.get_initial_state()
            states = {k: v.repeat_interleave(self.n, 0) for k, v in init.items()}
        chunks = [Episode(self.ids[i]) for i in range(self.n)]
        for i, c in enumerate(chunks):
            c.add_initial_observation(initial_observation=self.obs()[i])
        done_chunks = []
        for _ in range(steps):
            actions, states = dm.forward_inference(states, self.obs(), self.first)
            a = actions.argmax(-1).cpu().numpy()
            self.first[:] = 0.0
            self.v = 0.8 * self.v + 0.2 * (a.astype(np.float32) - 1.0)
            self.x = self.x + self.v
            o = self.obs()
            for i in range(self.n):
                term = bool(abs(self.x[i]) > 3.0)
                chunks[i].add_timestep(o[i], int(a[i]), float(-abs(self.x[i])), is_terminated=term)
                if term:
                    done_chunks.append(chunks[i])
                    self.x[i], self.v[i], self.first[i] = self.rng.uniform(-1, 1), 0.0, 1.0
                    self.ids[i] = f"{self.ids[i].rsplit('-', 1)[0]}-{self.count}"
                    self.count += 1
                    chunks[i] = Episode(self.ids[i])
                    chunks[i].add_initial_observation(initial_observation=self.obs()[i])
        self.states = states
        return done_chunks + [c for c in chunks if len(c) > 0]


def _build(seed=11, use_curiosity=False):
    from dreamer_v3_b200.models.components import MLP, VectorDecoder
    from dreamer_v3_b200.models.dreamer_model import DreamerModel
    from dreamer_v3_b200.models.world_model import WorldModel
    from dreamer_v3_b200.spec import Box, Discrete
    from dreamer_v3_b200.utils.episode_replay_buffer import EpisodeReplayBuffer
    B, T, H = 4, 8, 3
    wm = WorldModel(model_dimension="XXS", action_space=Discrete(3), batch_length_T=T, encoder=MLP(model_dimension="XXS"),
                    decoder=VectorDecoder(model_dimension="XXS", observation_space=Box(-5, 5, (3,))), symlog_obs=True, seed=seed)
    dm = DreamerModel(model_dimension="XXS", action_space=Discrete(3), batch_size_B=B, batch_length_T=T, horizon_H=H,
                      world_model=wm, use_curiosity=use_curiosity)
    return dm, EpisodeReplayBuffer(capacity=2000), dict(batch_size_B=B, batch_length_T=T, horizon_H=H, training_ratio=4)


@pytest.mark.parametrize("use_curiosity", (False, True))
def test_driver_loop_and_exact_resume(tmp_path, use_curiosity):
    from dreamer_v3_b200.training.driver import run_training
    seen = []

    def run(dm, buf, envs, n, **kw):
        return run_training(dreamer_model=dm, buffer=buf, collect=lambda it: envs.collect(dm, 16), num_iterations=n,
                            on_iteration=lambda it, c, wmr, acr: seen.append((it, float(wmr["WORLD_MODEL_L_total"]), float(acr["CRITIC_L_total"]))),
                            **kw)

    # ---- uninterrupted: 4 iterations, checkpoint after the second ----
    dm, buf, kw = _build(use_curiosity=use_curiosity)
    envs = ToyEnvs(4, seed=1)
    np.random.seed(0)
    c2 = run(dm, buf, envs, 2, checkpoint_path=str(tmp_path), model_save_frequency_main_iters=2, **kw)
    ck = os.path.join(str(tmp_path), "2")
    for f in ("world_model.npz", "actor.npz", "critic.npz", "world_model_optimizer.npz", "buffer.npz", "total_env_steps.npy",
              "total_replayed_steps.npy", "total_train_steps.npy", "iteration.npy") + (("disagree_nets.npz",) if use_curiosity else ()):
        assert os.path.exists(os.path.join(ck, f)), f
    assert c2["iteration"] == 2 and c2["total_env_steps"] >= 2 * 4 * 16
    # training ratio 4 with B*T = 32 replayed steps per update: at least 4 * env_steps / 32 updates
    assert c2["total_train_steps"] >= 4 * c2["total_env_steps"] // 32 and c2["total_replayed_steps"] == 32 * c2["total_train_steps"]
    assert dm.world_model.optimizer.iterations == c2["total_train_steps"]
    rng_state = np.random.get_state()
    envs_state = (envs.x.copy(), envs.v.copy(), envs.first.copy(), list(envs.ids), envs.count, envs.rng.bit_generator.state,
                  {k: v.clone() for k, v in envs.states.items()})
    c4 = run(dm, buf, envs, 2, counters=c2, **kw)
    assert c4["iteration"] == 4
    want = dm.engine.get_params()
    want_sizes = (buf.get_num_episodes(), buf.get_num_timesteps())
    assert all(np.isfinite(x[1]) and np.isfinite(x[2]) for x in seen)
    tail = seen[-2:]

    # ---- resumed: fresh model (other seed) and empty buffer, restored from the checkpoint, same env / sampler state ----
    dm2, buf2, _ = _build(seed=99, use_curiosity=use_curiosity)
    envs2 = ToyEnvs(4, seed=1)
    envs2.x, envs2.v, envs2.first, envs2.ids, envs2.count = envs_state[0], envs_state[1], envs_state[2], envs_state[3], envs_state[4]
    envs2.rng.bit_generator.state = envs_state[5]
    envs2.states = {k: v.to(dm2.engine.device) for k, v in envs_state[6].items()}
    np.random.set_state(rng_state)
    seen.clear()
    r4 = run(dm2, buf2, envs2, 2, resume_from=ck, **kw)
    assert r4 == c4
    assert (buf2.get_num_episodes(), buf2.get_num_timesteps()) == want_sizes
    got = dm2.engine.get_params()
    # Same kernels, same inputs, same noise stream: the continuation is the uninterrupted run up to the summation order of
    # the gradient atomics (LayerNorm gains, biases, split-K). That is the last ulp of a summand - but Adam divides by
    # sqrt(v): where a gradient component cancels to ~0 its sign, and with it a whole step of size lr, follows the noise. So:
    # the distance between the two runs is measured against the distance both travelled since the checkpoint (a forgotten
    # optimizer slot, step counter, buffer entry or noise position would make it O(1)), and is exactly 0 for most tensors.
    base = {k[len("param/"):]: v for k, v in np.load(os.path.join(ck, "dv3_state.npz")).items() if k.startswith("param/")}
    num = np.sqrt(sum(float(((got[n] - want[n]).astype(np.float64) ** 2).sum()) for n in want))
    den = np.sqrt(sum(float(((want[n] - base[n]).astype(np.float64) ** 2).sum()) for n in want))
    assert den > 0 and num <= 0.02 * den, (num, den)
    worst = max(float(np.abs(got[n] - want[n]).max()) for n in want)
    assert worst <= 16 * 2 * 1e-4, worst          # no component further apart than the updates since the checkpoint allow
    # (measured on B200: 0.0 without curiosity, i.e. bit-identical; 1.9e-4 on a few components of the 8 extra nets with it)
    np.testing.assert_allclose(np.array(seen[-2:]), np.array(tail), rtol=1e-3)
    dm.engine.close(); dm2.engine.close()

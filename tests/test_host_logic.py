"""Host-side logic that needs no GPU: the checkpoint weight orders cover the parameter table exactly (every optimizer group,
with and without curiosity), and the training loop's bookkeeping (run_experiment.py:349-472, :692-714: training ratio, counters,
save cadence, resume) with the device work stubbed out."""
import numpy as np
import pytest

from dreamer_v3_b200.spec import EngineConfig, param_spec
from dreamer_v3_b200.training import driver
from dreamer_v3_b200.utils import checkpoint
from dreamer_v3_b200.utils.episode import Episode

CONFIGS = [
    dict(model_dimension="XS", A=2, discrete=True, obs_dim=4),
    dict(model_dimension="S", A=6, discrete=True, image_obs=True),
    dict(model_dimension="M", A=6, discrete=False, image_obs=True),
    dict(model_dimension="XL", A=18, discrete=True, image_obs=True, use_curiosity=True),
    dict(model_dimension="nano", A=3, discrete=False, obs_dim=5, use_curiosity=True, num_curiosity_nets=3),
]


@pytest.mark.parametrize("kw", CONFIGS)
def test_checkpoint_weight_orders_cover_the_parameter_table(kw):
    cfg = EngineConfig.make(**kw)
    spec = param_spec(cfg)
    by_group = {}
    for p in spec:
        by_group.setdefault(p.group, []).append(p.name)
    models = ["world_model", "actor", "critic"] + (["disagree"] if cfg.num_curiosity_nets else [])
    assert [m for m, _ in checkpoint._models(cfg)] == models
    for m in models:
        order = checkpoint.weight_order(cfg, m)
        assert len(order) == len(set(order))
        want = set(by_group[m]) | (set(by_group["critic_ema"]) if m == "critic" else set())
        assert set(order) == want, (m, sorted(set(order) ^ want)[:5])
        # the optimizer's slots exist for the trainable ones only (no slots for the EMA critic)
        assert set(checkpoint._trainable(cfg, m)) == set(by_group[m])
    if cfg.num_curiosity_nets:   # disagree_networks.py:28-40: every MLP first, then every RepresentationLayer
        order = checkpoint.weight_order(cfg, "disagree")
        first_repr = min(i for i, n in enumerate(order) if ".repr." in n)
        assert all(".repr." in n for n in order[first_repr:]) and not any(".repr." in n for n in order[:first_repr])
        assert sum(n.endswith(".repr.w") for n in order) == cfg.num_curiosity_nets
    with pytest.raises(KeyError):
        checkpoint.weight_order(cfg, "nonexistent")


class _FakeBuffer:
    def __init__(self):
        self.timesteps, self.added = 0, []

    def add(self, episodes):
        self.added.append(len(episodes))
        self.timesteps += sum(len(e) for e in episodes)

    def get_num_timesteps(self):
        return self.timesteps


class _FakeCritic:
    def __init__(self):
        self.ema_inits = 0

    def init_ema(self):
        self.ema_inits += 1


class _FakeModel:
    use_curiosity = False

    def __init__(self):
        self.critic = _FakeCritic()


def _chunk(n):
    e = Episode("e")
    e.add_initial_observation(initial_observation=np.zeros(2, np.float32))
    for _ in range(n):
        e.add_timestep(np.zeros(2, np.float32), 0, 0.0)
    return e


def test_training_loop_bookkeeping(monkeypatch, tmp_path):
    updates, saves, loads = [], [], []
    monkeypatch.setattr(driver, "update_from_buffer", lambda **kw: (updates.append(kw["batch_size_B"]) or ({"wm": 1}, {"ac": 2})))
    monkeypatch.setattr(driver.checkpoint, "save_checkpoint", lambda path, dm, buffer=None, **c: saves.append((path, dict(c))))
    monkeypatch.setattr(driver.checkpoint, "load_checkpoint",
                        lambda path, dm, buffer=None: loads.append(path) or dict(total_env_steps=70, total_replayed_steps=640,
                                                                                  total_train_steps=20, iteration=2))
    dm, buf, seen = _FakeModel(), _FakeBuffer(), []
    kw = dict(batch_size_B=4, batch_length_T=8, horizon_H=3, training_ratio=16)
    # 10 env steps per collect call; the first iteration has to collect until one batch (32 timesteps) is in the buffer
    c = driver.run_training(dreamer_model=dm, buffer=buf, collect=lambda it: [_chunk(6), _chunk(4)], num_iterations=3,
                            checkpoint_path=str(tmp_path), model_save_frequency_main_iters=2,
                            on_iteration=lambda it, cc, w, a: seen.append((it, cc["total_train_steps"], w, a)), **kw)
    # iteration 0: 4 collect calls (40 steps) -> ceil(16 * 40 / 32) = 20 updates; iterations 1, 2: 10 steps -> 5 updates each
    assert buf.added == [2] * 6
    assert c == dict(total_env_steps=60, total_replayed_steps=32 * 30, total_train_steps=30, iteration=3)
    assert len(updates) == 30 and dm.critic.ema_inits == 1                      # EMA initialised once, before the first update
    assert [s[0] for s in seen] == [0, 1, 2] and seen[-1][1:] == (30, {"wm": 1}, {"ac": 2})
    assert len(saves) == 1 and saves[0][0].endswith("/2") and saves[0][1]["iteration"] == 2 and saves[0][1]["total_train_steps"] == 25
    # continuing in the same process keeps counting and does not re-initialise the EMA
    c2 = driver.run_training(dreamer_model=dm, buffer=buf, collect=lambda it: [_chunk(10)], num_iterations=1, counters=c, **kw)
    assert c2["iteration"] == 4 and c2["total_train_steps"] == 35 and dm.critic.ema_inits == 1
    # resume: counters come from the checkpoint, the EMA came with the critic's weights
    dm3, buf3 = _FakeModel(), _FakeBuffer()
    buf3.timesteps = 100
    c3 = driver.run_training(dreamer_model=dm3, buffer=buf3, collect=lambda it: [_chunk(2)], num_iterations=2, resume_from="ck", **kw)
    assert loads == ["ck"] and dm3.critic.ema_inits == 0
    assert c3 == dict(total_env_steps=74, total_replayed_steps=640 + 2 * 32, total_train_steps=22, iteration=4)
    # train_critic=False never touches the EMA
    dm4 = _FakeModel()
    driver.run_training(dreamer_model=dm4, buffer=_FakeBuffer(), collect=lambda it: [_chunk(32)], num_iterations=1, train_critic=False, **kw)
    assert dm4.critic.ema_inits == 0

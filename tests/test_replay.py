"""Replay sampler: the restatement `oracle/replay_oracle.py` and the HBM-resident
`dreamer_v3_b200.utils.episode_replay_buffer.EpisodeReplayBuffer` (host plan + `dv3_replay_gather` kernel) against the
outputs of the reference's own `EpisodeReplayBuffer` (tests/golden/ref_replay.npz, tests/golden/make_replay_golden.py).
Everything here is index / byte work: the bar is bit-exact."""
import os

import numpy as np
import pytest

from _replay_scenario import CAPACITY, SAMPLES, make_adds
from oracle.replay_oracle import ReplayOracle

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_replay.npz"))
KEYS = ("obs", "actions", "rewards", "is_first", "is_last", "is_terminated")


def oracle_buffer(image, capacity=CAPACITY, **kw):
    o = ReplayOracle(capacity)
    for id_, obs, actions, rewards, term in make_adds(image=image, **kw):
        o.add(id_, list(obs), list(actions), list(rewards), term)
    return o


@pytest.mark.parametrize("kind", ("vec", "img"))
def test_oracle_matches_reference_buffer(kind):
    o = oracle_buffer(kind == "img")
    assert len(o.eps) == int(GOLD[f"{kind}/num_episodes"]) and len(o.indices) == int(GOLD[f"{kind}/num_timesteps"])
    for i, (B, T, seed) in enumerate(SAMPLES):
        np.random.seed(seed)
        s = o.sample(B, T)
        for k in KEYS:
            want = GOLD[f"{kind}/s{i}/{k}"]
            assert s[k].shape == want.shape and np.array_equal(s[k], want), (kind, i, k)


@pytest.mark.parametrize("kind", ("vec", "img"))
def test_native_plan_matches_reference_buffer(kind):
    """Host logic without a GPU: the bookkeeping of `add` (pages, chunk index space, ejection) and the native plan builder
    `dv3_replay_plan`; the plan is applied with NumPy to a host copy of the store and compared with the reference's output."""
    buf = device_buffer(kind == "img", _host_index_only=True)
    assert buf.get_num_episodes() == int(GOLD[f"{kind}/num_episodes"])
    assert buf.get_num_timesteps() == int(GOLD[f"{kind}/num_timesteps"])
    obs_s, act_s, rew_s = buf._obs_store.numpy(), buf._act_store.numpy(), buf._rew_store.numpy()
    for i, (B, T, seed) in enumerate(SAMPLES):
        np.random.seed(seed)
        plan = buf.plan(B, T)
        want = {k: GOLD[f"{kind}/s{i}/{k}"] for k in KEYS}
        assert np.array_equal(obs_s[plan[:, 0]].reshape(want["obs"].shape), want["obs"])
        assert np.array_equal(act_s[plan[:, 1]].reshape(want["actions"].shape), want["actions"])
        assert np.array_equal(np.where(plan[:, 2] >= 0, rew_s[np.maximum(plan[:, 2], 0)], 0.0).reshape(B, T), want["rewards"])
        for bit, k in ((1, "is_first"), (2, "is_last"), (4, "is_terminated")):
            assert np.array_equal((plan[:, 3] & bit).reshape(B, T) != 0, want[k]), k
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        buf.sample(4, 8)


def device_buffer(image, capacity=CAPACITY, normalize_uint8=False, **kw):
    from dreamer_v3_b200.utils.episode import Episode
    from dreamer_v3_b200.utils.episode_replay_buffer import EpisodeReplayBuffer
    buf = EpisodeReplayBuffer(capacity=capacity, normalize_uint8=normalize_uint8, _host_index_only=kw.pop("_host_index_only", False))
    for id_, obs, actions, rewards, term in make_adds(image=image, **kw):
        buf.add(Episode(id_, observations=list(obs), actions=list(actions), rewards=list(rewards), is_terminated=term))
    return buf


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ("vec", "img"))
def test_device_buffer_matches_reference_buffer(kind):
    buf = device_buffer(kind == "img")
    assert buf.get_num_episodes() == int(GOLD[f"{kind}/num_episodes"])
    assert buf.get_num_timesteps() == int(GOLD[f"{kind}/num_timesteps"])
    for i, (B, T, seed) in enumerate(SAMPLES):
        np.random.seed(seed)
        s = buf.sample(batch_size_B=B, batch_length_T=T)
        for k in KEYS:
            want = GOLD[f"{kind}/s{i}/{k}"].astype(np.float32)
            got = s[k].cpu().numpy()
            assert got.shape == want.shape and np.array_equal(got, want), (kind, i, k)
    assert buf.get_sampled_timesteps() == sum(B * T for B, T, _ in SAMPLES)


@pytest.mark.gpu
def test_device_buffer_matches_oracle_full_size_and_normalises_uint8():
    """B16 x T64 rows of 64x64x3 uint8 frames at the Atari shape (BASELINE config 3): many ejections, page growth,
    frames stored as uint8 and normalised x/128-1 in the gather (env_runner_v2.py:53-54) - bit-exact in float32."""
    import torch
    from dreamer_v3_b200.utils.episode import Episode
    from dreamer_v3_b200.utils.episode_replay_buffer import EpisodeReplayBuffer
    rng = np.random.default_rng(5)
    o, buf = ReplayOracle(3000), EpisodeReplayBuffer(capacity=3000, normalize_uint8=True)
    for e in range(40):
        n = int(rng.integers(5, 400))
        obs = rng.integers(0, 256, size=(n + 1, 64, 64, 3), dtype=np.uint8)
        act = np.eye(6, dtype=np.float32)[rng.integers(0, 6, size=n)]
        rew = rng.choice([-1.0, 0.0, 1.0], size=n).astype(np.float32)
        term = bool(rng.random() < 0.7)
        o.add(f"e{e}", list(obs), list(act), list(rew), term)
        buf.add(Episode(f"e{e}", observations=list(obs), actions=list(act), rewards=list(rew), is_terminated=term))
    assert buf.get_num_timesteps() == len(o.indices) <= 3000
    out = {"obs": torch.empty((16, 64, 64, 64, 3), device="cuda")}  # gathers straight into a caller-owned tensor
    for seed in (1, 2):
        np.random.seed(seed)
        want = o.sample(16, 64)
        np.random.seed(seed)
        got = buf.sample(16, 64, out=out)
        assert got["obs"].data_ptr() == out["obs"].data_ptr()
        assert np.array_equal(got["obs"].cpu().numpy(), want["obs"].astype(np.float32) / 128.0 - 1.0)
        for k in KEYS[1:]:
            assert np.array_equal(got[k].cpu().numpy(), want[k].astype(np.float32)), k


@pytest.mark.gpu
def test_device_buffer_properties_of_the_reference_main():
    """The property test of utils/episode_replay_buffer.py:265-313 (obs == timestep, action == timestep, reward ==
    0.1 * timestep) on the device buffer."""
    from dreamer_v3_b200.utils.episode import Episode
    from dreamer_v3_b200.utils.episode_replay_buffer import EpisodeReplayBuffer
    rng = np.random.default_rng(3)
    buf = EpisodeReplayBuffer(capacity=10000)
    for _ in range(200):
        eps = Episode(observations=[np.float32(0.0)])
        for t in range(int(rng.integers(1, 500))):
            eps.add_timestep(observation=np.float32(t + 1), action=np.float32(t), reward=np.float32(0.1) * np.float32(t + 1))
        eps.is_terminated = bool(rng.random() > 0.5)
        buf.add(eps)
    np.random.seed(0)
    for _ in range(50):
        s = {k: v.cpu().numpy() for k, v in buf.sample(batch_size_B=16, batch_length_T=64).items()}
        obs, actions, rewards = s["obs"], s["actions"], s["rewards"]
        is_first, is_last, is_term = s["is_first"] > 0, s["is_last"] > 0, s["is_terminated"] > 0
        assert is_last[:, -1].all() and is_first[:, 0].all()
        assert obs.shape[:2] == rewards.shape == actions.shape[:2] == is_first.shape == is_last.shape == is_term.shape
        assert np.array_equal(obs.astype(np.float32) * np.float32(0.1), rewards)
        assert np.where(is_last, True, obs == actions).all()
        assert np.where(is_term[:, 1:], actions[:, 1:] == actions[:, :-1], True).all()
        assert np.where(is_term[:, :-1], rewards[:, 1:] == 0.0, True).all()


def test_no_cpu_fallback_for_the_replay_buffer():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from dreamer_v3_b200.utils.episode_replay_buffer import EpisodeReplayBuffer
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        EpisodeReplayBuffer(capacity=10)


@pytest.mark.parametrize("kind", ("vec", "img"))
def test_buffer_state_round_trip(kind, tmp_path):
    """get_state / set_state in the reference's key layout (utils/episode_replay_buffer.py:240-264; written as buffer.npz by
    run_experiment.py:708, read back at :228): a buffer restored from the saved file reproduces the reference's own sample
    vectors bit for bit (host bookkeeping + native plan; the store is a host tensor here)."""
    from dreamer_v3_b200.utils.episode_replay_buffer import EpisodeReplayBuffer
    buf = device_buffer(kind == "img", _host_index_only=True)
    np.random.seed(5)
    buf.plan(4, 8)
    buf.sampled_timesteps = 32
    np.savez(tmp_path / "buffer", state=buf.get_state())
    state = np.load(tmp_path / "buffer.npz", allow_pickle=True)["state"]
    assert [state[i][0] for i in range(6)] == ["episodes", "episode_id_to_index", "_num_episodes_ejected", "_indices", "size",
                                               "sampled_timesteps"]
    new = EpisodeReplayBuffer(capacity=CAPACITY, normalize_uint8=False, _host_index_only=True)
    new.set_state(state)
    assert new.get_num_episodes() == buf.get_num_episodes() == int(GOLD[f"{kind}/num_episodes"])
    assert new.get_num_timesteps() == buf.get_num_timesteps() == int(GOLD[f"{kind}/num_timesteps"])
    assert new.size == buf.size and new.get_sampled_timesteps() == 32
    assert new._num_episodes_ejected == buf._num_episodes_ejected and new.episode_id_to_index == buf.episode_id_to_index
    obs_s, act_s, rew_s = new._obs_store.numpy(), new._act_store.numpy(), new._rew_store.numpy()
    for i, (B, T, seed) in enumerate(SAMPLES):
        np.random.seed(seed)
        plan = new.plan(B, T)
        want = {k: GOLD[f"{kind}/s{i}/{k}"] for k in KEYS}
        assert np.array_equal(obs_s[plan[:, 0]].reshape(want["obs"].shape), want["obs"])
        assert np.array_equal(act_s[plan[:, 1]].reshape(want["actions"].shape), want["actions"])
        assert np.array_equal(np.where(plan[:, 2] >= 0, rew_s[np.maximum(plan[:, 2], 0)], 0.0).reshape(B, T), want["rewards"])
    # a restored buffer keeps accepting chunks of its ongoing episodes and new episodes
    from dreamer_v3_b200.utils.episode import Episode
    n0 = new.get_num_timesteps()
    new.add(Episode("fresh", observations=list(np.zeros((4,) + new._obs_shape, new._obs_store.numpy().dtype)),
                    actions=list(np.zeros((3,) + new._act_shape, np.float32)), rewards=[0.0, 1.0, 2.0], is_terminated=True))
    assert new.get_num_timesteps() == n0 + 3 or new.size <= new.capacity   # (an ejection may have made room)


def test_reference_buffer_loads_our_state():
    """Cross-check against the reference's own code where it is available (this container only): the state written by our
    buffer is accepted by the reference's `EpisodeReplayBuffer.set_state` and both then draw identical samples."""
    import importlib.util
    import sys
    ref = "/root/reference/utils/episode_replay_buffer.py"
    if not os.path.exists(ref):
        pytest.skip("reference sources not present on this machine")
    sys.path.insert(0, "/root/reference")
    try:
        for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[m]
        spec = importlib.util.spec_from_file_location("ref_erb", ref)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove("/root/reference")
        for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[m]
    buf = device_buffer(False, _host_index_only=True)
    rb = mod.EpisodeReplayBuffer(capacity=CAPACITY)
    rb.set_state(buf.get_state())
    assert rb.get_num_episodes() == buf.get_num_episodes() and rb.get_num_timesteps() == buf.get_num_timesteps()
    obs_s, act_s, rew_s = buf._obs_store.numpy(), buf._act_store.numpy(), buf._rew_store.numpy()
    for B, T, seed in SAMPLES[:2]:
        np.random.seed(seed)
        want = rb.sample(batch_size_B=B, batch_length_T=T)
        np.random.seed(seed)
        plan = buf.plan(B, T)
        assert np.array_equal(obs_s[plan[:, 0]].reshape(np.asarray(want["obs"]).shape), np.asarray(want["obs"]))
        assert np.array_equal(np.where(plan[:, 2] >= 0, rew_s[np.maximum(plan[:, 2], 0)], 0.0).reshape(B, T), np.asarray(want["rewards"]))

"""Golden vectors of the replay sampler from the REFERENCE'S OWN buffer (utils/episode_replay_buffer.py and
utils/episode.py are pure NumPy and import cleanly): the seeded scenario of tests/_replay_scenario.py is pushed through
`EpisodeReplayBuffer.add`, then `sample` is called under fixed `np.random` seeds.
    python tests/golden/make_replay_golden.py      # needs /root/reference; writes tests/golden/ref_replay.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("DV3_REFERENCE", "/root/reference"))
sys.path.insert(1, os.path.dirname(HERE))
from _replay_scenario import CAPACITY, SAMPLES, make_adds  # noqa: E402
from utils.episode import Episode  # noqa: E402  (reference)
from utils.episode_replay_buffer import EpisodeReplayBuffer  # noqa: E402  (reference)

out = {}
for kind, image in (("vec", False), ("img", True)):
    buf = EpisodeReplayBuffer(capacity=CAPACITY)
    for id_, obs, actions, rewards, term in make_adds(image=image):
        buf.add(Episode(id_, observations=list(obs), actions=list(actions), rewards=list(rewards), is_terminated=term))
    out[f"{kind}/num_episodes"] = np.int64(buf.get_num_episodes())
    out[f"{kind}/num_timesteps"] = np.int64(buf.get_num_timesteps())
    for i, (B, T, seed) in enumerate(SAMPLES):
        np.random.seed(seed)
        s = buf.sample(batch_size_B=B, batch_length_T=T)
        for k, v in s.items():
            out[f"{kind}/s{i}/{k}"] = v
path = os.path.join(HERE, "ref_replay.npz")
np.savez_compressed(path, **out)
print(path, len(out), "arrays", os.path.getsize(path) // 1024, "KiB")

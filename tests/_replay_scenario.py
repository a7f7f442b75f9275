"""Seeded replay scenario shared by tests/golden/make_replay_golden.py (drives the reference's own buffer) and
tests/test_replay.py (drives the oracle restatement and the device buffer): a stream of `add` calls with whole
episodes, chunks of ongoing episodes (same id, first observation repeating the previous chunk's last) and a capacity
small enough that old episodes are ejected, followed by seeded `sample` calls."""
import numpy as np

CAPACITY = 900
SAMPLES = ((4, 8, 11), (16, 64, 12), (3, 5, 13), (16, 64, 14))  # (B, T, np.random seed)


def make_adds(seed=7, image=False, n_adds=60):
    """-> list of (id, observations [n+1, ...], actions [n, A], rewards [n], is_terminated)."""
    rng = np.random.default_rng(seed)
    adds, ongoing, next_id = [], {}, 0
    for step in range(n_adds):
        # only recent episodes continue: the reference's ejection loop (:68-95) indexes past the end of `_indices` when
        # the ejected (oldest) episode owns the newest index entries, which a long-lived ongoing episode would trigger
        ongoing = {k: v for k, v in ongoing.items() if v[1] >= step - 4}
        if ongoing and rng.random() < 0.4:
            id_ = rng.choice(sorted(ongoing))
            first_obs = ongoing.pop(id_)[0]
        else:
            id_, first_obs = f"eps{next_id}", None
            next_id += 1
        n = int(rng.integers(1, 70))
        if image:
            obs = rng.integers(0, 256, size=(n + 1, 8, 8, 3), dtype=np.uint8)
        else:
            obs = rng.standard_normal((n + 1, 3)).astype(np.float32)
        if first_obs is not None:
            obs[0] = first_obs
        actions = rng.uniform(-1, 1, size=(n, 2)).astype(np.float32)
        rewards = rng.standard_normal(n).astype(np.float32)
        terminated = bool(rng.random() < 0.5)
        if not terminated:
            ongoing[id_] = (obs[-1].copy(), step)
        adds.append((id_, obs, actions, rewards, terminated))
    return adds

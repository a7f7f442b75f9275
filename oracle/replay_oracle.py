"""ORACLE (test infrastructure, not product code) -- plain-Python restatement of the reference's replay sampler,
utils/episode_replay_buffer.py:27-211 (+ utils/episode.py for the chunk concatenation).

Parity pin: the reference's buffer is pure NumPy and IS importable in the build container; tests/golden/
make_replay_golden.py runs the reference's own `EpisodeReplayBuffer.add / sample` on a seeded scenario (chunked ongoing
episodes, ejection at capacity) and commits its outputs (tests/golden/ref_replay.npz); tests/test_replay.py holds this
restatement to them bit for bit.  Only tests/ may import this module.
"""
import numpy as np


class ReplayOracle:
    def __init__(self, capacity=10000):
        self.capacity = capacity
        self.eps = []          # dicts: id, obs(list), act(list), rew(list), term, abs index
        self.ejected = 0
        self.indices = []      # (abs episode index, ts) in arrival order  (:44-56)
        self.size = 0

    def _find(self, id_):
        for e in self.eps:
            if e["id"] == id_:
                return e
        return None

    def add(self, id_, observations, actions, rewards, is_terminated):
        n = len(actions)
        self.size += n
        e = self._find(id_)
        if e is not None:  # chunk of an ongoing episode (:36-44, Episode.concat_episode)
            old = len(e["act"])
            assert np.all(np.asarray(observations[0]) == np.asarray(e["obs"][-1]))
            e["obs"] = e["obs"][:-1] + list(observations)
            e["act"] += list(actions)
            e["rew"] += list(rewards)
            e["term"] = e["term"] or bool(is_terminated)
            self.indices += [(e["abs"], old + i) for i in range(n)]
        else:
            e = dict(id=id_, obs=list(observations), act=list(actions), rew=list(rewards), term=bool(is_terminated),
                     abs=len(self.eps) + self.ejected)
            self.eps.append(e)
            self.indices += [(e["abs"], i) for i in range(n)]
        while self.size > self.capacity:  # :60-96: the oldest episode leaves, with all of its index entries
            out = self.eps.pop(0)
            self.size -= len(out["act"])
            self.indices = [it for it in self.indices if it[0] != out["abs"]]
            self.ejected += 1

    def sample(self, B, T):
        obs = [[] for _ in range(B)]
        act = [[] for _ in range(B)]
        rew = [[] for _ in range(B)]
        first = np.zeros((B, T), bool)
        last = np.zeros((B, T), bool)
        term = np.zeros((B, T), bool)
        draws, cursor, b, fill = [], 0, 0, 0
        while b < B:
            if len(draws) <= cursor:
                draws.extend(list(np.random.randint(0, len(self.indices), size=B * 10)))  # :118-121
            abs_idx, ts = self.indices[draws[cursor]]
            cursor += 1
            e = self.eps[abs_idx - self.ejected]
            first[b, fill] = True
            if len(rew[b]) == 0:
                rew[b].append(0.0 if ts == 0 else e["rew"][ts - 1])  # :134-138
            else:
                ts = 0  # :139-141
                rew[b].append(0.0)
            obs[b] += e["obs"][ts:]
            act[b] += e["act"][ts:] + [e["act"][-1]]  # :157-158
            rew[b] += e["rew"][ts:]
            fill = min(len(obs[b]), T)
            last[b, fill - 1] = True
            if e["term"] and fill == len(obs[b]):
                term[b, fill - 1] = True
            if fill == T:
                obs[b], act[b], rew[b] = obs[b][:T], act[b][:T], rew[b][:T]
                b, fill = b + 1, 0
        return {"obs": np.array(obs), "actions": np.array(act), "rewards": np.array(rew), "is_first": first,
                "is_last": last, "is_terminated": term}

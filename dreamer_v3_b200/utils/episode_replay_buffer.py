"""EpisodeReplayBuffer with the reference's surface (utils/episode_replay_buffer.py:9-238: `add`, `sample`,
`get_num_episodes`, `get_num_timesteps`, `get_sampled_timesteps`), re-designed for the GPU: the timesteps live in HBM in
fixed-size *pages* owned by episodes, `sample()` builds a 16-byte-per-timestep gather plan on the host with the
reference's index arithmetic and ONE CUDA kernel (`dv3_replay_gather`, csrc/dv3_replay.cu) packs the `[B, T, ...]`
sample dict on the device - image frames stay uint8 in the store (1 B / pixel) and are normalised in the gather.

Semantics kept from the reference (line numbers of utils/episode_replay_buffer.py):
  * every stored timestep (episode, ts), ts < len(episode), is equally likely to start a row (:118-126); the order of that
    index space is the arrival order of the added chunks (:44-56), so the same `np.random.randint` stream selects the
    same timesteps;
  * a row that runs past its episode's end continues with timestep 0 of the episode of the next drawn index (:139-141);
  * the first reward of a row is the one that led to its first observation, 0.0 at an episode start (:134-138, :141);
  * the action at an episode's final observation repeats the last action (:157-158);
  * is_first at every segment start, is_last at every segment end, is_terminated only where a terminated episode's
    final observation made it into the row (:128, :166-170);
  * capacity counts timesteps; the oldest episodes are ejected first (:60-96).
There is no CPU fallback: the store and the returned tensors are CUDA tensors.
"""
import ctypes as C
from collections import deque

import numpy as np
import torch

from .. import _lib
from .episode import Episode

PAGE = 64  # timesteps per page


class _Stored:
    __slots__ = ("id_", "pages", "length", "is_terminated")

    def __init__(self, id_):
        self.id_, self.pages, self.length, self.is_terminated = id_, [], 0, False


class EpisodeReplayBuffer:
    def __init__(self, capacity: int = 10000, device=None, normalize_uint8=True, _host_index_only=False):
        # _host_index_only (tests of the host logic on a machine without a GPU): bookkeeping and `plan()` work, the store is
        # a host tensor nobody gathers from - `sample()` raises
        self._host_only = bool(_host_index_only)
        if not self._host_only and not torch.cuda.is_available():
            raise RuntimeError("EpisodeReplayBuffer keeps its store in HBM; no CUDA device found (there is no CPU fallback)")
        self.lib = _lib.load()
        self.capacity = int(capacity)
        if self._host_only:
            self.device = torch.device("cpu")
        else:
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.normalize_uint8 = bool(normalize_uint8)
        self.episodes = deque()                # _Stored, oldest first
        self.episode_id_to_index = {}          # id_ -> absolute episode index (ejected ones counted)
        self._num_episodes_ejected = 0
        self._chunks = []                      # arrival-ordered (abs episode index, ts0, count): the reference's `_indices`
        self._chunk_cum = np.zeros(1, np.int64)
        self.size = 0
        self.sampled_timesteps = 0
        self._obs_store = self._act_store = self._rew_store = None
        self._free_pages = []
        self._num_pages = 0
        self._tab = None

    # ------------------------------------------------------------------ device store
    def _alloc(self, obs0, act0):
        obs0, act0 = np.asarray(obs0), np.asarray(act0, dtype=np.float32)
        self._obs_shape = tuple(obs0.shape)
        self._obs_u8 = obs0.dtype == np.uint8
        self._act_shape = tuple(act0.shape)
        self._obs_elems = int(np.prod(self._obs_shape)) if self._obs_shape else 1
        self._A = int(np.prod(self._act_shape)) if self._act_shape else 1
        # every episode wastes < 1 page (+1 slot for its final observation); short episodes are the worst case
        # grown on demand (x1.5) up to what the capacity can need
        self._num_pages = min(2 * ((self.capacity + PAGE - 1) // PAGE) + 64, 1024)
        slots = self._num_pages * PAGE
        self._obs_store = torch.empty((slots, self._obs_elems), dtype=torch.uint8 if self._obs_u8 else torch.float32, device=self.device)
        self._act_store = torch.zeros((slots, self._A), dtype=torch.float32, device=self.device)
        self._rew_store = torch.zeros((slots,), dtype=torch.float32, device=self.device)
        self._free_pages = list(range(self._num_pages - 1, -1, -1))

    def _slots(self, st, ts0, count):
        """Store slots of timesteps ts0 .. ts0+count-1 of a stored episode (allocating pages as needed)."""
        ts = np.arange(ts0, ts0 + count)
        need = (ts0 + count + PAGE - 1) // PAGE
        while len(st.pages) < need:
            if not self._free_pages:
                self._grow()
            st.pages.append(self._free_pages.pop())
        return torch.from_numpy(np.asarray(st.pages, np.int64)[ts // PAGE] * PAGE + ts % PAGE).to(self.device)

    def _grow(self):
        more = max(64, self._num_pages // 2)
        for name in ("_obs_store", "_act_store", "_rew_store"):
            old = getattr(self, name)
            new = torch.empty((old.shape[0] + more * PAGE,) + tuple(old.shape[1:]), dtype=old.dtype, device=self.device)
            new[:old.shape[0]] = old
            setattr(self, name, new)
        self._free_pages = list(range(self._num_pages + more - 1, self._num_pages - 1, -1)) + self._free_pages
        self._num_pages += more

    def _upload(self, st, obs, actions, rewards, obs_ts0, ts0):
        """obs -> observation indices obs_ts0.., actions / rewards -> timestep indices ts0.. of the stored episode."""
        if len(obs):
            o = torch.from_numpy(np.ascontiguousarray(np.asarray(obs)).reshape(len(obs), self._obs_elems))
            self._obs_store[self._slots(st, obs_ts0, len(obs))] = o.to(self.device, self._obs_store.dtype, non_blocking=True)
        if len(actions):
            idx = self._slots(st, ts0, len(actions))
            a = torch.from_numpy(np.asarray(actions, dtype=np.float32).reshape(len(actions), self._A))
            self._act_store[idx] = a.to(self.device, non_blocking=True)
            self._rew_store[idx] = torch.from_numpy(np.asarray(rewards, dtype=np.float32)).to(self.device, non_blocking=True)

    # ------------------------------------------------------------------ reference API
    def add(self, episodes):
        """:27-96. Episodes (or chunks of ongoing ones, same `id_`) are copied into the device store."""
        if isinstance(episodes, Episode) or not isinstance(episodes, (list, tuple)):
            episodes = [episodes]
        for eps in episodes:
            n = len(eps)
            if self._obs_store is None:
                self._alloc(eps.observations[0], eps.actions[0] if n else 0.0)
            self.size += n
            if eps.id_ in self.episode_id_to_index:
                abs_idx = self.episode_id_to_index[eps.id_]
                st = self.episodes[abs_idx - self._num_episodes_ejected]
                if st.is_terminated:
                    raise AssertionError("chunk added to a terminated episode")
                old = st.length
                # the chunk's first observation repeats the stored last one (:Episode.concat_episode)
                self._upload(st, eps.observations[1:], eps.actions, eps.rewards, old + 1, old)
                self._chunks.append((abs_idx, old, n))
            else:
                st = _Stored(eps.id_)
                self.episodes.append(st)
                abs_idx = len(self.episodes) - 1 + self._num_episodes_ejected
                self.episode_id_to_index[eps.id_] = abs_idx
                self._upload(st, eps.observations, eps.actions, eps.rewards, 0, 0)
                self._chunks.append((abs_idx, 0, n))
            st.length += n
            st.is_terminated = st.is_terminated or bool(eps.is_terminated)
            while self.size > self.capacity:
                ej = self.episodes.popleft()
                self.size -= ej.length
                ej_idx = self.episode_id_to_index.pop(ej.id_)
                self._free_pages.extend(ej.pages)
                self._chunks = [c for c in self._chunks if c[0] != ej_idx]
                self._num_episodes_ejected += 1
        self._chunk_cum = np.concatenate([[0], np.cumsum([c[2] for c in self._chunks], dtype=np.int64)]).astype(np.int64)
        self._tab = None

    def _tables(self):
        """Episode / chunk tables for `dv3_replay_plan` (rebuilt after `add`)."""
        if self._tab is None:
            eps = list(self.episodes)
            off = np.zeros(len(eps) + 1, np.int64)
            np.cumsum([len(e.pages) for e in eps], out=off[1:])
            self._tab = dict(
                ep_len=np.asarray([e.length for e in eps], np.int32),
                ep_term=np.asarray([e.is_terminated for e in eps], np.uint8),
                ep_page_off=off[:-1].copy(),
                pages=np.asarray([p for e in eps for p in e.pages], np.int32),
                chunk_ep=np.asarray([c[0] - self._num_episodes_ejected for c in self._chunks], np.int32),
                chunk_ts0=np.asarray([c[1] for c in self._chunks], np.int32),
                chunk_cum=np.ascontiguousarray(self._chunk_cum[:-1]),
            )
        return self._tab

    def plan(self, batch_size_B, batch_length_T):
        """The reference's sampling loop (:98-189) as index arithmetic (native: `dv3_replay_plan`): int32 [B*T, 4] = obs
        slot, action slot, reward slot (-1: 0.0), flags (1 is_first | 2 is_last | 4 is_terminated). Draws from `np.random`
        exactly as the reference does (`randint(0, n, size=B*10)` blocks, one index consumed per segment)."""
        B, T = int(batch_size_B), int(batch_length_T)
        n_idx = int(self._chunk_cum[-1])
        if n_idx == 0:
            raise ValueError("cannot sample from an empty buffer")
        t = self._tables()
        plan = np.empty((B * T, 4), np.int32)
        draws = np.random.randint(0, n_idx, size=B * 10).astype(np.int64)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        while True:
            rc = self.lib.dv3_replay_plan(B, T, PAGE, ptr(t["chunk_cum"]), ptr(t["chunk_ep"]), ptr(t["chunk_ts0"]), len(t["chunk_ep"]),
                                          ptr(t["ep_len"]), ptr(t["ep_term"]), ptr(t["ep_page_off"]), ptr(t["pages"]),
                                          ptr(draws), len(draws), ptr(plan))
            if rc >= 0:
                return plan
            if rc < -(1 << 40):
                raise RuntimeError("dv3_replay_plan rejected its arguments")
            draws = np.concatenate([draws, np.random.randint(0, n_idx, size=B * 10).astype(np.int64)])

    def sample(self, batch_size_B: int = 16, batch_length_T: int = 64, out=None):
        """:98-211 -> dict of float32 CUDA tensors obs [B,T,...], actions [B,T(,A...)], rewards, is_first, is_last,
        is_terminated [B,T]. `out` may hold preallocated tensors (e.g. the engine's in.obs / in.actions / in.rewards /
        in.is_first / in.is_terminated) to gather straight into."""
        if self._host_only:
            raise RuntimeError("the sample is gathered by a CUDA kernel; there is no CPU fallback")
        B, T = int(batch_size_B), int(batch_length_T)
        plan = torch.from_numpy(self.plan(B, T)).to(self.device, non_blocking=True)
        out = dict(out or {})
        shapes = {"obs": (B, T) + self._obs_shape, "actions": (B, T) + self._act_shape, "rewards": (B, T),
                  "is_first": (B, T), "is_last": (B, T), "is_terminated": (B, T)}
        for k, shp in shapes.items():
            t = out.get(k)
            if t is None:
                out[k] = torch.empty(shp, dtype=torch.float32, device=self.device)
            elif not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == int(np.prod(shp))):
                raise ValueError(f"out[{k!r}] must be a contiguous float32 CUDA tensor of {int(np.prod(shp))} elements")
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        rc = self.lib.dv3_replay_gather(
            p(self._obs_store), int(self._obs_u8), int(self.normalize_uint8), self._obs_elems, p(self._act_store), self._A,
            p(self._rew_store), p(plan), B * T, p(out["obs"]), p(out["actions"]), p(out["rewards"]), p(out["is_first"]),
            p(out["is_last"]), p(out["is_terminated"]), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError(f"dv3_replay_gather failed ({rc})")
        self._last_plan = plan  # keeps the plan alive until the kernel has run
        self.sampled_timesteps += B * T
        return out

    # ------------------------------------------------------------------ checkpointing (run_experiment.py:228, :708)
    def _download(self, st):
        """A stored episode's pages read back from the store as a host `Episode` (observations in their original dtype)."""
        n = st.length
        pg = np.asarray(st.pages, np.int64)
        ts = np.arange(n + 1)
        slots = torch.from_numpy(pg[ts // PAGE] * PAGE + ts % PAGE).to(self.device)
        obs = self._obs_store[slots].cpu().numpy().reshape((n + 1,) + self._obs_shape)
        act = self._act_store[slots[:n]].cpu().numpy().reshape((n,) + self._act_shape)
        rew = self._rew_store[slots[:n]].cpu().numpy()
        return Episode(st.id_, observations=list(obs), actions=list(act), rewards=list(rew), is_terminated=st.is_terminated)

    def get_state(self):
        """utils/episode_replay_buffer.py:240-248, same keys in the same order (`np.savez(.../buffer, state=...)` of
        run_experiment.py:708 stores it as an object array): the episodes are read back from the HBM pages; `_indices` is the
        reference's arrival-ordered list of (absolute episode index, timestep) pairs, expanded from the chunk table."""
        eps = [self._download(st).get_state() for st in self.episodes]
        indices = [(int(c[0]), int(c[1]) + i) for c in self._chunks for i in range(int(c[2]))]
        state = list({
            "episodes": eps,
            "episode_id_to_index": list(self.episode_id_to_index.items()),
            "_num_episodes_ejected": self._num_episodes_ejected,
            "_indices": indices,
            "size": self.size,
            "sampled_timesteps": self.sampled_timesteps,
        }.items())
        out = np.empty((len(state), 2), dtype=object)   # np.array(list(...)) of the reference, without ragged broadcasting
        for i, (k, v) in enumerate(state):
            out[i, 0], out[i, 1] = k, v
        return out

    def set_state(self, state):
        """utils/episode_replay_buffer.py:250-264: positional keys as written by `get_state` (also of the reference's own
        buffer). The episodes are uploaded into fresh pages, `_indices` is run-length encoded back into the chunk table."""
        names = ("episodes", "episode_id_to_index", "_num_episodes_ejected", "_indices", "size", "sampled_timesteps")
        for i, k in enumerate(names):
            assert state[i][0] == k, (i, state[i][0], k)
        self.episodes = deque()
        self._chunks = []
        self._obs_store = self._act_store = self._rew_store = None
        self._free_pages, self._num_pages = [], 0
        self.episode_id_to_index = {k: int(v) for k, v in dict(state[1][1]).items()}
        self._num_episodes_ejected = int(state[2][1])
        for eps_state in state[0][1]:
            eps = Episode.from_state(eps_state)
            n = len(eps)
            if self._obs_store is None:
                self._alloc(eps.observations[0], eps.actions[0] if n else 0.0)
            st = _Stored(eps.id_)
            self._upload(st, eps.observations, eps.actions, eps.rewards, 0, 0)
            st.length, st.is_terminated = n, bool(eps.is_terminated)
            self.episodes.append(st)
        chunks = []
        for e_idx, ts in state[3][1]:
            e_idx, ts = int(e_idx), int(ts)
            if chunks and chunks[-1][0] == e_idx and chunks[-1][1] + chunks[-1][2] == ts:
                chunks[-1][2] += 1
            else:
                chunks.append([e_idx, ts, 1])
        self._chunks = [tuple(c) for c in chunks]
        self._chunk_cum = np.concatenate([[0], np.cumsum([c[2] for c in self._chunks], dtype=np.int64)]).astype(np.int64)
        self.size = int(state[4][1])
        self.sampled_timesteps = int(state[5][1])
        self._tab = None

    def get_num_episodes(self):
        return len(self.episodes)

    def get_num_timesteps(self):
        return int(self._chunk_cum[-1])

    def get_sampled_timesteps(self):
        return self.sampled_timesteps

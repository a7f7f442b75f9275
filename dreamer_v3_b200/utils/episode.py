"""Episode container with the reference's surface (utils/episode.py): `observations` holds len+1 entries (the reset
observation first), `actions` / `rewards` len entries; chunks of an ongoing episode share `id_` and are joined with
`concat_episode` (the chunk's first observation repeats the previous chunk's last)."""
import uuid

import numpy as np


class Episode:
    def __init__(self, id_=None, *, observations=None, actions=None, rewards=None, states=None, is_terminated=False,
                 render_images=None):
        self.id_ = id_ or uuid.uuid4().hex
        self.observations = list(observations) if observations is not None else []
        self.actions = list(actions) if actions is not None else []
        self.rewards = list(rewards) if rewards is not None else []
        self.states = states
        self.is_terminated = bool(is_terminated)
        self.render_images = list(render_images) if render_images is not None else []

    def add_initial_observation(self, *, initial_observation, initial_state=None, initial_render_image=None):
        if self.is_terminated or len(self.observations) != 0:
            raise AssertionError("initial observation of a started / terminated episode")
        self.observations.append(initial_observation)
        self.states = initial_state
        if initial_render_image is not None:
            self.render_images.append(initial_render_image)

    def add_timestep(self, observation, action, reward, *, state=None, is_terminated=False, render_image=None):
        if self.is_terminated:
            raise AssertionError("episode already terminated")
        self.observations.append(observation)
        self.actions.append(action)
        self.rewards.append(reward)
        self.states = state
        if render_image is not None:
            self.render_images.append(render_image)
        self.is_terminated = bool(is_terminated)
        self.validate()

    def concat_episode(self, episode_chunk):
        if episode_chunk.id_ != self.id_ or self.is_terminated:
            raise AssertionError("chunk does not continue this episode")
        episode_chunk.validate()
        if not np.all(np.asarray(episode_chunk.observations[0]) == np.asarray(self.observations[-1])):
            raise AssertionError("chunk does not start at this episode's last observation")
        self.observations = list(self.observations[:-1]) + list(episode_chunk.observations)
        self.actions = list(self.actions) + list(episode_chunk.actions)
        self.rewards = list(self.rewards) + list(episode_chunk.rewards)
        self.states = episode_chunk.states
        self.is_terminated = self.is_terminated or episode_chunk.is_terminated
        self.validate()

    def validate(self):
        if not (len(self.observations) == len(self.rewards) + 1 == len(self.actions) + 1):
            raise AssertionError("observations must hold one more entry than actions / rewards")

    def get_return(self):
        return sum(self.rewards)

    def get_state(self):
        """utils/episode.py:110-118: the same key order (`from_state` and the buffer's `set_state` index it positionally)."""
        return list({
            "id_": self.id_,
            "observations": self.observations,
            "actions": self.actions,
            "rewards": self.rewards,
            "states": self.states,
            "is_terminated": self.is_terminated,
        }.items())

    @staticmethod
    def from_state(state):
        """utils/episode.py:120-128."""
        eps = Episode(id_=state[0][1])
        eps.observations = list(state[1][1])
        eps.actions = list(state[2][1])
        eps.rewards = list(state[3][1])
        eps.states = state[4][1]
        eps.is_terminated = bool(state[5][1])
        return eps

    def __len__(self):
        if len(self.observations) == 0:
            raise AssertionError("episode has no initial observation yet")
        return len(self.observations) - 1

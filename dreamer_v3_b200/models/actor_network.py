"""ActorNetwork surface (models/actor_network.py:16-118): weights, optimizer info and the percentile EMAs live in the engine."""
from .world_model import _OptimizerInfo, _Variables


class ActorNetwork:
    def __init__(self, *, action_space, model_dimension="XS", return_normalization_decay=0.99, engine_getter=None):
        self.action_space = action_space
        self.model_dimension = model_dimension
        self.return_normalization_decay = return_normalization_decay
        self._e = engine_getter
        self.optimizer = _OptimizerInfo(engine_getter, "actor", 3e-5, 1e-5)  # actor_network.py:60
        self.trainable_variables = _Variables(engine_getter, "actor")

    @property
    def ema_value_target_pct5(self):
        return self._e().tensor("ac.pct_ema")[0]

    @property
    def ema_value_target_pct95(self):
        return self._e().tensor("ac.pct_ema")[1]

    def get_weights(self):
        """`model.get_weights()` as saved by run_experiment.py:705 (order: utils/checkpoint.py)."""
        from ..utils import checkpoint
        return checkpoint.get_weights(self._e(), "actor")

    def set_weights(self, weights):
        from ..utils import checkpoint
        checkpoint.set_weights(self._e(), "actor", weights)

"""CriticNetwork surface (models/critic_network.py:15-100)."""
from .world_model import _OptimizerInfo, _Variables


class CriticNetwork:
    def __init__(self, *, model_dimension="XS", num_buckets=255, lower_bound=-20.0, upper_bound=20.0, ema_decay=0.98,
                 engine_getter=None):
        assert (num_buckets, lower_bound, upper_bound) == (255, -20.0, 20.0)
        self.model_dimension = model_dimension
        self.ema_decay = ema_decay
        self._e = engine_getter
        self.optimizer = _OptimizerInfo(engine_getter, "critic", 3e-5, 1e-5)  # critic_network.py:58
        self.trainable_variables = _Variables(engine_getter, "critic")
        self.ema_variables = _Variables(engine_getter, "critic_ema")

    def init_ema(self):  # critic_network.py:84-91
        self._e().critic_init_ema()

    def update_ema(self):  # critic_network.py:93-100
        self._e().critic_update_ema(self.ema_decay)

    def get_weights(self):
        """`model.get_weights()` as saved by run_experiment.py:705 (order: utils/checkpoint.py)."""
        from ..utils import checkpoint
        return checkpoint.get_weights(self._e(), "critic")

    def set_weights(self, weights):
        from ..utils import checkpoint
        checkpoint.set_weights(self._e(), "critic", weights)

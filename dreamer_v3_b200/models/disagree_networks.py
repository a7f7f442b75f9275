"""DisagreeNetworks surface (models/disagree_networks.py:12-95): the Plan2Explore ensemble that predicts the next z from
stop_gradient([h, z, a]); its disagreement is the intrinsic reward of the dreamed steps. The nets, their loss and their Adam run
inside libdv3 (csrc/dv3_dream.cu `disagree_forward` / `disagree_loss_backward`, csrc/dv3_disagree.cu); this class is the view
the reference's callers use (`.optimizer`, `.trainable_variables`, `get_weights()` for run_experiment.py:495,702)."""
from ..utils.model_dimensions import get_num_curiosity_nets
from .world_model import _OptimizerInfo, _Variables


class DisagreeNetworks:
    def __init__(self, *, num_networks=None, model_dimension="XS", intrinsic_rewards_scale=0.1, engine_getter=None):
        self.model_dimension = model_dimension
        self.num_networks = get_num_curiosity_nets(model_dimension, override=num_networks)
        self.intrinsic_rewards_scale = intrinsic_rewards_scale
        self._e = engine_getter
        self.optimizer = _OptimizerInfo(engine_getter, "disagree", 1e-4, 1e-5)  # disagree_networks.py:43
        self.trainable_variables = _Variables(engine_getter, "disagree")

    def forward_train_outs(self):
        """{"z_predicted_probs_N_HxB": [N x [(H+1)*B, Zc, Zk]]} of the last dream (disagree_networks.py:78-95)."""
        return {"z_predicted_probs_N_HxB": list(self._e().tensor("disagree.z_predicted_probs"))}

    def get_weights(self):
        from ..utils import checkpoint
        return checkpoint.get_weights(self._e(), "disagree")

    def set_weights(self, weights):
        from ..utils import checkpoint
        checkpoint.set_weights(self._e(), "disagree", weights)

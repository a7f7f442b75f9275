"""Stand-ins for the reference's Keras component constructors that callers pass into WorldModel
(run_experiment.py:185-201). They only carry configuration; the weights live in the CUDA engine."""
from ...utils.model_dimensions import get_cnn_multiplier


class MLP:
    """Vector-observation encoder (reference components/mlp.py:17-67 used as encoder, run_experiment.py:190-193)."""

    def __init__(self, *, model_dimension="XS", num_dense_layers=None, dense_hidden_units=None, output_layer_size=None,
                 trainable=True):
        if num_dense_layers is not None or dense_hidden_units is not None or output_layer_size:
            raise NotImplementedError("the engine builds the encoder MLP from model_dimension only")
        self.model_dimension = model_dimension
        self.image = False


class CNNAtari:
    """Image encoder descriptor (reference components/cnn_atari.py:14-76)."""

    def __init__(self, *, model_dimension="XS", cnn_multiplier=None):
        self.model_dimension = model_dimension
        self.cnn_multiplier = get_cnn_multiplier(model_dimension, override=cnn_multiplier)
        self.image = True


class ConvTransposeAtari:
    """Image decoder descriptor (reference components/conv_transpose_atari.py:19-92)."""

    def __init__(self, *, model_dimension="XS", cnn_multiplier=None, gray_scaled=True):
        self.model_dimension = model_dimension
        self.cnn_multiplier = get_cnn_multiplier(model_dimension, override=cnn_multiplier)
        self.gray_scaled = gray_scaled
        self.image = True
        self.img_c = 1 if gray_scaled else 3


class VectorDecoder:
    """Vector decoder descriptor (reference components/vector_decoder.py:15-33)."""

    def __init__(self, *, model_dimension="XS", observation_space):
        assert len(observation_space.shape) == 1
        self.model_dimension = model_dimension
        self.observation_space = observation_space
        self.obs_dim = int(observation_space.shape[0])
        self.image = False

"""Model size table (nano ... XL) with the reference's accessor API.

Mirrors the public surface of the reference `utils/model_dimensions.py:9-90`
(`get_cnn_multiplier`, `get_dense_hidden_units`, `get_gru_units`,
`get_num_z_categoricals`, `get_num_z_classes`, `get_num_curiosity_nets`,
`get_num_dense_layers`; every accessor takes `(model_dimension, override=None)`).
The numbers are the DreamerV3 paper's appendix-B table as used by the reference.
"""
from collections import namedtuple

SizeRow = namedtuple(
    "SizeRow", "cnn_multiplier dense_units gru_units z_categoricals z_classes dense_layers curiosity_nets"
)

# size -> (cnn mult, dense units, GRU units, z categoricals, z classes, MLP layers, curiosity nets)
SIZE_TABLE = {
    "nano": SizeRow(2, 16, 16, 4, 4, 1, 8),
    "micro": SizeRow(4, 32, 32, 8, 8, 1, 8),
    "XXS": SizeRow(16, 128, 128, 32, 32, 1, 8),
    "XS": SizeRow(24, 256, 256, 32, 32, 1, 8),
    "S": SizeRow(32, 512, 512, 32, 32, 2, 8),
    "M": SizeRow(48, 640, 1024, 32, 32, 3, 8),
    "L": SizeRow(64, 768, 2048, 32, 32, 4, 8),
    "XL": SizeRow(96, 1024, 4096, 32, 32, 5, 8),
}
_ALLOWED_DIMS = list(SIZE_TABLE)


def _lookup(model_dimension, field, override):
    if override is not None:
        return override
    assert model_dimension in SIZE_TABLE, f"unknown model_dimension {model_dimension!r}"
    return getattr(SIZE_TABLE[model_dimension], field)


def get_cnn_multiplier(model_dimension, override=None):
    return _lookup(model_dimension, "cnn_multiplier", override)


def get_dense_hidden_units(model_dimension, override=None):
    return _lookup(model_dimension, "dense_units", override)


def get_gru_units(model_dimension, override=None):
    return _lookup(model_dimension, "gru_units", override)


def get_num_z_categoricals(model_dimension, override=None):
    return _lookup(model_dimension, "z_categoricals", override)


def get_num_z_classes(model_dimension, override=None):
    return _lookup(model_dimension, "z_classes", override)


def get_num_curiosity_nets(model_dimension, override=None):
    return _lookup(model_dimension, "curiosity_nets", override)


def get_num_dense_layers(model_dimension, override=None):
    return _lookup(model_dimension, "dense_layers", override)

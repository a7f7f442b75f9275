"""Engine configuration and the parameter table of the DreamerV3 training hot path.

The table below is the single Python statement of which weights exist, in which
optimizer group, in which order and with which shape; the CUDA engine
(`csrc/dv3_engine.cu`, `Engine::build_param_table`) builds the same table and
`tests/test_spec.py` / `tests/test_gpu_engine.py` check that both agree.

Shapes follow the reference's Keras layers (all `[in, out]` row-major):
  * MLP hidden layer = Dense(no bias) -> LayerNorm -> SiLU   (components/mlp.py:43-84)
  * SequenceModel = pre-GRU MLP(1 layer) + GRU               (components/sequence_model.py:37-52)
  * RepresentationLayer = Dense(D -> Zc*Zk, bias)            (components/representation_layer.py:41-46)
  * reward / critic head = zero-initialised Dense(255)       (components/reward_predictor_layer.py:49-59)
  * CNNAtari / ConvTransposeAtari                            (components/cnn_atari.py:33-76,
                                                              components/conv_transpose_atari.py:41-92)
Concatenation orders: posterior input = [encoder_out, h] (world_model.py:259),
heads / actor / critic input = [h, z] (e.g. reward_predictor.py:50), pre-GRU input
= [z, a] (sequence_model.py:67).
"""
from dataclasses import dataclass, field, asdict
from typing import List, Optional, Tuple

import numpy as np

from .utils.model_dimensions import (
    get_cnn_multiplier,
    get_dense_hidden_units,
    get_gru_units,
    get_num_curiosity_nets,
    get_num_dense_layers,
    get_num_z_categoricals,
    get_num_z_classes,
)

NUM_BUCKETS = 255  # reward_predictor.py:20-22, critic_network.py:20-22
GROUPS = ("world_model", "actor", "critic", "critic_ema", "disagree")


class Discrete:
    """Minimal stand-in for gymnasium.spaces.Discrete (gymnasium is not a dependency)."""

    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


class Box:
    """Minimal stand-in for gymnasium.spaces.Box."""

    def __init__(self, low=-1.0, high=1.0, shape=(1,)):
        self.low, self.high = low, high
        self.shape = tuple(int(s) for s in shape)


def action_space_info(space) -> Tuple[bool, int]:
    """(is_discrete, A) from a gymnasium-like space (duck typed)."""
    if hasattr(space, "n"):
        return True, int(space.n)
    return False, int(np.prod(space.shape))


@dataclass
class EngineConfig:
    model_dimension: str = "XS"
    G: int = 256  # GRU units
    D: int = 256  # dense hidden units
    L: int = 1  # MLP layers of encoder/decoder/reward/continue/actor/critic
    cnn_mult: int = 24
    Zc: int = 32  # categoricals
    Zk: int = 32  # classes per categorical
    A: int = 2
    discrete: bool = True
    image_obs: bool = False
    obs_dim: int = 4  # vector observations
    img_hw: int = 64
    img_c: int = 3
    B: int = 16
    T: int = 64
    H: int = 15
    symlog_obs: bool = True
    # Plan2Explore curiosity (models/disagree_networks.py): 0 = off (use_curiosity=False), else the number of disagree nets
    num_curiosity_nets: int = 0
    intrinsic_rewards_scale: float = 0.1

    @property
    def Z(self):
        return self.Zc * self.Zk

    @property
    def N(self):
        return self.B * self.T

    @property
    def E(self):  # encoder output width
        return 4 * 4 * 8 * self.cnn_mult if self.image_obs else self.D

    @property
    def obs_flat(self):
        return self.img_hw * self.img_hw * self.img_c if self.image_obs else self.obs_dim

    @staticmethod
    def make(model_dimension="XS", *, action_space=None, A=None, discrete=None, image_obs=False,
             obs_dim=4, img_c=3, B=16, T=64, H=15, symlog_obs=None, num_gru_units=None, use_curiosity=False,
             num_curiosity_nets=None, intrinsic_rewards_scale=0.1):
        if action_space is not None:
            discrete, A = action_space_info(action_space)
        assert A is not None and discrete is not None
        if symlog_obs is None:
            symlog_obs = not image_obs  # run_experiment.py:164-166
        return EngineConfig(
            model_dimension=model_dimension,
            G=get_gru_units(model_dimension, override=num_gru_units),
            D=get_dense_hidden_units(model_dimension),
            L=get_num_dense_layers(model_dimension),
            cnn_mult=get_cnn_multiplier(model_dimension),
            Zc=get_num_z_categoricals(model_dimension),
            Zk=get_num_z_classes(model_dimension),
            A=int(A), discrete=bool(discrete), image_obs=bool(image_obs), obs_dim=int(obs_dim),
            img_c=int(img_c), B=int(B), T=int(T), H=int(H), symlog_obs=bool(symlog_obs),
            num_curiosity_nets=(get_num_curiosity_nets(model_dimension, override=num_curiosity_nets) if use_curiosity else 0),
            intrinsic_rewards_scale=float(intrinsic_rewards_scale),
        )

    def asdict(self):
        return asdict(self)


@dataclass
class ParamInfo:
    group: str
    name: str
    shape: Tuple[int, ...]
    init: str  # glorot | orthogonal | zeros | ones
    fan: Optional[Tuple[int, int]] = None  # (fan_in, fan_out) for glorot

    @property
    def size(self):
        return int(np.prod(self.shape))


def _mlp(out: List[ParamInfo], group, prefix, in_dim, D, L):
    k = in_dim
    for l in range(L):
        out.append(ParamInfo(group, f"{prefix}.mlp{l}.w", (k, D), "glorot", (k, D)))
        out.append(ParamInfo(group, f"{prefix}.mlp{l}.g", (D,), "ones"))
        out.append(ParamInfo(group, f"{prefix}.mlp{l}.b", (D,), "zeros"))
        k = D
    return k


def _dense(out, group, prefix, in_dim, out_dim, init="glorot"):
    out.append(ParamInfo(group, f"{prefix}.w", (in_dim, out_dim), init, (in_dim, out_dim)))
    out.append(ParamInfo(group, f"{prefix}.bias", (out_dim,), "zeros"))


def param_spec(cfg: EngineConfig) -> List[ParamInfo]:
    G, D, L, Z, A, m = cfg.G, cfg.D, cfg.L, cfg.Z, cfg.A, cfg.cnn_mult
    p: List[ParamInfo] = []
    wm = "world_model"
    # --- encoder -------------------------------------------------------------------
    if cfg.image_obs:
        cin = cfg.img_c
        for l in range(4):
            cout = m * (1 << l)
            p.append(ParamInfo(wm, f"enc.conv{l}.w", (4, 4, cin, cout), "glorot", (16 * cin, 16 * cout)))
            p.append(ParamInfo(wm, f"enc.conv{l}.g", (cout,), "ones"))
            p.append(ParamInfo(wm, f"enc.conv{l}.b", (cout,), "zeros"))
            cin = cout
    else:
        _mlp(p, wm, "enc", cfg.obs_dim, D, L)
    E = cfg.E
    # --- posterior / prior / initial state / sequence model ---------------------------
    _mlp(p, wm, "post", E + G, D, 1)
    _dense(p, wm, "post.repr", D, Z)
    _mlp(p, wm, "dyn", G, D, 1)
    _dense(p, wm, "dyn.repr", D, Z)
    p.append(ParamInfo(wm, "initial_h", (G,), "zeros"))
    _mlp(p, wm, "seq.pre", Z + A, D, 1)
    p.append(ParamInfo(wm, "seq.gru.kernel", (D, 3 * G), "glorot", (D, 3 * G)))
    p.append(ParamInfo(wm, "seq.gru.recurrent", (G, 3 * G), "orthogonal"))
    p.append(ParamInfo(wm, "seq.gru.bias", (2, 3 * G), "zeros"))
    # --- heads ---------------------------------------------------------------------
    _mlp(p, wm, "rew", G + Z, D, L)
    _dense(p, wm, "rew.head", D, NUM_BUCKETS, init="zeros")
    _mlp(p, wm, "cont", G + Z, D, L)
    _dense(p, wm, "cont.out", D, 1)
    if cfg.image_obs:
        c0 = 8 * m
        _dense(p, wm, "dec.dense", G + Z, 16 * c0)
        cin = c0
        for l in range(3):
            cout = cin // 2
            # Keras Conv2DTranspose kernel layout [kh, kw, out_ch, in_ch].
            p.append(ParamInfo(wm, f"dec.convT{l}.w", (4, 4, cout, cin), "glorot", (16 * cout, 16 * cin)))
            p.append(ParamInfo(wm, f"dec.convT{l}.g", (cout,), "ones"))
            p.append(ParamInfo(wm, f"dec.convT{l}.b", (cout,), "zeros"))
            cin = cout
        p.append(ParamInfo(wm, "dec.out.w", (4, 4, cfg.img_c, cin), "glorot", (16 * cfg.img_c, 16 * cin)))
        p.append(ParamInfo(wm, "dec.out.bias", (cfg.img_c,), "zeros"))
    else:
        _mlp(p, wm, "dec", G + Z, D, L)
        _dense(p, wm, "dec.out", D, cfg.obs_dim)
    # --- actor ----------------------------------------------------------------------
    _mlp(p, "actor", "actor", G + Z, D, L)
    _dense(p, "actor", "actor.out", D, A)
    if not cfg.discrete:
        _mlp(p, "actor", "actor.std", G + Z, D, L)
        _dense(p, "actor", "actor.std_out", D, A)
    # --- critic (+ EMA copy) -----------------------------------------------------------
    for grp, pre in (("critic", "critic"), ("critic_ema", "critic_ema")):
        _mlp(p, grp, pre, G + Z, D, L)
        _dense(p, grp, f"{pre}.head", D, NUM_BUCKETS, init="zeros")
    # --- curiosity: N x (MLP(L) -> RepresentationLayer) on [h, z, a] (models/disagree_networks.py:31-40) ----
    for n in range(cfg.num_curiosity_nets):
        _mlp(p, "disagree", f"disagree.net{n}", G + Z + A, D, L)
        _dense(p, "disagree", f"disagree.net{n}.repr", D, Z)
    return p


def _orthogonal(rng, rows, cols):
    """Keras `Orthogonal` initialiser restated: QR of a normal matrix, sign-fixed."""
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return q[:rows, :cols]


def init_params(cfg: EngineConfig, seed=0, mode="keras"):
    """name -> float32 ndarray.

    mode="keras": the default Keras initialisers the reference relies on (glorot-uniform
    kernels, orthogonal GRU recurrent kernel, zero biases, unit LayerNorm gains, zero
    reward/critic heads). mode="random": every tensor (incl. biases,
    LayerNorm offsets and the zero-initialised heads) gets small random values so that
    every gradient path is exercised by the parity tests.
    """
    rng = np.random.default_rng(seed)
    out = {}
    for pi in param_spec(cfg):
        if pi.group == "critic_ema":
            continue
        if pi.init == "glorot":
            lim = np.sqrt(6.0 / (pi.fan[0] + pi.fan[1]))
            w = rng.uniform(-lim, lim, size=pi.shape)
        elif pi.init == "orthogonal":
            G = pi.shape[0]
            w = _orthogonal(rng, pi.shape[0], pi.shape[1])
            assert w.shape == (G, pi.shape[1])
        elif pi.init == "ones":
            w = np.ones(pi.shape)
            if mode == "random":
                w = w + 0.1 * rng.standard_normal(pi.shape)
        else:
            w = np.zeros(pi.shape)
            if mode == "random":
                scale = 0.05
                w = scale * rng.standard_normal(pi.shape)
        out[pi.name] = np.ascontiguousarray(w, dtype=np.float32)
    # critic EMA starts as a copy of the fast critic (critic_network.py:84-91, run_experiment.py:440-454)
    for name in [n for n in out if n.startswith("critic.")]:
        out["critic_ema." + name[len("critic."):]] = out[name].copy()
    return out

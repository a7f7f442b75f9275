"""Checkpoints in the reference's on-disk layout (run_experiment.py:692-714 saves, :489-509 and :224-233 load):

    <dir>/{world_model,actor,critic}.npz            weights = model.get_weights()
    <dir>/{world_model,actor,critic}_optimizer.npz  weights = model.optimizer.get_weights()
    <dir>/{total_env_steps,total_replayed_steps,total_train_steps,iteration}.npy

`get_weights()` order is Keras' (Keras 2.x `Layer.weights`: trainable before non-trainable; inside each, a layer's own
variables first, then its sub-layers in attribute-assignment order, list attributes where the list was assigned) applied
to the reference's constructors; the optimizer list is the Keras-2 `OptimizerV2` one: [iterations, m of every variable,
v of every variable].  **[3P - restated from the Keras sources' documented behaviour; TensorFlow is not installable
here, so a file written by a real TF run has not been read back yet]** - the orders are data (`weight_order`), easy to
correct against the first real checkpoint.  Nothing else in the engine depends on them.

A second, self-describing file `<dir>/dv3_state.npz` (every tensor keyed by its spec name) makes resume exact
regardless of that order.
"""
import os

import numpy as np
import torch

from ..spec import param_spec


def _mlp(prefix, L, out=None):
    """components/mlp.py:43-67: dense_layers list, layer_normalizations list, output_layer."""
    names = [f"{prefix}.mlp{l}.w" for l in range(L)]
    for l in range(L):
        names += [f"{prefix}.mlp{l}.g", f"{prefix}.mlp{l}.b"]
    if out:
        names += [f"{out}.w", f"{out}.bias"]
    return names


def weight_order(cfg, model):
    """spec names in `model.get_weights()` order, model in {"world_model", "actor", "critic", "disagree"}."""
    L = cfg.L
    if model == "world_model":
        n = ["initial_h"]  # the model's own tf.Variable (world_model.py:98-101) precedes its sub-layers
        if cfg.image_obs:  # cnn_atari.py:33-76: conv_layers list, then layer_normalizations list
            n += [f"enc.conv{l}.w" for l in range(4)]
            for l in range(4):
                n += [f"enc.conv{l}.g", f"enc.conv{l}.b"]
        else:
            n += _mlp("enc", L)
        n += _mlp("post", 1) + ["post.repr.w", "post.repr.bias"]
        n += _mlp("dyn", 1) + ["dyn.repr.w", "dyn.repr.bias"]
        n += _mlp("seq.pre", 1) + ["seq.gru.kernel", "seq.gru.recurrent", "seq.gru.bias"]
        n += _mlp("rew", L) + ["rew.head.w", "rew.head.bias"]
        n += _mlp("cont", L, "cont.out")
        if cfg.image_obs:  # conv_transpose_atari.py:41-92
            n += ["dec.dense.w", "dec.dense.bias"] + [f"dec.convT{l}.w" for l in range(3)]
            for l in range(3):
                n += [f"dec.convT{l}.g", f"dec.convT{l}.b"]
            n += ["dec.out.w", "dec.out.bias"]
        else:
            n += _mlp("dec", L, "dec.out")
        return n
    if model == "actor":  # + the two non-trainable percentile EMAs at the end (actor_network.py:31-36)
        n = _mlp("actor", L, "actor.out")
        if not cfg.discrete:
            n += _mlp("actor.std", L, "actor.std_out")
        return n
    if model == "critic":  # trainable (mlp, return_layer) then non-trainable (mlp_ema, return_layer_ema)
        return (_mlp("critic", L) + ["critic.head.w", "critic.head.bias"]
                + _mlp("critic_ema", L) + ["critic_ema.head.w", "critic_ema.head.bias"])
    if model == "disagree":  # disagree_networks.py:28-40: the mlps list, then the representation_layers list
        N = cfg.num_curiosity_nets
        n = []
        for i in range(N):
            n += _mlp(f"disagree.net{i}", L)
        for i in range(N):
            n += [f"disagree.net{i}.repr.w", f"disagree.net{i}.repr.bias"]
        return n
    raise KeyError(model)


def get_weights(engine, model):
    t = engine.tensor
    w = [t(n).detach().cpu().numpy().copy() for n in weight_order(engine.cfg, model)]
    if model == "actor":
        pct = t("ac.pct_ema").detach().cpu().numpy()
        w += [np.float32(pct[0]), np.float32(pct[1])]
    return w


def set_weights(engine, model, weights):
    names = weight_order(engine.cfg, model)
    extra = 2 if model == "actor" else 0
    if len(weights) != len(names) + extra:
        raise ValueError(f"{model}: expected {len(names) + extra} arrays, got {len(weights)}")
    for n, w in zip(names, weights):
        dst = engine.tensor(n)
        w = np.asarray(w, dtype=np.float32)
        if tuple(w.shape) != tuple(dst.shape):
            raise ValueError(f"{model}: {n} has shape {tuple(dst.shape)}, checkpoint holds {tuple(w.shape)}")
        dst.copy_(torch.from_numpy(w))
    if extra:
        engine.tensor("ac.pct_ema").copy_(torch.tensor([float(weights[-2]), float(weights[-1])]))


def _trainable(cfg, model):
    return [n for n in weight_order(cfg, model) if not n.startswith("critic_ema.")]


def get_optimizer_weights(engine, model):
    """[iterations, m..., v...] (Keras-2 OptimizerV2 Adam); empty before the first step, like Keras."""
    t = engine.tensor
    it = int(t(f"opt.{model}.step").item())
    if it == 0:
        return []
    names = _trainable(engine.cfg, model)
    return ([np.int64(it)] + [t("adam_m." + n).detach().cpu().numpy().copy() for n in names]
            + [t("adam_v." + n).detach().cpu().numpy().copy() for n in names])


def set_optimizer_weights(engine, model, weights):
    if len(weights) == 0:
        return
    names = _trainable(engine.cfg, model)
    if len(weights) != 1 + 2 * len(names):
        raise ValueError(f"{model} optimizer: expected {1 + 2 * len(names)} arrays, got {len(weights)}")
    engine.tensor(f"opt.{model}.step").fill_(int(weights[0]))
    for i, n in enumerate(names):
        engine.tensor("adam_m." + n).copy_(torch.from_numpy(np.asarray(weights[1 + i], dtype=np.float32)))
        engine.tensor("adam_v." + n).copy_(torch.from_numpy(np.asarray(weights[1 + len(names) + i], dtype=np.float32)))


MODELS = ("world_model", "actor", "critic")


def _models(cfg):
    """(engine group, file stem) pairs; run_experiment.py:495,702 names the curiosity files `disagree_nets`."""
    return [(m, m) for m in MODELS] + ([("disagree", "disagree_nets")] if cfg.num_curiosity_nets > 0 else [])


def save_checkpoint(path, dreamer_model, buffer=None, **counters):
    """run_experiment.py:692-714 (+ dv3_state.npz). counters: total_env_steps, total_replayed_steps, total_train_steps, iteration.
    buffer: the EpisodeReplayBuffer, written as buffer.npz with key `state` exactly like run_experiment.py:708."""
    e = dreamer_model.engine
    os.makedirs(path, exist_ok=True)
    if buffer is not None:
        np.savez(os.path.join(path, "buffer"), state=buffer.get_state())
    for m, stem in _models(e.cfg):
        np.savez(os.path.join(path, stem), weights=np.array(get_weights(e, m), dtype=object))
        np.savez(os.path.join(path, stem + "_optimizer"), weights=np.array(get_optimizer_weights(e, m), dtype=object))
    for k, v in counters.items():
        np.save(os.path.join(path, k), v)
    state = {}
    for pi in param_spec(e.cfg):
        state["param/" + pi.name] = e.tensor(pi.name).detach().cpu().numpy()
        if pi.group != "critic_ema":
            state["adam_m/" + pi.name] = e.tensor("adam_m." + pi.name).detach().cpu().numpy()
            state["adam_v/" + pi.name] = e.tensor("adam_v." + pi.name).detach().cpu().numpy()
    for m, _ in _models(e.cfg):
        state["step/" + m] = e.tensor(f"opt.{m}.step").detach().cpu().numpy()
    state["pct_ema"] = e.tensor("ac.pct_ema").detach().cpu().numpy()
    # position of the device noise stream (counter-based Philox: seed + step), so that a resumed run draws what the
    # uninterrupted one would have drawn
    state["noise_step"] = e.tensor("noise_step").detach().cpu().numpy()
    wm = dreamer_model.world_model
    state["py_noise_step"] = np.int64(getattr(wm, "_noise_step", 0))
    state["noise_seed"] = np.int64(getattr(wm, "noise_seed", 0))
    state["inference_calls"] = np.int64(e.lib.dv3_get_inference_calls(e.h))
    np.savez(os.path.join(path, "dv3_state"), **state)


def load_checkpoint(path, dreamer_model, buffer=None):
    """Restores from dv3_state.npz when present (exact), else from the reference-layout files (EXPERIMENTAL: the Keras
    get_weights() / optimizer-slot order is restated from Keras' documented tracking order and has not been validated against a
    file written by TensorFlow). buffer: restored from buffer.npz when given (run_experiment.py:226-228). Returns the counters."""
    e = dreamer_model.engine
    if buffer is not None:
        buffer.set_state(np.load(os.path.join(path, "buffer.npz"), allow_pickle=True)["state"])
    exact = os.path.join(path, "dv3_state.npz")
    if os.path.exists(exact):
        z = np.load(exact)
        for k in z.files:
            kind, _, name = k.partition("/")
            if kind == "param":
                e.tensor(name).copy_(torch.from_numpy(z[k]))
            elif kind in ("adam_m", "adam_v"):
                e.tensor(f"{kind}.{name}").copy_(torch.from_numpy(z[k]))
            elif kind == "step":
                e.tensor(f"opt.{name}.step").copy_(torch.from_numpy(z[k]))
        e.tensor("ac.pct_ema").copy_(torch.from_numpy(z["pct_ema"]))
        if "noise_step" in z.files:
            e.tensor("noise_step").copy_(torch.from_numpy(z["noise_step"]))
            dreamer_model.world_model._noise_step = int(z["py_noise_step"])
        if "noise_seed" in z.files:
            dreamer_model.world_model.noise_seed = int(z["noise_seed"])
            e._check(e.lib.dv3_set_inference_calls(e.h, int(z["inference_calls"])), "dv3_set_inference_calls")
    else:
        for m, stem in _models(e.cfg):
            set_weights(e, m, list(np.load(os.path.join(path, stem + ".npz"), allow_pickle=True)["weights"]))
            set_optimizer_weights(e, m, list(np.load(os.path.join(path, stem + "_optimizer.npz"), allow_pickle=True)["weights"]))
    out = {}
    for k in ("total_env_steps", "total_replayed_steps", "total_train_steps", "iteration"):
        f = os.path.join(path, k + ".npy")
        if os.path.exists(f):
            out[k] = int(np.load(f))
    return out

"""Python handle on one libdv3 engine (one per device / rank).

PyTorch is used only for device memory views, streams and (in the data-parallel path)
`torch.distributed`; every computation of the hot path runs in libdv3's CUDA kernels.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .spec import EngineConfig, param_spec

# 4: gradient span of actor + critic; 5: the Plan2Explore disagree nets (curiosity)
# 6 / 7 / 8: the world-model gradient buffer as three buckets in the order the backward pass finalises them (tail = reward /
# continue / decoder heads | recurrent part | encoder)
GROUP_IDS = {"world_model": 0, "actor": 1, "critic": 2, "critic_ema": 3, "actor_critic": 4, "disagree": 5,
             "world_model_heads": 6, "world_model_recurrent": 7, "world_model_encoder": 8}


class _CudaView:
    """Exposes an engine-owned device buffer through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False),
            "version": 3, "strides": None}


class Dv3Error(RuntimeError):
    pass


class Engine:
    def __init__(self, cfg: EngineConfig, compute_mode=0, world_size=1, rank=0, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise Dv3Error("dreamer_v3_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        c = _lib.Dv3Config()
        c.abi_version = _lib.ABI_VERSION
        c.G, c.D, c.L, c.cnn_mult, c.Zc, c.Zk = cfg.G, cfg.D, cfg.L, cfg.cnn_mult, cfg.Zc, cfg.Zk
        c.A, c.discrete, c.image_obs, c.obs_dim = cfg.A, int(cfg.discrete), int(cfg.image_obs), cfg.obs_dim
        c.img_hw, c.img_c, c.B, c.T, c.H = cfg.img_hw, cfg.img_c, cfg.B, cfg.T, cfg.H
        c.symlog_obs, c.compute_mode = int(cfg.symlog_obs), int(compute_mode)
        c.world_size, c.rank, c.device = int(world_size), int(rank), self.device_index
        c.num_curiosity_nets, c.intrinsic_rewards_scale = int(cfg.num_curiosity_nets), float(cfg.intrinsic_rewards_scale)
        self.world_size, self.rank, self.compute_mode = int(world_size), int(rank), int(compute_mode)
        h = C.c_void_p()
        rc = self.lib.dv3_create(C.byref(c), C.byref(h))
        if rc != 0:
            raise Dv3Error(f"dv3_create failed ({rc}): {self.lib.dv3_last_error(None).decode()}")
        self.h = h
        self._views = {}
        self._desc = {}
        d = _lib.Dv3TensorDesc()
        for i in range(self.lib.dv3_num_tensors(self.h)):
            self.lib.dv3_tensor_info(self.h, i, C.byref(d))
            self._desc[d.name.decode()] = (int(d.ptr), tuple(d.shape[j] for j in range(d.ndim)), int(d.dtype),
                                           int(d.kind), int(d.group))
        self.hp = _lib.Dv3Hparams()
        self.lib.dv3_default_hparams(C.byref(self.hp))

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "h", None):
            torch.cuda.synchronize(self.device)
            self.lib.dv3_destroy(self.h)
            self.h = None
            self._views.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise Dv3Error(f"{what} failed ({rc}): {self.lib.dv3_last_error(self.h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def tensor_names(self, kind=None):
        return [n for n, d in self._desc.items() if kind is None or d[3] == kind]

    def tensor(self, name) -> torch.Tensor:
        """Zero-copy torch view of a named engine buffer (valid until the engine is closed)."""
        t = self._views.get(name)
        if t is None:
            ptr, shape, dtype, _, _ = self._desc[name]
            t = torch.as_tensor(_CudaView(ptr, shape, "<f4" if dtype == 0 else "<i4"), device=self.device)
            self._views[name] = t
        return t

    def group_buffer(self, group, kind="grad") -> torch.Tensor:
        """Flat fp32 buffer of a whole optimizer group (what the data-parallel all-reduce runs on)."""
        key = ("group", group, kind)
        t = self._views.get(key)
        if t is None:
            ptr, n = C.c_void_p(), C.c_int64()
            rc = self.lib.dv3_group_buffer(self.h, GROUP_IDS[group], {"param": 0, "grad": 1, "m": 2, "v": 3}[kind],
                                           C.byref(ptr), C.byref(n))
            self._check(rc, "dv3_group_buffer")
            t = torch.as_tensor(_CudaView(ptr.value, (n.value,), "<f4"), device=self.device)
            self._views[key] = t
        return t

    # ------------------------------------------------------------------ parameters
    def load_params(self, params):
        """name -> ndarray / tensor. Missing critic_ema.* entries are copied from critic.*"""
        for pi in param_spec(self.cfg):
            src = params.get(pi.name)
            if src is None and pi.group == "critic_ema":
                src = params.get("critic." + pi.name[len("critic_ema."):])
            if src is None:
                raise KeyError(f"missing parameter {pi.name}")
            src = torch.as_tensor(np.asarray(src) if not torch.is_tensor(src) else src, dtype=torch.float32)
            dst = self.tensor(pi.name)
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"{pi.name}: shape {tuple(src.shape)} != {tuple(dst.shape)}")
            dst.copy_(src)

    def get_params(self, prefix=""):
        return {pi.name: self.tensor(pi.name).detach().cpu().numpy().copy()
                for pi in param_spec(self.cfg) if pi.name.startswith(prefix)}

    def get_grads(self):
        return {pi.name: self.tensor("grad." + pi.name).detach().cpu().numpy().copy()
                for pi in param_spec(self.cfg) if pi.group != "critic_ema"}

    # ------------------------------------------------------------------ inputs
    def set_sample(self, sample, non_blocking=True):
        """sample: dict with obs [B,T,...], actions [B,T,A], rewards/is_first/is_terminated [B,T]
        (numpy, pinned CPU tensors or device tensors)."""
        for key, name in (("obs", "in.obs"), ("actions", "in.actions"), ("rewards", "in.rewards"),
                          ("is_first", "in.is_first"), ("is_terminated", "in.is_terminated")):
            if key in sample and sample[key] is not None:
                self.write(name, sample[key], non_blocking)

    def write(self, name, value, non_blocking=True):
        dst = self.tensor(name)
        src = value if torch.is_tensor(value) else torch.from_numpy(np.ascontiguousarray(value))
        dst.copy_(src.reshape(dst.shape), non_blocking=non_blocking)

    def set_noise(self, u_post=None, u_dyn=None, act_noise=None):
        if u_post is not None:
            self.write("in.u_post", u_post)
        if u_dyn is not None:
            self.write("in.u_dyn", u_dyn)
        if act_noise is not None:
            self.write("in.act_noise", act_noise)

    def phase_report(self, enable=-1):
        """Device time per labelled phase of the non-graph path since the last report: {label: (ms, count)} (bench.py)."""
        buf = C.create_string_buffer(1 << 14)
        n = self.lib.dv3_phase_report(self.h, int(enable), buf, len(buf))
        if n < 0:
            raise Dv3Error("dv3_phase_report failed")
        out = {}
        for line in buf.value.decode().splitlines():
            k, ms, cnt = line.rsplit(" ", 2)
            out[k] = (float(ms), int(cnt))
        return out

    def fill_noise(self, seed, step=0, which=3):
        self._check(self.lib.dv3_fill_noise(self.h, int(seed), int(step), int(which), self._stream()), "dv3_fill_noise")

    # ------------------------------------------------------------------ hot path
    def initial_state(self):
        self._check(self.lib.dv3_initial_state(self.h, self._stream()), "dv3_initial_state")

    def observe_forward(self):
        self._check(self.lib.dv3_observe_forward(self.h, self._stream()), "dv3_observe_forward")

    def wm_loss_backward(self):
        self._check(self.lib.dv3_wm_loss_backward(self.h, self._stream()), "dv3_wm_loss_backward")

    def optimizer_step(self, group, clip, lr, eps, ema_decay=-1.0):
        self._check(self.lib.dv3_optimizer_step(self.h, GROUP_IDS[group], clip, lr, eps, ema_decay, self._stream()),
                    "dv3_optimizer_step")

    def forward_inference(self, rows, noise_seed=0):
        self._check(self.lib.dv3_forward_inference(self.h, int(rows), int(noise_seed), self._stream()), "dv3_forward_inference")

    def dream_forward(self, gamma=None, start_from_observe=True):
        g = self.hp.gamma if gamma is None else float(gamma)
        self._check(self.lib.dv3_dream_forward(self.h, g, int(start_from_observe), self._stream()), "dv3_dream_forward")

    def dream_forward_actions(self, action_mode, gamma=None, start_from_observe=False):
        """Prior rollout with given actions (include/dv3.h: dv3_dream_forward_actions)."""
        g = float(self.hp.gamma if gamma is None else gamma)
        self._check(self.lib.dv3_dream_forward_actions(self.h, g, int(start_from_observe), int(action_mode), self._stream()),
                    "dv3_dream_forward_actions")

    def ac_loss_backward(self, phase=3):
        self._check(self.lib.dv3_ac_loss_backward(self.h, C.byref(self.hp), int(phase), self._stream()),
                    "dv3_ac_loss_backward")

    def disagree_loss_backward(self):
        """disagree_loss + its gradients (losses/disagree_loss.py, train_one_step.py:180-181, 209-212); curiosity engines only."""
        self._check(self.lib.dv3_disagree_loss_backward(self.h, self._stream()), "dv3_disagree_loss_backward")

    def critic_init_ema(self):
        self._check(self.lib.dv3_critic_init_ema(self.h, self._stream()), "dv3_critic_init_ema")

    def critic_update_ema(self, decay=0.98):
        self._check(self.lib.dv3_critic_update_ema(self.h, decay, self._stream()), "dv3_critic_update_ema")

    def train_world_model_step(self, noise_seed=0, use_graph=True):
        self._check(self.lib.dv3_train_world_model_step(self.h, C.byref(self.hp), int(noise_seed), int(use_graph),
                                                        self._stream()), "dv3_train_world_model_step")

    def train_actor_critic_step(self, noise_seed=0, use_graph=True):
        self._check(self.lib.dv3_train_actor_critic_step(self.h, C.byref(self.hp), int(noise_seed), int(use_graph),
                                                         self._stream()), "dv3_train_actor_critic_step")

    def train_update(self, noise_seed=0, use_graph=True):
        self._check(self.lib.dv3_train_update(self.h, C.byref(self.hp), int(noise_seed), int(use_graph), self._stream()),
                    "dv3_train_update")

    def dp_piece(self, piece, noise_seed=0, use_graph=True):
        self._check(self.lib.dv3_dp_piece(self.h, C.byref(self.hp), int(piece), int(noise_seed), int(use_graph), self._stream()),
                    "dv3_dp_piece")

    def kernel_launch_count(self):
        return int(self.lib.dv3_kernel_launch_count(self.h))

    def sync(self):
        self._check(self.lib.dv3_sync(self.h, self._stream()), "dv3_sync")

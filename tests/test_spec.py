"""Host-side checks (no GPU): size table, parameter table, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from dreamer_v3_b200 import _lib
from dreamer_v3_b200.spec import Box, Discrete, EngineConfig, init_params, param_spec
from dreamer_v3_b200.utils import model_dimensions as md

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_size_table_matches_reference_values():
    # utils/model_dimensions.py:14-17,26-29,38-41,50-53,62-65,86-89
    want = {"XS": (24, 256, 256, 1), "S": (32, 512, 512, 2), "M": (48, 640, 1024, 3), "L": (64, 768, 2048, 4),
            "XL": (96, 1024, 4096, 5), "XXS": (16, 128, 128, 1), "nano": (2, 16, 16, 1), "micro": (4, 32, 32, 1)}
    for size, (m, d, g, l) in want.items():
        assert md.get_cnn_multiplier(size) == m and md.get_dense_hidden_units(size) == d
        assert md.get_gru_units(size) == g and md.get_num_dense_layers(size) == l
        assert md.get_num_z_categoricals(size) == md.get_num_z_classes(size) == (32 if size not in ("nano", "micro") else (4 if size == "nano" else 8))
    assert md.get_gru_units("XS", override=77) == 77
    assert md.get_num_curiosity_nets("S") == 8
    with pytest.raises(AssertionError):
        md.get_gru_units("XXL")


def test_param_counts_match_survey():
    # WM / actor / critic parameter counts per config
    def count(cfg, group):
        return sum(p.size for p in param_spec(cfg) if p.group == group)
    c1 = EngineConfig.make("XS", action_space=Discrete(2), obs_dim=4)
    assert abs(count(c1, "world_model") / 2.4e6 - 1) < 0.05
    assert abs(count(c1, "actor") / 0.33e6 - 1) < 0.05 and abs(count(c1, "critic") / 0.39e6 - 1) < 0.05
    c2 = EngineConfig.make("XS", action_space=Box(-2, 2, (1,)), obs_dim=3)
    assert abs(count(c2, "actor") / 0.66e6 - 1) < 0.05
    c3 = EngineConfig.make("S", action_space=Discrete(6), image_obs=True)
    assert abs(count(c3, "world_model") / 15.7e6 - 1) < 0.05
    c5 = EngineConfig.make("XL", action_space=Discrete(18), image_obs=True)
    assert abs(count(c5, "world_model") / 181.5e6 - 1) < 0.05
    assert abs(count(c5, "critic") / 9.70e6 - 1) < 0.05
    assert count(c5, "critic") == count(c5, "critic_ema")


def test_keras_default_init():
    cfg = EngineConfig.make("XXS", action_space=Discrete(3), obs_dim=5)
    p = init_params(cfg, seed=0)
    assert set(p) == {pi.name for pi in param_spec(cfg)}
    assert np.all(p["rew.head.w"] == 0) and np.all(p["critic.head.w"] == 0)  # reward_predictor_layer.py:58-59
    assert np.all(p["initial_h"] == 0) and np.all(p["post.mlp0.g"] == 1) and np.all(p["seq.gru.bias"] == 0)
    U = p["seq.gru.recurrent"]  # Keras orthogonal recurrent kernel: orthonormal rows
    np.testing.assert_allclose(U @ U.T, np.eye(U.shape[0]), atol=1e-5)
    w = p["post.mlp0.w"]
    lim = np.sqrt(6.0 / sum(w.shape))
    assert np.abs(w).max() <= lim and np.abs(w).max() > 0.9 * lim  # glorot uniform
    np.testing.assert_array_equal(p["critic_ema.mlp0.w"], p["critic.mlp0.w"])  # init_ema, critic_network.py:84-91


def test_capi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dv3.h")).read()
    declared = set(re.findall(r"\b(dv3_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dv3_abi_version() == 1
    # struct layouts agree with the header (field order / count)
    assert ctypes.sizeof(_lib.Dv3Config) == 4 * (21 + 8)
    assert ctypes.sizeof(_lib.Dv3Hparams) == 4 * (12 + 1 + 2 + 5)
    hp = _lib.Dv3Hparams()
    lib.dv3_default_hparams(ctypes.byref(hp))
    assert abs(hp.gamma - 0.997) < 1e-7 and abs(hp.lambda_ - 0.95) < 1e-7 and abs(hp.wm_lr - 1e-4) < 1e-10
    assert abs(hp.critic_ema_decay - 0.98) < 1e-7 and hp.train_actor == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    c = _lib.Dv3Config()
    c.abi_version = 1
    h = ctypes.c_void_p()
    assert lib.dv3_create(ctypes.byref(c), ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.dv3_last_error(None)
    from dreamer_v3_b200.engine import Dv3Error, Engine
    with pytest.raises(Dv3Error):
        Engine(EngineConfig.make("nano", action_space=Discrete(2)))
    from dreamer_v3_b200 import ops
    with pytest.raises(RuntimeError):
        ops.symlog(torch.zeros(3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dreamer_v3_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(d, f)).read()
                assert "oracle" not in src.lower() or f.endswith((".cu", ".cuh", ".h")) and "import" not in src, f


@pytest.mark.parametrize("size", ("S", "M", "L", "XL"))
@pytest.mark.parametrize("ctas", (148, 132, 96, 37))
def test_scan_tc_stream_k_schedule(size, ctas):
    """Host-only walk of the stream-K schedules of the tensor-core observe-scan kernels (csrc/dv3_scan_tc.cu; the work of
    world_model.py:244-273 and its BPTT cut into per-CTA segment / staging tables): built with the very code the kernels run, every
    (contraction, tile, k-block) unit must be covered exactly once on any grid size, within the table bounds."""
    cfg = EngineConfig.make(size, action_space=Discrete(6), image_obs=True)
    lib = _lib.load()
    rc = lib.dv3_scan_tc_schedule_check(16, 64, cfg.G, cfg.D, cfg.Z, ctas)
    assert rc in (0, 1), rc      # 1: this size / grid is routed to the other scan kernels (launcher bounds), never a defect
    if ctas == 148:
        assert rc == 0, (size, rc)   # the B200 grid must be schedulable at every size


def test_scan_tc_schedule_check_rejects_other_sizes():
    lib = _lib.load()
    assert lib.dv3_scan_tc_schedule_check(16, 64, 256, 256, 1024, 148) == 1     # XS: cluster kernel
    assert lib.dv3_scan_tc_schedule_check(17, 64, 512, 512, 1024, 148) == 1     # more than 16 rows
    assert lib.dv3_scan_tc_schedule_check(16, 1024, 512, 512, 1024, 148) == 1   # is_first flags do not fit shared memory


def test_ctypes_structs_match_the_c_header(tmp_path):
    """dv3_config / dv3_hparams / dv3_tensor_desc are mirrored by hand in _lib.py: compile a probe against include/dv3.h (gcc, no
    CUDA needed) and compare sizes and the offsets of the fields added last (curiosity, the critic's / disagree nets' Adam)."""
    import ctypes
    import shutil
    import subprocess
    from dreamer_v3_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "probe.c"
    src.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include "dv3.h"\n'
        'int main(void) {\n'
        '  printf("%zu %zu %zu\\n", sizeof(dv3_config), sizeof(dv3_hparams), sizeof(dv3_tensor_desc));\n'
        '  printf("%zu %zu %zu\\n", offsetof(dv3_config, device), offsetof(dv3_config, num_curiosity_nets), offsetof(dv3_config, intrinsic_rewards_scale));\n'
        '  printf("%zu %zu %zu %zu\\n", offsetof(dv3_hparams, train_actor), offsetof(dv3_hparams, critic_lr), offsetof(dv3_hparams, disagree_grad_clip), offsetof(dv3_hparams, disagree_eps));\n'
        '  printf("%zu %zu %zu\\n", offsetof(dv3_tensor_desc, shape), offsetof(dv3_tensor_desc, kind), offsetof(dv3_tensor_desc, group));\n'
        '  return 0;\n}\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    C, H, D = _lib.Dv3Config, _lib.Dv3Hparams, _lib.Dv3TensorDesc
    want = [ctypes.sizeof(C), ctypes.sizeof(H), ctypes.sizeof(D),
            C.device.offset, C.num_curiosity_nets.offset, C.intrinsic_rewards_scale.offset,
            H.train_actor.offset, H.critic_lr.offset, H.disagree_grad_clip.offset, H.disagree_eps.offset,
            D.shape.offset, D.kind.offset, D.group.offset]
    assert got == want, (got, want)

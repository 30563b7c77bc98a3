"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/sacx.h declares,
the Python mirror of its structs matches, config translation and error mapping follow the reference."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "sacx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sacx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from sac import _engine as E
    lib = E.load()
    declared = _header_functions()
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sacx.h but not exported by libsacx.so"
    assert sorted(E.SYMBOLS) == declared, "Python binding table and header disagree"
    assert lib.sacx_version() == 100


def test_struct_mirrors_match_the_library():
    from sac import _engine as E
    lib = E.load()
    assert lib.sacx_sizeof_config() == C.sizeof(E.SacxConfig)
    assert lib.sacx_sizeof_metrics() == C.sizeof(E.SacxMetrics)
    assert lib.sacx_sizeof_tensor_desc() == C.sizeof(E.SacxTensorDesc)


def _cfg(**over):
    cfg = {
        "sac": {"gamma": 0.99, "tau": 0.005, "alpha": 0.1, "auto_entropy_tuning": True, "actor_lr": 3e-4, "critic_lr": 3e-4, "alpha_lr": 3e-4},
        "q_net": {"hidden_sizes": [256, 256], "hidden_layers_act": "relu", "output_activation": "identity"},
        "policy_net": {"hidden_sizes": [256, 256], "hidden_layers_act": "relu", "output_activation": "identity",
                       "log_std_min": -20, "log_std_max": 2, "action_scale": 1.0},
        "buffer": {"capacity": 1000}, "train": {"batch_size": 256, "seed": 3, "device": "cuda", "warming_steps": 10},
    }
    for k, v in over.items():
        sec, key = k.split("__")
        cfg[sec][key] = v
    return cfg


def test_config_translation_and_arena_size():
    from sac import _engine as E
    lib = E.load()
    c = E.make_config(24, 4, _cfg())
    assert (c.obs_dim, c.act_dim, c.n_hidden_pi, c.n_hidden_q, c.batch_size) == (24, 4, 2, 2, 256)
    assert list(c.hidden_pi)[:2] == [256, 256] and c.act_hidden_pi == 1 and c.act_out_q == 0 and c.seed == 3
    floats = C.c_int64()
    assert lib.sacx_agent_arena_floats(C.byref(c), C.byref(floats)) == 0
    n_on = 74248 + 2 * 73473            # policy 74 248, Q 73 473 each
    assert floats.value * 4 > (4 * n_on + 2 * 73473) * 4          # params + m + v + g + targets, plus scratch
    assert floats.value % 128 == 0


def test_error_conventions_follow_the_reference():
    from sac import _engine as E
    lib = E.load()
    with pytest.raises(KeyError):                                  # sac/models.py:138-139
        E.make_config(3, 1, _cfg(q_net__hidden_layers_act="swish"))
    with pytest.raises(ValueError, match="hidden_sizes cannot be empty"):      # sac/models.py:135-136
        E.make_config(3, 1, _cfg(policy_net__hidden_sizes=[]))
    assert lib.sacx_activation_id(b"gelu") == 5 and lib.sacx_activation_id(b"nope") == E.SACX_ERR_ACTIVATION
    bad = E.make_config(3, 1, _cfg())
    bad.act_dim = 99
    floats = C.c_int64()
    rc = lib.sacx_agent_arena_floats(C.byref(bad), C.byref(floats))
    assert rc == E.SACX_ERR_INVALID
    with pytest.raises(ValueError):
        E.check(rc)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sac import _engine as E
    from sac.agent import SAC
    from sac.replay_buffer import ReplayBuffer
    lib = E.load()
    assert lib.sacx_device_count() == 0
    h = C.c_void_p()
    c = E.make_config(3, 1, _cfg())
    assert lib.sacx_agent_create(C.byref(c), None, C.byref(h)) == E.SACX_ERR_CUDA
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ReplayBuffer(10, 3, 1)

    class Env:
        class S:
            shape = (3,)
        observation_space = action_space = S()
    cfg = _cfg()
    cfg["train"]["device"] = "cpu"
    cfg["logger"] = {"enabled": False}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SAC(Env(), cfg)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "soft-actor-critic_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} touches oracle/"


def test_models_match_reference_init_and_state_dict_layout(golden_dir):
    """Same torch calls in the same order as the reference => bit-identical initial weights (F10)."""
    import numpy as np
    from helpers import Golden
    from sac.models import PolicyNetwork, QNetwork, build_mlp
    g = Golden("acts_gelu")
    c = g.cfg
    seed = c["train"]["seed"]
    pi = PolicyNetwork(g.obs, g.act, c["policy_net"]["hidden_sizes"], hidden_activations="gelu", seed=seed)
    q1 = QNetwork(g.obs, g.act, c["q_net"]["hidden_sizes"], "gelu", "identity", seed=seed)
    q2 = QNetwork(g.obs, g.act, c["q_net"]["hidden_sizes"], "gelu", "identity", seed=seed + 1)
    for tag, net in (("pi", pi), ("q1", q1), ("q2", q2)):
        sd = net.state_dict()
        ref = g.sd(f"init/{tag}")
        assert sorted(sd) == sorted(ref)
        for k in ref:
            assert np.array_equal(sd[k].numpy(), ref[k]), (tag, k)
    with pytest.raises(ValueError):
        build_mlp(3, [], 1)
    with pytest.raises(KeyError):
        build_mlp(3, [4], 1, "swish")


def test_bulk_index_stream_is_the_stdlib_stream():
    """sac.replay_buffer.sample_range (the host-RNG path's index draw) == random.sample(range(n), k) of the reference's
    replay_buffer.py:39, value for value, and it leaves the global generator in the same state -- checked against the
    stdlib and against the index streams recorded from the real reference (tests/golden/sampling.json)."""
    import json
    import random

    from sac.replay_buffer import sample_range
    cases = [(0, 1_000_000, 256), (1, 20000, 256), (2, 1500, 100), (3, 1025, 1024), (4, 5000, 1), (5, 300, 256),
             (6, 2 ** 20, 4096), (7, 2 ** 32 - 5, 64), (8, 2 ** 33, 16), (9, 4200, 1024), (10, 7, 7)]
    for seed, n, k in cases:
        random.seed(seed)
        want = [random.sample(range(n), k) for _ in range(3)]
        state = random.getstate()
        random.seed(seed)
        got = [list(sample_range(n, k)) for _ in range(3)]
        assert got == want, (seed, n, k)
        assert random.getstate() == state, (seed, n, k)
    with pytest.raises(ValueError):
        sample_range(5, 6)
    with open(os.path.join(os.path.dirname(__file__), "golden", "sampling.json")) as f:
        g = json.load(f)
    for c in g["cases"]:
        random.seed(c["seed"])
        n = min(c["pushes"], c["capacity"])
        oldest = max(c["pushes"] - c["capacity"], 0)
        for ids in c["push_ids"]:
            assert [oldest + j for j in sample_range(n, c["k"])] == ids


def test_one_normal_draw_equals_two_consecutive_draws():
    """SAC._normal_pair: torch's CPU normal_ on 2*n elements == two consecutive normal_ draws of n elements when n % 16 == 0
    (what the reference consumes per update: models.py:82-83 called twice, agent.py:204,241)."""
    import torch
    for rows, act in ((256, 4), (8, 2), (1024, 2), (64, 1), (48, 3)):
        assert (rows * act) % 16 == 0
        torch.manual_seed(5)
        a, b = torch.empty(rows, act).normal_(), torch.empty(rows, act).normal_()
        after = torch.get_rng_state()
        torch.manual_seed(5)
        both = torch.empty(2, rows, act).normal_()
        assert torch.equal(both[0], a) and torch.equal(both[1], b) and torch.equal(torch.get_rng_state(), after)

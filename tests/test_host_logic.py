"""CPU tests of the host-side logic and of the C-ABI library's surface (no compute calls:
this container has no GPU).  Reference-signal construction is pinned bit-exactly against
vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

import f16_model_b200 as f16
from f16_model_b200 import _capi
from f16_model_b200.sharding import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- reference signals
@pytest.mark.parametrize("scenario", ["step", "cos", "pure_step", "combo"])
def test_reference_signal_deterministic_bit_exact(golden, scenario):
    sig = f16.ReferenceSignal(0.01, 10, determenistic=True, scenario=scenario)
    assert np.array_equal(sig.theta_ref, golden[f"ref_det_{scenario}"])
    assert np.array_equal(sig.wz_ref, golden[f"ref_det_{scenario}"] / 2)


@pytest.mark.parametrize("scenario", ["step", "cos", "pure_step", "combo"])
@pytest.mark.parametrize("seed", [1, 2, 3, 11, 12345])
def test_reference_signal_random_follows_reference_draw_order(golden, scenario, seed):
    random.seed(seed)
    sig = f16.ReferenceSignal(0.01, 10, determenistic=False, scenario=scenario)
    assert np.array_equal(sig.theta_ref, golden[f"ref_rnd_{scenario}_{seed}"])


@pytest.mark.parametrize("seed", [4, 5, 6, 7, 8, 9])
def test_reference_signal_random_scenario_choice(golden, seed):
    random.seed(seed)
    sig = f16.ReferenceSignal(0.01, 10, determenistic=False, scenario=None)
    assert np.array_equal(sig.theta_ref, golden[f"ref_rnd_None_{seed}"])


def test_reference_signal_other_horizons(golden):
    assert np.array_equal(f16.ReferenceSignal(0.01, 5, True, "pure_step").theta_ref, golden["ref_det_pure_step_tn5"])
    assert np.array_equal(f16.ReferenceSignal(0.02, 20, True, "step").theta_ref, golden["ref_det_step_tn20_dt02"])
    assert np.array_equal(f16.ReferenceSignal(0.01, 20, True, "combo").theta_ref, golden["ref_det_combo_tn20"])


def test_reference_signal_known_shapes():
    """det. pure_step -> wz_ref[100:] = radians(5); det. combo plateaus."""
    ps = f16.ReferenceSignal(0.01, 10, True, "pure_step")
    assert np.all(ps.wz_ref[:100] == 0) and np.all(ps.wz_ref[100:] == np.radians(10) / 2)
    cb = f16.ReferenceSignal(0.01, 10, True, "combo")
    assert np.all(cb.theta_ref[50:250] == np.radians(10)) and np.all(cb.theta_ref[:50] == 0)
    assert np.all(cb.theta_ref[700:800] == np.radians(20)) and np.all(cb.theta_ref[900:] == 0)


def test_combo_needs_ten_seconds():
    with pytest.raises(ValueError):    # the reference fails the same way (np.concatenate of nothing)
        f16.ReferenceSignal(0.01, 5, True, "combo")


def test_reference_bank_enumerates_discrete_space():
    bank = f16.build_reference_bank(0.01, 10, "pure_step", False)
    assert bank.shape == (24, 1000)
    assert len({row.tobytes() for row in bank}) == 24
    bank = f16.build_reference_bank(0.01, 10, None, False)
    # scenario chosen uniformly, then parameters uniformly: equal mass per scenario
    assert bank.shape[0] % 3 == 0
    random.seed(3)
    sig = f16.ReferenceSignal(0.01, 10, False, None)
    assert any(np.array_equal(sig.wz_ref, row) for row in bank)
    bank = f16.build_reference_bank(0.01, 10, "combo", False, bank_size=16, seed=5)
    assert bank.shape == (16, 1000)
    assert np.array_equal(bank, f16.build_reference_bank(0.01, 10, "combo", False, bank_size=16, seed=5))
    assert f16.build_reference_bank(0.01, 10, "combo", True).shape == (1, 1000)
    full = f16.build_reference_bank(0.01, 10, "combo", False, bank_size="full")
    assert full.shape == (6 * 6 * 5 * 6 * 6, 1000)
    for seed in (1, 2, 3, 11, 12345):          # every seeded draw of the reference is one of the rows
        random.seed(seed)
        sig = f16.ReferenceSignal(0.01, 10, False, "combo").wz_ref
        assert np.any(np.all(full == sig, axis=1))


# ---------------------------------------------------------------- env statics
def test_rescale_action_matches_reference(golden):
    out = np.array([float(f16.F16.rescale_action(np.array([v]))) for v in golden["rescale_in"]])
    assert np.array_equal(out, golden["rescale_out"])


def test_device_action_constants_are_the_float32_cast_bounds():
    lo, hi = np.float32(-np.radians(25)), np.float32(np.radians(25))
    src = open(os.path.join(ROOT, "f16-model_b200", "csrc", "f16_device.cuh")).read()
    assert repr(float(lo)) in src and repr(float(hi - lo)) in src


def test_normalize_denormalize_roundtrip():
    raw = [1234.5, 0.3, 210.0]
    n = f16.F16.normalize(raw)
    assert np.allclose(n, [1234.5 / 35000, (0.3 + np.radians(150)) / (2 * np.radians(150)), 0.42])
    back = f16.F16.denormalize(np.array(n + [0.01, 0.02]))
    assert np.allclose(back[:3], raw) and np.allclose(back[3:], [0.01, 0.02])


def test_trimmed_state_and_thrust_position(golden):
    x0, u0 = f16.get_trimmed_state_control()
    assert np.array_equal(x0, golden["trim_x0"]) and np.array_equal(u0, golden["trim_u0"])
    mine = np.array([f16.find_correct_thrust_position(t) for t in golden["thrpos_in"]])
    assert np.max(np.abs(mine - golden["thrpos_out"])) < 1e-9     # closed form vs the reference's fsolve


def test_get_random_state_draw_order(golden):
    for seed, ref in zip(golden["random_state_seeds"], golden["random_state_out"]):
        np.random.seed(int(seed))
        assert np.array_equal(f16.get_random_state().to_array(), ref)


def test_plane_constants_match_reference(golden):
    from f16_model_b200 import plane

    names = ["m", "l", "S", "b_a", "Jz", "rcgx", "Tstab", "Xistab", "maxabsstab", "maxabsdstab", "minthrottle", "maxthrottle", "lef", "sb"]
    assert np.array_equal(np.array([float(getattr(plane, k)) for k in names]), golden["plane_scalars"])
    assert list(plane.state_bound) == ["Oy", "wz", "V"]          # dict order = observation order
    assert np.array_equal(np.array([plane.state_bound[k] for k in ("Oy", "wz", "V")], dtype=float), golden["plane_state_bound"])
    rb = np.array([plane.random_state_bound[k] for k in ("Oy", "wz", "V", "alpha", "maxabsstab")], dtype=float)
    assert np.array_equal(rb, golden["plane_random_state_bound"])


def test_states_arithmetic():
    a = f16.States(*range(9))
    b = 0.5 * a + a
    assert np.allclose(b.to_array(), 1.5 * np.arange(9))
    assert np.allclose((f16.Control(0.1, 0.5) * 2).to_array(), [0.2, 1.0])


# ---------------------------------------------------------------- sharding
def test_shard_range_covers_batch():
    for total, world in [(1 << 20, 8), (4097, 4), (10, 3), (7, 8)]:
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1


def test_shard_range_properties():
    """Any (total, world): blocks are contiguous, ordered, cover [0, total) exactly and differ by at most one env."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 1 << 24), st.integers(1, 64))
    def check(total, world):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(b <= e for b, e in spans) and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)

    check()
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def test_trim_grid_sharded_single_rank_needs_no_process_group():
    from f16_model_b200.sharding import trim_grid_sharded

    V, H = np.meshgrid(np.linspace(100, 300, 5), np.linspace(1000, 9000, 3))

    def fake_trim(V, H, device=None, **kw):
        return {"stab": -V, "throttle": H, "alpha": V + H, "cost": np.where(V > 200, 1e-3, 1e-7), "iterations": np.full(V.size, 7, np.int32),
                "n_starts": 3}

    g = trim_grid_sharded(V, H, 0, 1, device="cpu", trim_fn=fake_trim)
    assert np.array_equal(g["stab"], -V.ravel()) and np.array_equal(g["alpha"], (V + H).ravel()) and g["n_starts"] == 3
    assert np.array_equal(g["accepted"], V.ravel() <= 200) and g["iterations"].dtype == np.int32
    with pytest.raises(ValueError):
        trim_grid_sharded(V, H[:2], 0, 1, device="cpu", trim_fn=fake_trim)
    with pytest.raises(RuntimeError):                      # more than one rank without an initialised process group
        trim_grid_sharded(V, H, 0, 2, device="cpu", trim_fn=fake_trim)


def test_combo_segment_bank_is_the_reference_distribution():
    """f3: the default device bank for random `combo` episodes (segment mode) holds EXACTLY the reference's distribution
    (env/task.py:65-103): 6480 distinct, equally weighted signals; every signal rebuilt from its packed code is
    bit-identical to the enumerated whole row, and every draw the reference's own procedure makes (random.choice x 4 +
    random.shuffle, any seed) is one of them."""
    import random

    from f16_model_b200 import task

    bank, codes, L = task.build_combo_segments(0.01, 10)
    assert bank.shape == (43, 300) and L == 300 and codes.dtype == np.uint32
    assert codes.size == 6480 == len(set(codes.tolist())) and int(codes.max()) < 2 ** 31
    assert np.all(bank[42] == 0) and np.all((codes >> 24) == 42)
    rows = task.signals_from_codes(bank, codes, 1000)
    full = task.build_reference_bank(0.01, 10, "combo", False, bank_size="full")
    assert np.array_equal(rows, full)
    assert len({r.tobytes() for r in rows}) == 6480                       # all distinct: uniform over codes == uniform over signals
    where = {r.tobytes(): i for i, r in enumerate(rows)}
    seen = set()
    for seed in range(300):
        sig = task.ReferenceSignal(0.01, 10, False, "combo", rng=random.Random(seed)).wz_ref
        assert sig.tobytes() in where, seed
        seen.add(where[sig.tobytes()])
    assert len(seen) > 280                                                # draws spread over the space (300 of 6480: ~7 collisions expected)
    # the parameter marginals of the enumeration are uniform, as random.choice makes them
    params = list(task.combo_parameter_sets())
    for k, n_values in ((0, 6), (1, 6), (2, 5), (3, 6), (4, 6)):
        counts = {}
        for p_ in params:
            counts[p_[k]] = counts.get(p_[k], 0) + 1
        assert len(counts) == n_values and len(set(counts.values())) == 1
    # other step sizes factor the same way; multi-block episodes do not (they keep the row bank)
    b2, c2, L2 = task.build_combo_segments(0.02, 10)
    assert L2 == 150 and np.array_equal(task.signals_from_codes(b2, c2, 500), task.build_reference_bank(0.02, 10, "combo", False, bank_size="full"))
    assert task.build_combo_segments(0.01, 20) is None
    assert task.build_reference_bank(0.01, 20, "combo", False).shape == (256, 2000)


# ---------------------------------------------------------------- C ABI surface
def _header_functions():
    hdr = open(os.path.join(ROOT, "include", "f16_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(f16_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    f16.build()
    lib = ctypes.CDLL(f16.library_path())
    names = _header_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/f16_b200.h but not exported"
    assert sorted(_capi.EXPORTS) == names
    lib.f16_abi_version.restype = ctypes.c_int
    assert lib.f16_abi_version() == _capi.ABI_VERSION == 2


def test_struct_layouts_match_header():
    """ctypes mirrors of the ABI structs: sizes as the C compiler lays them out."""
    assert ctypes.sizeof(_capi.Config) == 4 + 4 + 8 + 8 + 8 * 4 + 8 + 8 + 9 * 8 + 8   # 152 bytes
    assert ctypes.sizeof(_capi.EnvBuffers) == 11 * 8
    assert ctypes.sizeof(_capi.StepIO) == 7 * 8 + 8 + 8


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header compiles as strict C99 on its own (no C++, no CUDA, no torch types)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "hdr.c"
    src.write_text('#include "f16_b200.h"\nint main(void) { f16_config c; f16_env_buffers b; f16_step_io io; (void)c; (void)b; (void)io;'
                   ' return F16_ABI_VERSION == 0; }\n')
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                          "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_integration_md_struct_mirrors_match_the_binding():
    """The ctypes snippets a maintainer would copy from INTEGRATION.md lay the structs out like _capi.py (and the header)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = text[text.index("class Config(C.Structure)"):text.index("N, dev = 1 << 20")]
    ns = {"C": ctypes}
    exec(code, ns)   # noqa: S102 (our own documentation snippet)
    for name in ("Config", "EnvBuffers", "StepIO"):
        doc, ref = ns[name], getattr(_capi, name)
        assert ctypes.sizeof(doc) == ctypes.sizeof(ref), name
        assert [(f[0], ctypes.sizeof(f[1])) for f in doc._fields_] == [(f[0], ctypes.sizeof(f[1])) for f in ref._fields_], name


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _capi.load_library()
    assert lib.f16_init(0) != 0
    assert b"no CPU fallback" in lib.f16_last_error()
    cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "combo"}
    with pytest.raises(f16.F16Error):
        f16.F16VecEnv(cfg, 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "f16-model_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.lower() or fn == "f16_tables_data.h", f"{fn} mentions the oracle"


def test_split_k_linear_gradients_match_plain_linear():
    """The split-K backward of the PPO update's Linear layers is the same fp32 arithmetic in another summation order."""
    import torch

    import f16_model_b200.policy as pol   # importing the module needs neither a GPU nor a loaded library
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(), torch.nn.Linear(64, 1))
    x = torch.rand(16384, 5) * 2 - 1
    tgt = torch.randn(16384)
    loss_a = ((net(x).squeeze(-1) - tgt) ** 2).mean()
    ga = torch.autograd.grad(loss_a, list(net.parameters()))
    out_b = pol.mlp_forward(net, x).squeeze(-1)
    loss_b = ((out_b - tgt) ** 2).mean()
    gb = torch.autograd.grad(loss_b, list(net.parameters()))
    assert torch.allclose(loss_a, loss_b, rtol=1e-6)
    for a, b in zip(ga, gb):
        assert torch.allclose(a, b, rtol=2e-4, atol=1e-7), (a - b).abs().max().item()
    # small or ragged batches fall back to F.linear
    assert pol.split_k_linear(x[:100], net[0].weight, net[0].bias).shape == (100, 64)


def test_every_abi_struct_is_laid_out_like_the_c_compiler_does(tmp_path):
    """sizeof and every field offset of the ctypes mirrors against a C program compiled from the header itself (gcc): covers the
    structs the size test above does not spell out (f16_rollout_io, f16_peer_group, f16_ppo_step16_args)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = [("f16_config", _capi.Config), ("f16_env_buffers", _capi.EnvBuffers), ("f16_step_io", _capi.StepIO), ("f16_rollout_io", _capi.RolloutIO),
             ("f16_peer_group", _capi.PeerGroup), ("f16_ppo_step16_args", _capi.PpoStep16)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "f16_b200.h"', 'int main(void) {']
    for cname, mirror in pairs:
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    res = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr          # a field named in the mirror but not in the header fails right here
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, mirror in pairs:
        assert int(got[cname]) == ctypes.sizeof(mirror), cname
        for fname, _ in mirror._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(mirror, fname).offset, f"{cname}.{fname}"


def test_scatter_table_inverts_the_gather_permutations():
    """policy.scatter_table (the inverse maps f16_ppo_step16 scatters updated parameters through): every blob word that gathers parameter j
    is listed under j, nothing else is, padding words (perm 0) are never written; more copies than the table holds is an error."""
    import torch

    from f16_model_b200.policy import scatter_table

    g = torch.Generator().manual_seed(3)
    n = 500
    fwd = torch.zeros(700, dtype=torch.long)
    fwd[torch.randperm(700, generator=g)[:n]] = torch.randperm(n, generator=g) + 1                 # every parameter once, 200 padding words
    tail = torch.zeros(120, dtype=torch.long)
    tail[torch.randperm(120, generator=g)[:80]] = torch.randperm(n, generator=g)[:80] + 1          # a subset of the parameters once
    half = torch.zeros(900, dtype=torch.long)
    twice = torch.randperm(n, generator=g)[:300] + 1
    half[torch.randperm(900, generator=g)[:600]] = torch.cat([twice, twice])                       # 300 parameters twice (W2 and W2^T)
    t = scatter_table(n, fwd, tail, half)
    assert t.shape == (n, 4) and t.dtype == torch.int32
    for col, perm, blob_cols in ((0, fwd, (0,)), (1, tail, (1,)), (2, half, (2, 3))):
        rebuilt = torch.zeros_like(perm)
        for c in blob_cols:
            j = torch.nonzero(t[:, c] >= 0).reshape(-1)
            assert torch.all(rebuilt[t[j, c].long()] == 0)                                         # no word listed twice
            rebuilt[t[j, c].long()] = j + 1
        assert torch.equal(rebuilt, perm)
    with pytest.raises(ValueError):
        scatter_table(n, torch.cat([fwd, fwd]), tail, half)

"""Offline model of the shared-memory wavefronts of the table lookups (LDS.128) in the env-step kernel.

A warp's LDS.128 is served quarter-warp by quarter-warp (8 lanes x 16 B = 128 B per wavefront): lanes that read the
same 16-byte chunk are broadcast, lanes that read DIFFERENT chunks of the same 16-byte bank group (chunk index mod 8)
serialise.  The cells a lane reads depend on its env's state, so the cost depends on the state distribution: this
script rolls the CPU oracle under the bench's stick distribution (i.i.d. U(-1,1)), takes states from the stationary
part of the episodes, and counts wavefronts per lookup for a given table layout (strides / bases in 16-byte chunks).
Test infrastructure (it uses oracle/, hence it lives under tests/); run on the CPU:  python tests/studies/bank_conflict_model.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import f16_oracle as orc  # noqa: E402
import f16_model_b200 as f16  # noqa: E402

NA, NF1, NF3 = 19, 4, 6


def sample_states(n=8192, T=400, seed=3, stick_period=None):
    """States of n envs over T steps under a U(-1,1) stick: i.i.d. per step, or (stick_period = P) a P-step random
    pattern repeated -- what a bench that cycles a pool of P action rows feeds every env."""
    rng = np.random.default_rng(seed)
    d5 = np.radians(5)
    al = rng.uniform(-d5, d5, n)
    x0 = np.zeros((n, 9))
    x0[:, 1], x0[:, 2], x0[:, 3] = rng.uniform(2000, 4000, n), rng.uniform(-d5, d5, n), al
    x0[:, 4], x0[:, 5] = rng.uniform(90, 250, n), al
    if stick_period:
        acts = np.tile(rng.uniform(-1, 1, (stick_period, n)), (T // stick_period + 1, 1))[:T]
    else:
        acts = rng.uniform(-1, 1, (T, n))
    bank = f16.build_reference_bank(0.01, 10, "combo", True)
    r = orc.rollout(x0, acts, bank, np.zeros(n, dtype=np.int32), want_obs=False)
    return r["states"], r["ep_len"]


def stationary_cells(states, ep_len, lo=100, hi=None, seed=0):
    """cell indices of the env-steps lo..hi of still-running episodes, shuffled (lanes of a warp are independent envs)."""
    hi = hi or states.shape[0]
    xs = np.concatenate([states[t][ep_len > t] for t in range(lo, hi, 7)])
    xs = xs[np.random.default_rng(seed).permutation(xs.shape[0])]
    return cells(xs)


def cells(x):
    """cell indices of one step for states x [n, 9] (csrc/f16_device.cuh aero_coeffs / thrust)."""
    rad = np.pi / 180
    alpha, fi, H, V, Pa = x[:, 5], x[:, 6], x[:, 1], x[:, 4], x[:, 8]
    ac = np.clip(alpha, -20 * rad, 90 * rad)
    j = np.floor((ac + 20 * rad) / (5 * rad)).astype(int)
    j = np.where(j < 16, j, 16 + ((j - 16) >> 1))
    ia = np.clip(j, 0, NA - 1)
    fc = np.clip(fi, -25 * rad, 25 * rad)
    i1 = (fc >= -10 * rad).astype(int) + (fc >= 0) + (fc >= 10 * rad)
    i3 = i1 + (fc >= 15 * rad) + (fc >= 20 * rad)
    Hg = 6356766.0 * H / (6356766.0 + H)
    a = np.sqrt(1.4 * 287.05287 * (288.15 - 6.5e-3 * Hg))
    ih = np.clip(np.floor(H / 3048.0).astype(int), 0, 4)
    im = np.clip(np.floor(V / a / 0.2).astype(int), 0, 4)
    ip = (Pa >= 50).astype(int)
    return ia, i1, i3, ip, ih, im


def wavefronts(chunk):
    """chunk [n] = 16-byte chunk index each lane reads; lanes grouped 8 by 8 -> mean wavefronts per quarter-warp."""
    q = chunk[: chunk.size // 8 * 8].reshape(-1, 8)
    out = np.zeros(q.shape[0], dtype=int)
    for g in range(8):
        m = (q % 8) == g
        qs = np.where(m, q, -1)
        qs.sort(axis=1)
        distinct = ((qs[:, 1:] != qs[:, :-1]) & (qs[:, 1:] >= 0)).sum(axis=1) + (qs[:, 0] >= 0)
        out = np.maximum(out, distinct)
    return out.mean()


def layout_cost(ia, i1, i3, ip, ih, im, L, verbose=False):
    """L: strides of the table image in 16-byte fp32 chunks.  Mean wavefronts per quarter-warp summed over the 11
    loads of one env-step (11.0 = conflict-free)."""
    parts = {}
    for c in range(3):   # cell3: Cx, Cy, mz polynomials
        parts[f"cell3.{c}"] = wavefronts(L["cell3_base"] + i1 * L["cell3_f"] + ia * L["cell3_a"] + c)
    for c in range(4):   # arow: Cxwz, Cywz, mzwz cubics, dmz1 line
        parts[f"arow.{c}"] = wavefronts(L["arow_base"] + ia * L["arow_a"] + c)
    parts["dmzds"] = wavefronts(L["dmzds_base"] + ia * L["dmzds_a"] + i3)
    parts["eta"] = wavefronts(L["eta_base"] + i1)
    for c in range(2):
        parts[f"thrust.{c}"] = wavefronts(L["thrust_base"] + (ip * 5 + ih) * L["thrust_h"] + im * 2 + c)
    if verbose:
        print("  " + "  ".join(f"{k} {v:.2f}" for k, v in parts.items()))
    return sum(parts.values())


# the image the first round-1 builds used: rows of 16 / cells of 12 / 8 floats back to back
ORIGINAL = dict(cell3_base=0, cell3_f=NA * 3, cell3_a=3, arow_base=228, arow_a=4, dmzds_base=304, dmzds_a=NF3,
                eta_base=418, thrust_base=422, thrust_h=10)


def layout_from_header(path=None):
    """Strides of the image csrc/f16_tables.h declares (parsed from its constants and array bounds)."""
    import re

    path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "f16-model_b200", "csrc", "f16_tables.h")
    src = open(path).read()
    const = {k: int(eval(v, {}, {})) for k, v in re.findall(r"constexpr int (k\w+) = ([^;]+);", src)}   # noqa: S307 (integer expressions)

    def dims(name):
        m = re.search(r"R %s((?:\[[^\]]+\])+);" % name, src)
        return [int(eval(d, {}, dict(const))) for d in re.findall(r"\[([^\]]+)\]", m.group(1))]

    c3, ar, dz, et, th = dims("cell3"), dims("arow"), dims("dmzds"), dims("eta"), dims("thrust")
    L = dict(cell3_base=0, cell3_f=c3[1] * c3[2] // 4, cell3_a=c3[2] // 4)
    L["arow_base"] = c3[0] * c3[1] * c3[2] // 4
    L["arow_a"] = ar[1] // 4
    L["dmzds_base"] = L["arow_base"] + ar[0] * ar[1] // 4
    L["dmzds_a"] = dz[1] * dz[2] // 4
    L["eta_base"] = L["dmzds_base"] + dz[0] * dz[1] * dz[2] // 4
    L["thrust_base"] = L["eta_base"] + et[0] * et[1] // 4
    L["thrust_h"] = th[2] // 4
    return L


if __name__ == "__main__":
    CURRENT = layout_from_header()
    print("layout in f16_tables.h:", CURRENT)
    for name, period in (("i.i.d. stick", None), ("stick repeating every 8 steps (bench --pool 8)", 8), ("every 64 steps (bench default)", 64)):
        states, ep_len = sample_states(stick_period=period)
        idx = stationary_cells(states, ep_len)
        ia = idx[0]
        print(f"\n{name}: {ia.size} env-steps, {np.mean([len(set(q)) for q in ia[:8000].reshape(-1, 8)]):.2f} distinct alpha cells per quarter-warp")
        print(" original image :", f"{layout_cost(*idx, ORIGINAL, verbose=True):.2f}")
        print(" f16_tables.h   :", f"{layout_cost(*idx, CURRENT, verbose=True):.2f}")
        i1, i3 = idx[1], idx[2]
        print("  cell3 by plane stride mod 8:", {f: round(sum(wavefronts(i1 * (800 + f) + ia * 3 + c) for c in range(3)) / 3, 2) for f in range(8)})
        print("  dmzds by row stride mod 8  :", {a: round(wavefronts(ia * (800 + a) + i3), 2) for a in range(8)})
        print("  arow by row stride         :", {a: round(wavefronts(ia * a), 2) for a in (4, 5, 6, 7)})

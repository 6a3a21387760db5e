"""Offline model of the shared-memory wavefronts of the table lookups (LDS.128) in the env-step kernel.

A warp's LDS.128 is served quarter-warp by quarter-warp (8 lanes x 16 B = 128 B per wavefront): lanes that read the
same 16-byte chunk are broadcast, lanes that read DIFFERENT chunks of the same 16-byte bank group (chunk index mod 8)
serialise.  The cells a lane reads depend on its env's state, so the cost depends on the state distribution: this
script rolls the CPU oracle under the bench's stick distribution (i.i.d. U(-1,1)), takes states from the stationary
part of the episodes, and counts wavefronts per lookup for a given table layout (strides / bases in 16-byte chunks).
Test infrastructure (uses oracle/); run here on the CPU:  python tools/probes/bank_conflict_model.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import f16_oracle as orc  # noqa: E402
import f16_model_b200 as f16  # noqa: E402

NA, NF1, NF3 = 19, 4, 6


def sample_states(n=8192, T=400, seed=3):
    rng = np.random.default_rng(seed)
    d5 = np.radians(5)
    al = rng.uniform(-d5, d5, n)
    x0 = np.zeros((n, 9))
    x0[:, 1], x0[:, 2], x0[:, 3] = rng.uniform(2000, 4000, n), rng.uniform(-d5, d5, n), al
    x0[:, 4], x0[:, 5] = rng.uniform(90, 250, n), al
    # bench.py cycles a pool of 8 action rows per batch: every env sees a PERIODIC 8-step stick (DC offset + 12.5 Hz
    # harmonics), which spreads stab and alpha much wider than an i.i.d. stick would
    acts = np.tile(rng.uniform(-1, 1, (8, n)), (T // 8 + 1, 1))[:T]
    bank = f16.build_reference_bank(0.01, 10, "combo", True)
    r = orc.rollout(x0, acts, bank, np.zeros(n, dtype=np.int32), want_obs=False)
    return r["states"], r["ep_len"]


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
    """L: dict of base / strides (in chunks).  Returns total mean wavefronts per quarter-warp over the 12 loads."""
    tot = 0.0
    parts = {}
    for c in range(3):   # cell3: px, py, pm
        parts[f"cell3.{c}"] = wavefronts(L["cell3_base"] + i1 * L["cell3_f"] + ia * L["cell3_a"] + c)
    for c in range(4):   # arow: kx, ky, km, r3
        parts[f"arow.{c}"] = wavefronts(L["arow_base"] + ia * L["arow_a"] + c)
    parts["dmzds"] = wavefronts(L["dmzds_base"] + ia * L["dmzds_a"] + i3)
    parts["eta"] = wavefronts(L["eta_base"] + i1)
    cell = (ip * 5 + ih) * 5 + im
    for c in range(2):
        parts[f"thrust.{c}"] = wavefronts(L["thrust_base"] + cell * L["thrust_c"] + c)
    tot = sum(parts.values())
    if verbose:
        print("  " + "  ".join(f"{k} {v:.2f}" for k, v in parts.items()))
    return tot


ORIGINAL = dict(cell3_base=0, cell3_f=NA * 3, cell3_a=3, arow_base=228, arow_a=4, dmzds_base=304, dmzds_a=6,
                eta_base=418, thrust_base=422, thrust_c=2)
CURRENT = ORIGINAL

if __name__ == "__main__":
    states, ep_len = sample_states()
    for lo, hi in ((0, 20), (100, 400)):
        xs = np.concatenate([states[t][ep_len > t] for t in range(lo, hi, 7)])
        rng = np.random.default_rng(0)
        xs = xs[rng.permutation(xs.shape[0])]          # lanes of a warp are independent envs
        idx = cells(xs)
        print(f"steps {lo}..{hi}: {xs.shape[0]} samples; distinct ia per quarter-warp "
              f"{np.mean([len(set(q)) for q in idx[0][: 8000].reshape(-1, 8)]):.2f}, ideal total = 12 wavefronts")
        print(" current layout:", f"{layout_cost(*idx, CURRENT, verbose=True):.2f}")
        for arow_a in (5, 6, 7):
            L = dict(CURRENT, arow_a=arow_a)
            print(f" arow stride {arow_a} chunks:", f"{layout_cost(*idx, L, verbose=True):.2f}")

    # ---- search: strides only matter mod 8 (and must be >= the cell size) ----------------------------
    print("\nsearch on the stationary sample")
    ia, i1, i3, ip, ih, im = idx
    best = {}
    res = {(a, f): sum(wavefronts(i1 * (800 + f) + ia * a + c) for c in range(3)) for a in (3, 5, 7) for f in range(8)}   # 800 = 0 mod 8: distinct cells stay distinct
    for k in sorted(res, key=res.get)[:5]:
        print("  cell3 a-stride %d, f-stride = %d mod 8: %.3f" % (k[0], k[1], res[k] / 3))
    res = {a: wavefronts(ia * a) for a in (5, 7)}
    print("  arow:", res)
    res = {(a, b): wavefronts(ia * (800 + a) + i3 * b) for a in range(8) for b in (1, 3, 5, 7)}
    for k in sorted(res, key=res.get)[:5]:
        print("  dmzds a-stride = %d mod 8, f-stride %d: %.3f" % (k[0], k[1], res[k]))
    cell = (ip * 5 + ih) * 5 + im
    res = {(s, hs): wavefronts((ip * 5 + ih) * (800 + hs) + im * s) for s in (2, 3, 5, 7) for hs in range(8)}
    for k in sorted(res, key=res.get)[:6]:
        print("  thrust m-stride %d, h-stride = %d mod 8: %.3f" % (k[0], k[1], res[k]))
    print("\nall residues, stationary sample")
    print("  cell3 (a-stride 3) by f-stride mod 8:", {f: round(sum(wavefronts(i1 * (800 + f) + ia * 3 + c) for c in range(3)) / 3, 3) for f in range(8)})
    print("  dmzds (f-stride 1) by a-stride mod 8:", {a: round(wavefronts(ia * (800 + a) + i3), 3) for a in range(8)})
    print("  arow by a-stride:", {a: round(wavefronts(ia * a), 3) for a in (4, 5, 6, 7, 9, 11, 13)})

"""Pitch reference signals -- host-side mirror of ``ReferenceSignal`` (F16model/env/task.py:6-118).

Signals are built once on the host (NumPy, fp64) and uploaded as a bank ``[R, T]`` of
``wz_ref = theta_ref / 2``; every env follows one row of the bank and the step kernel reads
``wz_ref[ref_idx][episode_length]``.  Random draws go through Python's ``random`` module in
the reference's order, so ``random.seed(s)`` reproduces the reference's signal for seed s."""
import itertools
import random as _random

import numpy as np

SCENARIOS = ("step", "cos", "pure_step", "combo")
_AMPS6 = (20, 10, 5, -5, -10, -20)


class ReferenceSignal:
    def __init__(self, dt, tn, determenistic=False, scenario=None, rng=None):
        self.dt = dt
        self.tn = tn
        self.t = np.arange(0, tn, dt)
        self.determenistic = determenistic
        self._rng = _random if rng is None else rng
        if scenario is None:
            scenario = self._rng.choice(["step", "cos", "pure_step"])
        if scenario not in SCENARIOS:
            raise ValueError(f"unknown scenario {scenario!r}")
        self.scenario = scenario
        theta = getattr(self, f"_build_{scenario}")()
        self.theta_ref = theta
        self.wz_ref = theta / 2

    # -- building blocks ---------------------------------------------------
    def _nearest(self, time):
        return int(np.abs(self.t - time).argmin())

    def cosstep(self, start_time, w):
        """Raised-cosine ramp 0 -> 1 of width w starting at start_time (task.py:105-118)."""
        i0, i1 = self._nearest(start_time), self._nearest(start_time + w)
        ramp = -(np.cos(1 / w * np.pi * (self.t - start_time)) - 1) / 2
        if start_time >= 0:
            ramp[:i0] = 0.0
        ramp[i1:] = 1.0
        return ramp

    # -- scenarios ---------------------------------------------------------
    def _build_step(self):
        if self.determenistic:
            amp = 10 * np.pi / 180
        else:
            amp = np.radians(self._rng.choice(list(_AMPS6)))
        sig = amp * self.cosstep(-1, 2) * 2 - amp
        sig -= amp * self.cosstep(self.tn * 0.25, 1)
        sig -= amp * self.cosstep(self.tn * 0.50, 1)
        sig += amp * self.cosstep(self.tn * 0.75, 1)
        return sig[::-1]

    def _build_cos(self):
        if self.determenistic:
            amp, freq = 20 * np.pi / 180, 0.25
        else:
            amp = np.radians(self._rng.choice([10, 5, -5, -10]))
            freq = self._rng.choice([0.125, 0.25, 0.45, 0.5])
        return np.sin(2 * np.pi * freq * self.t) * amp

    def _build_pure_step(self):
        if self.determenistic:
            amp, t_step = np.radians(10), 1
        else:
            amp = np.radians(self._rng.choice(list(_AMPS6)))
            t_step = self._rng.choice([1, 2, 3, 4])
        sig = np.zeros(self.t.shape)
        sig[self._nearest(t_step):] = amp
        return sig

    def _build_combo(self):
        if self.determenistic:
            a_step, a_sin, freq, a_cos = np.radians(10), np.radians(10), 0.25, np.radians(20)
        else:
            a_step = np.radians(self._rng.choice(list(_AMPS6)))
            a_sin = np.radians(self._rng.choice(list(_AMPS6)))
            freq = self._rng.choice([0.125, 0.25, 0.45, 0.5, 0.75])
            a_cos = np.radians(self._rng.choice(list(_AMPS6)))
        return self._combo_from(a_step, a_sin, freq, a_cos, order=None)

    def _combo_segments(self, a_step, a_sin, freq, a_cos):
        n1 = int(1 / self.dt)
        rect = np.concatenate(([0] * int(n1 / 2), [a_step] * int(n1 * 2), [0] * int(n1 / 2)))
        sine = np.sin(2 * np.pi * freq * np.arange(0, 3, self.dt)) * a_sin
        ramp = self.cosstep(0, 1)[:n1]
        bump = np.concatenate((ramp * a_cos, [a_cos] * n1, ramp[::-1] * a_cos))
        return [rect, sine, bump]

    def _combo_from(self, a_step, a_sin, freq, a_cos, order):
        segs = []
        for _ in range(int(self.tn / 10)):
            segs += self._combo_segments(a_step, a_sin, freq, a_cos)
        if order is not None:
            segs = [segs[j] for j in order]
        elif not self.determenistic:
            self._rng.shuffle(segs)
        flat = np.concatenate(segs)   # tn < 10 -> ValueError, exactly like the reference
        sig = np.zeros(self.t.shape)
        sig[: len(flat)] = flat
        return sig


_FULL_COMBO_CACHE = {}
_COMBO_PARAMS = (_AMPS6, _AMPS6, (0.125, 0.25, 0.45, 0.5, 0.75), _AMPS6)   # step amplitude, sine amplitude, sine frequency, bump amplitude


def combo_parameter_sets():
    """The reference's discrete space of random ``combo`` signals (env/task.py:73-76, :100-101) in the order both device
    banks enumerate it: (step amplitude [deg], sine amplitude [deg], frequency [Hz], bump amplitude [deg], segment order)
    -- 6 x 6 x 5 x 6 parameter sets x 3! orders of the (pulse, sine, bump) segments = 6480 equally likely signals."""
    for a_s, a_n, f, a_c in itertools.product(*_COMBO_PARAMS):
        for order in itertools.permutations(range(3)):
            yield a_s, a_n, f, a_c, order


def build_combo_segments(dt, tn):
    """The random ``combo`` scenario as a SEGMENT bank (the device's segment mode, include/f16_b200.h ``seg_len``).

    Every combo signal of a 10 s episode is three 3 s segments in shuffled order followed by zeros, and each segment
    depends on one parameter group only: 6 pulses, 6 x 5 sines, 6 bumps.  So the 6480 signals (26 MB as fp32 rows) are
    43 rows of ``L = 3 / dt`` samples -- rows 0-5 pulses, 6-35 sines, 36-41 bumps, 42 zeros -- plus one packed code per
    signal: row of segment s in bits [8s, 8s+8), the zero row in the top byte.  Values are the very numbers
    ``ReferenceSignal`` computes (the segments are slices of its arrays, halved exactly), so a signal rebuilt from its
    code is bit-identical to the reference's.  Returns (bank [43, L] float64, codes uint32 [6480], L), or None when
    the scenario does not factor this way (multi-block episodes, a dt for which the three segments differ in length).
    """
    if int(tn / 10) != 1:
        return None
    proto = ReferenceSignal(dt, tn, True, "pure_step")
    T = proto.t.size
    rows, index = [], {}
    for a in _AMPS6:
        rect, _, _ = proto._combo_segments(np.radians(a), 0.0, 0.25, 0.0)
        index["rect", a] = len(rows)
        rows.append(np.asarray(rect, dtype=np.float64) / 2)
    for a, f in itertools.product(_AMPS6, _COMBO_PARAMS[2]):
        _, sine, _ = proto._combo_segments(0.0, np.radians(a), f, 0.0)
        index["sine", a, f] = len(rows)
        rows.append(np.asarray(sine, dtype=np.float64) / 2)
    for a in _AMPS6:
        _, _, bump = proto._combo_segments(0.0, 0.0, 0.25, np.radians(a))
        index["bump", a] = len(rows)
        rows.append(np.asarray(bump, dtype=np.float64) / 2)
    L = rows[0].size
    if any(r.size != L for r in rows) or not (3 * L <= T <= 4 * L) or len(rows) + 1 > 256:
        return None
    zero = len(rows)
    rows.append(np.zeros(L))
    codes = []
    for a_s, a_n, f, a_c, order in combo_parameter_sets():
        seg = (index["rect", a_s], index["sine", a_n, f], index["bump", a_c])
        codes.append(seg[order[0]] | seg[order[1]] << 8 | seg[order[2]] << 16 | zero << 24)
    return np.ascontiguousarray(np.stack(rows)), np.array(codes, dtype=np.uint32), L


def signals_from_codes(bank, codes, T):
    """Whole ``wz_ref`` rows [len(codes), T] rebuilt from packed segment codes (host side; tests, ``call('ref_signal')``)."""
    codes = np.atleast_1d(np.asarray(codes)).astype(np.uint32)
    L = bank.shape[1]
    out = np.concatenate([bank[(codes >> np.uint32(8 * s)) & np.uint32(0xFF)] for s in range(4)], axis=1)
    assert out.shape[1] == 4 * L
    return np.ascontiguousarray(out[:, :T])


def build_reference_bank(dt, tn, scenario, determenistic, bank_size=None, seed=0):
    """Bank of ``wz_ref`` rows for a batch of envs, shape [R, T] (fp64).

    ``bank_size=None`` (the default) serves the reference's own distribution wherever its parameter space is
    finite: every member of it, equally weighted -- for the random ``combo`` scenario of a 10 s episode that is
    ``bank_size="full"`` (6480 rows); only multi-block combo episodes (tn >= 20 s, 97,200 distinct signals) fall
    back to 256 draws.  An integer asks for that many draws where the space is larger than it.

    Deterministic configs give one row (three when ``scenario`` is None, which the
    reference still draws at random).  Random configs enumerate the reference's whole
    discrete parameter space when it has at most ``bank_size`` members (pure_step 24,
    cos 16, step 6) and otherwise (combo: 6*6*5*6 parameter sets x segment orders) hold
    ``bank_size`` draws made exactly as the reference makes them, from ``random.Random(seed)``;
    ``bank_size="full"`` enumerates all 6480 combo signals instead.
    A reset picks a row uniformly, so resets are distributed like the reference's.
    """
    rows = []
    if bank_size is None:
        bank_size = "full" if (scenario == "combo" and not determenistic and int(tn / 10) == 1) else 256
    if determenistic:
        for sc in ([scenario] if scenario is not None else ["step", "cos", "pure_step"]):
            rows.append(ReferenceSignal(dt, tn, True, sc).wz_ref)
        return np.ascontiguousarray(np.stack(rows))
    scen = [scenario] if scenario is not None else ["step", "cos", "pure_step"]
    proto = ReferenceSignal(dt, tn, True, "pure_step")
    enumerable = {"step": 6, "cos": 16, "pure_step": 24}
    if all(s in enumerable for s in scen):   # at most 3 x 48 rows: always enumerated in full
        # equal weight per scenario: replicate to the lcm so that uniform row choice == choice of
        # scenario followed by uniform choice of parameters
        per = int(np.lcm.reduce([enumerable[s] for s in scen]))
        for sc in scen:
            sigs = []
            if sc == "step":
                for a in _AMPS6:
                    amp = np.radians(a)
                    s_ = amp * proto.cosstep(-1, 2) * 2 - amp
                    s_ -= amp * proto.cosstep(tn * 0.25, 1)
                    s_ -= amp * proto.cosstep(tn * 0.50, 1)
                    s_ += amp * proto.cosstep(tn * 0.75, 1)
                    sigs.append(s_[::-1] / 2)
            elif sc == "cos":
                for a, f in itertools.product([10, 5, -5, -10], [0.125, 0.25, 0.45, 0.5]):
                    sigs.append(np.sin(2 * np.pi * f * proto.t) * np.radians(a) / 2)
            else:
                for a, ts in itertools.product(_AMPS6, [1, 2, 3, 4]):
                    s_ = np.zeros(proto.t.shape)
                    s_[proto._nearest(ts):] = np.radians(a)
                    sigs.append(s_ / 2)
            rows += sigs * (per // len(sigs))
        return np.ascontiguousarray(np.stack(rows))
    if bank_size == "full":
        key = (float(dt), float(tn))
        if key in _FULL_COMBO_CACHE:          # 0.6 s to enumerate: built once per (dt, tn), handed out read-only
            return _FULL_COMBO_CACHE[key]
        # the whole discrete space of the random `combo` scenario: 6 x 6 x 5 x 6 parameter sets x 3! segment orders
        # = 6480 rows (26 MB in fp32).  Exact distribution equivalence, at the price of a bank that no longer stays
        # hot in L2 when several large batches stream through it.
        if scenario != "combo" or int(tn / 10) != 1:
            raise ValueError('bank_size="full" enumerates the single-block combo scenario only')
        for a_s, a_n, f, a_c, order in combo_parameter_sets():
            rows.append(proto._combo_from(np.radians(a_s), np.radians(a_n), f, np.radians(a_c), order) / 2)
        bank = np.ascontiguousarray(np.stack(rows))
        bank.setflags(write=False)
        _FULL_COMBO_CACHE[key] = bank
        return bank
    rng = _random.Random(seed)
    for _ in range(bank_size):
        rows.append(ReferenceSignal(dt, tn, False, scenario, rng=rng).wz_ref)
    return np.ascontiguousarray(np.stack(rows))

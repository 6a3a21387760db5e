"""The packed table image is laid out for the shared-memory banks (csrc/f16_tables.h): the lookups are LDS.128 at
lane-dependent addresses, and neighbouring cells must fall into different 16-byte bank groups.  This CPU test keeps
the header's strides honest with the offline wavefront model (tests/studies/bank_conflict_model.py), which rolls the
oracle to get a realistic state distribution -- test infrastructure only, no GPU."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    spec = importlib.util.spec_from_file_location("bank_conflict_model", os.path.join(ROOT, "tests", "studies", "bank_conflict_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_header_layout_is_conflict_free_for_neighbouring_cells():
    m = _model()
    L = m.layout_from_header()
    # strides in 16-byte chunks, as documented in f16_tables.h
    assert L["arow_a"] % 2 == 1                    # odd: eight consecutive alpha rows hit eight different bank groups
    assert L["cell3_a"] == 3 and L["cell3_f"] % 8 == 4
    assert L["thrust_h"] % 2 == 1
    states, ep_len = m.sample_states(n=2048, T=200, seed=11)
    idx = m.stationary_cells(states, ep_len, lo=100)
    new, old = m.layout_cost(*idx, L), m.layout_cost(*idx, m.ORIGINAL)
    assert new < 12.2, new                         # 11 loads per env-step, 11.0 = one wavefront each
    assert old - new > 4.0, (old, new)             # the back-to-back image paid ~6 extra wavefronts per quarter-warp

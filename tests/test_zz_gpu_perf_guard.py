"""GPU (-m gpu), runs last: relative performance guards -- ratios measured in one process on one GPU, so they do not depend
on clocks or on the box.  They exist because a scheduling change once cost the small-molecule workloads 70 % without any
correctness test noticing."""
import numpy as np
import pytest
from conftest import golden_input

pytestmark = pytest.mark.gpu


def _best_ms(h, P, reps=6):
    best = 1e30
    for _ in range(reps):
        h.fock_rhf(P)
        best = min(best, h.stats()["last_fock_ms"])
    return best


def test_ket_slices_keep_small_molecules_fast():
    """SF6/TZ2P: the generic kernel's (bra, ket slice) work items against one warp per bra (option bra_split=0);
    measured 5.9 against 8.8 ms"""
    from unomol_b200 import capi
    from unomol_b200.basis import Basis
    b = Basis.from_patin(golden_input("tz2p.sf6"))
    h = capi.Handle(b)
    rng = np.random.default_rng(3)
    P = rng.standard_normal(b.no2)
    G1 = h.fock_rhf(P).copy()
    t_split = _best_ms(h, P)
    h.set_option("bra_split", 0)
    G0 = h.fock_rhf(P).copy()
    t_whole = _best_ms(h, P)
    assert np.max(np.abs(G1 - G0)) < 1e-12 * np.max(np.abs(G0))
    # best of 6 on each side; measured ratio 0.67 -- the guard only has to catch the 70 % regression it was written for
    assert t_split < 0.97 * t_whole, (t_split, t_whole)

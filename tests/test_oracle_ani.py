"""Pin the ANI oracle: golden TorchANI vectors of the reference's own C++ test, the compiled reference CPU class,
and the fp64 arbiter.  CPU only."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from systems import ANI2X, lattice, rel_err, cubic_box

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ani_water18.json")))


def close(expected, found, atol, rtol):
    # assertEqual of TestANISymmetryFunctions.h:8-12: fails only if BOTH criteria fail
    expected = np.asarray(expected, np.float64).ravel()
    found = np.asarray(found, np.float64).ravel()
    diff = np.abs(expected - found)
    with np.errstate(divide="ignore", invalid="ignore"):
        bad = (diff > atol) & (diff / expected > rtol)
    return not bad.any()


@pytest.mark.parametrize("case", ["nonperiodic", "periodic", "triclinic"])
@pytest.mark.parametrize("impl", ["oracle", "ref"])
def test_golden_torchani(case, impl):
    if impl == "ref" and O.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    c = G["cases"][case]
    pos = np.array(G["positions"], np.float32).reshape(-1, 3)
    r, a = O.ani_forward(pos, G["species"], 2, G["rcr"], G["rca"], G["radial_fn"], G["angular_fn"], box=c["box"], impl=impl)
    assert close(c["radial"], r, 1e-4, 1e-3)
    assert close(c["angular"], a, 1e-4, 1e-3)


def _ani2x_case(n, periodic, seed):
    pos, L = lattice(n, 2.154, 0.3, seed)
    species = np.random.default_rng(seed + 1).integers(0, 7, n).astype(np.int32)
    rfn, afn = O.fn_tables(ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])
    box = cubic_box(L) if periodic else None
    return pos, species, rfn, afn, box


@pytest.mark.parametrize("periodic", [False, True])
def test_oracle_matches_reference_and_fp64(periodic):
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    n = 216 if periodic else 60
    pos, species, rfn, afn, box = _ani2x_case(n, periodic, 1001)
    kw = dict(box=box)
    r0, a0 = O.ani_forward(pos, species, 7, 5.1, 3.5, rfn, afn, impl="ref", **kw)
    r1, a1 = O.ani_forward(pos, species, 7, 5.1, 3.5, rfn, afn, **kw)
    r2, a2 = O.ani_forward(pos, species, 7, 5.1, 3.5, rfn, afn, bits=64, **kw)
    assert rel_err(r1, r0) < 2e-6 and rel_err(a1, a0) < 2e-6
    assert rel_err(r0, r2) < 3e-6 and rel_err(a0, a2) < 3e-6
    rng = np.random.default_rng(7)
    rg = rng.standard_normal(r0.shape).astype(np.float32)
    ag = rng.standard_normal(a0.shape).astype(np.float32)
    g0 = O.ani_backward(pos, species, 7, 5.1, 3.5, rfn, afn, rg, ag, impl="ref", **kw)
    g1 = O.ani_backward(pos, species, 7, 5.1, 3.5, rfn, afn, rg, ag, **kw)
    g2 = O.ani_backward(pos, species, 7, 5.1, 3.5, rfn, afn, rg, ag, bits=64, **kw)
    assert rel_err(g1, g0) < 1e-5
    assert rel_err(g0, g2) < 1e-5 and rel_err(g1, g2) < 1e-5


def test_oracle_backward_is_gradient_of_forward():
    """Finite-difference check of the oracle's analytic backward on the fp64 build (procedure of TestANISymmetryFunctions.h:14-58)."""
    pos, species, rfn, afn, _ = _ani2x_case(24, False, 5)
    pos = (pos * 0.6).astype(np.float32)   # denser, so that angular terms are populated
    rng = np.random.default_rng(3)
    r, a = O.ani_forward(pos, species, 7, 5.1, 3.5, rfn, afn, bits=64)
    rg = rng.standard_normal(r.shape)
    ag = rng.standard_normal(a.shape)
    g = O.ani_backward(pos, species, 7, 5.1, 3.5, rfn, afn, rg, ag, bits=64)
    direction = rng.standard_normal(pos.shape)
    h = 2.0 ** -8   # exactly representable so that pos +- h*dir is what the float32 interface sees (approximately)
    def energy(p):
        rr, aa = O.ani_forward(p.astype(np.float32), species, 7, 5.1, 3.5, rfn, afn, bits=64)
        return float((rr * rg).sum() + (aa * ag).sum())
    ep = energy(pos.astype(np.float64) + h * direction)
    em = energy(pos.astype(np.float64) - h * direction)
    pp = (pos.astype(np.float64) + h * direction).astype(np.float32).astype(np.float64)
    pm = (pos.astype(np.float64) - h * direction).astype(np.float32).astype(np.float64)
    fd = ep - em
    an = float((g * (pp - pm)).sum())
    assert abs(fd - an) <= 2e-3 * abs(an) + 1e-6

"""Pin the CFConv oracle: SchNetPack golden outputs of the reference's own C++ test + the compiled reference class.  CPU only."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from systems import lattice, rel_err, cubic_box

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cfconv_water18.json")))


def golden_inputs():
    pos = np.array(G["positions"], np.float32).reshape(-1, 3)
    x = (0.1 * np.arange(8 * 18)).astype(np.float32).reshape(18, 8)
    return pos, x, np.array(G["w1"], np.float32), np.array(G["b1"], np.float32), np.array(G["w2"], np.float32), np.array(G["b2"], np.float32)


@pytest.mark.parametrize("case", ["nonperiodic", "periodic", "triclinic", "tanh"])
@pytest.mark.parametrize("impl", ["oracle", "ref"])
def test_golden_schnetpack(case, impl):
    if impl == "ref" and O.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    c = G["cases"][case]
    pos, x, w1, b1, w2, b2 = golden_inputs()
    y = O.cfconv(pos, 8, 5, 2.0, 0.5, c["activation"], w1, b1, w2, b2, x, box=c["box"], impl=impl)
    exp = np.array(c["output"]).reshape(18, 8)
    diff = np.abs(exp - y)
    assert not ((diff > 1e-4) & (diff / np.abs(exp) > 1e-3)).any()   # assertEqual of TestCFConv.h:8-12, tolerance :134


@pytest.mark.parametrize("periodic", [False, True])
def test_oracle_matches_reference(periodic):
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    n, W, Gn = 150, 32, 20
    pos, L = lattice(n, 2.154, 0.3, 44)
    box = cubic_box(L) if periodic else None
    w1 = rng.normal(0, 0.3, (W, Gn)); b1 = rng.normal(0, 0.3, W); w2 = rng.normal(0, 0.2, (W, W)); b2 = rng.normal(0, 0.2, W)
    x = rng.standard_normal((n, W)); g = rng.standard_normal((n, W))
    a = O.cfconv(pos, W, Gn, 5.0, 0.25, "ssp", w1, b1, w2, b2, x, box=box, out_grad=g, impl="ref")
    b = O.cfconv(pos, W, Gn, 5.0, 0.25, "ssp", w1, b1, w2, b2, x, box=box, out_grad=g)
    c = O.cfconv(pos, W, Gn, 5.0, 0.25, "ssp", w1, b1, w2, b2, x, box=box, out_grad=g, bits=64)
    assert a[3] == b[3] == c[3]
    for k in range(3):
        assert rel_err(b[k], a[k]) < 5e-6 and rel_err(a[k], c[k]) < 5e-6

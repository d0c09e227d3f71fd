"""Pin the numpy oracles for getNeighborPairs and PME against the reference's own known answers.  CPU only."""
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("neighbors_pme_oracle", os.path.join(HERE, "..", "oracle", "neighbors_pme_oracle.py"))
NP = importlib.util.module_from_spec(spec)
spec.loader.exec_module(NP)
G = json.load(open(os.path.join(HERE, "golden", "pme_openmm.json")))


def test_neighbor_doctest_cases():
    """The four examples of getNeighborPairs.py:104-138."""
    pos = np.array([[0.0, 0, 0], [1.0, 0, 0], [2.0, 0, 0]], np.float32)
    nb, d, r, f = NP.neighbor_pairs(pos, 3.0)
    assert nb.tolist() == [[1, 2, 2], [0, 0, 1]] and r.tolist() == [1.0, 2.0, 1.0]
    nb, d, r, f = NP.neighbor_pairs(pos, 1.5)
    assert nb.tolist() == [[1, -1, 2], [0, -1, 1]] and np.isnan(r[1]) and np.isnan(d[1]).all()
    nb, d, r, f = NP.neighbor_pairs(pos, 3.0, 6)
    assert nb.tolist() == [[1, 2, 2, -1, -1, -1], [0, 0, 1, -1, -1, -1]]
    nb, d, r, f = NP.neighbor_pairs(pos, 1.5, 6)
    assert nb.tolist() == [[1, 2, -1, -1, -1, -1], [0, 1, -1, -1, -1, -1]] and f == 2


@pytest.mark.parametrize("case", ["rectangular", "triclinic", "exclusions"])
def test_pme_openmm_golden(case):
    """Energies and forces 'computed with OpenMM' (TestPme.py:38-63, 86-112, 145-171), rtol 1e-4."""
    c = G["cases"][case]
    gx, gy, gz, order, alpha, coulomb = c["pme_args"]
    pos = np.array(c["pos"]); q = np.array([(i - 4) * 0.1 for i in range(9)]); box = np.array(c["box"], np.float64)
    e_d, f_d, _ = NP.pme_direct(pos, q, box, c["cutoff"], alpha, coulomb, c.get("excl"))
    e_r, f_r, _ = NP.pme_reciprocal(pos, q, box, (int(gx), int(gy), int(gz)), int(order), alpha, coulomb)
    assert np.allclose(c["edirect"], e_d, rtol=1e-4)
    assert np.allclose(c["erecip"], e_r, rtol=1e-4)
    assert np.allclose(c["expected_ddirect"], f_d, rtol=1e-4, atol=1e-4)
    assert np.allclose(c["expected_drecip"], f_r, rtol=1e-4, atol=1e-3)

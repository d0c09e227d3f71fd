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


@pytest.mark.parametrize("order", [4, 5])
def test_vectorised_pme_oracle_equals_loop_version(order):
    """pme_reciprocal_vec (used at BASELINE config 5 size) is the same arithmetic as the loop restatement pinned above, and
    pme_direct_sampled equals pme_direct on the sampled atoms of a cubic box."""
    rng = np.random.default_rng(order)
    n = 40
    box = np.array([[3.0, 0, 0], [0.4, 3.1, 0], [-0.3, 0.5, 2.9]])
    pos = rng.uniform(-0.5, 1.5, (n, 3)) @ box
    q = rng.uniform(-0.5, 0.5, n); q -= q.mean()
    a = NP.pme_reciprocal(pos, q, box, (12, 10, 14), order, 3.2, 138.935)
    b = NP.pme_reciprocal_vec(pos, q, box, (12, 10, 14), order, 3.2, 138.935)
    assert abs(a[0] - b[0]) <= 1e-12 * abs(a[0]) and np.allclose(a[1], b[1], rtol=1e-10, atol=1e-10) and np.allclose(a[2], b[2], rtol=1e-10, atol=1e-10)
    L = 3.0
    e, dp, dq = NP.pme_direct(pos, q, np.eye(3) * L, 1.2, 3.2, 138.935)
    s = np.arange(0, n, 5)
    dps, dqs, es = NP.pme_direct_sampled(pos, q, L, 1.2, 3.2, 138.935, s)
    assert np.allclose(dps, dp[s], rtol=1e-10, atol=1e-10) and np.allclose(dqs, dq[s], rtol=1e-10, atol=1e-10)

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib as O
from systems import ANI2X, lattice, protein_species, rel_err
from nnpops_b200.SymmetryFunctions import Holder
for n, a in ((600, 2.0), (5000, 2.0), (5000, 2.154)):
    pos, _ = lattice(n, a, 0.3, 11)
    species = protein_species(n)
    rfn, afn = O.fn_tables(ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])
    r_o, a_o = O.ani_forward(pos, species, 7, 5.2, 3.5, rfn, afn)
    p = torch.tensor(pos, device="cuda")
    ref = O.RefCudaANI(species, 7, 5.2, 3.5, rfn, afn, False)
    r0, a0 = ref.forward(p, None)
    ours = Holder(7, 5.2, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], species.tolist())
    r1, a1 = ours.forward(p, None)
    torch.cuda.synchronize()
    print(n, a, "ours vs oracle", rel_err(r1.cpu().numpy(), r_o), rel_err(a1.cpu().numpy(), a_o), " refcuda vs oracle", rel_err(r0.cpu().numpy(), r_o), rel_err(a0.cpu().numpy(), a_o),
          "caps", ours.caps, "max|a_o|", np.abs(a_o).max(), "max|a0|", float(a0.abs().max()), "max|a1|", float(a1.abs().max()))
    ref.close()

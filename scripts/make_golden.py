#!/usr/bin/env python
"""Transcribe the reference's own golden vectors into tests/golden/*.json.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python scripts/make_golden.py

Sources (relative to /root/reference/src):
  ani/TestANISymmetryFunctions.h:61-94      inputs of the 18-atom water cluster
  ani/TestANISymmetryFunctions.h:111-252    AEVs "computed with TorchANI" (non-periodic / cubic 9 A / triclinic)
  schnet/TestCFConv.h:83-127                CFConv inputs + weights
  schnet/TestCFConv.h:140-248               outputs "computed with SchNetPack" (4 cases)
  pytorch/pme/TestPme.py:17-171             energies and forces "computed with OpenMM" (3 cases)
The numbers are parsed from the reference text, nothing is typed by hand.
"""
import ast
import json
import os
import re

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def numbers(text):
    return [float(x) for x in re.findall(NUM, text)]


def c_function(src, name):
    """Body of `void name(...) { ... }` by brace matching."""
    start = src.index("void %s(" % name)
    i = src.index("{", start)
    depth, j = 0, i
    while True:
        if src[j] == "{":
            depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0:
                return src[i + 1:j]
        j += 1


def c_array(body, name):
    m = re.search(r"%s(?:\[[^\]]*\])+\s*=\s*\{" % re.escape(name), body)
    i = m.end() - 1
    depth, j = 0, i
    while True:
        if body[j] == "{":
            depth += 1
        elif body[j] == "}":
            depth -= 1
            if depth == 0:
                return numbers(body[i:j])
        j += 1


def make_ani():
    src = open(os.path.join(REF, "ani", "TestANISymmetryFunctions.h")).read()
    water = c_function(src, "testWater")
    pos = c_array(water, "positions")
    species = [int(x) for x in c_array(water, "species =")] if False else [int(v) for v in numbers(re.search(r"species\s*=\s*\{([^}]*)\}", water).group(1))]
    radial_fn = numbers(re.search(r"radialFunctions\s*=\s*\{(.*?)\};", water, re.S).group(1))
    angular_fn = numbers(re.search(r"angularFunctions\s*=\s*\{(.*?)\};", water, re.S).group(1))
    m = re.search(r"createSymmetryCalculator\(numAtoms, numSpecies, (%s), (%s)" % (NUM, NUM), water)
    cases = {}
    for fn, key in (("testWaterNonperiodic", "nonperiodic"), ("testWaterPeriodic", "periodic"), ("testWaterTriclinic", "triclinic")):
        body = c_function(src, fn)
        case = {"radial": c_array(body, "expectedRadial"), "angular": c_array(body, "expectedAngular")}
        case["box"] = c_array(body, "periodicVectors") if "periodicVectors[]" in body else None
        cases[key] = case
    data = {"source": "ani/TestANISymmetryFunctions.h", "n_atoms": 18, "n_species": 2, "positions": pos, "species": species,
            "radial_fn": radial_fn, "angular_fn": angular_fn, "rcr": float(m.group(1)), "rca": float(m.group(2)),
            "tolerance": {"atol": 1e-4, "rtol": 1e-3}, "cases": cases}
    assert len(pos) == 54 and len(species) == 18 and len(radial_fn) == 4 and len(angular_fn) == 16
    for c in cases.values():
        assert len(c["radial"]) == 18 * 2 * 2 and len(c["angular"]) == 18 * 3 * 4
    json.dump(data, open(os.path.join(OUT, "ani_water18.json"), "w"), indent=1)


def make_cfconv():
    src = open(os.path.join(REF, "schnet", "TestCFConv.h")).read()
    water = c_function(src, "testWater")
    data = {"source": "schnet/TestCFConv.h", "n_atoms": 18, "width": 8, "n_gaussians": 5, "cutoff": 2.0, "gaussian_width": 0.5,
            "positions": c_array(water, "positions"), "w1": c_array(water, "w1"), "w2": c_array(water, "w2"),
            "b1": numbers(re.search(r"b1\s*=\s*\{([^}]*)\}", water).group(1)),
            "b2": numbers(re.search(r"b2\s*=\s*\{([^}]*)\}", water).group(1)),
            "input": "x[i] = 0.1*i", "tolerance": {"atol": 1e-4, "rtol": 1e-3}, "cases": {}}
    for fn, key, act in (("testWaterNonperiodic", "nonperiodic", "ssp"), ("testWaterPeriodic", "periodic", "ssp"),
                         ("testWaterTriclinic", "triclinic", "ssp"), ("testWaterTanh", "tanh", "tanh")):
        body = c_function(src, fn)
        case = {"output": c_array(body, "expectedOutput"), "activation": act}
        case["box"] = c_array(body, "periodicVectors") if "periodicVectors[]" in body else None
        assert len(case["output"]) == 144
        data["cases"][key] = case
    assert len(data["w1"]) == 40 and len(data["w2"]) == 64
    json.dump(data, open(os.path.join(OUT, "cfconv_water18.json"), "w"), indent=1)


def make_pme():
    src = open(os.path.join(REF, "pytorch", "pme", "TestPme.py")).read()
    tree = ast.parse(src)
    out = {"source": "pytorch/pme/TestPme.py", "cases": {}}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("test_rectangular", "test_triclinic", "test_exclusions"):
            seg = ast.get_source_segment(src, node)
            case = {}
            case["pme_args"] = numbers(re.search(r"PME\(([^)]*?)(?:,\s*torch|,\s*excl)", seg).group(1))
            for key in ("pos", "expected_ddirect", "expected_drecip", "excl"):
                m = re.search(r"^\s*%s\s*=\s*(\[\[.*?\]\])" % key, seg, re.S | re.M)
                if m:
                    case[key] = ast.literal_eval(m.group(1))
            case["box"] = ast.literal_eval(re.search(r"box_vectors\s*=\s*torch\.tensor\((\[\[.*?\]\])", seg, re.S).group(1))
            case["cutoff"] = float(re.search(r"compute_direct\(positions, charges, (%s)" % NUM, seg).group(1))
            e = re.findall(r"np\.allclose\((%s), e(direct|recip)" % NUM, seg)
            for val, which in e:
                case["e" + which] = float(val)
            case["charges"] = "(i-4)*0.1"
            case["rtol"] = 1e-4
            out["cases"][node.name[5:]] = case
    assert set(out["cases"]) == {"rectangular", "triclinic", "exclusions"}
    json.dump(out, open(os.path.join(OUT, "pme_openmm.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    make_ani()
    make_cfconv()
    make_pme()
    print("golden vectors written to", os.path.normpath(OUT))

"""Generates tests/golden/rdf_reference.npz by IMPORTING the reference's own analysis script (scripts/rdf.py) and running its
rdfPair / rdfChains on walks written by the unmodified reference (the committed walk_*.npz fixtures). matplotlib is absent here
and only used by rdf.py's __main__ block, so an empty stand-in module is registered before the import.
Run from the repo root:  python tests/golden/make_golden_rdf.py   (needs /root/reference)"""
import importlib.util
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SARW_REFERENCE", "/root/reference")


def load_reference_rdf():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("reference_rdf", os.path.join(REF, "scripts", "rdf.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def counts_like_rdfChains(rdf, atoms, rMin, dr):
    """The accumulation loop of rdf.py:46-53 with the reference's rdfPair, without the normalisation (integer counts)."""
    chainIDs = np.unique(atoms[:, 0])
    atomChains = [atoms[np.where(atoms[:, 0] == c)[0], 1:] for c in chainIDs]
    intra = inter = None
    for i1, a1 in enumerate(atomChains):
        for i2, a2 in enumerate(atomChains):
            if i1 == i2:
                h = rdf.rdfPair(a1, a2, rMin, dr)[0]
                intra = h if intra is None else intra + h
            elif i1 > i2:
                h = rdf.rdfPair(a1, a2, rMin, dr)[0] * 2
                inter = h if inter is None else inter + h
    return intra, inter


def main():
    rdf = load_reference_rdf()
    out = {}
    cases = [("walk_pe_in_seed1.npz", (10.0, 10.0, 40.0), 2.0, 0.5), ("walk_dense_tiny_seed2.npz", (14.0, 14.0, 14.0), 2.0, 0.5),
             ("walk_graft_file.npz", (24.0, 24.0, 30.0), 0.0, 0.3), ("walk_small_defaults_seed1.npz", None, 1.0, 0.7)]
    n = 0
    for fname, box, rMin, dr in cases:
        w = np.load(os.path.join(HERE, fname))
        if box is None:
            sys.path.insert(0, os.path.dirname(HERE))
            import cases as C
            box = tuple(C.SMALL_DEFAULTS["boxSize"])
        atoms = np.column_stack([w["chainID"].astype(float), w["xyz"]])
        atoms = atoms[atoms[:, 0].argsort(kind="stable")]            # rdf.py:77
        rdf.L = np.array(box)
        intra, inter = counts_like_rdfChains(rdf, atoms, rMin, dr)
        # the normalised curves of rdfChains itself (rdf.py:39-57) for one case, to pin the x2 / ordering conventions end to end
        if n == 0:
            gi, ge, den, edges = rdf.rdfChains(atoms, rMin, dr)
            assert np.allclose(gi * den, intra) and np.allclose(ge * den, inter)
        edges = rdf.rdfPair(atoms[:1, 1:], atoms[:1, 1:], rMin, dr)[1]
        out[f"file{n}"] = fname; out[f"box{n}"] = np.array(box); out[f"rMin{n}"] = rMin; out[f"dr{n}"] = dr
        out[f"edges{n}"] = edges; out[f"intra{n}"] = intra.astype(np.int64); out[f"inter{n}"] = inter.astype(np.int64)
        print(fname, "atoms", len(atoms), "bins", len(edges) - 1, "intra", int(intra.sum()), "inter", int(inter.sum()))
        n += 1
    out["n"] = n
    np.savez_compressed(os.path.join(HERE, "rdf_reference.npz"), **out)


if __name__ == "__main__":
    main()

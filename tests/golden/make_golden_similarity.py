"""Golden vectors for SURVEY 8 f2 (second half): parse_formula_elements / element_similarity of the REFERENCE
(scripts/holdout/holdout_search.py:109-182, which parses with data/canonical_ordering.CanonicalOrderer).  Run in the build
container only:   PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_similarity.py
Writes tests/golden/similarity.pt: the formulas, every parsed composition, and the full similarity matrix."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, "/root/reference/scripts/holdout")
sys.argv = ["holdout_search.py"]
import holdout_search as H     # noqa: E402  (re-opens sys.stdout line buffered at import: needs a real stdout)
import torch                   # noqa: E402

FORMULAS = [
    "YBa2Cu3O7", "La(7/10)Sr(3/10)CuO4", "Mg0.9Al0.1B2", "MgB2", "Nb3Sn", "Nb(3)Sn", "LaH10", "H3S", "FeSe", "FeSe(1/2)Te(1/2)",
    "Ba(3/5)K(2/5)Fe2As2", "Bi2Sr2CaCu2O8", "Bi2Sr2Ca2Cu3O10", "HgBa2Ca2Cu3O8", "Tl2Ba2CuO6", "CuO", "Cu0O2", "Cu", "", "xyz",
    "La(7/10)La(3/10)CuO4", "O4CuSr(3/10)La(7/10)", "Pb", "PbBi", "Pb(99/100)Bi(1/100)", "Y(1/3)Ba(2/3)CuO(7/3)", "Sr2RuO4",
    "{381}Og{381}Og", "Og(1/7)Xx3", "C60K3", "K3C60", "NbN", "NbTi", "V3Si", "(1/2)Cu", "Cu(1/0)O", "Na0.35CoO2H2.6O1.3",
]


def main():
    parsed = [H.parse_formula_elements(f) for f in FORMULAS]
    n = len(FORMULAS)
    sim = torch.zeros((n, n), dtype=torch.float64)
    for i, a in enumerate(FORMULAS):
        for j, b in enumerate(FORMULAS):
            sim[i, j] = H.element_similarity(a, b)
    torch.save({"formulas": FORMULAS, "parsed": parsed, "similarity": sim, "meta": {"torch": torch.__version__}},
               os.path.join(HERE, "similarity.pt"))
    print(f"wrote similarity.pt: {n} formulas, {int((sim > 0).sum())} non-zero pairs")


if __name__ == "__main__":
    main()

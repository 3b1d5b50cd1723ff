"""Stand-in for biopython's PDBParser (reference: mdgen/protein.py:24). Never called on the
hot path; importing the reference only needs the name to exist."""


class PDBParser:  # pragma: no cover
    def __init__(self, *a, **k):
        pass

    def get_structure(self, *a, **k):
        raise NotImplementedError("biopython is not installed; PDB parsing is out of scope")

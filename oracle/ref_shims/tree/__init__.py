"""Stand-in for dm-tree: only `map_structure` over nested lists/tuples/dicts is needed
(reference: mdgen/residue_constants.py:1082)."""


def map_structure(fn, structure):
    if isinstance(structure, dict):
        return {k: map_structure(fn, v) for k, v in structure.items()}
    if isinstance(structure, (list, tuple)):
        return type(structure)(map_structure(fn, v) for v in structure)
    return fn(structure)

"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference (`/root/reference/src/mlconfgen`) in this
container so the oracle port (`oracle/edm_oracle.py`) can be validated against it and golden vectors
generated (`oracle/make_golden.py`).  `/root/reference` does not exist on the GPU box; there the
loader falls back to the staged unmodified copy `oracle/_ref` (recipe: oracle/build_ref.py, git-ignored).  Only tests,
`smoke()` and the reference / cpu_baseline legs of `bench.py` may import this module -- never the product.

rdkit is not installed here; the hot-path modules (egnn.py, equivariant_diffusion.py, adj_mat_seer.py,
the tensor half of utils/mol_utils.py) are pure torch, so rdkit is stubbed in sys.modules.
"""
import os
import sys
from unittest.mock import MagicMock

_STUBS = [
    "rdkit", "rdkit.Chem", "rdkit.Chem.rdDetermineBonds", "rdkit.Chem.rdmolops", "rdkit.Chem.AllChem",
    "rdkit.Chem.MolStandardize", "rdkit.Chem.MolStandardize.rdMolStandardize",
    "rdkit.Chem.rdFingerprintGenerator", "rdkit.DataStructs", "rdkit.DataStructs.cDataStructs",
    "rdkit.Geometry", "rdkit.Chem.rdMolTransforms", "rdkit.Chem.rdMolAlign", "rdkit.Chem.Descriptors",
    "rdkit.Chem.rdchem", "rdkit.Chem.rdShapeHelpers", "rdkit.Chem.rdMolDescriptors",
]


_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_path(path: str = None) -> str:
    """/root/reference/src in the build container; the staged unmodified copy oracle/_ref (oracle/build_ref.py) on the
    GPU box, where /root/reference does not exist."""
    if path:
        return path
    if os.environ.get("MLCG_REF"):
        return os.environ["MLCG_REF"]
    for cand in ("/root/reference/src", _STAGED):
        if os.path.isdir(os.path.join(cand, "mlconfgen")):
            return cand
    return "/root/reference/src"


def reference_available(path: str = None) -> bool:
    return os.path.isdir(os.path.join(reference_path(path), "mlconfgen"))


def load_reference(path: str = None):
    """Returns the imported `mlconfgen` package of the reference."""
    path = reference_path(path)
    if not reference_available(path):
        raise RuntimeError("reference not present at %s (run `python -m oracle.build_ref` in the build container)" % path)
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    if path not in sys.path:
        sys.path.insert(0, path)
    import mlconfgen  # noqa
    import mlconfgen.egnn  # noqa
    import mlconfgen.equivariant_diffusion  # noqa
    import mlconfgen.adj_mat_seer  # noqa
    import mlconfgen.utils.mol_utils  # noqa
    return mlconfgen

from . import ChemicalComposition  # noqa: F401  (pickle fixture references this path)

"""pyqms_b200 -- B200-native implementation of pyQms's two hot paths behind the pyQms API.

    import pyqms_b200 as pyqms
    lib = pyqms.IsotopologueLibrary(molecules=[...], charges=[...])
    results = lib.match_all(mz_i_list=..., file_name=..., spec_id=..., spec_rt=..., results=None)

Exports mirror pyqms/__init__.py:38-42 of the reference.
"""
from . import adaptors, knowledge_base
from .params import params
from .results import Results, match, m_key
from .isotopologue_library import IsotopologueLibrary

__version__ = "0.1.0"
__all__ = ["IsotopologueLibrary", "Results", "match", "m_key", "params", "knowledge_base", "adaptors"]

"""Stand-in for the un-vendored PyPI dependency ``unimod_mapper`` (TEST INFRASTRUCTURE ONLY).

`/root/reference/pyqms/adaptors.py:32` imports it; ``parse_evidence`` only needs ``id2first_name`` for
mzTab input (adaptors.py:364-366).  Never imported by the product."""
import xml.etree.ElementTree as ET
from pathlib import Path

_DEFAULT_UNIMOD = Path("/root/reference/pyqms/kb/ext/unimod.xml")


class UnimodMapper:
    def __init__(self, xml_file_list=None, add_default_files=True, **kw):
        files = [str(f) for f in (xml_file_list or [])]
        if add_default_files and str(_DEFAULT_UNIMOD) not in files:
            files.append(str(_DEFAULT_UNIMOD))
        self._first_name = {}
        for path in files:
            for element in ET.parse(path).getroot().iter():
                if element.tag.endswith("}mod") and element.get("record_id") is not None:
                    self._first_name.setdefault(element.get("record_id"), element.get("title"))

    def id2first_name(self, unimod_id):
        return self._first_name.get(str(unimod_id))

"""Stub so that `/root/reference/pyqms/adaptors.py:32` imports; adaptors are out of scope."""


class UnimodMapper:
    def __init__(self, *a, **kw):
        pass

"""TEST INFRASTRUCTURE.  /root/reference/src/python/seulext.py:4 imports `LU` from janus-linear and never uses it."""


class LU:
    pass

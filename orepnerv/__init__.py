"""orepnerv — importable alias of the package directory
`boosting-neural-video-representation-via-online-structural-reparameteration_b200/` (whose name is not a
Python identifier).  `import orepnerv.model`, `orepnerv.utils`, ... resolve to the files in that directory.
"""
import os as _os

_PKG_DIR = _os.path.abspath(_os.path.join(_os.path.dirname(__file__), "..",
    "boosting-neural-video-representation-via-online-structural-reparameteration_b200"))
__path__.append(_PKG_DIR)
PKG_DIR = _PKG_DIR

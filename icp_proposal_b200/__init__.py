"""Import shim: the product package lives in ``icp-proposal_b200/`` (a name Python cannot import
directly); this module exposes it as ``icp_proposal_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "icp-proposal_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))

"""bfm_b200 - B200-native libbfm: the FEM hot path of obiwac/bfm (assembly + global solve) on sm_100a.

The product is ``lib/libbfm.so`` (C ABI, see include/); this package is the thin Python mirror of the
reference's pybfm object model (``api``) plus the bfmx_* additions (``ext``).
"""

from .api import (  # noqa: F401
	Binding, CInstance, CMaterial, CObj, CRule, CSim, Condition, Ez_lepl1110, Force, Force_funky, Force_linear, Force_none,
	Instance, Material, Mesh, Mesh_lepl1110, Mesh_wavefront, Obj, Rule, Rule_gauss_legendre, Sim, System, Vec, default_binding,
)
from . import ext  # noqa: F401

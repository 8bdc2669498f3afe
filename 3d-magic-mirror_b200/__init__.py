"""3d-magic-mirror_b200 -- Blackwell-native differentiable render-and-compare path of
layumi/3D-Magic-Mirror (reference: networks.DiffRender + the Kaolin DIB-R calls under it).

The directory name is not a Python identifier; load the package through
`__graft_entry__.load_package()` (registers it as `magic_mirror_b200`).
"""
from .diffrender import DiffRender                                            # noqa: F401
from .camera import camera_position_from_spherical_angles, generate_transformation_matrix   # noqa: F401
from .mesh import TemplateMesh, load_obj, save_obj, icosphere                # noqa: F401
from ._lib import MagicMirrorError, LIB_PATH, lib                            # noqa: F401
from .template_em import template_update, sharded_template_update          # noqa: F401

__all__ = ["DiffRender", "camera_position_from_spherical_angles", "generate_transformation_matrix",
           "TemplateMesh", "load_obj", "save_obj", "icosphere", "MagicMirrorError", "LIB_PATH", "lib",
           "template_update", "sharded_template_update"]

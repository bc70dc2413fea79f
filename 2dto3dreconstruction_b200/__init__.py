"""B200-native registration hot path for 3370sohail/2dto3dreconstruction.

Hand-written sm_100a CUDA kernels behind a C ABI (include/r3d_b200.h), with Python modules that
keep the reference's call signatures:

    utils.utils.depth_to_voxel / depth_to_voxel_ld          (back-projection)
    utils.icp.icp / nearest_neighbor / best_fit_transform   (ICP)
    utils.r3d.rigid_transform_3D                            (Kabsch)
    utils.transformation3d.homo_rigid_transform_3d / get_transformed_points / register_imgs
    utils.homography_utils.q9.ransac_loop / get_ssd_v2 / get_transformed_points / make_binned_array
    reconstruct.depth_images_to_3d_pts[_ld] / chain_prefix

plus ``pipeline`` for batched, device-resident registration of many frame pairs and its
multi-GPU sharding.  There is no CPU fallback: without the compiled library or without an
sm_100 device every compute call raises.
"""
from . import _capi  # noqa: F401

__all__ = ["_capi", "ops", "pipeline", "reconstruct", "utils"]
__version__ = "0.1.0"

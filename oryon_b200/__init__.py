"""oryon_b200 -- B200 (sm_100a) implementation of the Oryon inference hot path.

Layout mirrors the reference's module paths for the path it replaces:

    oryon_b200.utils.pcd           nn_correspondences, lift_pcd          (reference utils/pcd.py)
    oryon_b200.utils.coordinates   scale_coords, get_valid_coords        (reference utils/coordinates.py)
    oryon_b200.utils.pointdsc.init get_pointdsc_solver, get_pointdsc_pose (reference utils/pointdsc/init.py)
    oryon_b200.pipeline            FPM_Pipeline                          (reference pipeline.py)

All arithmetic runs in liboryon_b200.so (hand-written CUDA, C ABI in include/oryon_b200.h) reached
through ctypes (``oryon_b200._lib``).  There is no CPU or PyTorch fallback: a missing library or a
non-sm_100 device raises.
"""
__version__ = "0.1.0"

"""texpose_b200 -- B200 (sm_100a) kernels behind TexPose's per-ray NeRF render hot path.

Sub-modules mirror the reference's files for that path: `camera`, `tools.ray_sampler`,
`layers.nerf_static_transient_light`, `layers.nerf`, `model.nerf_adapt_st_gan`, `compute_box`,
`compute_surfelinfo`.  All arithmetic runs in `libtexpose_b200.so` (C-ABI: include/texpose_b200.h).
"""
__version__ = "0.1.0"

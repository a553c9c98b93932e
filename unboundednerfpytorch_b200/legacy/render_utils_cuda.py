"""Drop-in for the reference extension module ``render_utils_cuda`` (FourierGrid/cuda/render_utils.cpp:170-184):
same 13 function names and signatures, served by libubnerf_b200.so."""
from unboundednerfpytorch_b200.ops import (  # noqa: F401
    infer_t_minmax, infer_n_samples, infer_ray_start_dir, sample_pts_on_rays, sample_ndc_pts_on_rays,
    sample_bg_pts_on_rays, maskcache_lookup, raw2alpha, raw2alpha_backward, raw2alpha_nonuni,
    raw2alpha_nonuni_backward, alpha2weight, alpha2weight_backward)

"""lemo_b200 -- B200-native (sm_100a) temporal body-fitting engine behind sanweiliti/LEMO's call surface.

Sub-modules mirror the reference's import paths for the fitting hot path:
    lemo_b200.smplx                      <- smplx.create / body_model(**params)
    lemo_b200.vposer                     <- human_body_prior load_vposer(...).decode
    lemo_b200.models.AE / models.AE_sep  <- models/AE.py, models/AE_sep.py
    lemo_b200.temp_prox.dist_chamfer     <- temp_prox/dist_chamfer.py
    lemo_b200.utils.utils                <- utils/utils.py (6D/aa conversions, gen_body_mesh_v1, ...)
    lemo_b200.fit                        <- the Adam inner loops of opt_amass_perframe.py / opt_amass_temp.py
    lemo_b200.shard                      <- round-robin sequence sharding over the GPUs of one box
Everything computes in liblemo_b200.so (hand-written CUDA, C ABI in include/lemo_b200.h); no CPU fallback.
"""
__version__ = '0.1.0'

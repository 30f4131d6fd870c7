"""oracle/ -- CPU restatement of the reference's fitting hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under lemo_b200/ imports this package.  The only legal importers are tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, and there only as the
checker or the timed CPU baseline -- never as a fallback for the CUDA path.

What is pinned and what is not:
  * LBS core (blend shapes, joint regression, Rodrigues, rigid chain, skinning): restated from
    /root/reference/human_body_prior/body_model/lbs.py:34-263 and PINNED -- oracle/make_golden.py runs
    the vendored reference lbs() in the build container on the same synthetic model and the outputs
    are committed under tests/golden/ (tests/test_oracle_golden.py re-checks them on every run).
  * Enc / AE networks: restated from /root/reference/models/AE_sep.py:11-99 and models/AE.py:11-108,
    PINNED against the reference modules run with the shipped weights (runs/15217, runs/59547).
  * Index tables: taken verbatim (data) from loader/SSM2*.json, body_segments, foot_verts_id.
  * smplx==0.1.26 wrapper semantics, torchgeometry==0.1.2 conversions, the external `chamfer`
    extension: third-party code ABSENT from /root/reference and from this image.  Restated from
    their published algorithms -- "parity unpinned" for exactly these three.
"""

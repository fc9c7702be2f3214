"""oracle/ -- CPU restatement of the LSENeRF per-ray training/render path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lsenerf_b200/`` imports this package.  The only
legitimate importers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the *checker* or as the timed CPU
baseline, never as the thing shipped.

What it restates (file:line relative to /root/reference):

* ``poses.py``    -- lse_nerf/interpolation_utils.py:14-246, lse_nerf/ns_camera_optimizer.py:130-197,258-329
* ``cameras.py``  -- lse_nerf/lse_cameras.py:52-73,339-395,487-493,539-586, lse_nerf/lse_ray_generator.py
* ``field.py``    -- lse_nerf/lse_field.py:43-91,264-360 (+ nerfstudio 0.3.2 HashEncoding/MLP/SHEncoding/
                     SceneContraction/trunc_exp, torch "non-tcnn" path)
* ``nerfacc.py``  -- nerfacc 0.5.2 traverse_grids / visibility / pack_info / render_weight_from_density /
                     accumulate_along_rays / OccGridEstimator update (C part in ``nerfacc_ref.c``)
* ``model.py``    -- lse_nerf/lsenerf.py:278-439, lse_nerf/lse_grid_estimator.py:15-144,
                     lse_nerf/lse_renderer.py, lse_nerf/lse_embeddings.py, lse_nerf/intensity_mappers.py
* ``data.py``     -- lse_nerf/lse_pixel_sampler.py, lse_nerf/utils.py:153-194, lse_nerf/data_components.py,
                     lse_nerf/lse_datamanager.py:135-144,337-372, lse_nerf/lse_pipeline.py:110-145

Pinning status
--------------
* SE(3) interpolation (``poses.py``): PINNED.  ``tests/golden/make_pose_golden.py`` imports the reference's
  own ``lse_nerf/interpolation_utils.py`` in the build container and stores its outputs (and the scipy
  ground truth its self-tests use, interpolation_utils.py:392-458) as ``tests/golden/pose_*.npz``;
  ``tests/test_oracle_poses.py`` checks the restatement against them.
* Everything else: **PARITY UNPINNED**.  The arithmetic lives in nerfstudio 0.3.2 / nerfacc 0.5.2 /
  tinycudann 1.7, none of which is vendored under /root/reference or installed in this image, and the
  reference has no test touching it.  The restatement follows the published
  semantics recorded in SURVEY.md Appendix A and is frozen here as this repo's oracle; every function
  names the upstream symbol it restates.
"""

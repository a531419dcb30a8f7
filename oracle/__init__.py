"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's algorithms for the hot path
(ngp_oracle.c, tcnn_oracle.py), the recipe that builds the reference's own CUDA kernels in place
(build_ref.py -> oracle/_ref/) and an end-to-end CPU pipeline made of those pieces (cpu_pipeline.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (instant-ngp-pp_b200/) never does."""

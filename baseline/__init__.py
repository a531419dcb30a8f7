"""Reference-arm harness (TEST / BENCH INFRASTRUCTURE, never imported by the product package).

  install_ref.py   installs the reference's own Python glue, byte for byte, under baseline/_ref/ (git-ignored)
  shims/           import shims for the reference's missing third-party deps (torch_scatter)
  tcnn_standin.py  plain-PyTorch-op `tinycudann` stand-in (tiny-cuda-nn is un-vendored and absent offline)
  ref_harness.py   loads the unmodified reference modules with switchable `vren` / `tinycudann` back ends
  ref_train.py     the reference's training_step (train.py:268-345) without Lightning, for bench + tests
"""

"""ORACLE — test infrastructure only.

CPU restatement of the reference's sampling hot path (schedulers, sampling loop, denoiser forward).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this package; the
product package (`flow-matching-and-diffusion-models_b200/`, alias `fmdm_b200`) never does.
"""

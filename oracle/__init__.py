"""CPU oracles for the dae kernels — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker or the timed CPU baseline.  Nothing under
dynamic-asr-eval_b200/ imports it; the product path has no CPU fallback.
"""

"""Offline (CPU, numpy + the oracle as the checker) models of the gather kernels' work counts.
They live under tests/ because they use the oracle, which only test code may touch; they are
scripts (`python tests/model/<name>.py`), not collected tests."""

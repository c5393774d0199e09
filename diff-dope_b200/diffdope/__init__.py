"""B200-native Diff-DOPE: same Python API as NVlabs/diff-dope's `diffdope` package
(`diffdope/__init__.py:1-7`), hot path in libddope_b200.so (hand-written sm_100a CUDA)."""

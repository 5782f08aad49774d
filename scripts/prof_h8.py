"""One submission of the tensor-core Hamming kernel (8 images x 8000 AKAZE-like 61-byte descriptors, 16 symmetric pairs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opensfm_b200 import matching, synthetic as syn  # noqa: E402

pm = matching.PairMatcher()
for i in range(8):
    pm.add(i, syn.binary_descriptors(8000, 50 + i, 61))
pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)][:16]
for _ in range(2):
    pm.submit(pairs, 0.8, True)
    pm.sync()
print(pm.last_kernel(), pm.device_ms())

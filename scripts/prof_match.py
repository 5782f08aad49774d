"""One tcgen05 matching batch (for ncu): 24 images x 4000 integer-valued descriptors, 64 symmetric pairs."""
import sys
sys.path.insert(0, ".")
import numpy as np
from opensfm_b200 import matching, synthetic as syn
n_img, n_desc = 24, 4000
pm = matching.PairMatcher()
for i in range(n_img):
    pm.add(i, syn.hahog_like_descriptors(n_desc, 100 + i))
pairs = [(i, j) for i in range(n_img) for j in range(i + 1, n_img)][:64]
for _ in range(2):
    pm.submit(pairs, 0.8, True)
    pm.sync()
print("kernel", pm.last_kernel(), "ms", pm.device_ms(), "pairs/s", 2 * len(pairs) * n_desc * n_desc / (pm.device_ms()[1] * 1e-3))

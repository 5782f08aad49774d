"""Throughput of the non-tensor-core matcher paths (for bench / ncu): general float32 (cv2-order exact kernel),
Hamming 61-byte AKAZE and 32-byte ORB.  8 images x 8000 descriptors, 16 symmetric pairs each."""
import sys
sys.path.insert(0, ".")
import numpy as np
from opensfm_b200 import matching, synthetic as syn

n_img, n_desc = 8, 8000
pairs = [(i, j) for i in range(n_img) for j in range(i + 1, n_img)][:16]


def run(name, make):
    pm = matching.PairMatcher()
    for i in range(n_img):
        pm.add(i, make(i))
    for _ in range(3):
        pm.submit(pairs, 0.8, True)
        pm.sync()
    tot, ker = pm.device_ms()
    print("%-12s kernel %d  %.3f ms  %.3e descriptor-pairs/s" % (name, pm.last_kernel(), ker,
                                                               2 * len(pairs) * n_desc * n_desc / (ker * 1e-3)))


run("float128", lambda i: np.random.RandomState(i).rand(n_desc, 128).astype(np.float32))
run("akaze61", lambda i: syn.binary_descriptors(n_desc, 50 + i, 61))
run("orb32", lambda i: syn.binary_descriptors(n_desc, 70 + i, 32))
run("hahog_tc", lambda i: syn.hahog_like_descriptors(n_desc, 100 + i))

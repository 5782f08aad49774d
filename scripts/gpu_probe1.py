"""First GPU bring-up: each stage isolated, most dangerous (tcgen05) last."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, ".")
from oracle import match_oracle as mo, ba_lm as oracle
from opensfm_b200 import matching, bundle, synthetic as syn, ba_problem as bp

def stage(name, fn):
    t = time.time()
    try:
        fn(); print("[OK ] %s (%.2fs)" % (name, time.time() - t), flush=True)
    except Exception:
        print("[ERR] %s" % name, flush=True); traceback.print_exc(); sys.stdout.flush()

cfg = {"lowes_ratio": 0.8}
which = sys.argv[1]
if which == "simt":
    def f():
        a = syn.hahog_like_descriptors(1500, 1); b = syn.hahog_like_descriptors(1700, 2)
        b[:800] = np.clip(a[:800] + np.random.RandomState(3).randint(-6, 7, (800, 128)), 0, 255)
        pm = matching.PairMatcher(kernel=1); pm.add("a", a); pm.add("b", b)
        got = pm.match_pairs([("a", "b")], cfg, symmetric=False)[("a", "b")]
        ref = mo.match_brute_force(a, b, cfg)
        print("simt matches", len(got), len(ref), [tuple(x) for x in got.tolist()] == ref, "ms", pm.device_ms())
        u1 = syn.binary_descriptors(700, 1); u2 = syn.binary_descriptors(900, 2); u2[:300] = u1[:300]; u2[:300, :3] ^= 5
        print("hamming", matching.match_brute_force(u1, u2, cfg) == mo.match_brute_force(u1, u2, cfg))
    stage("simt matcher", f)
elif which == "ba":
    def g():
        pt=[1,2,3]; rt=[.1,.2,.3,.4,.5,.6]
        for t, c in {0:[0.3,0.1,-0.03], 5:[0.3,1.0,0.001,-0.02,0.1,-0.03,0.001,-0.005,0.01,0.006,0.02,0.003,0.001,-0.009,-0.01,0.03], 6:[0.0]}.items():
            got = bundle.eval_observation(t, c, rt, rt, True, pt, [.5,.5], 10.0)
            ref = oracle.reprojection(t, c, rt, rt, True, pt, [.5,.5], 10.0, autodiff=True)
            print("eval", t, max(np.abs(x-y).max() for x,y in zip(got,ref)))
    stage("ba eval_observation", g)
    def h():
        sc = syn.cube_scene(10, 1000, 1.0, with_descriptors=False); pb = syn.scene_to_problem(sc)
        ref = oracle.solve(pb)
        t=time.time(); got = bundle.solve(pb); dt=time.time()-t
        print("ref", ref["initial_cost"], ref["final_cost"], ref["iterations"], ref["message"])
        print("gpu", got["summary"], dt)
        print("dpts", np.abs(got["points"]-ref["points"]).max(), "dinst", np.abs(got["inst"]-ref["inst"]).max(), "dcam", np.abs(got["cam_params"]-ref["cam_params"]).max())
    stage("ba solve small", h)
elif which == "tc":
    def k():
        a = syn.hahog_like_descriptors(1500, 1); b = syn.hahog_like_descriptors(1700, 2)
        b[:800] = np.clip(a[:800] + np.random.RandomState(3).randint(-6, 7, (800, 128)), 0, 255)
        pm = matching.PairMatcher(kernel=2); pm.add("a", a); pm.add("b", b)
        got = pm.match_pairs([("a", "b")], cfg, symmetric=False)[("a", "b")]
        ref = mo.match_brute_force(a, b, cfg)
        g = [tuple(x) for x in got.tolist()]
        print("tc matches", len(g), len(ref), g == ref, "ms", pm.device_ms(), "kernel", pm.last_kernel())
        if g != ref:
            print("first diffs", [x for x in g if x not in set(ref)][:10], [x for x in ref if x not in set(g)][:10])
    stage("tc matcher", k)

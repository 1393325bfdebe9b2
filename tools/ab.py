"""A/B two builds of libmvtm.so on the same box: python tools/ab.py libA.so[@ENV=VAL,...] libB.so [workload[:docs] ...]
Each (library, workload) runs in its own process; prints the mean device ms per view pass of sweeps 5..12."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import ctypes, mvtopicmodel_b200._lib as L
L.SO_PATH = sys.argv[1]
_so = ctypes.CDLL(L.SO_PATH)          # an older build may lack entry points added since: bind what it has
L.SIGNATURES = {k: v for k, v in L.SIGNATURES.items() if hasattr(_so, k)}
from mvtopicmodel_b200 import Engine, corpus
wl, docs = sys.argv[2], (int(sys.argv[3]) if sys.argv[3] != "0" else None)
if wl == "uniform_k1000": K, Vs, views = corpus.generate_uniform(docs or 100_000, 1000, 400_000, 200)     # HBM-bound: no word reuse
elif wl == "acmtext": K, Vs, views = corpus.generate(dict(D=docs or 200_000, K=1000, views=[(100_000, 120, 0.5, 1.0, 1024)]))   # single view, K = 1000
else: K, Vs, views = corpus.generate(wl, docs=docs)
e = Engine(K, Vs, views, seed=1); e.init_assignments()
acc = [0.0] * len(views); n = 0
for it in range(1, 13):
    e.sweep(it)
    if it >= 5:
        st = e.stats(); n += 1
        for m in range(len(views)): acc[m] += st["ms_view"][m]
assert e.check_invariants() == 0
tok = sum(e.ntok)
print("%%s %%s ms/view %%s  total %%.3f ms  %%.3f Gtok/s" %% (os.path.basename(sys.argv[1]), wl, [round(a / n, 3) for a in acc], sum(acc) / n, tok / (sum(acc) / n) / 1e6))
''' % ROOT
libs = [a for a in sys.argv[1:] if ".so" in a]
wls = [a for a in sys.argv[1:] if ".so" not in a] or ["lda_100k"]
for wl in wls:
    name, _, docs = wl.partition(":")
    for rep in range(int(os.environ.get("AB_REPS", "2"))):
        for lib in libs:
            path, _, envs = lib.partition("@")
            env = dict(os.environ)
            for kv in filter(None, envs.split(",")):
                k, _, v = kv.partition("="); env[k] = v
            out = subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(path), name, docs or "0"], capture_output=True, text=True, env=env)
            print((envs + " " if envs else "") + (out.stdout.strip() or out.stderr[-400:]), flush=True)

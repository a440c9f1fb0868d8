"""Build experimental variants of the CUDA library (tuning runs only): NAME=DEF1,DEF2 ..."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from active_particle_jamming_b200 import _build
from concurrent.futures import ThreadPoolExecutor
def one(spec):
    name, _, defs = spec.partition("=")
    out = os.path.join(_build.PKG, "lib", "libapj_%s.so" % name)
    _build.build_library(force=True, defines=[d for d in defs.split(",") if d], out=out)
    return out
with ThreadPoolExecutor(4) as ex:
    for o in ex.map(one, sys.argv[1:]):
        print(o)

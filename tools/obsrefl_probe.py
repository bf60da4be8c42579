import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from solvcon_b200.apps.obsrefl import ObliqueShock, measure_zones
for ct in ("quad", "unstructured"):
    sh = ObliqueShock(); sh.build_constant(); sh.build_numerical(cell_type=ct)
    sh.svr.resident(True)
    for n in (1500, 1500, 3000):
        sh.march(n)
        print(ct, "after +", n, "cfl", sh.svr.reduce()[:2])
        for mg in (0.12, 0.2):
            print("  margin", mg)
            for g, w in zip(measure_zones(sh, mg), sh.zone_states()):
                print("   ", g[0], ["%.4f" % v for v in g[1:]], ["%.4f" % v for v in w])
print(sh.shock_path())

"""Diagnostic: solve_esd(kktsolver='qr') on the band test problem, device vs oracle, with the trace tails."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from smcp_b200 import solvers
import test_gpu_solver as T
fo, fd = T._factories()
P = T._make("band")
for nm, fac in (("oracle", fo), ("device", fd)):
    solvers.options["show_progress"] = False
    solvers.set_backend_factory(fac)
    sol = P.solve_esd(kktsolver="qr")
    tr = sol.get("trace", [])
    print(nm, sol["status"], sol["iterations"], sol["primal objective"], sol["dual objective"])
    for r in tr[-12:]:
        print("   ", {k: ("%.3e" % v if isinstance(v, float) else v) for k, v in r.items() if k in ("it", "pres", "dres", "gap", "step", "pcost")})

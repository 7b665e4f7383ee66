import sys, time
sys.path.insert(0, '.')
import numpy as np
from guacho_b200.config import *
from guacho_b200.solver import Block
from tests.util import *
from tests.oracle_lib import U
for strict in (True, False):
    p = ot_shipped(nxtot=128, nytot=128, nztot=2, zmax=2.0/128, MPI_NBX=1, strict_fp=strict)
    g = global_ic(p); o = oracle_from_ic(p, g)
    b = Block(p); b.set_state(g)
    dto,_ = o.get_timestep(1,10,0.0,1e300); dtg,_ = b.get_timestep(1,10,0.0,1e300)
    print("dt", dto, dtg, dto-dtg)
    o.tstep(dto); b.tstep(dto)
    ug = b.get_state(); uo = o.get_block(0, U)
    print("strict" if strict else "fast", rel_err_per_var(interior(ug), interior(uo)))
    d = np.abs(interior(ug)-interior(uo)); print(" max abs", d.max(), "nonzero frac", (d>0).mean())

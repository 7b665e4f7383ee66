"""The host-side time loop (guacho_b200.solver.Simulation = src/main.f90:94-125) on CPU: an adapter drives it with
the oracle in place of the GPU block (tests only), and the dump / clipping / termination logic is compared with the
loop written out as the reference has it."""
import numpy as np

from guacho_b200.config import ot_shipped
from guacho_b200.solver import Simulation
from tests.oracle_lib import U
from tests.util import global_ic, oracle_from_ic


class OracleAsBlock:
    """Duck-typed stand-in for guacho_b200.solver.Block (get_timestep / set_time / tstep / p)."""

    def __init__(self, oracle):
        self.o, self.p = oracle, oracle.p

    def set_state(self, u):
        self.o.scatter_u(u); self.o.start()

    def get_timestep(self, it, n_iter, time, tprint):
        return self.o.get_timestep(it, n_iter, time, tprint)

    def set_time(self, t):
        self.o.time = t

    def tstep(self, dt):
        assert self.o.tstep(dt) == 0


def test_simulation_loop_matches_main_f90():
    p = ot_shipped(nxtot=32, nytot=32, nztot=2, zmax=2.0 / 32, MPI_NBX=1, tmax=0.02, dtprint=0.005)
    g = global_ic(p, "ot")
    # the loop exactly as src/main.f90:94-125 has it
    o = oracle_from_ic(p, g, threads=2)
    time, tprint, itprint, it, dumps, dts = 0.0, p.dtprint, 1, 1, [], []
    while time <= p.tmax:
        dt, dump = o.get_timestep(it, 10, time, tprint)
        o.time = time
        assert o.tstep(dt) == 0
        time += dt
        dts.append(dt)
        if dump:
            dumps.append((itprint, time))
            tprint += p.dtprint
            itprint += 1
        it += 1
    ref_u = o.get_block(0, U)
    # the same through Simulation
    o2 = oracle_from_ic(p, g, threads=2)
    sim = Simulation(OracleAsBlock(o2))
    sim.itprint = 1
    seen = []
    sim.on_output = lambda s: seen.append((s.itprint, s.time))
    n = sim.run()
    assert n == it - 1 and sim.iteration == it and sim.time == time
    assert seen == dumps and len(dumps) == 4                       # outputs at t = 0.005, 0.01, 0.015, 0.02 ...
    assert all(abs(t - k * p.dtprint) < 1e-15 for k, (_i, t) in enumerate(dumps, start=1))   # ... hit exactly (dt clipped, hydro_core.f90:691-694)
    assert time > p.tmax                                            # one extra step after the last dump (`do while (time <= tmax)`)
    assert np.array_equal(o2.get_block(0, U), ref_u)
    assert dts[0] < dts[9] < dts[10] and abs(dts[9] / dts[8] - 2.0) < 0.05      # the 10-step CFL ramp doubles dt each iteration (:677-682)

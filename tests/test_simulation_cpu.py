"""The host-side time loop (guacho_b200.solver.Simulation = src/main.f90:94-125) on CPU: an adapter drives it with
the oracle in place of the GPU block (tests only), and the dump / clipping / termination logic is compared with the
loop written out as the reference has it."""
import os

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


def test_warm_start_from_a_bin_dump(tmp_path):
    """iwarm: a run restarted from dump k reads the reference-format file of this block and carries on with
    time = k*dtprint, tprint = time + dtprint and the CFL ramp restarted (src/init.f90:134-142, 436-471)."""
    from guacho_b200.bin_io import write_bin
    p = ot_shipped(nxtot=32, nytot=32, nztot=2, zmax=2.0 / 32, MPI_NBX=1, tmax=0.0125, dtprint=0.005)
    g = global_ic(p, "ot")
    o = oracle_from_ic(p, g, threads=2)
    blk = OracleAsBlock(o); blk.rank = 0
    sim = Simulation(blk)
    sim.itprint = 1
    sim.on_output = lambda s: write_bin(str(tmp_path) + "/", o.get_block(0, U), p, (0, 0, 0), 0, s.itprint)
    sim.run()
    assert sorted(os.listdir(tmp_path / "BIN")) == ["points000.001.bin", "points000.002.bin"]
    # restart from dump 2 (t = 0.01) and run to the same tmax
    o2 = oracle_from_ic(p, g, threads=2)
    blk2 = OracleAsBlock(o2); blk2.rank = 0
    sim2 = Simulation(blk2)
    sim2.warm_start(str(tmp_path) + "/", 2)
    assert sim2.time == 2 * p.dtprint and sim2.tprint == 3 * p.dtprint and sim2.itprint == 3 and sim2.iteration == 1
    u_restart = o2.get_block(0, U)
    from guacho_b200.bin_io import read_bin
    u_file, _ = read_bin(str(tmp_path / "BIN" / "points000.002.bin"))
    assert np.array_equal(u_restart[..., 2:-2, 2:-2, 2:-2], u_file[..., 2:-2, 2:-2, 2:-2])
    n = sim2.run()
    assert n >= 10 and sim2.time > p.tmax          # the restarted run ramps its time step again, then finishes
    assert np.isfinite(o2.get_block(0, U)).all()

"""Run configuration: the runtime mirror of Guacho's compile-time `parameters.f90`.

Every field of :class:`GxConfig` corresponds to a Fortran ``parameter`` the hydro/MHD
step reads (reference ``OT/parameters.f90:48-227``) plus the block decomposition that
replaces ``MPI_NBX/NBY/NBZ`` and ``mpi_cart_coords`` (``src/init.f90:103-110``).
The ctypes layout is exactly ``gx_config`` in ``include/guacho_gx.h``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, asdict

# ---- named constants, values identical to src/constants.f90:56-98 ----
SOLVER_HLL, SOLVER_HLLC, SOLVER_HLLE, SOLVER_HLLD = 1, 2, 3, 4
SOLVER_HLLE_SPLIT_ALL = 7            # src/constants.f90:62 (HLLE on fluctuations about a background state, src/hlle_split_all.f90)
EOS_ADIABATIC, EOS_SINGLE_SPECIE, EOS_H_RATE, EOS_CHEM = 1, 2, 3, 4
TC_OFF, TC_ISOTROPIC, TC_ANISOTROPIC = 0, 1, 2
BC_OUTFLOW, BC_CLOSED, BC_PERIODIC, BC_OTHER = 1, 2, 3, 4
COOL_NONE, COOL_H = 0, 1
LIMITER_NO_AVERAGE, LIMITER_NO_LIMIT, LIMITER_MINMOD, LIMITER_VAN_LEER = -1, 0, 1, 2
LIMITER_VAN_ALBADA, LIMITER_UMIST, LIMITER_WOODWARD, LIMITER_SUPERBEE = 3, 4, 5, 6

SOLVER_NAMES = {SOLVER_HLL: "HLL", SOLVER_HLLC: "HLLC", SOLVER_HLLE: "HLLE", SOLVER_HLLD: "HLLD", SOLVER_HLLE_SPLIT_ALL: "HLLE_SPLIT_ALL"}
ALL_LIMITERS = (LIMITER_NO_AVERAGE, LIMITER_NO_LIMIT, LIMITER_MINMOD, LIMITER_VAN_LEER,
                LIMITER_VAN_ALBADA, LIMITER_UMIST, LIMITER_WOODWARD, LIMITER_SUPERBEE)

NGHOST = 2  # parameters.f90:193


class GxConfig(C.Structure):
    """ctypes image of ``gx_config`` (include/guacho_gx.h)."""
    _fields_ = [
        ("struct_bytes", C.c_int32), ("device", C.c_int32),
        ("nxtot", C.c_int32), ("nytot", C.c_int32), ("nztot", C.c_int32),
        ("nbx", C.c_int32), ("nby", C.c_int32), ("nbz", C.c_int32),
        ("cx", C.c_int32), ("cy", C.c_int32), ("cz", C.c_int32),
        ("nghost", C.c_int32),
        ("neq", C.c_int32), ("neqdyn", C.c_int32), ("npas", C.c_int32),
        ("mhd", C.c_int32), ("pmhd", C.c_int32), ("passives", C.c_int32),
        ("riemann_solver", C.c_int32), ("slope_limiter", C.c_int32), ("eq_of_state", C.c_int32),
        ("enable_flux_cd", C.c_int32), ("eight_wave", C.c_int32), ("user_source_terms", C.c_int32),
        ("bc_left", C.c_int32), ("bc_right", C.c_int32), ("bc_bottom", C.c_int32),
        ("bc_top", C.c_int32), ("bc_out", C.c_int32), ("bc_in", C.c_int32),
        ("bc_user", C.c_int32), ("strict_fp", C.c_int32), ("cooling", C.c_int32),
        ("th_cond", C.c_int32), ("tc_saturation", C.c_int32), ("pad_", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
        ("cv", C.c_double), ("gamma", C.c_double), ("Tempsc", C.c_double),
        ("cfl", C.c_double), ("eta", C.c_double), ("tsc", C.c_double),
        ("rsc", C.c_double), ("rhosc", C.c_double), ("vsc2", C.c_double), ("bsc", C.c_double), ("mu", C.c_double),
    ]


@dataclass
class Params:
    """Python-side parameters (same names as parameters.f90)."""
    nxtot: int = 512
    nytot: int = 512
    nztot: int = 2
    MPI_NBX: int = 1
    MPI_NBY: int = 1
    MPI_NBZ: int = 1
    xmax: float = 1.0
    ymax: float = 1.0
    zmax: float = 2.0 / 512.0
    mhd: bool = True
    pmhd: bool = False
    npas: int = 0
    riemann_solver: int = SOLVER_HLLD
    slope_limiter: int = LIMITER_MINMOD
    eq_of_state: int = EOS_ADIABATIC
    enable_flux_cd: bool = True
    eight_wave: bool = False
    user_source_terms: bool = False
    bc_left: int = BC_PERIODIC
    bc_right: int = BC_PERIODIC
    bc_bottom: int = BC_PERIODIC
    bc_top: int = BC_PERIODIC
    bc_out: int = BC_PERIODIC
    bc_in: int = BC_PERIODIC
    bc_user: bool = False
    cv: float = 1.5
    Tempsc: float = 1.0
    cfl: float = 0.2
    eta: float = 0.0
    cooling: int = COOL_NONE
    tsc: float = 1.0
    th_cond: int = 0                 # TC_OFF | TC_ISOTROPIC | TC_ANISOTROPIC (src/constants.f90:96-98)
    tc_saturation: bool = False
    rsc: float = 1.0                 # scalings to cgs (parameters.f90:159-170); thermal conduction only
    rhosc: float = 1.0
    vsc2: float = 1.0
    bsc: float = 1.0
    mu: float = 1.0
    tmax: float = 0.5
    dtprint: float = 0.1
    strict_fp: bool = False
    device: int = -1
    extra: dict = field(default_factory=dict)

    # ---- derived parameters (parameters.f90:186-227) ----
    @property
    def bfield(self) -> bool:
        return bool(self.mhd or self.pmhd)

    @property
    def neqdyn(self) -> int:
        return 8 if self.bfield else 5

    @property
    def neq(self) -> int:
        return self.neqdyn + self.npas

    @property
    def passives(self) -> bool:
        return self.npas > 0

    @property
    def gamma(self) -> float:
        return (self.cv + 1.0) / self.cv            # parameters.f90:155

    @property
    def nx(self) -> int:
        return self.nxtot // self.MPI_NBX

    @property
    def ny(self) -> int:
        return self.nytot // self.MPI_NBY

    @property
    def nz(self) -> int:
        return self.nztot // self.MPI_NBZ

    @property
    def dx(self) -> float:
        return self.xmax / self.nxtot               # init.f90:120

    @property
    def dy(self) -> float:
        return self.ymax / self.nytot

    @property
    def dz(self) -> float:
        return self.zmax / self.nztot

    @property
    def nblocks(self) -> int:
        return self.MPI_NBX * self.MPI_NBY * self.MPI_NBZ

    def block_shape(self):
        """Fortran shape (neq, nx+4, ny+4, nz+4) of one block's arrays."""
        g = 2 * NGHOST
        return (self.neq, self.nx + g, self.ny + g, self.nz + g)

    def validate(self) -> None:
        if self.nxtot % self.MPI_NBX or self.nytot % self.MPI_NBY or self.nztot % self.MPI_NBZ:
            raise ValueError("grid is not divisible by the block decomposition")
        if min(self.nx, self.ny, self.nz) < NGHOST:
            raise ValueError("each block needs at least nghost=2 cells per direction")
        if self.riemann_solver in (SOLVER_HLLE, SOLVER_HLLD, SOLVER_HLLE_SPLIT_ALL) and not self.mhd:
            raise ValueError("HLLE/HLLD need mhd=True (they use cfastX)")
        if self.riemann_solver in (SOLVER_HLL, SOLVER_HLLC) and self.mhd:
            raise ValueError("HLL/HLLC use the hydro sound speed: run them with mhd=False")
        if self.cooling == COOL_H and self.npas < 1:
            raise ValueError("COOL_H evolves the neutral-H passive u(neqdyn+1): needs npas >= 1 (cooling_h.f90:30)")
        if self.enable_flux_cd and not self.mhd:
            raise ValueError("flux-CD without B field updates nothing (hydro_solver.f90:103-113)")

    def to_c(self, coords=(0, 0, 0)) -> GxConfig:
        c = GxConfig()
        c.struct_bytes = C.sizeof(GxConfig)
        c.device = self.device
        c.nxtot, c.nytot, c.nztot = self.nxtot, self.nytot, self.nztot
        c.nbx, c.nby, c.nbz = self.MPI_NBX, self.MPI_NBY, self.MPI_NBZ
        c.cx, c.cy, c.cz = coords
        c.nghost = NGHOST
        c.neq, c.neqdyn, c.npas = self.neq, self.neqdyn, self.npas
        c.mhd, c.pmhd, c.passives = int(self.mhd), int(self.pmhd), int(self.passives)
        c.riemann_solver, c.slope_limiter, c.eq_of_state = self.riemann_solver, self.slope_limiter, self.eq_of_state
        c.enable_flux_cd, c.eight_wave = int(self.enable_flux_cd), int(self.eight_wave)
        c.user_source_terms = int(self.user_source_terms)
        c.bc_left, c.bc_right, c.bc_bottom = self.bc_left, self.bc_right, self.bc_bottom
        c.bc_top, c.bc_out, c.bc_in = self.bc_top, self.bc_out, self.bc_in
        c.bc_user = int(self.bc_user)
        c.strict_fp = int(self.strict_fp)
        c.dx, c.dy, c.dz = self.dx, self.dy, self.dz
        c.cv, c.gamma, c.Tempsc = self.cv, self.gamma, self.Tempsc
        c.cfl, c.eta = self.cfl, self.eta
        c.cooling, c.tsc = self.cooling, self.tsc
        c.th_cond, c.tc_saturation = self.th_cond, int(self.tc_saturation)
        c.rsc, c.rhosc, c.vsc2, c.bsc, c.mu = self.rsc, self.rhosc, self.vsc2, self.bsc, self.mu
        return c

    def replace(self, **kw) -> "Params":
        d = asdict(self)
        d.update(kw)
        return Params(**d)


def ot_shipped(**kw) -> Params:
    """The Orszag-Tang run exactly as shipped (OT/parameters.f90): 512x512x2, HLLD +
    flux-CD + minmod, periodic, cfl 0.2, eta 0, 4x1x1 blocks."""
    p = Params(MPI_NBX=4)
    return p.replace(**kw) if kw else p


def ot_3d(n: int = 256, **kw) -> Params:
    """OT initial conditions extruded along z on an n^3 grid, zmax=1."""
    p = Params(nxtot=n, nytot=n, nztot=n, zmax=1.0)
    return p.replace(**kw) if kw else p

"""Worker of the multi-rank GPU parity test: launched with torch.distributed.run, one rank per GPU.
Each rank advances the 3D Taylor-Green problem on its slab and compares its local fields with the
single-process CPU oracle at the same global dofs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from oasisx_b200.comm import HostComm  # noqa: E402
from problems import TaylorGreen, TaylorGreenRot, make_cpu_port, make_mesh, make_oracle, make_solver, relerr, vscale  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
krylov = len(sys.argv) > 3 and sys.argv[3] in ("krylov", "mg", "bench")
bench_opts = len(sys.argv) > 3 and sys.argv[3] == "bench"  # the settings bench.py times (extrapolated guesses, multigrid)
use_mg = len(sys.argv) > 3 and sys.argv[3] == "mg"
comm = HostComm.from_env()
if len(sys.argv) > 3 and sys.argv[3] == "pbc":
    # open channel of test/test_tentative_velocity.py on several ranks: inlet callable + walls (two DirichletBCs per
    # component), outlet PressureBC: natural boundary term, pressure Dirichlet rows/columns across the slab interface
    from test_gpu_tentative import build

    kry = {"ksp_type": "bcgs", "pc_type": "jacobi", "ksp_rtol": 1e-12}
    cg = {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-12}
    s, o, inlet, _ = build(2, True, solver_options={"tentative": kry, "pressure": cg, "scalar": cg}, comm=comm,
                           device=int(os.environ.get("LOCAL_RANK", "0")))
    lp = s._lp
    dt, nu = 0.01, 0.5
    inlet.t = 0.0
    worst = 0.0
    for n in range(steps):
        inlet.t += dt
        s.solve(dt, nu, max_iter=2, max_error=1e-30)
        o.solve(dt, nu, max_iter=2, max_error=1e-30)
        for i in range(2):
            worst = max(worst, relerr(s._u[i].x.array_ro(), o.u[i][lp.V.l2g], vscale(o.u)))
        worst = max(worst, relerr(s._p.x.array_ro(), o.p[lp.Q.l2g]))
    assert max(np.abs(p).max() for p in o.p_surf) > 1e-3  # the natural pressure term is there and matters
    worst = comm.allreduce(worst, "max")
    print(f"rank {comm.rank}/{comm.size}: channel with PressureBC, max rel err {worst:.2e}, local facets {len(s._bcs_p[0]._facet_cells)}", flush=True)
    assert worst <= 1e-7, worst
    comm.Barrier()
    print("MR_OK", comm.rank, flush=True)
    sys.exit(0)
dt, nu = 0.005, 0.01
use_cpu_port = N >= 12  # the LU oracle stops being practical: compare with the CPU port (pinned against the oracle on CPU)
Field = TaylorGreenRot if use_cpu_port else TaylorGreen
tg = Field(nu, 3)
msh = make_mesh(3, N, comm)
opts = None
if krylov:
    opts = {k: {"ksp_type": t, "pc_type": "jacobi", "ksp_rtol": 1e-11}
            for k, t in (("tentative", "bcgs"), ("pressure", "cg"), ("scalar", "cg"))}
    if use_mg:
        opts["pressure"]["pc_type"] = "mg"
        opts["scalar"]["ksp_type"] = "chebyshev"  # reduction-free mass solves
if bench_opts:
    import copy

    import bench

    opts = copy.deepcopy(bench.KRYLOV)
    for o_ in opts.values():
        o_["ksp_rtol"] = 1e-11
s = make_solver(msh, 2, tg, dt, solver_options=opts, device=int(os.environ.get("LOCAL_RANK", "0")))
tg2 = Field(nu, 3)
if use_cpu_port:
    from oracle import ipcs_cpu as cpu

    class _PortView:  # the three attributes the comparison below reads, taken from the C++ port
        def __init__(self, c):
            self.c = c

        def solve(self, dt, nu, max_iter=1):
            self.c.solve(dt, nu)
            self.u = [self.c.get(cpu.U, i) for i in range(3)]
            self.u1 = [self.c.get(cpu.U1, i) for i in range(3)]
            self.p = self.c.get(cpu.P, 0)
            return None

    o = _PortView(make_cpu_port(make_mesh(3, N), 2, tg2, dt, rtol=1e-12))
else:
    o = make_oracle(make_mesh(3, N), 2, tg2, dt)
lp = s._lp
tg.t_u = tg2.t_u = 0.0
tg.t_p = tg2.t_p = -dt / 2
worst = 0.0
for n in range(steps):
    for t in (tg, tg2):
        t.t_u += dt
        t.t_p += dt
    d1 = s.solve(dt, nu, max_iter=1)
    d2 = o.solve(dt, nu, max_iter=1)
    for i in range(3):
        worst = max(worst, relerr(s._u[i].x.array_ro(), o.u[i][lp.V.l2g], vscale(o.u)))
        worst = max(worst, relerr(s._u1[i].x.array_ro(), o.u1[i][lp.V.l2g], vscale(o.u1)))
    worst = max(worst, relerr(s._p.x.array_ro(), o.p[lp.Q.l2g]))
    assert d2 is None or abs(d1 - d2) <= 1e-6 * d2, (d1, d2)
st = s.stats()
worst = comm.allreduce(worst, "max")
peer = s._ctx.peer_enabled()
print(f"rank {comm.rank}/{comm.size}: max rel err {worst:.2e} halos {st.halo_exchanges} nccl allreduces {st.allreduces} "
      f"peer kernels {st.peer_kernels} (peer path {'on' if peer else 'off'}) owned V {lp.V.n_owned} ghosts {lp.V.n_ghost} "
      f"its {list(st.its_tentative)}/{st.its_pressure}/{list(st.its_update)}", flush=True)
assert worst <= (1e-7 if krylov else 1e-8), worst
assert st.halo_exchanges > 0 and (st.peer_kernels > 0 if peer else st.allreduces > 0)
if os.environ.get("B200_PEER") == "0":
    assert not peer
if os.environ.get("B200_REQUIRE_PEER") == "1":
    assert peer, "peer-memory path expected on this box"
comm.Barrier()
print("MR_OK", comm.rank, flush=True)

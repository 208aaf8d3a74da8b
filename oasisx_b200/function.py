"""``Projector`` with the surface of ``/root/reference/src/oasisx/function.py:13-143``.

On the IPCS hot path the projector is used for exactly one thing: the rotational pressure update
``ps = Proj_Q(p + dp - xi nu div u)`` (``fracstep.py:237-247,593-602``).  That fixed expression is
assembled and solved on the device inside ``b2_pressure_solve``.  This class exposes the same
mass-matrix solve for user data: ``function`` is either a :class:`oasisx_b200.fem.Function` in the
pressure space or a callable returning nodal values there; the right-hand side ``(f, v)`` is then
``MQ f`` (exact for f in the space).  Arbitrary UFL expressions need a form compiler and are out of
scope.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import fem as _fem

__all__ = ["Projector"]


class Projector:
    def __init__(self, function, space: _fem.FunctionSpace, bcs=None, petsc_options: dict | None = None,
                 jit_options: dict | None = None, form_compiler_options: dict | None = None,
                 metadata: dict | None = None, solver=None):
        if bcs:
            raise NotImplementedError("Projector with Dirichlet BCs is not on the B200 hot path")
        if solver is None or not solver._rotational:
            raise NotImplementedError(
                "Projector needs the Q mass matrix of a FractionalStep_AB_CN built with rotational=True "
                "(pass solver=...)"
            )
        if space is not solver._Q:
            raise NotImplementedError("Projector is available on the pressure space only")
        self._function, self._space, self._solver = function, space, solver
        self._ctx: L.Context = solver._ctx
        for k, v in (petsc_options or {}).items():
            if k in ("ksp_type", "pc_type", "ksp_rtol", "ksp_atol", "ksp_max_it"):
                self._ctx.set_solver_option(L.SOLVER_PROJECTOR, k, v)
        self._x = _fem.Function(space)
        self._b = _fem.Function(space)
        self.assemble_rhs()

    def assemble_rhs(self):
        """``function.py:108-119``: b = (f, v) = MQ f_h."""
        f = self._function
        nodal = f.x.array_ro() if isinstance(f, _fem.Function) else np.asarray(
            f(self._space.tabulate_dof_coordinates().T), dtype=np.float64)
        n = self._space.num_dofs
        self._b.x.array[:] = self._ctx.mat_mult(L.MAT_MQ, 0, nodal, n)

    def solve(self, assemble_rhs: bool = True):
        """``function.py:121-133``; returns the KSP converged reason."""
        if assemble_rhs:
            self.assemble_rhs()
        x, reason = self._ctx.project_q(self._b.x.array_ro())
        self._x.x.array[:] = x
        return reason

    @property
    def x(self):
        return self._x

"""Host-side mirror of the reference's equation front-end and IR.

    #[equation]          prestige_macros/src/lib.rs:8-58     -> @equation
    SliceVisitor         prestige_macros/src/lib.rs:64-135   -> _SliceVisitor
    EquationIR           prestige/src/equations/ir.rs:3-14   -> EquationIR
    fuse / FusedEquations prestige/src/equations/fuse.rs:4-40 -> fuse / FusedEquations
    debug_equation       prestige/src/equations/debug.rs:3-22 -> debug_equation

Same names, same argument meaning, same (un)ordering guarantees: `reads` / `writes` are
sets turned into lists (order unspecified, lib.rs:26-27, fuse.rs:35-36); `bodies` keep
input order (fuse.rs:18,30).  Like the macro, the decorator DROPS the original function
and leaves an object whose only method is `ir()` (lib.rs:36-55), every written array is
also listed in reads (the visitor recurses into the left-hand side, lib.rs:84,119), and
only bare-name bases count (`self.x[i]` is not detected, lib.rs:73-75,103-105,125).

One extension, needed by any executing back-end: FusedEquations also carries `names`
(the reference's fuse() drops them; its only back-end just prints bodies).
"""
from __future__ import annotations

import ast
import inspect
import textwrap
from dataclasses import dataclass, field


@dataclass
class EquationIR:
    name: str
    reads: list
    writes: list
    body: str

    def clone(self) -> "EquationIR":          # #[derive(Clone)], ir.rs:3
        return EquationIR(self.name, list(self.reads), list(self.writes), self.body)


class _SliceVisitor(ast.NodeVisitor):
    def __init__(self):
        self.reads, self.writes = set(), set()

    @staticmethod
    def _base(node):
        return node.value.id if isinstance(node, ast.Subscript) and isinstance(node.value, ast.Name) else None

    def visit_Assign(self, node):             # force[i] = ...      (lib.rs:71-85)
        for t in node.targets:
            b = self._base(t)
            if b:
                self.writes.add(b)
        self.generic_visit(node)

    def visit_AugAssign(self, node):          # force[i] += ...     (lib.rs:88-120)
        b = self._base(node.target)
        if b:
            self.writes.add(b)
        self.generic_visit(node)

    def visit_Subscript(self, node):          # every indexed bare name is a read (lib.rs:123-133)
        b = self._base(node)
        if b:
            self.reads.add(b)
        self.generic_visit(node)


class Equation:
    """What `#[equation] fn name(...)` expands to: a unit struct with `ir()`."""

    def __init__(self, name, reads, writes, body):
        self.__name__ = name
        self._ir = EquationIR(name, list(reads), list(writes), body)

    def ir(self) -> EquationIR:
        return self._ir.clone()

    def __call__(self, *a, **k):
        raise TypeError(f"equation '{self.__name__}' is not callable: #[equation] drops the function and keeps only ir() "
                        "(prestige_macros/src/lib.rs:36-55)")


def equation(fn) -> Equation:
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fdef = next(n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef,)))
    v = _SliceVisitor()
    for stmt in fdef.body:
        v.visit(stmt)
    body = "{ " + " ".join(ast.unparse(s) + " ;" for s in fdef.body
                           if not (isinstance(s, ast.Expr) and isinstance(s.value, ast.Constant))) + " }"
    return Equation(fdef.name, v.reads, v.writes, body)


def declare(name: str, reads, writes, body: str) -> Equation:
    """Declare an equation whose body is implemented by a hand-written kernel of the same name."""
    return Equation(name, set(reads) | set(writes), set(writes), body)


@dataclass
class FusedEquations:
    reads: list
    writes: list
    bodies: list
    names: list = field(default_factory=list)   # extension, see module docstring


def fuse(eqs) -> FusedEquations:
    reads, writes, bodies, names = set(), set(), [], []
    for eq in eqs:
        reads.update(eq.reads)
        writes.update(eq.writes)
        bodies.append(str(eq.body))
        names.append(eq.name)
    return FusedEquations(list(reads), list(writes), bodies, names)


def debug_equation(eq: EquationIR) -> None:
    print("----------------------")
    print(f"Equation : {eq.name}")
    print("Reads:")
    for r in eq.reads:
        print(f"  {r}")
    print("Writes:")
    for w in eq.writes:
        print(f"  {w}")
    print("Body:")
    print(eq.body)
    print("----------------------")


# ---------------------------------------------------------------------------------------------
# The equations this back-end has hand-written kernels for.
# ---------------------------------------------------------------------------------------------
@equation
def eq1(i, j, force, mass):
    # the reference's sample, prestige/src/lib.rs:7-12
    force[i] += mass[j]


tait_eos = declare("tait_eos", ["rho"], ["p"], "{ p[i] = B * ((rho[i] / rho0).powf(gamma) - 1.0) ; }")
continuity = declare("continuity", ["x", "y", "z", "u", "v", "w", "m", "h"], ["arho"],
                     "{ arho[i] += m[j] * dot(v_ij, grad_w(x_ij, h[i])) ; }")
momentum = declare("momentum", ["x", "y", "z", "u", "v", "w", "m", "h", "rho", "p"], ["au", "av", "aw"],
                   "{ a[i] -= m[j] * (p[i]/rho[i]^2 + p[j]/rho[j]^2 + visc_ij) * grad_w(x_ij, h[i]) ; }")
dem_contact = declare("dem_contact", ["x", "y", "z", "u", "v", "w", "wx", "wy", "wz", "rad", "m", "hist_id", "hist_x", "hist_y", "hist_z", "hist_n"],
                      ["fx", "fy", "fz", "tx", "ty", "tz", "hist_id", "hist_x", "hist_y", "hist_z", "hist_n"],
                      "{ (F[i], T[i], xi[i][j]) += spring_dashpot(x_ij, v_ij, w, rad, xi[i][j]) ; }")
# gather over the FLUID neighbours of every dummy (non-fluid) particle: extrapolated pressure and the density the EOS maps
# to it.  Applied after tait_eos; the pair equations then read the extrapolated values.
wall_pressure = declare("wall_pressure", ["x", "y", "z", "h", "tag", "rho", "p"], ["p", "rho"],
                        "{ if tag[i] != 0 && tag[j] == 0 { num[i] += (p[j] + rho[j] * dot(g, x_ij)) * w(x_ij, h[i]) ; den[i] += w(x_ij, h[i]) ; } }")
# per-particle (no j): sums the force loop's results over the members of each rigid body
body_reduce = declare("body_reduce", ["x", "y", "z", "m", "body", "fx", "fy", "fz", "tx", "ty", "tz", "au", "av", "aw"],
                      ["body_force", "body_torque"],
                      "{ body_force[body[i]] += f_total(i) ; body_torque[body[i]] += cross(x[i] - cm[body[i]], f_total(i)) + t[i] ; }")

"""Back-ends, mirroring prestige::codegen (prestige/src/codegen/mod.rs:1).

  simple_cpu.generate_simple_cpu   the reference's only back-end, restated: returns the all-pairs loop
                                   as TEXT (prestige/src/codegen/simple_cpu.rs:3-22).  Never executed.
  b200.generate_b200 / b200.run    the sibling back-end this repo adds: maps the fused equation set to
                                   the hand-written sm_100a kernels behind pst_apply and runs them.
"""
from __future__ import annotations

from .equations import FusedEquations


class simple_cpu:
    @staticmethod
    def generate_simple_cpu(ir: FusedEquations) -> str:
        code = "for i in 0..n {\n"
        code += "    for j in 0..n {\n"
        for b in ir.bodies:
            code += "        " + b + "\n"
        code += "    }\n"
        code += "}\n"
        return code


class b200:
    KERNELS = ("eq1", "tait_eos", "wall_pressure", "continuity", "momentum", "dem_contact", "body_reduce")

    @staticmethod
    def generate_b200(ir: FusedEquations) -> list:
        """The launch plan for a fused set: the equation names pst_apply receives, in body order."""
        if not ir.names:
            raise ValueError("FusedEquations.names is empty: fuse() must keep equation names for an executing back-end")
        unknown = [n for n in ir.names if n not in b200.KERNELS]
        if unknown:
            raise ValueError(f"no hand-written kernel for equation(s) {unknown}; known: {b200.KERNELS}")
        return list(ir.names)

    @staticmethod
    def run(ctx, ir: FusedEquations) -> None:
        """Execute the fused set on the context's device-resident arrays (one fused pair kernel)."""
        ctx.apply(b200.generate_b200(ir))


generate_simple_cpu = simple_cpu.generate_simple_cpu

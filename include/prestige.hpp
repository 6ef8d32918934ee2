// prestige.hpp -- C++ host-side mirror of the reference's equation API, above the C ABI.
//
// The reference is Rust and its toolchain is absent from this image, so the host side
// above include/prestige_b200.h is restated here in C++17 with the reference's names, argument meaning
// and (un)ordering guarantees:
//
//   prestige::equations::ir::EquationIR          prestige/src/equations/ir.rs:3-14
//   prestige::equations::fuse::{FusedEquations, fuse}   prestige/src/equations/fuse.rs:4-40
//   prestige::equations::debug::debug_equation   prestige/src/equations/debug.rs:3-22
//   prestige::codegen::simple_cpu::generate_simple_cpu  prestige/src/codegen/simple_cpu.rs:3-22
//   prestige::codegen::b200::{generate_b200, run}       NEW sibling back-end (prestige/src/codegen/mod.rs:1)
//   prestige::eq1                                prestige/src/lib.rs:7-12 (what #[equation] expands it to,
//                                                prestige_macros/src/lib.rs:36-55: a unit struct with ir())
//
// `reads` / `writes` are sets turned into vectors: their order is unspecified (lib.rs:26-27, fuse.rs:35-36);
// `bodies` keep input order (fuse.rs:18,30).  One extension: FusedEquations also carries `names`, which an
// executing back-end needs and the reference's fuse() drops.
#pragma once

#include <cstdio>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "prestige_b200.h"

namespace prestige {

namespace equations {
namespace ir {
struct EquationIR {
    std::string name;
    std::vector<std::string> reads;
    std::vector<std::string> writes;
    std::string body;   // the reference keeps a proc_macro2::TokenStream; its Display form is what fuse() stores
};
}  // namespace ir

namespace fuse {
struct FusedEquations {
    std::vector<std::string> reads;
    std::vector<std::string> writes;
    std::vector<std::string> bodies;
    std::vector<std::string> names;   // extension (see header comment)
};

inline FusedEquations fuse(const std::vector<ir::EquationIR>& eqs) {
    std::set<std::string> reads, writes;
    FusedEquations f;
    for (const auto& eq : eqs) {
        reads.insert(eq.reads.begin(), eq.reads.end());
        writes.insert(eq.writes.begin(), eq.writes.end());
        f.bodies.push_back(eq.body);
        f.names.push_back(eq.name);
    }
    f.reads.assign(reads.begin(), reads.end());
    f.writes.assign(writes.begin(), writes.end());
    return f;
}
}  // namespace fuse

namespace debug {
inline void debug_equation(const ir::EquationIR& eq) {
    std::printf("----------------------\nEquation : %s\nReads:\n", eq.name.c_str());
    for (const auto& r : eq.reads) std::printf("  %s\n", r.c_str());
    std::printf("Writes:\n");
    for (const auto& w : eq.writes) std::printf("  %s\n", w.c_str());
    std::printf("Body:\n%s\n----------------------\n", eq.body.c_str());
}
}  // namespace debug
}  // namespace equations

// What `#[equation] fn eq1(i, j, force: &mut [f64], mass: &[f64]) { force[i] += mass[j]; }` expands to.
struct eq1 {
    static equations::ir::EquationIR ir() { return {"eq1", {"force", "mass"}, {"force"}, "{ force [i] += mass [j] ; }"}; }
};
// The equations this back-end has hand-written kernels for (names = pst_apply names).
struct tait_eos {
    static equations::ir::EquationIR ir() { return {"tait_eos", {"rho", "p"}, {"p"}, "{ p [i] = B * ((rho [i] / rho0) . powf (gamma) - 1.0) ; }"}; }
};
struct continuity {
    static equations::ir::EquationIR ir() {
        return {"continuity", {"x", "y", "z", "u", "v", "w", "m", "h", "arho"}, {"arho"}, "{ arho [i] += m [j] * dot (v_ij , grad_w (x_ij , h [i])) ; }"};
    }
};
struct momentum {
    static equations::ir::EquationIR ir() {
        return {"momentum", {"x", "y", "z", "u", "v", "w", "m", "h", "rho", "p", "au", "av", "aw"}, {"au", "av", "aw"},
                "{ a [i] -= m [j] * (p [i] / rho [i] ^ 2 + p [j] / rho [j] ^ 2 + visc_ij) * grad_w (x_ij , h [i]) ; }"};
    }
};
struct dem_contact {
    static equations::ir::EquationIR ir() {
        return {"dem_contact", {"x", "y", "z", "u", "v", "w", "wx", "wy", "wz", "rad", "m", "hist_n", "hist_id", "hist_x", "hist_y", "hist_z"},
                {"fx", "fy", "fz", "tx", "ty", "tz", "hist_n", "hist_id", "hist_x", "hist_y", "hist_z"},
                "{ (F [i] , T [i] , xi [i] [j]) += spring_dashpot (x_ij , v_ij , w , rad , xi [i] [j]) ; }"};
    }
};
// gather over the FLUID neighbours of every dummy (non-fluid) particle: extrapolated pressure + matching density
struct wall_pressure {
    static equations::ir::EquationIR ir() {
        return {"wall_pressure", {"x", "y", "z", "h", "tag", "rho", "p"}, {"p", "rho"},
                "{ if tag [i] != 0 && tag [j] == 0 { num [i] += (p [j] + rho [j] * dot (g , x_ij)) * w (x_ij , h [i]) ; den [i] += w (x_ij , h [i]) ; } }"};
    }
};
// per-particle (no j): sums the force loop's results over the members of each rigid body
struct body_reduce {
    static equations::ir::EquationIR ir() {
        return {"body_reduce", {"x", "y", "z", "m", "body", "fx", "fy", "fz", "tx", "ty", "tz", "au", "av", "aw"}, {"body_force", "body_torque"},
                "{ body_force [body [i]] += f_total (i) ; body_torque [body [i]] += cross (x [i] - cm [body [i]] , f_total (i)) + t [i] ; }"};
    }
};

namespace codegen {
namespace simple_cpu {
inline std::string generate_simple_cpu(const equations::fuse::FusedEquations& ir) {
    std::string code;
    code += "for i in 0..n {\n";
    code += "    for j in 0..n {\n";
    for (const auto& b : ir.bodies) {
        code += "        ";
        code += b;
        code += "\n";
    }
    code += "    }\n";
    code += "}\n";
    return code;
}
}  // namespace simple_cpu

namespace b200 {
// The launch plan for a fused set: the names pst_apply receives, in body order.
inline std::vector<std::string> generate_b200(const equations::fuse::FusedEquations& ir) {
    static const std::set<std::string> kernels = {"eq1", "tait_eos", "wall_pressure", "continuity", "momentum", "dem_contact", "body_reduce"};
    if (ir.names.empty()) throw std::invalid_argument("FusedEquations.names is empty");
    for (const auto& n : ir.names)
        if (!kernels.count(n)) throw std::invalid_argument("no hand-written kernel for equation '" + n + "'");
    return ir.names;
}
// Execute the fused set on a context's device-resident arrays.
inline pst_status run(pst_ctx* ctx, const equations::fuse::FusedEquations& ir) {
    const std::vector<std::string> names = generate_b200(ir);
    std::vector<const char*> c;
    for (const auto& n : names) c.push_back(n.c_str());
    return pst_apply(ctx, c.data(), (int)c.size());
}
}  // namespace b200
}  // namespace codegen

}  // namespace prestige

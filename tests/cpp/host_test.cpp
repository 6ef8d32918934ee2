// host_test.cpp -- the reference's own two tests (prestige/src/lib.rs:20-27, :32-52) restated against the C++
// host mirror (include/prestige.hpp), WITH the assertions the reference lacks; plus, with --gpu, eq1 executed
// through the C ABI and compared bit for bit with the loop generate_simple_cpu emits (simple_cpu.rs:7-16).
//   g++ -std=c++17 -Iinclude tests/cpp/host_test.cpp -Lprestige_b200 -lprestige_b200 -Wl,-rpath,$PWD/prestige_b200
#include <cassert>
#include <cstdio>
#include <cstring>
#include <set>
#include <vector>

#include "prestige.hpp"

using namespace prestige;
using prestige::equations::fuse::fuse;

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); return 1; } } while (0)

static int test_equation_ir() {
    auto ir = eq1::ir();
    equations::debug::debug_equation(ir);
    CHECK(ir.name == "eq1");
    CHECK((std::set<std::string>(ir.writes.begin(), ir.writes.end()) == std::set<std::string>{"force"}));
    CHECK((std::set<std::string>(ir.reads.begin(), ir.reads.end()) == std::set<std::string>{"force", "mass"}));
    return 0;
}

static int test_fusion() {
    std::vector<equations::ir::EquationIR> eqs = {eq1::ir()};
    auto fused = fuse(eqs);
    CHECK(fused.bodies.size() == 1 && fused.names[0] == "eq1");
    const std::string code = codegen::simple_cpu::generate_simple_cpu(fused);
    std::printf("Generated CPU Code:\n%s", code.c_str());
    CHECK(code == "for i in 0..n {\n    for j in 0..n {\n        { force [i] += mass [j] ; }\n    }\n}\n");
    auto plan = codegen::b200::generate_b200(fuse({tait_eos::ir(), continuity::ir(), momentum::ir()}));
    CHECK(plan.size() == 3 && plan[1] == "continuity");
    auto walled = fuse({tait_eos::ir(), wall_pressure::ir(), continuity::ir(), momentum::ir()});
    CHECK(codegen::b200::generate_b200(walled).size() == 4 && walled.names[1] == "wall_pressure");
    bool writes_rho = false;
    for (const auto& w : walled.writes) writes_rho = writes_rho || w == "rho";      // the equation slaves the dummy density
    CHECK(writes_rho);
    bool threw = false;
    try { codegen::b200::generate_b200(fuse({{"nope", {}, {}, "{}"}})); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    return 0;
}

static int test_abi_without_gpu_is_loud() {
    pst_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.dim = 3; cfg.real = PST_F64; cfg.capacity = 8; cfg.cell_size = 0.5; cfg.hi[0] = cfg.hi[1] = cfg.hi[2] = 1.0;
    pst_ctx* ctx = nullptr;
    cfg.struct_size = 4;                      // ABI guard
    CHECK(pst_create(&cfg, &ctx) == PST_EINVAL && ctx == nullptr);
    cfg.struct_size = sizeof cfg; cfg.dim = 5;
    CHECK(pst_create(&cfg, &ctx) == PST_EINVAL);
    std::printf("pst_version: %s | last error: %s\n", pst_version(), pst_last_error(nullptr));
    return 0;
}

static int test_eq1_on_gpu() {
    const int n = 1537;
    std::vector<double> mass(n), force(n), ref(n);
    for (int i = 0; i < n; ++i) { mass[i] = 0.5 + 1e-3 * ((i * 7919) % 1000); force[i] = ref[i] = 1e-2 * (i % 13); }
    for (int i = 0; i < n; ++i)               // the reference loop, literally
        for (int j = 0; j < n; ++j) ref[i] += mass[j];
    pst_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.dim = 3; cfg.real = PST_F64; cfg.capacity = n; cfg.cell_size = 0.5; cfg.hi[0] = cfg.hi[1] = cfg.hi[2] = 1.0;
    pst_ctx* ctx = nullptr;
    if (pst_create(&cfg, &ctx) != PST_OK) { std::fprintf(stderr, "pst_create: %s\n", pst_last_error(nullptr)); return 1; }
    CHECK(pst_set_count(ctx, n) == PST_OK);
    CHECK(pst_array_create(ctx, "force", PST_REAL, PST_ARRAY_PERSISTENT) == PST_OK);
    CHECK(pst_array_create(ctx, "mass", PST_REAL, PST_ARRAY_PERSISTENT) == PST_OK);
    CHECK(pst_upload(ctx, "mass", mass.data(), n) == PST_OK);
    CHECK(pst_upload(ctx, "force", force.data(), n) == PST_OK);
    CHECK(codegen::b200::run(ctx, fuse({eq1::ir()})) == PST_OK);
    CHECK(pst_download(ctx, "force", force.data(), n) == PST_OK);
    for (int i = 0; i < n; ++i) CHECK(force[i] == ref[i]);
    double launches = 0;
    CHECK(pst_get_stat(ctx, "launches", &launches) == PST_OK && launches >= 1);
    pst_destroy(ctx);
    std::printf("eq1 on GPU: bit-exact over %d particles\n", n);
    return 0;
}

int main(int argc, char** argv) {
    int rc = test_equation_ir() | test_fusion() | test_abi_without_gpu_is_loud();
    if (argc > 1 && std::strcmp(argv[1], "--gpu") == 0) rc |= test_eq1_on_gpu();
    std::printf(rc ? "FAILED\n" : "OK\n");
    return rc;
}

// Compile-and-run check of include/sgufp_b200.hpp and sgufp_b200_explorer.hpp against the reference's own headers and
// Network/DD objects (tests/test_cpp_adapter.py; built where /root/reference exists, into oracle/_ref/ so that the binary
// travels to the GPU box).
//   adapter_check <instance.txt>                 no device: structure + model through the C++ adapter, compute must fail
//   adapter_check <instance.txt> gpu <device> [device_count] [width]
//        with a device: solveSubProblem cuts printed for the Python side to compare with the ctypes path, the SAME cuts applied
//        to the reference's RelaxedDDNew (Oracle A, linked in here) and to the adapter's — bounds, paths and cut-sets must be
//        identical —, then the Benders loop of sgufp_b200_explorer.hpp (one node at a time, and `width` nodes side by side).
#include "DD.h"
#include "sgufp_b200.hpp"
#include "sgufp_b200_explorer.hpp"
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static int structure(const std::shared_ptr<Network> &net, sgufp::GuroSolver &solver) {
    int L = 0, T = 0, nv = 0;
    sgufp_dims(solver.handle(), &L, &T, &nv);
    if ((unsigned)L != net->totalLayers) { std::printf("L mismatch %d %u\n", L, net->totalLayers); return 1; }
    std::vector<int32_t> order(L);
    sgufp_processing_order(solver.handle(), order.data());
    for (int l = 0; l < L; l++) if (order[l] != net->processingOrder[l].second) { std::printf("processingOrder mismatch at %d\n", l); return 1; }
    // reference diagram vs adapter diagram: same layers, same cut-set
    Inavap::RelaxedDDNew ref(net.get());
    Inavap::Node root;
    ref.buildTree(root);
    sgufp::RelaxedDDNewT<Inavap::Node> ours(solver);
    ours.buildTree(root);
    if (ours.isTreeExact() != ref.isTreeExact()) { std::printf("exactness mismatch\n"); return 1; }
    if (ours.getSolution() != ref.getSolution()) { std::printf("initial solution mismatch\n"); return 1; }
    if (!ref.isTreeExact()) {
        auto a = ref.getCutset(5.0);
        auto b = ours.getCutset(5.0);
        if (a.size() != b.size()) { std::printf("cutset size mismatch\n"); return 1; }
        for (size_t i = 0; i < a.size(); i++)
            if (a[i].states != b[i].states || a[i].solutionVector != b[i].solutionVector || a[i].globalLayer != b[i].globalLayer || a[i].ub != b[i].ub) { std::printf("cutset node %zu mismatch\n", i); return 1; }
    }
    std::printf("adapter ok L=%d T=%d\n", L, T);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    auto net = std::make_shared<Network>(std::string(argv[1]));
    if (argc < 4 || std::strcmp(argv[2], "gpu") != 0) {
        sgufp::GuroSolver solver(net, SGUFP_DEVICE_NONE);
        if (int rc = structure(net, solver)) return rc;
        // compute must fail loudly without a device
        sgufp::RelaxedDDNewT<Inavap::Node> ours(solver);
        Inavap::Node root;
        ours.buildTree(root);
        try { solver.solveSubProblem(ours.getSolution()); std::printf("compute did not fail\n"); return 1; } catch (const std::runtime_error &e) {}
        return 0;
    }
    const int device = std::atoi(argv[3]), device_count = argc > 4 ? std::atoi(argv[4]) : 1, width = argc > 5 ? std::atoi(argv[5]) : 4;
    using Explorer = sgufp::NodeExplorerT<Inavap::Node, Inavap::Container, Inavap::cut_node_t>;
    Explorer ex(net, device, width, device_count);
    sgufp::GuroSolver &solver = ex.solver();
    if (int rc = structure(net, solver)) return rc;
    // 1. the Benders inner loop on the root by hand: every cut goes to the reference's diagram (Oracle A) and to ours
    Inavap::RelaxedDDNew ref(net.get());
    sgufp::RelaxedDDNewT<Inavap::Node> ours(solver);
    Inavap::Node root;
    ref.buildTree(root); ours.buildTree(root);
    if (ref.isTreeExact()) {
        double ub_ref = 1e300, ub_our = 1e300;
        for (int it = 0; it < 12; it++) {
            const auto pr = ref.getSolution(), po = ours.getSolution();
            if (pr != po) { std::printf("iteration %d: argmax paths differ\n", it); return 1; }
            auto res = solver.solveSubProblem(po);
            const Inavap::Cut &cut = res.second;
            std::printf("CUT %d %d %.17g %" PRIu64 " %zu", it, (int)res.first, cut.RHS, (uint64_t)cut.hash_val, cut.coeff.size());
            for (auto v : po) std::printf(" %d", (int)v);
            std::printf(" |");
            for (const auto &kv : cut.coeff) std::printf(" %" PRIu64 ":%.17g", (uint64_t)kv.first, kv.second);
            std::printf("\n");
            if (res.first == FEASIBILITY) {
                const auto a = ref.applyFeasibilityCut(cut), b = ours.applyFeasibilityCut(cut);
                if (a != b) { std::printf("iteration %d: feasibility flags differ\n", it); return 1; }
                if (!a) break;
            } else {
                ub_ref = ref.applyOptimalityCut(cut, -1e300, ub_ref);
                ub_our = ours.applyOptimalityCut(cut, -1e300, ub_our);
                if (ub_ref != ub_our) { std::printf("iteration %d: bounds differ %.17g %.17g\n", it, ub_ref, ub_our); return 1; }   // bit-exact
                std::printf("BOUND %d %.17g\n", it, ub_our);
            }
        }
    }
    // 2. the explorer: one node at a time, then `width` nodes side by side: same optimum
    long n1 = 0, nw = 0;
    const double opt1 = sgufp::solve_frontier(ex, 1, 400, &n1);
    const long cuts1 = ex.cuts_generated, calls1 = ex.k1_calls;
    Explorer exw(net, device, width, device_count);
    const double optw = sgufp::solve_frontier(exw, width, 400, &nw);
    std::printf("EXPLORER width 1: optimum %.17g nodes %ld cuts %ld k1_calls %ld\n", opt1, n1, cuts1, calls1);
    std::printf("EXPLORER width %d: optimum %.17g nodes %ld cuts %ld k1_calls %ld\n", width, optw, nw, exw.cuts_generated, exw.k1_calls);
    if (n1 < 400 && nw < 400 && opt1 != optw) { std::printf("optimum depends on the width\n"); return 1; }
    std::printf("adapter gpu ok\n");
    return 0;
}

// Compile-and-run check of include/sgufp_b200.hpp against the reference's own headers and
// Network/DD objects (tests/test_cpp_adapter.py; only where /root/reference exists).
// Uses SGUFP_DEVICE_NONE: structure + model through the C++ adapter, no compute.
#include "DD.h"
#include "sgufp_b200.hpp"
#include <cstdio>

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    auto net = std::make_shared<Network>(std::string(argv[1]));
    sgufp::GuroSolver solver(net, SGUFP_DEVICE_NONE);
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
    // compute must fail loudly without a device
    try { solver.solveSubProblem(ours.getSolution()); std::printf("compute did not fail\n"); return 1; } catch (const std::runtime_error &e) {}
    std::printf("adapter ok L=%d T=%d\n", L, T);
    return 0;
}

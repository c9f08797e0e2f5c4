// sgufp_b200.hpp — header-only C++ adapters with the reference's own signatures.
//
// Include this AFTER the reference's "Network.h" and "Cut.h" (it uses `Network`, `CutType`,
// `Inavap::Cut`, `Inavap::Node`-like data exactly as declared there) and link libsgufp_b200.so.
//
//   sgufp::GuroSolver    replaces  ::GuroSolver                 (/root/reference/grb.h:17-104)
//   sgufp::RelaxedDDNew  replaces  Inavap::RelaxedDDNew          (/root/reference/DD.h:734-810)
//   sgufp::RestrictedDDNew         Inavap::RestrictedDDNew       (/root/reference/DD.h:653-730)
//
// so that NodeExplorer (/root/reference/NodeExplorer.h:115-116, NodeExplorer.cpp:915-986) compiles
// against them unchanged apart from the two member types.  See INTEGRATION.md.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "sgufp_b200.h"
#include "sgufp_b200_dd.h"

namespace sgufp {

// Error convention.  The reference has none: its Release build is -fno-exceptions (CMakeLists.txt:27), where a GRBException
// terminates the process.  With exceptions enabled an SGUFP_ERR_* code becomes a std::runtime_error; without them
// (-fno-exceptions: __cpp_exceptions is not defined) the message goes to stderr and the process aborts, as the reference's would.
[[noreturn]] inline void raise(const std::string &msg) {
#if defined(__cpp_exceptions) && !defined(SGUFP_B200_NO_EXCEPTIONS)
    throw std::runtime_error(msg);
#else
    std::fprintf(stderr, "%s\n", msg.c_str());
    std::abort();
#endif
}
inline void check(int rc, const sgufp_ctx *ctx) {
    if (rc < 0) raise(std::string("sgufp_b200: ") + sgufp_last_error(ctx));
}

// One handle per host thread, like one GuroSolver per NodeExplorer (NodeExplorer.h:115).
class GuroSolver {
public:
    // `GuroSolver(const shared_ptr<Network>&, const GRBEnv&)` (grb.h:36) minus the Gurobi environment.
    // device_count > 1: ONE host thread's solver spreads the scenarios over devices device .. device+device_count-1 and the
    // all-reduce of the partial cuts runs inside the library (sgufp_create_sharded); solveSubProblem is unchanged.
    explicit GuroSolver(const std::shared_ptr<Network> &net, int device = 0, int device_count = 1) : net_(net) {
        const int n = (int)net->n, m = (int)net->edges, S = (int)net->nScenarios;
        std::vector<int32_t> tail(m), head(m), r0(m), up((size_t)m * S), lo((size_t)m * S), vbar;
        for (int a = 0; a < m; a++) {
            const NetworkArc &arc = net->networkArcs[a];
            tail[a] = (int32_t)arc.tailId; head[a] = (int32_t)arc.headId; r0[a] = arc.rewards[0];
            for (int s = 0; s < S; s++) { up[(size_t)a * S + s] = arc.upperCapacities[s]; lo[(size_t)a * S + s] = arc.lowerCapacities[s]; }
        }
        // the handle applies shuffleVBarNodes itself; it is idempotent on an already shuffled list
        for (auto v : net->Vbar) vbar.push_back((int32_t)v);
        int rc;
        if (device_count > 1) {
            std::vector<int> devs(device_count);
            for (int i = 0; i < device_count; i++) devs[i] = device + i;
            rc = sgufp_create_sharded(&ctx_, n, m, S, tail.data(), head.data(), up.data(), lo.data(), r0.data(), vbar.data(), (int)vbar.size(), devs.data(), device_count);
        } else
            rc = sgufp_create(&ctx_, n, m, S, tail.data(), head.data(), up.data(), lo.data(), r0.data(), vbar.data(), (int)vbar.size(), device, 0, S);
        if (rc < 0) raise(std::string("sgufp_create: ") + sgufp_last_error(nullptr));
        sgufp_dims(ctx_, &L_, &T_, nullptr);
    }
    // another solver on the SAME device-resident capacities (one per worker thread, NodeExplorer.h:115 / DDSolver.cpp:675)
    struct CloneTag {};
    GuroSolver(const GuroSolver &other, CloneTag) : net_(other.net_), L_(other.L_), T_(other.T_) {
        if (sgufp_clone(other.ctx_, &ctx_) < 0) raise(std::string("sgufp_clone: ") + sgufp_cache_last_error());
    }
    ~GuroSolver() { sgufp_destroy(ctx_); }
    GuroSolver(const GuroSolver &) = delete;
    GuroSolver &operator=(const GuroSolver &) = delete;

    // grb.h:75
    std::pair<CutType, Inavap::Cut> solveSubProblem(const std::vector<int16_t> &path) {
        std::vector<uint64_t> keys(T_ > 0 ? T_ : 1);
        std::vector<double> vals(T_ > 0 ? T_ : 1);
        int type = 0, nnz = 0;
        double rhs = 0;
        check(sgufp_solve_path(ctx_, path.data(), (int)path.size(), &type, &rhs, keys.data(), vals.data(), &nnz, nullptr, nullptr, nullptr, nullptr), ctx_);
        std::vector<std::pair<uint64_t, double>> coeff(nnz);
        for (int k = 0; k < nnz; k++) coeff[k] = {keys[k], vals[k]};
        return std::make_pair(type == SGUFP_CUT_FEASIBILITY ? FEASIBILITY : OPTIMALITY, Inavap::Cut{rhs, coeff});   // Inavap::Cut has explicit copy/move ctors (Cut.h:258-259)
    }

    // K candidates in one launch
    std::vector<std::pair<CutType, Inavap::Cut>> solveSubProblems(const std::vector<std::vector<int16_t>> &paths) {
        const int K = (int)paths.size(), Tn = T_ > 0 ? T_ : 1;
        std::vector<int16_t> flat((size_t)K * L_, -1);
        for (int k = 0; k < K; k++) {
            if ((int)paths[k].size() > L_) raise("sgufp_b200: path " + std::to_string(k) + " has more than totalLayers entries");
            std::copy(paths[k].begin(), paths[k].end(), flat.begin() + (size_t)k * L_);
        }
        std::vector<uint64_t> keys((size_t)K * Tn);
        std::vector<double> vals((size_t)K * Tn), rhs(K);
        std::vector<int> type(K), nnz(K);
        check(sgufp_solve_paths(ctx_, flat.data(), K, L_, type.data(), rhs.data(), keys.data(), vals.data(), nnz.data(), nullptr, nullptr, nullptr, nullptr), ctx_);
        std::vector<std::pair<CutType, Inavap::Cut>> out;
        for (int k = 0; k < K; k++) {
            std::vector<std::pair<uint64_t, double>> coeff(nnz[k]);
            for (int t = 0; t < nnz[k]; t++) coeff[t] = {keys[(size_t)k * Tn + t], vals[(size_t)k * Tn + t]};
            out.emplace_back(type[k] == SGUFP_CUT_FEASIBILITY ? FEASIBILITY : OPTIMALITY, Inavap::Cut{rhs[k], coeff});
        }
        return out;
    }
    sgufp_ctx *handle() const { return ctx_; }

private:
    std::shared_ptr<Network> net_;
    sgufp_ctx *ctx_ = nullptr;
    int L_ = 0, T_ = 0;
};

namespace detail {
struct FlatCut {
    std::vector<uint64_t> keys;
    std::vector<double> vals;
    explicit FlatCut(const Inavap::Cut &c) : keys(c.coeff.size()), vals(c.coeff.size()) {
        for (size_t i = 0; i < c.coeff.size(); i++) { keys[i] = c.coeff[i].first; vals[i] = c.coeff[i].second; }
    }
};
template <class NodeT>
std::vector<NodeT> unpack_nodes(const std::vector<int32_t> &w, int k, double ub) {
    std::vector<NodeT> out;
    for (int i = 0; i < k;) {
        const int gl = w[i], ns = w[i + 1];
        std::vector<int16_t> st(w.begin() + i + 2, w.begin() + i + 2 + ns);
        i += 2 + ns;
        const int nl = w[i];
        std::vector<int16_t> sol(w.begin() + i + 1, w.begin() + i + 1 + nl);
        i += 1 + nl;
        out.emplace_back(st, sol, std::numeric_limits<double>::lowest(), ub, (uint16_t)gl);
    }
    return out;
}
}  // namespace detail

// NodeT is the reference's Inavap::Node (DD.h:456-479): (states, solutionVector, lb, ub, globalLayer).
template <class NodeT>
class RelaxedDDNewT {
public:
    explicit RelaxedDDNewT(GuroSolver &solver) : ctx_(solver.handle()) { check(sgufp_dd_create(ctx_, SGUFP_DD_RELAXED, 0, &dd_), ctx_); }
    ~RelaxedDDNewT() { sgufp_dd_destroy(dd_); }
    void buildTree(const NodeT &root) {   // DD.h:798
        check(sgufp_dd_build(dd_, root.states.data(), (int)root.states.size(), root.solutionVector.data(), (int)root.solutionVector.size(), root.globalLayer, nullptr), ctx_);
    }
    std::vector<int16_t> getSolution() const {   // DD.h:800
        std::vector<int16_t> p(1 << 15);
        const int k = sgufp_dd_solution(dd_, p.data(), (int)p.size());
        check(k, ctx_);
        p.resize(k);
        return p;
    }
    bool isTreeExact() const noexcept { return sgufp_dd_is_exact(dd_) == 1; }   // DD.h:802
    uint8_t applyFeasibilityCut(const Inavap::Cut &cut) {                        // DD.h:804
        detail::FlatCut f(cut);
        int ok = 0;
        check(sgufp_dd_apply_feasibility(dd_, cut.RHS, f.keys.data(), f.vals.data(), (int)f.keys.size(), &ok), ctx_);
        return (uint8_t)ok;
    }
    double applyOptimalityCut(const Inavap::Cut &cut, double optimal, double upperbound) {   // DD.h:805
        detail::FlatCut f(cut);
        double b = 0;
        check(sgufp_dd_apply_optimality(dd_, cut.RHS, f.keys.data(), f.vals.data(), (int)f.keys.size(), optimal, upperbound, &b), ctx_);
        return b;
    }
    // The loops of NodeExplorer::process over the GLOBAL cuts (NodeExplorer.cpp:935-944, 975-983) as one device call each
    // (sgufp_dd_apply_sequence): same results as the one-by-one calls, stops where the caller's loop returns.
    // Returns false if some cut was infeasible (the loop `return PRUNED_BY_FEASIBILITY_CUT`).
    bool applyFeasibilityCuts(const std::vector<const Inavap::Cut *> &cuts) {
        if (cuts.empty()) return true;
        Packed p(cuts);
        std::vector<int> ok(cuts.size(), 1);
        int applied = 0;
        check(sgufp_dd_apply_sequence(dd_, 1, p.rhs.data(), p.keys.data(), p.vals.data(), p.ptr.data(), (int)cuts.size(), 0.0, nullptr, ok.data(), &applied), ctx_);
        return applied == 0 || ok[applied - 1] != 0;
    }
    // Returns how many cuts were applied; `bound`: the upper bound the loop holds when it ends — on an exact diagram the last
    // call's value (:941), on a non-exact one the minimum over the calls (:981).
    int applyOptimalityCuts(const std::vector<const Inavap::Cut *> &cuts, double optimal, bool exact, double &bound) {
        if (cuts.empty()) return 0;
        Packed p(cuts);
        std::vector<double> b(cuts.size(), 0.0);
        int applied = 0;
        check(sgufp_dd_apply_sequence(dd_, 0, p.rhs.data(), p.keys.data(), p.vals.data(), p.ptr.data(), (int)cuts.size(), optimal, b.data(), nullptr, &applied), ctx_);
        if (applied > 0) bound = exact ? b[applied - 1] : *std::min_element(b.begin(), b.begin() + applied);
        return applied;
    }
    std::vector<NodeT> getCutset(double ub) {   // DD.h:807
        std::vector<int32_t> w(1 << 22);
        const int k = sgufp_dd_cutset(dd_, ub, w.data(), (int)w.size());
        check(k, ctx_);
        return detail::unpack_nodes<NodeT>(w, k, ub);
    }

private:
    struct Packed {   // a run of cuts flattened for the C ABI
        std::vector<double> rhs, vals;
        std::vector<uint64_t> keys;
        std::vector<int32_t> ptr;
        explicit Packed(const std::vector<const Inavap::Cut *> &cuts) : ptr(cuts.size() + 1, 0) {
            for (size_t i = 0; i < cuts.size(); i++) {
                rhs.push_back(cuts[i]->RHS);
                for (const auto &kv : cuts[i]->coeff) { keys.push_back(kv.first); vals.push_back(kv.second); }
                ptr[i + 1] = (int32_t)keys.size();
            }
            if (keys.empty()) { keys.push_back(0); vals.push_back(0.0); }
        }
    };
    sgufp_ctx *ctx_;
    sgufp_dd *dd_ = nullptr;
};

}  // namespace sgufp

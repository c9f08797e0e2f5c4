"""The caller of the hot path: the Benders loop on one branch-and-bound node.

`NodeExplorer.process` mirrors `Inavap::NodeExplorer::process`
(`/root/reference/NodeExplorer.cpp:915-986`) line for line on top of the
B200-backed `GuroSolver` (K1) and `RelaxedDDNew` (K2): build the relaxed diagram of the node, apply
the global cuts, then — on an exact diagram — alternate argmax path -> scenario cuts -> cut
application until a path repeats.  `Container` mirrors the reference's cut list (`Cut.h:448-485`,
new cuts are pushed to the FRONT).  `solve` is a minimal sequential driver over the cut-set nodes:
the reference's parallel `DDSolver` (queues, threads) is out of scope.
"""
from __future__ import annotations

from typing import List, Optional

from .dd import DOUBLE_MAX, DOUBLE_MIN, Node, RelaxedDDNew
from .solver import FEASIBILITY, GuroSolver

SUCCESS, PRUNED_BY_FEASIBILITY_CUT, PRUNED_BY_OPTIMALITY_CUT = 0, 1, 2   # OutObj::STATUS_OP


class Container:
    """`Inavap::Container` (Cut.h:448-485): singly linked, push-front; iteration starts at the newest cut."""

    def __init__(self):
        self._cuts: List = []

    def add(self, cut) -> None:
        self._cuts.insert(0, cut)

    def __iter__(self):
        return iter(list(self._cuts))

    def __len__(self):
        return len(self._cuts)


class OutObject:
    """`Inavap::OutObject` (NodeExplorer.h:86-103)."""

    def __init__(self, lb, ub, nodes, status):
        self.lb, self.ub, self.nodes, self.status = lb, ub, nodes, status


class NodeExplorer:
    def __init__(self, solver: GuroSolver, batch_global_cuts: bool = False, device_sequences: bool = True):
        # device_sequences: every loop over the GLOBAL cuts (NodeExplorer.cpp:935-944,
        # 975-983) is one device call (`sgufp_dd_apply_sequence`) that returns what the one-by-one calls
        # return, stops where the loop returns, and works on non-exact diagrams and feasibility cuts too.
        # False replays the reference's loop call by call.
        self.device_sequences = device_sequences
        # batch_global_cuts: on an EXACT diagram the loop over the global optimality
        # cuts (NodeExplorer.cpp:940-944) has no side effect but the terminal minima, so all of them are
        # applied in ONE K2 launch pair; the bound and the prune decision are the sequential ones.
        self.batch_global_cuts = batch_global_cuts
        self.solver = solver                      # NodeExplorer.h:115
        self.relaxedDD = RelaxedDDNew(solver)     # NodeExplorer.h:116 (one diagram, rebuilt per node)
        self.cuts_generated = 0

    def process(self, node: Node, optimalLB: float, globalFeasCuts: Container, globalOptCuts: Container) -> OutObject:
        upperBound = node.ub
        dd = self.relaxedDD
        dd.buildTree(node)                                                          # :922
        feas = list(globalFeasCuts)
        opt = list(globalOptCuts)
        if dd.isTreeExact():                                                        # :931
            if self.device_sequences:
                flags, n = dd.applyFeasibilityCuts(feas)                            # :935-938 in one call
                if n and not flags[n - 1]:
                    return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_FEASIBILITY_CUT)
                feas_done = True
            else:
                feas_done = False
            for cut in ([] if feas_done else feas):                                 # :935-938
                if not dd.applyFeasibilityCut(cut):
                    return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_FEASIBILITY_CUT)
            if self.device_sequences and opt:
                bounds, n = dd.applyOptimalityCuts(opt, optimalLB)                  # :940-944 in one call
                upperBound = float(bounds[n - 1])
                if upperBound <= optimalLB:
                    return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_OPTIMALITY_CUT)
                opt = []
            if self.batch_global_cuts and len(opt) > 1:
                from .dd import apply_optimality_batch
                upperBound = float(apply_optimality_batch([dd], opt)[0])
                if upperBound <= optimalLB:
                    return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_OPTIMALITY_CUT)
                opt = []
            for cut in opt:                                                         # :940-944
                upperBound = dd.applyOptimalityCut(cut, optimalLB, upperBound)
                if upperBound <= optimalLB:
                    return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_OPTIMALITY_CUT)
            allSolutions = []
            while True:                                                             # :949-971
                path = dd.getSolution().tolist()
                if path in allSolutions:
                    return OutObject(upperBound, upperBound, [], SUCCESS)           # :951-953
                allSolutions.append(path)
                cutType, cut = self.solver.solveSubProblem(path)                    # :957  (K1)
                self.cuts_generated += 1
                if cutType == FEASIBILITY:
                    globalFeasCuts.add(cut)
                    if not dd.applyFeasibilityCut(cut):                             # (K2)
                        return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_FEASIBILITY_CUT)
                else:
                    globalOptCuts.add(cut)
                    upperBound = dd.applyOptimalityCut(cut, optimalLB, upperBound)  # (K2)
                    if upperBound <= optimalLB:
                        return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_OPTIMALITY_CUT)
        if self.device_sequences:
            flags, n = dd.applyFeasibilityCuts(feas)                                # :975-978 in one call
            if n and not flags[n - 1]:
                return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_FEASIBILITY_CUT)
            bounds, n = dd.applyOptimalityCuts(opt, optimalLB)                      # :980-983 in one call
            if n:
                upperBound = min(upperBound, float(bounds.min()))
                if upperBound <= optimalLB:
                    return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_OPTIMALITY_CUT)
            feas, opt = [], []
        for cut in feas:                                                            # :975-978
            if not dd.applyFeasibilityCut(cut):
                return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_FEASIBILITY_CUT)
        for cut in opt:                                                             # :980-983
            upperBound = min(dd.applyOptimalityCut(cut, optimalLB, upperBound), upperBound)
            if upperBound <= optimalLB:
                return OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_OPTIMALITY_CUT)
        return OutObject(DOUBLE_MIN, upperBound, dd.getCutset(upperBound), SUCCESS)  # :985


class FrontierExplorer:
    """W explorers in lock step: `process_many` takes up to W branch-and-bound nodes, one diagram each, and
    every round hands the current argmax path of every node still in its cut loop (NodeExplorer.cpp:949-971) to ONE
    `solve_paths` call — K candidates per K1 launch instead of 1.  Per node nothing changes: it reads the global containers when
    it starts (:928-929), applies ITS OWN new cut to its own diagram, takes its own prune / repeat decisions.  This is what the
    reference's N_WORKERS explorers do on N nodes at the same time (DDSolver.cpp:701-712), in lock step.  The C++ form is
    include/sgufp_b200_explorer.hpp."""

    def __init__(self, solver: GuroSolver, width: int):
        self.solver = solver
        self.dds = [RelaxedDDNew(solver) for _ in range(max(1, width))]
        self.cuts_generated = 0
        self.k1_calls = 0

    def process_many(self, nodes, optimalLB: float, globalFeasCuts: Container, globalOptCuts: Container) -> List[OutObject]:
        import numpy as np
        n = len(nodes)
        assert n <= len(self.dds)
        out: List[Optional[OutObject]] = [None] * n
        ub = [nd.ub for nd in nodes]
        looping = [False] * n
        seen = [[] for _ in range(n)]
        feas, opt = list(globalFeasCuts), list(globalOptCuts)
        for i, node in enumerate(nodes):
            dd = self.dds[i]
            dd.buildTree(node)                                                      # :922
            exact = dd.isTreeExact()                                                # :931
            flags, k = dd.applyFeasibilityCuts(feas)                                # :935-938 / :975-978
            if k and not flags[k - 1]:
                out[i] = OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_FEASIBILITY_CUT)
                continue
            bounds, k = dd.applyOptimalityCuts(opt, optimalLB)                      # :940-944 / :980-983
            if k:
                ub[i] = float(bounds[k - 1]) if exact else min(ub[i], float(bounds[:k].min()))
                if ub[i] <= optimalLB:
                    out[i] = OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_OPTIMALITY_CUT)
                    continue
            if exact:
                looping[i] = True
            else:
                out[i] = OutObject(DOUBLE_MIN, ub[i], dd.getCutset(ub[i]), SUCCESS)  # :985
        while True:
            who, paths = [], []
            for i in range(n):
                if not looping[i]:
                    continue
                path = self.dds[i].getSolution().tolist()                           # :950
                if path in seen[i]:                                                 # :951-953
                    out[i] = OutObject(ub[i], ub[i], [], SUCCESS)
                    looping[i] = False
                    continue
                seen[i].append(path)
                who.append(i)
                paths.append(path)
            if not who:
                break
            L = max(len(p) for p in paths)
            batch = np.full((len(paths), L), -1, dtype=np.int16)
            for k, p in enumerate(paths):
                batch[k, :len(p)] = p
            res = self.solver.solve_paths(batch, want_obj=False, want_status=False, want_dense=False)   # :957 for every node at once (K1)
            self.k1_calls += 1
            self.cuts_generated += len(who)
            for k, i in enumerate(who):
                cut = res.cut(k)
                if int(res.cut_type[k]) == FEASIBILITY:
                    globalFeasCuts.add(cut)
                    if not self.dds[i].applyFeasibilityCut(cut):                    # :961-962
                        out[i] = OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_FEASIBILITY_CUT)
                        looping[i] = False
                else:
                    globalOptCuts.add(cut)
                    ub[i] = self.dds[i].applyOptimalityCut(cut, optimalLB, ub[i])   # :966
                    if ub[i] <= optimalLB:                                          # :967-968
                        out[i] = OutObject(DOUBLE_MIN, DOUBLE_MIN, [], PRUNED_BY_OPTIMALITY_CUT)
                        looping[i] = False
        return out


def solve_frontier(solver: GuroSolver, width: int = 8, known_lb: float = DOUBLE_MIN, max_nodes: int = 100000, max_cuts: Optional[int] = None):
    """`solve` taking up to `width` nodes of the stack per round (FrontierExplorer.process_many).
    Returns (optimum, nodes processed, cuts generated, K1 calls).  `max_cuts` ends the search early (profiling)."""
    ex = FrontierExplorer(solver, width)
    feas, opt = Container(), Container()
    best = known_lb
    stack = [Node(ub=DOUBLE_MAX)]
    processed = 0
    while stack and processed < max_nodes and (max_cuts is None or ex.cuts_generated < max_cuts):
        batch = []
        while stack and len(batch) < width:
            node = stack.pop()
            if node.ub <= best:                                 # DDSolver.cpp:707-711
                continue
            batch.append(node)
        if not batch:
            break
        outs = ex.process_many(batch, best, feas, opt)          # DDSolver.cpp:712
        processed += len(outs)
        for o in outs:
            if o.status != SUCCESS:
                continue
            if o.lb > best:                                     # DDSolver.cpp:723-731
                best = o.lb
            for child in o.nodes:                               # DDSolver.cpp:744-748
                if child.ub > best:
                    stack.append(child)
    return best, processed, ex.cuts_generated, ex.k1_calls


def solve(solver: GuroSolver, known_lb: float = DOUBLE_MIN, max_nodes: int = 100000, batch_global_cuts: bool = False,
          device_sequences: bool = True):
    """Sequential depth-first branch and bound over cut-set nodes (stand-in for DDSolver.cpp:658-776:
    pop a node, prune on ub <= incumbent, process, raise the incumbent, push the children).
    Returns (optimum, nodes processed, cuts generated)."""
    explorer = NodeExplorer(solver, batch_global_cuts, device_sequences)
    feas, opt = Container(), Container()
    best = known_lb
    root = Node(ub=DOUBLE_MAX)
    stack = [root]
    processed = 0
    while stack and processed < max_nodes:
        node = stack.pop()
        if node.ub <= best:                                   # DDSolver.cpp:707-711
            continue
        out = explorer.process(node, best, feas, opt)          # DDSolver.cpp:712
        processed += 1
        if out.status != SUCCESS:
            continue
        if out.lb > best:                                      # DDSolver.cpp:723-731
            best = out.lb
        for child in out.nodes:                                # DDSolver.cpp:744-748
            if child.ub > best:
                stack.append(child)
    return best, processed, explorer.cuts_generated

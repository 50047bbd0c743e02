/* dmpc_test.h — TEST HOOKS of libdmpc.so: host-side scheduling without a device.
 *
 * Not part of the drop-in boundary (include/dmpc.h): nothing here is bound by the R shim.  The CPU test-suite uses these
 * entry points to run the engine's own scheduler (dmphyclus_b200/csrc/host_tree.hpp) on a box without a GPU and to
 * execute the plans it hands out with numpy (tests/plan_interpreter.py) against the oracle.
 */
#ifndef DMPC_TEST_H
#define DMPC_TEST_H

#include "dmpc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Host-only self-check of the plan builder (needs no GPU; used by the CPU test-suite): builds the topology, applies the
 * cluster flags, plans a full evaluation with the given fusion threshold and verifies the invariants the kernels rely
 * on — every internal vertex computed exactly once (as a stored task or a fused token), children before parents across
 * levels, fused clades no larger than the threshold, stack depth within FUSE_DEPTH, tip lists and token ranges
 * consistent; the trailing chain handed to the walk kernel is one task per level, each forwarding into the next, with
 * the fused clades hanging off it materialised in level 0.  dirtyVertex = 0 checks the plan of a full evaluation;
 * dirtyVertex = a 1-based internal vertex id checks the plan of the proposal that follows a full evaluation and
 * invalidates that vertex and its ancestors (a split / merge root path); -1: the plan after new between-cluster
 * matrices (the between part of the tree); -2, -3, ...: the plan after one NNI move (two root paths that meet), a
 * different pick for each value.
 * out[0..7] = stored vertices, fused vertices, levels, tokens, max stack depth, max tips of a task's programs, chain
 * levels, materialised clades.  Returns DMPC_OK or DMPC_ERR_STATE with the violated invariant in dmpc_last_error(). */
int dmpc_plan_selfcheck(const int32_t* edge, int nEdgeRows, int numTips, const double* clusterMRCAs, int nClusterMRCAs,
                        int fuseTips, int dirtyVertex, int* out);

/* Host-side scheduling WITHOUT a device, for the CPU test-suite (tests/plan_interpreter.py executes the plans it hands
 * out with numpy under the kernels' semantics and compares with the oracle).  No arithmetic happens here: a dmpc_hostplan
 * is the same topology / proposal state machine a dmpc_tree owns (host_tree.hpp), driven by the same steps the entry
 * points take before they evaluate.
 *   dmpc_hp_propose kind: 0 = InvalidateAll (new within-cluster matrices, recompute); 1 = new between-cluster matrices;
 *     2 = split, args = {mrca}; 3 = merge, args = the sibling mrcas; 4 = within-cluster NNI, args = {mrca, numMoves};
 *     5 = between-cluster NNI, args = {numMoves, clusterMRCAs...}.  All ids 1-based.
 *   dmpc_hp_plan: the pending evaluation as tasks (8 int32 each, struct Task), tokens (4 int32 each), tip list and level
 *     starts; counts = {tasks, tokens, tips, levels, chainLevels, materialised}.  Like an evaluation, it marks the planned
 *     vertices solved and flips their slots.  Fails with DMPC_ERR_ARG when a capacity is too small.
 *   dmpc_hp_path: one proposal of a dmpc_eval_proposals batch (kind = DMPC_PROPOSAL_SPLIT / _MERGE, a, b as in
 *     dmpc_proposal) as its run of tasks: the stored vertices of the path to the root, bottom-up, each forwarding into the
 *     next; nothing is mutated.  counts = {tasks, tokens, tips}.
 *   dmpc_hp_restore: the host part of dmpc_restore_previous_config.
 *   dmpc_hp_state: cur (slot of every vertex), within flags, and the edge matrix in BuildEdgeMatrix order. */
typedef struct dmpc_hostplan dmpc_hostplan;
int dmpc_hp_create(const int32_t* edge, int nEdgeRows, int numTips, const double* clusterMRCAs, int nClusterMRCAs,
                   int fuseTips, dmpc_hostplan** out);
int dmpc_hp_propose(dmpc_hostplan* h, int kind, const int32_t* args, int nArgs);
int dmpc_hp_plan(dmpc_hostplan* h, int32_t* tasks, int taskCap, int32_t* tokens, int tokenCap, int32_t* tips, int tipCap,
                 int32_t* levelStart, int levelCap, int32_t* counts);
int dmpc_hp_path(dmpc_hostplan* h, int kind, int a, int b, int32_t* tasks, int taskCap, int32_t* tokens, int tokenCap,
                 int32_t* tips, int tipCap, int32_t* counts);
int dmpc_hp_restore(dmpc_hostplan* h, const int32_t* edge, int nEdgeRows, int nniMove, const int32_t* clusterMRCAs,
                    int nClusterMRCAs, int splitMergeMove);
int dmpc_hp_state(dmpc_hostplan* h, uint8_t* cur, uint8_t* within, int32_t* edgeOut);
int dmpc_hp_destroy(dmpc_hostplan* h);

#ifdef __cplusplus
}
#endif
#endif /* DMPC_TEST_H */

// task.hpp — work descriptors shared by the host scheduler (host_tree.hpp) and the pruning kernels (kernels.cuh).
#pragma once
#include <stdint.h>

namespace dmpc {

// One STORED vertex of one level.  An operand (child) is a tip (its 4-bit mask row), a stored internal vertex (its
// partial in HBM), or a FUSED small clade: a vertex with at most `fuseTips` tips below it whose partial is never
// written to HBM but recomputed from the tip masks by whoever needs it, through a short token program.
struct Task {
  int32_t out;            // internal index (vertex id - N) to write
  int32_t c0, c1;         // children in the reference's order: tip index, internal index, or root of a fused clade
  int32_t flags;
  int32_t prog;           // first token of this task's fused-operand programs (child 0's, then child 1's)
  int32_t ntok;           // tokens of child 0's program | tokens of child 1's program << 16
  int32_t tipOff, nTips;  // the tips those programs read: slice of the tip list; tokens name them by local index
};
enum : int32_t {
  TF_OUTSLOT = 1, TF_C0TIP = 2, TF_C1TIP = 4, TF_C0SLOT = 8, TF_C1SLOT = 16,
  TF_WITHIN = 32,         // matrix set of this vertex = within-flag of its FIRST child (AugTree.cpp:125)
  TF_C0FUSED = 64, TF_C1FUSED = 128,
  TF_C0FWD = 256, TF_C1FWD = 512,  // proposal paths and walked chains: the operand is the output of the previous task
                                   // of the run, forwarded in registers
  TF_HEAD = 1024                   // first task of a walked chain: child 0 is read from memory, then plays the forwarded part
};
// A plan's trailing levels that hold ONE task each form a chain (every task consumes the previous one's output): a root
// path, or the spine of a caterpillar.  With at least CHAIN_MIN of them the chain is handed to k_walk, which keeps the
// running partial in registers and prefetches the off-chain operands many steps ahead, instead of paying a full
// load -> compute -> store -> load round trip per level.  Fused clades hanging off the chain are MATERIALISED first:
// evaluated once, in parallel, as ordinary stored tasks into their own (otherwise unused) rows.
constexpr int CHAIN_MIN = 4;

// One step of a fused-clade program.  The evaluator is a stack machine per thread (= per site): REG holds the
// partial of the vertex computed last, the stack lives in shared memory.
struct Token {
  int32_t op;             // opcode | set << 3 | slot << 4 | firstFlag << 5
  int32_t a, b;           // tips, as indices into the task's slice of the tip list
  int32_t vertex;         // internal index of the vertex this token computes
};
enum : int32_t {
  OP_CHERRY = 0,          // REG = L(tip a) % L(tip b)                       a = first child
  OP_REGTIP = 1,          // REG = (P REG) % L(tip b)                        firstFlag: the tip is the first child
  OP_SPILL = 2,           // push REG
  OP_SLOTREG = 3          // REG = (P pop) % (P REG)                         firstFlag: the popped one is the first child
};
constexpr int TOK_SET = 8, TOK_SLOT = 16, TOK_FIRST = 32;
constexpr int FUSE_DEPTH = 3;      // stack slots: 1 for the task level + 2 inside a clade of <= 15 tips
constexpr int FUSE_MAX_TIPS = 15;

}  // namespace dmpc

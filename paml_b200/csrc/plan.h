// plan.h -- host-side planner: turns a PAML tree (nodes[].nson/sons) into the straight-line "program"
// the pruning kernel walks, and assigns on-chip slots to the conditional-likelihood tiles.
#pragma once
#include <vector>
#include <string>

namespace pb200 {

// program: [nops, nslots, ntip, first_op_offset, nmat, matseq_offset] [tip rows in the order the walk consumes
// them] [matseq: node ids of the P(t) matrices in the order the walk consumes them (every non-partial child)] ops...
// op header: {nchild, out_slot, scale_idx, node, flags, ntip}; child record: {kind, idx, node};
// child order inside an op: the CH_DIRECT child (if any), then the ntip tips, then slot children.
enum { PROG_HDR = 6, OP_HDR = 6, CH_REC = 3 };
// CH_DIRECT: the internal child whose op ran immediately before this one; its vectors are still in the warp's
// registers and feed the DMMA A operand through shuffles (no shared-memory round trip).  It is always child 0.
enum { CH_TIP = 0, CH_INTERNAL = 1, CH_PARTIAL = 2, CH_DIRECT = 3 };
enum { OPF_PARTIAL_OUT = 1, OPF_HAS_DIRECT = 2 };
// out_slot: >= 0 store to that slot (parked), OUT_ROOT root reduction, OUT_REG keep in registers for the next op
enum { OUT_ROOT = -1, OUT_REG = -2 };

struct Plan {
    std::vector<int> prog;      // see layout above
    int nops = 0, nslots = 0, nscale = 0;
    long long n_internal_children = 0, n_tip_children = 0;   // per tree walk (for flop / byte accounting)
};

// ---- macro-op program of the 4-state register kernel (prune4v2.cuh) ----
// The 4-state kernel walks a flatter program: one 2-int macro-op per (almost every) internal node, so that the
// interpretive overhead per FP64 instruction stays below one.  A thread carries ONE running vector `acc`
// (4 doubles per pattern); parked partial products live in shared-memory slots.
//   M4_CHERRY   a,b   acc  = PT_a[code] * PT_b[code]                (node whose first two children are tips)
//   M4_MV_TIP   a,b   acc  = (P_a . acc) * PT_b[code]               (one internal child, then a tip)
//   M4_MV_PUSH  a,s   slot[s] = P_a . acc                           (first of several internal children: park)
//   M4_MV_POP   a,s   acc  = (P_a . acc) * slot[s]                  (next internal child joins the parked product)
//   M4_PUSH     s     slot[s] = acc                                 (park again: more internal children follow)
//   M4_TIP_MUL  a     acc *= PT_a[code]                             (further tip children)
//   M4_ROOT           root reduction with pi
// word 0 = opcode [0:3] | M4_SCALE_FLAG (NodeScale the finished node) | tip field 0 << 5 | tip field 1 << 11 |
// slot << 17; word 1 = a | b << 16.  Tips are consumed in the order of `tip_order`; the kernel reads their codes
// from a walk-ordered packing, 4 codes per 32-bit word, and everything about WHERE a tip's code sits is static, so
// it travels in the op: tip field = byte in the word [0:1] | advance [2:3] (0 stay in the current word, 1 move to
// the next word and prefetch the word four ahead of it, 2 move only) | word index mod 4 [4:5].  Field 0 belongs to the
// first tip an op consumes, field 1 to the second (M4_CHERRY).
enum { M4_CHERRY = 0, M4_MV_TIP = 1, M4_MV_PUSH = 2, M4_MV_POP = 3, M4_PUSH = 4, M4_TIP_MUL = 5, M4_ROOT = 6 };
enum { M4_SCALE_FLAG = 16, M4_TIPF0_SHIFT = 5, M4_TIPF1_SHIFT = 11, M4_SLOT_SHIFT = 17 };
struct Plan4 {
    std::vector<int> ops;        // 2 ints per macro-op
    std::vector<int> tip_order;  // tip node ids in consumption order
    int nslots = 0, nscale = 0;
};
std::string plan_tree4(int nnode, int ns, int root, const int* nson, const int* sons, int maxsons,
                       const unsigned char* scale_flags, Plan4& out);

// ---- programs of the P-stationary kernel (prune_ps.cuh) ----
// The kernel is warp-specialised: GATHER warps turn the tip children of a node into one vector per pattern (the product of the tips'
// P(t) rows) in a shared-memory "G slot"; CONTRACTION warps do the internal children on the tensor path and keep their results in
// "D slots".  Both walk the tree in the same post-order (Sethi-Ullman child order), each through its own op list:
//   prog = [n_gops, n_steps, nG, nD, off_gops, off_steps, nscale, 0] [W_s: writes per chunk of G slot s] gops... steps...
//   gop  = {gslot, k, ntip, node} + ntip tip node ids            -- the k-th write of the chunk into G slot gslot
//   step = one internal child of a node, 8 ints: {kind (PS_D / PS_G / PS_NONE) | PS_STEP_FIRST | PS_STEP_LAST, slot, child node
//          (selects the P(t) tile), k | W_slot << 16 (G slots: which write of the chunk), out (D slot of the node, -1 = root), scale_idx,
//          tp_slot (G slot with the node's tip product, -1 = none), tp_k | W_tp << 16}; the last four describe the node and are used by its
//          LAST step.  Fixed-size records: the kernel copies the table to shared memory and never chases a program counter.
// A node with tips only is a gop alone: its G slot IS the node's vector (kind PS_G in the parent's step).  A root with tips only
// gets one PS_NONE step.  The contraction warps hand a G slot back as soon as they have read it; the planner re-uses a G slot only
// after the step that read it, in program order, so the full / empty handshake of the kernel cannot deadlock.  The output D slot of
// a node never aliases one of its children.
enum { PS_HDR = 8, PS_GOP_HDR = 4, PS_STEP_INTS = 8 };
enum { PS_D = 0, PS_G = 1, PS_NONE = 2, PS_KIND_MASK = 3, PS_STEP_FIRST = 4, PS_STEP_LAST = 8 };
struct PlanPS {
    std::vector<int> prog;
    int n_gops = 0, n_dops = 0 /* steps */, nG = 0, nD = 0, nscale = 0, ntip = 0;
    long long n_internal_children = 0;
};
std::string plan_tree_ps(int nnode, int ns, int root, const int* nson, const int* sons, int maxsons,
                         const unsigned char* scale_flags, PlanPS& out);

// Returns "" on success, else an error message.
std::string plan_tree(int nnode, int ns, int root, const int* nson, const int* sons, int maxsons,
                      const unsigned char* scale_flags, Plan& out);

}  // namespace pb200

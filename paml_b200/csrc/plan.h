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

// Returns "" on success, else an error message.
std::string plan_tree(int nnode, int ns, int root, const int* nson, const int* sons, int maxsons,
                      const unsigned char* scale_flags, Plan& out);

}  // namespace pb200

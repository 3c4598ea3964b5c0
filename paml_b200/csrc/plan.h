// plan.h -- host-side planner: turns a PAML tree (nodes[].nson/sons) into the straight-line "program"
// the pruning kernel walks, and assigns on-chip slots to the conditional-likelihood tiles.
#pragma once
#include <vector>
#include <string>

namespace pb200 {

// program: [nops, nslots, ntip, first_op_offset] [tip rows in the order the walk consumes them] ops...
// op header: {nchild, out_slot, scale_idx, node, flags, ntip}; child record: {kind, idx, node};
// the ntip tip children of an op come first in its child list.
enum { PROG_HDR = 4, OP_HDR = 6, CH_REC = 3 };
enum { CH_TIP = 0, CH_INTERNAL = 1, CH_PARTIAL = 2 };
enum { OPF_PARTIAL_OUT = 1 };

struct Plan {
    std::vector<int> prog;      // see layout above
    int nops = 0, nslots = 0, nscale = 0;
    long long n_internal_children = 0, n_tip_children = 0;   // per tree walk (for flop / byte accounting)
};

// Returns "" on success, else an error message.
std::string plan_tree(int nnode, int ns, int root, const int* nson, const int* sons, int maxsons,
                      const unsigned char* scale_flags, Plan& out);

}  // namespace pb200

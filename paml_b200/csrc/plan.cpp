// plan.cpp -- see plan.h.
//
// The reference walks the tree recursively and keeps every internal node's conditional vectors in
// com.conP (src/codeml.c:3526-3582).  Here a warp owns a few site patterns and carries them through the
// WHOLE tree without touching HBM, so the planner's job is to order the post-order walk such that as few
// conditional-vector tiles as possible are alive at once (they live in shared memory):
//   * children are visited in decreasing order of their own slot need (Sethi-Ullman ordering);
//   * a node with more than two internal children is folded: after two of them the running product is
//     parked in a slot as a "partial product" and the remaining children are multiplied in one by one,
//     so a multifurcation (PAML's unrooted trees have a trifurcating root) needs no more slots than a
//     binary node.
// Scaling nodes keep the reference's numbering: scale_idx = rank of the node among flagged nodes in
// node-id order (NodeScale, src/treesub.c:7207-7211).
#include "plan.h"
#include <algorithm>
#include <functional>

namespace pb200 {

std::string plan_tree(int nnode, int ns, int root, const int* nson, const int* sons, int maxsons,
                      const unsigned char* scale_flags, Plan& out)
{
    if (nnode < 2 || ns < 1 || ns >= nnode) return "plan_tree: bad nnode/ns";
    if (root < ns || root >= nnode) return "plan_tree: root must be an internal node";
    for (int i = 0; i < nnode; i++) {
        if (nson[i] < 0 || nson[i] > maxsons) return "plan_tree: bad nson";
        if (i < ns && nson[i] != 0) return "plan_tree: tip with children";
        if (i >= ns && nson[i] == 0) return "plan_tree: internal node without children";
        for (int j = 0; j < nson[i]; j++) {
            int s = sons[i * maxsons + j];
            if (s < 0 || s >= nnode || s == i) return "plan_tree: bad son index";
        }
    }
    // every node except the root must be the child of exactly one node
    {
        std::vector<int> seen(nnode, 0);
        for (int i = 0; i < nnode; i++)
            for (int j = 0; j < nson[i]; j++) seen[sons[i * maxsons + j]]++;
        for (int i = 0; i < nnode; i++)
            if ((i == root && seen[i] != 0) || (i != root && seen[i] != 1)) return "plan_tree: not a tree";
    }
    std::vector<int> scale_idx(nnode, -1);
    out = Plan();
    if (scale_flags)
        for (int i = 0, k = 0; i < nnode; i++)
            if (scale_flags[i] && i != root && i >= ns) { scale_idx[i] = k++; out.nscale = k; }

    // slot need of every internal node (iterative post-order: trees can be 5000 tips deep)
    std::vector<int> need(nnode, 0), order;
    {
        std::vector<std::pair<int, int>> st;
        st.push_back({root, 0});
        while (!st.empty()) {
            int v = st.back().first, k = st.back().second;
            if (k < nson[v]) {
                st.back().second++;
                int c = sons[v * maxsons + k];
                if (nson[c] > 0) st.push_back({c, 0});
            }
            else { order.push_back(v); st.pop_back(); }
        }
        for (int v : order) {
            std::vector<int> cn;
            for (int j = 0; j < nson[v]; j++) { int c = sons[v * maxsons + j]; if (nson[c] > 0) cn.push_back(need[c]); }
            std::sort(cn.begin(), cn.end(), std::greater<int>());
            int nd = 1;
            for (size_t i = 0; i < cn.size(); i++) nd = std::max(nd, cn[i] + (i == 0 ? 0 : 1));
            need[v] = nd;
        }
    }

    std::vector<int> P;                      // ops only; the header and the tip order are prepended at the end
    std::vector<int> tip_order;
    std::vector<char> used;
    auto alloc = [&]() { for (size_t i = 0; i < used.size(); i++) if (!used[i]) { used[i] = 1; return (int)i; }
                         used.push_back(1); return (int)used.size() - 1; };
    auto release = [&](int s) { used[s] = 0; };

    struct Child { int kind, idx, node; };
    auto emit_op = [&](int node, const std::vector<Child>& ch, int out_slot, int sidx, int flags) {
        int ntip_op = 0;
        for (const Child& c : ch) ntip_op += (c.kind == CH_TIP);
        P.push_back((int)ch.size()); P.push_back(out_slot); P.push_back(sidx); P.push_back(node); P.push_back(flags);
        P.push_back(ntip_op);
        for (const Child& c : ch) { P.push_back(c.kind); P.push_back(c.idx); P.push_back(c.node);
            if (c.kind == CH_TIP) { out.n_tip_children++; tip_order.push_back(c.idx); } else if (c.kind == CH_INTERNAL) out.n_internal_children++; }
        out.nops++;
    };

    // explicit-stack emission.  result_slot[v] = slot that holds v's vectors once v is done.
    std::vector<int> result_slot(nnode, -1);
    struct Frame { int v; std::vector<int> ic; size_t next; int partial; std::vector<Child> pend; };
    std::vector<Frame> st;
    auto push = [&](int v) {
        Frame f; f.v = v; f.next = 0; f.partial = -1;
        for (int j = 0; j < nson[v]; j++) { int c = sons[v * maxsons + j]; if (nson[c] > 0) f.ic.push_back(c); }
        std::stable_sort(f.ic.begin(), f.ic.end(), [&](int a, int b) { return need[a] > need[b]; });
        st.push_back(std::move(f));
    };
    push(root);
    while (!st.empty()) {
        Frame& f = st.back();
        int v = f.v;
        size_t m = f.ic.size();
        // which children does the next op consume?  first op: tips + up to two internal children.
        if (f.next < m && result_slot[f.ic[f.next]] < 0) { int c = f.ic[f.next]; push(c); continue; }
        if (f.next < m) {
            // child f.next is ready
            int c = f.ic[f.next];
            f.pend.push_back({CH_INTERNAL, result_slot[c], c});
            f.next++;
            bool first_group = (f.partial < 0);
            // in the first group wait for a second internal child (if any) before emitting
            if (first_group && f.pend.size() == 1 && f.next < m) continue;
        }
        // emit an op with what is pending
        std::vector<Child> ch;
        if (f.partial < 0)
            for (int j = 0; j < nson[v]; j++) { int c = sons[v * maxsons + j]; if (nson[c] == 0) ch.push_back({CH_TIP, c, c}); }
        else ch.push_back({CH_PARTIAL, f.partial, v});
        for (const Child& c : f.pend) ch.push_back(c);
        bool last = (f.next >= m);
        for (const Child& c : f.pend) release(c.idx);
        if (f.partial >= 0) release(f.partial);
        f.pend.clear();
        if (!last) {
            int s = alloc();
            emit_op(v, ch, s, -1, OPF_PARTIAL_OUT);
            f.partial = s;
            continue;
        }
        if (v == root) emit_op(v, ch, -1, -1, 0);
        else { int s = alloc(); emit_op(v, ch, s, scale_idx[v], 0); result_slot[v] = s; }
        st.pop_back();
    }
    out.nslots = std::max<int>(1, (int)used.size());
    out.prog.clear();
    out.prog.push_back(out.nops); out.prog.push_back(out.nslots);
    out.prog.push_back((int)tip_order.size()); out.prog.push_back(PROG_HDR + (int)tip_order.size());
    out.prog.insert(out.prog.end(), tip_order.begin(), tip_order.end());
    out.prog.insert(out.prog.end(), P.begin(), P.end());
    return "";
}

}  // namespace pb200

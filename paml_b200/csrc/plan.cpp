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

static std::string check_tree(int nnode, int ns, int root, const int* nson, const int* sons, int maxsons)
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
    return "";
}

std::string plan_tree(int nnode, int ns, int root, const int* nson, const int* sons, int maxsons,
                      const unsigned char* scale_flags, Plan& out)
{
    {
        std::string e = check_tree(nnode, ns, root, nson, sons, maxsons);
        if (!e.empty()) return e;
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
            int nd = 0;      // a node whose children are all tips needs no slot: its result stays in registers
            for (size_t i = 0; i < cn.size(); i++) nd = std::max(nd, cn[i] + (i == 0 ? 0 : 1));
            need[v] = nd;
        }
    }

    std::vector<int> P;                      // ops only; the header and the tip order are prepended at the end
    std::vector<int> tip_order, matseq;
    std::vector<char> used;
    auto alloc = [&]() { for (size_t i = 0; i < used.size(); i++) if (!used[i]) { used[i] = 1; return (int)i; }
                         used.push_back(1); return (int)used.size() - 1; };
    auto release = [&](int s) { used[s] = 0; };

    struct Child { int kind, idx, node; };
    auto emit_op = [&](int node, const std::vector<Child>& ch, int out_slot, int sidx, int flags) {
        int ntip_op = 0;
        for (const Child& c : ch) ntip_op += (c.kind == CH_TIP);
        if (!ch.empty() && ch[0].kind == CH_DIRECT) flags |= OPF_HAS_DIRECT;
        int pos = (int)P.size();
        P.push_back((int)ch.size()); P.push_back(out_slot); P.push_back(sidx); P.push_back(node); P.push_back(flags);
        P.push_back(ntip_op);
        for (const Child& c : ch) { P.push_back(c.kind); P.push_back(c.idx); P.push_back(c.node);
            if (c.kind != CH_PARTIAL) matseq.push_back(c.node);
            if (c.kind == CH_TIP) { out.n_tip_children++; tip_order.push_back(c.idx); }
            else if (c.kind == CH_INTERNAL || c.kind == CH_DIRECT) out.n_internal_children++; }
        out.nops++;
        return pos;
    };

    // Explicit-stack emission.  A finished node leaves its vectors in registers (OUT_REG).  The parent either
    // consumes them at once (CH_DIRECT) or, when it still has to compute another subtree first, parks them: the
    // child's final op is patched to store into a slot.
    std::vector<int> final_op(nnode, -1);
    std::vector<char> done(nnode, 0);
    struct Frame { int v; std::vector<int> ic; size_t next; int parked, parked_kind, parked_node; bool tips_done; };
    std::vector<Frame> st;
    auto push = [&](int v) {
        Frame f; f.v = v; f.next = 0; f.parked = -1; f.parked_kind = 0; f.parked_node = -1; f.tips_done = false;
        for (int j = 0; j < nson[v]; j++) { int c = sons[v * maxsons + j]; if (nson[c] > 0) f.ic.push_back(c); }
        std::stable_sort(f.ic.begin(), f.ic.end(), [&](int a, int b) { return need[a] > need[b]; });
        st.push_back(std::move(f));
    };
    push(root);
    while (!st.empty()) {
        Frame& f = st.back();
        const int v = f.v;
        const size_t m = f.ic.size();
        std::vector<Child> ch;
        if (f.next < m) {
            const int c = f.ic[f.next];
            if (!done[c]) { push(c); continue; }
            const size_t i = f.next++;
            if (i == 0 && m >= 2) {                       // park the first (largest) child while the second is computed
                int s = alloc();
                P[final_op[c] + 1] = s;
                f.parked = s; f.parked_kind = CH_INTERNAL; f.parked_node = c;
                continue;
            }
            ch.push_back({CH_DIRECT, 0, c});
        }
        if (!f.tips_done) {
            for (int j = 0; j < nson[v]; j++) { int c = sons[v * maxsons + j]; if (nson[c] == 0) ch.push_back({CH_TIP, c, c}); }
            f.tips_done = true;
        }
        if (f.parked >= 0) { ch.push_back({f.parked_kind, f.parked, f.parked_node}); release(f.parked); f.parked = -1; }
        if (f.next < m) {                                  // more subtrees to come: park the running product
            int s = alloc();
            emit_op(v, ch, s, -1, OPF_PARTIAL_OUT);
            f.parked = s; f.parked_kind = CH_PARTIAL; f.parked_node = v;
            continue;
        }
        if (v == root) emit_op(v, ch, OUT_ROOT, -1, 0);
        else { final_op[v] = emit_op(v, ch, OUT_REG, scale_idx[v], 0); done[v] = 1; }
        st.pop_back();
    }
    out.nslots = std::max<int>(1, (int)used.size());
    out.prog.clear();
    out.prog.push_back(out.nops); out.prog.push_back(out.nslots);
    out.prog.push_back((int)tip_order.size()); out.prog.push_back(PROG_HDR + (int)tip_order.size() + (int)matseq.size());
    out.prog.push_back((int)matseq.size()); out.prog.push_back(PROG_HDR + (int)tip_order.size());
    out.prog.insert(out.prog.end(), tip_order.begin(), tip_order.end());
    out.prog.insert(out.prog.end(), matseq.begin(), matseq.end());
    out.prog.insert(out.prog.end(), P.begin(), P.end());
    return "";
}

// Programs of the P-stationary kernel (see plan.h): post-order with Sethi-Ullman child order; the tips of a node become a gather op
// (emitted right before the node's own contraction op, so its product is alive as briefly as possible), the internal children a
// contraction op.  G slots are handed out least-recently-released first, which lets the gather warps run ahead of the consumers.
std::string plan_tree_ps(int nnode, int ns, int root, const int* nson, const int* sons, int maxsons,
                         const unsigned char* scale_flags, PlanPS& out)
{
    {
        std::string e = check_tree(nnode, ns, root, nson, sons, maxsons);
        if (!e.empty()) return e;
    }
    out = PlanPS();
    std::vector<int> scale_idx(nnode, -1);
    if (scale_flags)
        for (int i = 0, k = 0; i < nnode; i++)
            if (scale_flags[i] && i != root && i >= ns) { scale_idx[i] = k++; out.nscale = k; }
    // D slots a subtree needs while it is computed, including the one its result occupies (0 for a node with tips only)
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
            if (cn.empty()) { need[v] = 0; continue; }
            std::sort(cn.begin(), cn.end(), std::greater<int>());
            int held = 0, nd = 0;                           // results of the children finished so far that sit in D slots
            for (size_t i = 0; i < cn.size(); i++) { nd = std::max(nd, cn[i] + held); if (cn[i] > 0) held++; }
            need[v] = std::max(nd, held + 1);               // all children's results + the node's own
        }
    }
    std::vector<char> d_used;
    auto d_alloc = [&]() { for (size_t i = 0; i < d_used.size(); i++) if (!d_used[i]) { d_used[i] = 1; return (int)i; }
                           d_used.push_back(1); return (int)d_used.size() - 1; };
    std::vector<int> g_free, g_writes;                       // free list in release order; writes per chunk of each G slot
    auto g_alloc = [&]() { if (g_free.empty()) { g_writes.push_back(0); return (int)g_writes.size() - 1; }
                           int s = g_free.front(); g_free.erase(g_free.begin()); return s; };
    struct Res { int kind = -1, slot = -1, k = 0; };
    std::vector<Res> res(nnode);
    std::vector<int> gops, dops;
    struct Frame { int v; std::vector<int> ic; size_t next; };
    std::vector<Frame> st;
    auto push = [&](int v) {
        Frame f; f.v = v; f.next = 0;
        for (int j = 0; j < nson[v]; j++) { int c = sons[v * maxsons + j]; if (nson[c] > 0) f.ic.push_back(c); }
        std::stable_sort(f.ic.begin(), f.ic.end(), [&](int a, int b) { return need[a] > need[b]; });
        st.push_back(std::move(f));
    };
    push(root);
    while (!st.empty()) {
        Frame& f = st.back();
        if (f.next < f.ic.size()) { int c = f.ic[f.next++]; push(c); continue; }
        const int v = f.v;
        std::vector<int> tips;
        for (int j = 0; j < nson[v]; j++) { int c = sons[v * maxsons + j]; if (nson[c] == 0) tips.push_back(c); }
        int gs = -1, gk = 0;
        if (!tips.empty()) {
            gs = g_alloc(); gk = g_writes[gs]++;
            gops.push_back(gs); gops.push_back(gk); gops.push_back((int)tips.size()); gops.push_back(v);
            for (int t : tips) gops.push_back(t);
            out.n_gops++; out.ntip += (int)tips.size();
        }
        if (f.ic.empty() && v != root) {
            if (scale_idx[v] >= 0) return "plan_tree_ps: a scaling node whose children are all tips";
            res[v].kind = PS_G; res[v].slot = gs; res[v].k = gk;
        }
        else {
            const int out_slot = (v == root) ? -1 : d_alloc();
            // one step per internal child (a lone PS_NONE step for a root with tips only); what happens when the op is complete
            // travels in its last step
            const int nst = std::max<int>(1, (int)f.ic.size());
            for (int i = 0; i < nst; i++) {
                const bool has = i < (int)f.ic.size();
                const int c = has ? f.ic[i] : -1;
                int st[PS_STEP_INTS] = {0, 0, 0, 0, 0, 0, 0, 0};
                st[0] = (has ? res[c].kind : PS_NONE) | (i == 0 ? PS_STEP_FIRST : 0) | (i == nst - 1 ? PS_STEP_LAST : 0);
                st[1] = has ? res[c].slot : 0;
                st[2] = has ? c : 0;
                st[3] = has ? res[c].k : 0;
                st[4] = out_slot; st[5] = scale_idx[v]; st[6] = gs; st[7] = gk;
                dops.insert(dops.end(), st, st + PS_STEP_INTS);
                if (has) out.n_internal_children++;
            }
            for (int c : f.ic) { if (res[c].kind == PS_D) d_used[res[c].slot] = 0; else g_free.push_back(res[c].slot); }
            if (gs >= 0) g_free.push_back(gs);
            res[v].kind = PS_D; res[v].slot = out_slot; res[v].k = 0;
            out.n_dops += std::max<int>(1, (int)f.ic.size());
        }
        st.pop_back();
    }
    out.nD = (int)d_used.size();
    out.nG = (int)g_writes.size();
    out.prog.assign(PS_HDR, 0);
    out.prog[0] = out.n_gops; out.prog[1] = out.n_dops; out.prog[2] = out.nG; out.prog[3] = out.nD;
    out.prog.insert(out.prog.end(), g_writes.begin(), g_writes.end());
    out.prog[4] = (int)out.prog.size();
    out.prog.insert(out.prog.end(), gops.begin(), gops.end());
    out.prog[5] = (int)out.prog.size();
    // the kernel needs W_s next to every use of a G slot: fold it into the upper half of the k fields
    for (size_t i = 0; i + PS_STEP_INTS <= dops.size(); i += PS_STEP_INTS) {
        if ((dops[i] & PS_KIND_MASK) == PS_G) dops[i + 3] |= g_writes[dops[i + 1]] << 16;
        if (dops[i + 6] >= 0) dops[i + 7] |= g_writes[dops[i + 6]] << 16;
    }
    out.prog.insert(out.prog.end(), dops.begin(), dops.end());
    out.prog[6] = out.nscale;
    return "";
}

// Macro-op program of the 4-state kernel (see plan.h).  Same Sethi-Ullman ordering as above; what is parked is
// always the running product of the PARENT (the child's vector already multiplied by its branch matrix), so one
// slot per pending node is enough and the kernel's register state is a single vector per pattern.
std::string plan_tree4(int nnode, int ns, int root, const int* nson, const int* sons, int maxsons,
                       const unsigned char* scale_flags, Plan4& out)
{
    {
        std::string e = check_tree(nnode, ns, root, nson, sons, maxsons);
        if (!e.empty()) return e;
    }
    if (nnode >= 65536) return "plan_tree4: more than 65535 nodes";
    out = Plan4();
    std::vector<char> scaled(nnode, 0);
    if (scale_flags)
        for (int i = 0; i < nnode; i++)
            if (scale_flags[i] && i != root && i >= ns) { scaled[i] = 1; out.nscale++; }

    std::vector<int> need(nnode, 0);
    {
        std::vector<int> order;
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
            int nd = 0;
            for (size_t i = 0; i < cn.size(); i++) nd = std::max(nd, cn[i] + (i == 0 ? 0 : 1));
            need[v] = nd;
        }
    }

    std::vector<char> used;
    auto alloc = [&]() { for (size_t i = 0; i < used.size(); i++) if (!used[i]) { used[i] = 1; return (int)i; }
                         used.push_back(1); return (int)used.size() - 1; };
    const int ntip_total = ns, ngroups = (ntip_total + 3) / 4;
    auto tip_field = [&](int ti) {          // where tip number ti of the walk sits in the packed code words
        const int g = ti >> 2, byte = ti & 3;
        int adv = 0;
        if (byte == 0 && g > 0) adv = (g + 4 < ngroups) ? 1 : 2;
        return byte | (adv << 2) | ((g & 3) << 4);
    };
    auto emit = [&](int opc, int slot, int a, int b) {
        int ntips = (opc == M4_CHERRY) ? 2 : (opc == M4_MV_TIP || opc == M4_TIP_MUL) ? 1 : 0;
        int ti = (int)out.tip_order.size();            // tips are pushed to tip_order right after their op is emitted
        int w0 = opc | (slot << M4_SLOT_SHIFT);
        if (ntips >= 1) w0 |= tip_field(ti) << M4_TIPF0_SHIFT;
        if (ntips >= 2) w0 |= tip_field(ti + 1) << M4_TIPF1_SHIFT;
        out.ops.push_back(w0);
        out.ops.push_back((a & 0xffff) | (b << 16));
    };
    struct Frame { int v; std::vector<int> ic, tips; size_t i; int slot; bool child_done; };
    std::vector<Frame> st;
    auto push = [&](int v) {
        Frame f; f.v = v; f.i = 0; f.slot = -1; f.child_done = false;
        for (int j = 0; j < nson[v]; j++) { int c = sons[v * maxsons + j]; (nson[c] > 0 ? f.ic : f.tips).push_back(c); }
        std::stable_sort(f.ic.begin(), f.ic.end(), [&](int a, int b) { return need[a] > need[b]; });
        st.push_back(std::move(f));
    };
    push(root);
    while (!st.empty()) {
        Frame& f = st.back();
        const size_t k = f.ic.size();
        size_t first_tip = 0;
        bool finished = false;
        if (k == 0) {
            if (f.tips.size() < 2) return "plan_tree4: internal node with fewer than two children";
            emit(M4_CHERRY, 0, f.tips[0], f.tips[1]);
            out.tip_order.push_back(f.tips[0]); out.tip_order.push_back(f.tips[1]);
            first_tip = 2; finished = true;
        }
        else if (!f.child_done) { f.child_done = true; int c = f.ic[f.i]; push(c); continue; }
        else if (k == 1) {
            if (f.tips.empty()) return "plan_tree4: internal node with a single child";
            emit(M4_MV_TIP, 0, f.ic[0], f.tips[0]);
            out.tip_order.push_back(f.tips[0]);
            first_tip = 1; finished = true;
        }
        else if (f.i == 0) {
            f.slot = alloc();
            emit(M4_MV_PUSH, f.slot, f.ic[0], 0);
            f.i = 1; f.child_done = false;
            continue;
        }
        else {
            emit(M4_MV_POP, f.slot, f.ic[f.i], 0);
            if (f.i + 1 < k) { emit(M4_PUSH, f.slot, 0, 0); f.i++; f.child_done = false; continue; }
            used[f.slot] = 0;
            finished = true;
        }
        if (finished) {
            for (size_t t = first_tip; t < f.tips.size(); t++) { emit(M4_TIP_MUL, 0, f.tips[t], 0); out.tip_order.push_back(f.tips[t]); }
            if (scaled[f.v]) out.ops[out.ops.size() - 2] |= M4_SCALE_FLAG;
            if (f.v == root) emit(M4_ROOT, 0, 0, 0);
            st.pop_back();
        }
    }
    out.nslots = std::max<int>(1, (int)used.size());
    return "";
}

}  // namespace pb200

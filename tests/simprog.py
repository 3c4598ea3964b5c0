"""Numpy interpreter of the pruning "program" (paml_b200/csrc/plan.h) -- used by the CPU tests to check the
planner (op order, slot reuse, folded multifurcations, scale indices) without a GPU.  It executes exactly
the op stream the CUDA kernel walks, with slots as arrays, and the arithmetic of the reference."""
import numpy as np

OP_HDR, CH_REC = 6, 3
CH_TIP, CH_INTERNAL, CH_PARTIAL, CH_DIRECT = 0, 1, 2, 3
OUT_ROOT, OUT_REG = -1, -2


def run_program(prog, nslots, p, pmat_fn):
    """Returns (neg_lnL, lnf) for Problem p by interpreting `prog`.  pmat_fn(cls, node) -> P[n,n]."""
    n, N, K = p.n, p.npatt, p.nclass
    nscale = int(p.scale_flags.sum()) if p.scale_flags is not None else 0
    fhK = np.zeros((K, N))
    for cls in range(K):
        slots = [None] * nslots
        prev = None                      # vectors left "in registers" by the previous op
        lnsc = np.zeros(N)
        pc = int(prog[3])
        for _ in range(prog[0]):
            nchild, out_slot, sidx, node, flags, ntip_op = prog[pc:pc + OP_HDR]
            pc += OP_HDR
            acc = np.ones((N, n))
            for _c in range(nchild):
                kind, idx, cnode = prog[pc:pc + CH_REC]
                pc += CH_REC
                if kind == CH_TIP:
                    P = pmat_fn(cls, cnode)
                    zs = p.z[idx]
                    contrib = np.zeros((N, n))
                    for code in np.unique(zs):
                        states = p.charamap[code, :p.nchara[code]]
                        contrib[zs == code] = P[:, states].sum(1)
                    acc *= contrib
                elif kind == CH_DIRECT:
                    assert prev is not None and _c == 0, "CH_DIRECT must be child 0 and follow an OUT_REG op"
                    acc *= prev @ pmat_fn(cls, cnode).T
                elif kind == CH_INTERNAL:
                    assert slots[idx] is not None, "read of an empty slot"
                    acc *= slots[idx] @ pmat_fn(cls, cnode).T
                    slots[idx] = None
                else:
                    assert slots[idx] is not None
                    acc *= slots[idx]
                    slots[idx] = None
            if sidx >= 0:
                mx = acc.max(1)
                tiny = mx < 1e-300
                acc[tiny] = 1.0
                acc[~tiny] /= mx[~tiny, None]
                lnsc += np.where(tiny, -800.0, np.log(np.where(tiny, 1.0, mx)))
            prev = None
            if out_slot >= 0:
                assert slots[out_slot] is None, "slot overwritten while live"
                slots[out_slot] = acc
            elif out_slot == OUT_REG:
                prev = acc
            else:
                fh = acc @ p.pi
                if p.semantics == 0:
                    fh = np.where(fh <= 0, 1e-80, fh)
                    fhK[cls] = np.log(fh) + lnsc
                else:
                    fh = np.where(fh <= 0, 1e-300, fh)
                    fhK[cls] = np.log(fh) + lnsc if nscale else fh
    if p.semantics == 0:
        lnf = fhK[0]
    elif nscale:
        t = fhK.max(0)
        lnf = t + np.log((p.freqK[:, None] * np.exp(fhK - t)).sum(0))
    else:
        lnf = np.log(np.maximum((p.freqK[:, None] * fhK).sum(0), 1e-300))
    lnf = np.where(p.fpatt > 0, lnf, 0.0)
    return -(lnf * p.fpatt).sum(), lnf


# ---- macro-op program of the 4-state kernel (paml_b200/csrc/plan.h, Plan4) ----
M4_CHERRY, M4_MV_TIP, M4_MV_PUSH, M4_MV_POP, M4_PUSH, M4_TIP_MUL, M4_ROOT = range(7)
M4_SCALE_FLAG = 16


def run_program4(ops, tip_order, nslots, p, pmat_fn):
    """Interprets the macro-op stream exactly as prune4v2_kernel does (one running vector, parked partial products,
    tips consumed in tip_order).  Returns (neg_lnL, lnf)."""
    n, N, K = p.n, p.npatt, p.nclass
    nscale = int(p.scale_flags.sum()) if p.scale_flags is not None else 0
    fhK = np.zeros((K, N))

    ngroups = (len(tip_order) + 3) // 4

    def tip_rows(cls, node, ti, field):
        assert tip_order[ti] == node, "tip consumed out of order"
        g, byte = ti >> 2, ti & 3                  # the static placement the kernel relies on
        adv = 0 if (byte != 0 or g == 0) else (1 if g + 4 < ngroups else 2)
        assert field == (byte | (adv << 2) | ((g & 3) << 4)), "bad tip field"
        P = pmat_fn(cls, node)
        zs = p.z[node]
        contrib = np.zeros((N, n))
        for code in np.unique(zs):
            states = p.charamap[code, :p.nchara[code]]
            contrib[zs == code] = P[:, states].sum(1)
        return contrib

    for cls in range(K):
        slots = [None] * nslots
        acc = None
        lnsc = np.zeros(N)
        ti = 0
        done = False
        for w0, w1 in ops:
            assert not done, "ops after ROOT"
            opc, slot = w0 & 15, (w0 >> 17) & 255
            a, b = w1 & 0xffff, (w1 >> 16) & 0xffff
            if opc == M4_CHERRY:
                acc = tip_rows(cls, a, ti, (w0 >> 5) & 63) * tip_rows(cls, b, ti + 1, (w0 >> 11) & 63); ti += 2
            elif opc == M4_MV_TIP:
                acc = (acc @ pmat_fn(cls, a).T) * tip_rows(cls, b, ti, (w0 >> 5) & 63); ti += 1
            elif opc == M4_MV_PUSH:
                assert slots[slot] is None, "slot overwritten while live"
                slots[slot] = acc @ pmat_fn(cls, a).T
                acc = None
            elif opc == M4_MV_POP:
                assert slots[slot] is not None, "read of an empty slot"
                acc = (acc @ pmat_fn(cls, a).T) * slots[slot]
                slots[slot] = None
            elif opc == M4_PUSH:
                assert slots[slot] is None
                slots[slot] = acc
                acc = None
            elif opc == M4_TIP_MUL:
                acc = acc * tip_rows(cls, a, ti, (w0 >> 5) & 63); ti += 1
            else:
                assert opc == M4_ROOT
                fh = acc @ p.pi
                if p.semantics == 0:
                    fh = np.where(fh <= 0, 1e-80, fh)
                    fhK[cls] = np.log(fh) + lnsc
                else:
                    fh = np.where(fh <= 0, 1e-300, fh)
                    fhK[cls] = np.log(fh) + lnsc if nscale else fh
                done = True
            if w0 & M4_SCALE_FLAG:
                mx = acc.max(1)
                tiny = mx < 1e-300
                acc[tiny] = 1.0
                acc[~tiny] /= mx[~tiny, None]
                lnsc += np.where(tiny, -800.0, np.log(np.where(tiny, 1.0, mx)))
        assert done and ti == len(tip_order) and all(s is None for s in slots)
    if p.semantics == 0:
        lnf = fhK[0]
    elif nscale:
        t = fhK.max(0)
        lnf = t + np.log((p.freqK[:, None] * np.exp(fhK - t)).sum(0))
    else:
        lnf = np.log(np.maximum((p.freqK[:, None] * fhK).sum(0), 1e-300))
    lnf = np.where(p.fpatt > 0, lnf, 0.0)
    return -(lnf * p.fpatt).sum(), lnf


# ---- programs of the P-stationary kernel (paml_b200/csrc/plan.h, PlanPS) ----
PS_HDR, PS_GOP_HDR, PS_STEP_INTS = 8, 4, 8
PS_D, PS_G, PS_NONE, PS_KIND_MASK, PS_STEP_FIRST, PS_STEP_LAST = 0, 1, 2, 3, 4, 8


def run_program_ps(prog, p, pmat_fn):
    """Interprets the gather-op list and the step table the way the kernel's two kinds of warps do: the gather ops in their order
    fill G slots, the steps in their order consume them.  Checks what the full / empty handshake relies on: a G slot is only
    rewritten after the step that read it (in program order), the k fields count the writes of a chunk, and a node's output D slot
    never aliases one of its children."""
    n, N, K = p.n, p.npatt, p.nclass
    nscale = int(p.scale_flags.sum()) if p.scale_flags is not None else 0
    n_gops, n_steps, nG, nD, off_g, off_s = (int(x) for x in prog[:6])
    W = [int(x) for x in prog[PS_HDR:PS_HDR + nG]]
    fhK = np.zeros((K, N))

    def tip_rows(cls, tip):
        P = pmat_fn(cls, tip)
        zs = p.z[tip]
        out = np.zeros((N, n))
        for code in np.unique(zs):
            states = p.charamap[code, :p.nchara[code]]
            out[zs == code] = P[:, states].sum(1)
        return out

    for cls in range(K):
        # the gather warps may run ahead as far as the slots allow; interleave lazily: run gather ops until the slot a step needs is full
        g = [None] * nG
        writes = [0] * nG
        d = [None] * nD
        state = {"pc": off_g, "done": 0}

        def gather_until(slot, k):
            while g[slot] is None or g[slot][0] != k:
                assert state["done"] < n_gops, "step waits for a G slot no gather op fills"
                pc = state["pc"]
                gs, gk, nt, _node = (int(x) for x in prog[pc:pc + PS_GOP_HDR])
                assert g[gs] is None, "gather op overwrites a G slot that has not been read"
                assert gk == writes[gs], "k field does not count the writes of the chunk"
                v = np.ones((N, n))
                for t in prog[pc + PS_GOP_HDR:pc + PS_GOP_HDR + nt]:
                    v = v * tip_rows(cls, int(t))
                g[gs] = (gk, v)
                writes[gs] += 1
                state["pc"] = pc + PS_GOP_HDR + nt
                state["done"] += 1

        lnsc = np.zeros(N)
        acc = None
        children_d = []
        for s in range(n_steps):
            st = [int(x) for x in prog[off_s + s * PS_STEP_INTS: off_s + (s + 1) * PS_STEP_INTS]]
            kind = st[0] & PS_KIND_MASK
            if st[0] & PS_STEP_FIRST:
                acc, children_d = None, []
            if kind != PS_NONE:
                if kind == PS_G:
                    assert (st[3] >> 16) == W[st[1]]
                    gather_until(st[1], st[3] & 0xffff)
                    v = g[st[1]][1]
                    g[st[1]] = None
                else:
                    assert d[st[1]] is not None, "read of an empty D slot"
                    v = d[st[1]]
                    children_d.append(st[1])
                c = v @ pmat_fn(cls, st[2]).T
                acc = c if acc is None else acc * c
            if not (st[0] & PS_STEP_LAST):
                continue
            out_slot, sidx, tp, tpk = st[4], st[5], st[6], st[7]
            if tp >= 0:
                assert (tpk >> 16) == W[tp]
                gather_until(tp, tpk & 0xffff)
                acc = g[tp][1] if acc is None else acc * g[tp][1]
                g[tp] = None
            if sidx >= 0:
                mx = acc.max(1)
                tiny = mx < 1e-300
                acc = acc.copy()
                acc[tiny] = 1.0
                acc[~tiny] /= mx[~tiny, None]
                lnsc += np.where(tiny, -800.0, np.log(np.where(tiny, 1.0, mx)))
            if out_slot >= 0:
                assert out_slot not in children_d and d[out_slot] is None, "output D slot aliases a live slot"
                d[out_slot] = acc
                for c in children_d:
                    d[c] = None
            else:
                fh = acc @ p.pi
                if p.semantics == 0:
                    fhK[cls] = np.log(np.where(fh <= 0, 1e-80, fh)) + lnsc
                else:
                    fh = np.where(fh <= 0, 1e-300, fh)
                    fhK[cls] = np.log(fh) + lnsc if nscale else fh
        assert state["done"] == n_gops and writes == W
    if p.semantics == 0:
        lnf = fhK[0]
    elif nscale:
        t = fhK.max(0)
        lnf = t + np.log((p.freqK[:, None] * np.exp(fhK - t)).sum(0))
    else:
        lnf = np.log(np.maximum((p.freqK[:, None] * fhK).sum(0), 1e-300))
    lnf = np.where(p.fpatt > 0, lnf, 0.0)
    return -(lnf * p.fpatt).sum(), lnf

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

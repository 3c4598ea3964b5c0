#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Needs the reference tree (/root/reference) and oracle/_ref/{codeml_hook,baseml_hook} (``make -C oracle ref``):
runs only in the authoring container.  Each case is one of the reference's shipped examples
(examples/*) driven by the reference's own .ctl parser and optimiser; at the
post-fit evaluation (com.print < 0, i.e. AT the MLE) the hook dumps every hot-path input together
with the reference's outputs (-lnL, per-pattern log f_h, fhK, some P(t) matrices) into a .pbfx file.
The known-answer lnL of each example (README values / SURVEY section 4) is asserted here, so the fixtures
are pinned to the reference's documented results, not just to "whatever the binary printed".
"""
import os, re, shutil, subprocess, sys, tempfile

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("PAML_REFERENCE", "/root/reference")
EX = os.path.join(REF, "examples")
OUT = os.path.join(REPO, "tests", "golden")
HOOK = {"codeml": os.path.join(REPO, "oracle", "_ref", "codeml_hook"),
        "baseml": os.path.join(REPO, "oracle", "_ref", "baseml_hook")}


def write_ctl(path, base, overrides):
    """Copy a reference .ctl, overriding / appending `key = value` lines."""
    lines = open(base).read().splitlines()
    seen = set()
    out = []
    for ln in lines:
        m = re.match(r"\s*\*?\s*([A-Za-z_]+)\s*=", ln)
        if m and m.group(1) in overrides and not ln.lstrip().startswith("*"):
            k = m.group(1)
            if k not in seen:
                out.append("%s = %s" % (k, overrides[k]))
                seen.add(k)
            continue
        out.append(ln)
    for k, v in overrides.items():
        if k not in seen:
            out.append("%s = %s" % (k, v))
    open(path, "w").write("\n".join(out) + "\n")


CASES = [
    # name, program, base ctl, files to copy, ctl overrides, dump call, expected lnL (or None), tolerance
    dict(name="brown_hky85", prog="baseml", ctl="baseml.ctl", files=["brown.nuc", "brown.trees"],
         ov=dict(seqfile="brown.nuc", treefile="brown.trees", model=4, fix_alpha=1, alpha=0, getSE=0, RateAncestor=0),
         lnL=-2665.422858),
    dict(name="brown_rev_g8", prog="baseml", ctl="baseml.ctl", files=["brown.nuc", "brown.trees"],
         ov=dict(seqfile="brown.nuc", treefile="brown.trees", model=7, fix_alpha=0, alpha=0.5, ncatG=8, getSE=0, RateAncestor=0),
         lnL=-2619.320596),
    dict(name="brown_k80_clock", prog="baseml", ctl="baseml.ctl", files=["brown.nuc", "brown.rooted.trees"],
         ov=dict(seqfile="brown.nuc", treefile="brown.rooted.trees", model=1, clock=1, fix_alpha=1, alpha=0, getSE=0, RateAncestor=0),
         lnL=None),
    dict(name="brown_tn93_g5_ambig", prog="baseml", ctl="baseml.ctl", files=["brown.nuc", "brown.trees"],
         ov=dict(seqfile="brown_ambig.nuc", treefile="brown.trees", model=6, fix_alpha=0, alpha=0.5, ncatG=5, cleandata=0, getSE=0, RateAncestor=0),
         lnL=None, ambig="nuc"),
    dict(name="lysozyme_m0", prog="codeml", ctl="lysozyme/lysozymeSmall.ctl", files=["lysozyme/lysozymeSmall.nuc", "lysozyme/lysozymeSmall.trees"],
         ov=dict(seqfile="lysozymeSmall.nuc", treefile="lysozymeSmall.trees", model=0, NSsites=0, CodonFreq=2, getSE=0, RateAncestor=0, noisy=2),
         lnL=-906.017440),
    dict(name="lysozyme_branch2", prog="codeml", ctl="lysozyme/lysozymeSmall.ctl", files=["lysozyme/lysozymeSmall.nuc", "lysozyme/lysozymeSmall.trees"],
         ov=dict(seqfile="lysozymeSmall.nuc", treefile="lysozymeSmall.trees", model=2, NSsites=0, CodonFreq=2, getSE=0, RateAncestor=0, noisy=2),
         lnL=None),
    dict(name="lysozyme_bsA", prog="codeml", ctl="lysozyme/lysozymeSmall.ctl", files=["lysozyme/lysozymeSmall.nuc", "lysozyme/lysozymeSmall.trees"],
         ov=dict(seqfile="lysozymeSmall.nuc", treefile="lysozymeSmall.trees", model=2, NSsites=2, CodonFreq=2, getSE=0, RateAncestor=0, noisy=2),
         lnL=None),
    dict(name="lysozyme_m3_ambig", prog="codeml", ctl="lysozyme/lysozymeSmall.ctl", files=["lysozyme/lysozymeSmall.nuc", "lysozyme/lysozymeSmall.trees"],
         ov=dict(seqfile="lysozymeSmall_ambig.nuc", treefile="lysozymeSmall.trees", model=0, NSsites=3, ncatG=3, CodonFreq=2, cleandata=0, getSE=0, RateAncestor=0, noisy=2),
         lnL=None, ambig="codon"),
    dict(name="stewart_wag_g4", prog="codeml", ctl="aaml.ctl", files=["stewart.aa", "stewart.trees"],
         ov=dict(seqfile="stewart.aa", treefile="stewart.trees", seqtype=2, model=2, aaRatefile=os.path.join(REF, "dat", "wag.dat"),
                 fix_alpha=0, alpha=0.5, ncatG=4, getSE=0, RateAncestor=0, noisy=2, cleandata=1),
         lnL=-1028.924417),
    dict(name="stewart_poisson", prog="codeml", ctl="aaml.ctl", files=["stewart.aa", "stewart.trees"],
         ov=dict(seqfile="stewart.aa", treefile="stewart.trees", seqtype=2, model=0, fix_alpha=1, alpha=0, getSE=0, RateAncestor=0, noisy=2, cleandata=1),
         lnL=None),
    dict(name="mhc_m0", prog="codeml", ctl="MHC.Swanson2002MBE/codeml.ctl", files=["MHC.Swanson2002MBE/bigmhc.phy", "MHC.Swanson2002MBE/bigmhc.trees"],
         ov=dict(seqfile="bigmhc.phy", treefile="bigmhc.trees", NSsites=0, method=0, getSE=0, RateAncestor=0, noisy=2),
         lnL=-8225.154790),
    dict(name="mhc_m2a", prog="codeml", ctl="MHC.Swanson2002MBE/codeml.ctl", files=["MHC.Swanson2002MBE/bigmhc.phy", "MHC.Swanson2002MBE/bigmhc.trees"],
         ov=dict(seqfile="bigmhc.phy", treefile="bigmhc.trees", NSsites=2, method=0, getSE=0, RateAncestor=0, noisy=2),
         lnL=-7231.154540),
    # model families beyond the BASELINE configurations: matexp (UNREST), spectral sets per branch (nhomo, free-ratio)
    dict(name="brown_unrest", prog="baseml", ctl="baseml.ctl", files=["brown.nuc", "brown.trees"],
         ov=dict(seqfile="brown.nuc", treefile="brown.trees", model=8, fix_alpha=1, alpha=0, getSE=0, RateAncestor=0),
         lnL=None),
    dict(name="brown_hky_nhomo4", prog="baseml", ctl="baseml.ctl", files=["brown.nuc", "brown.trees"],
         ov=dict(seqfile="brown.nuc", treefile="brown.trees", model=4, nhomo=4, fix_alpha=1, alpha=0, getSE=0, RateAncestor=0),
         lnL=None),
    dict(name="brown_k80_nhomo2_g4", prog="baseml", ctl="baseml.ctl", files=["brown.nuc", "brown.trees"],
         ov=dict(seqfile="brown.nuc", treefile="brown.trees", model=1, nhomo=2, fix_alpha=0, alpha=0.5, ncatG=4, getSE=0, RateAncestor=0),
         lnL=None),
    dict(name="lysozyme_freeratio", prog="codeml", ctl="lysozyme/lysozymeSmall.ctl", files=["lysozyme/lysozymeSmall.nuc", "lysozyme/lysozymeSmall.trees"],
         ov=dict(seqfile="lysozymeSmall.nuc", treefile="lysozymeSmall.trees", model=1, NSsites=0, CodonFreq=2, getSE=0, RateAncestor=0, noisy=2),
         lnL=None),
]


# Branch fixtures (tests/golden/branch/): `method = 1` runs stopped at the k-th lfuntdd / lfuntdd_SiteClass call with an internal
# b; the fixture holds the tree as minbranches had rerooted it and the reference's -lnL(t), d/dt, d2/dt2 of that branch.
BRANCH_CASES = [
    dict(name="lysozyme_m0_branch", prog="codeml", ctl="lysozyme/lysozymeSmall.ctl", files=["lysozyme/lysozymeSmall.nuc", "lysozyme/lysozymeSmall.trees"],
         ov=dict(seqfile="lysozymeSmall.nuc", treefile="lysozymeSmall.trees", model=0, NSsites=0, CodonFreq=2, getSE=0, RateAncestor=0, noisy=0, method=1), call=4),
    dict(name="lysozyme_m2a_branch", prog="codeml", ctl="lysozyme/lysozymeSmall.ctl", files=["lysozyme/lysozymeSmall.nuc", "lysozyme/lysozymeSmall.trees"],
         ov=dict(seqfile="lysozymeSmall.nuc", treefile="lysozymeSmall.trees", model=0, NSsites=2, CodonFreq=2, getSE=0, RateAncestor=0, noisy=0, method=1), call=9),
    dict(name="brown_hky_g4_branch", prog="baseml", ctl="baseml.ctl", files=["brown.nuc", "brown.trees"],
         ov=dict(seqfile="brown.nuc", treefile="brown.trees", model=4, fix_alpha=0, alpha=0.5, ncatG=4, getSE=0, RateAncestor=0, method=1, noisy=0), call=2),
    dict(name="stewart_wag_g4_branch", prog="codeml", ctl="aaml.ctl", files=["stewart.aa", "stewart.trees"],
         ov=dict(seqfile="stewart.aa", treefile="stewart.trees", seqtype=2, model=2, aaRatefile=os.path.join(REF, "dat", "wag.dat"),
                 fix_alpha=0, alpha=0.5, ncatG=4, getSE=0, RateAncestor=0, noisy=0, cleandata=1, method=1), call=3),
    dict(name="mhc_m0_branch", prog="codeml", ctl="MHC.Swanson2002MBE/codeml.ctl", files=["MHC.Swanson2002MBE/bigmhc.phy", "MHC.Swanson2002MBE/bigmhc.trees"],
         ov=dict(seqfile="bigmhc.phy", treefile="bigmhc.trees", NSsites=0, method=1, fix_blength=1, getSE=0, RateAncestor=0, noisy=0), call=150),
]


def run_branch_case(c, workroot):
    w = os.path.join(workroot, c["name"])
    os.makedirs(w, exist_ok=True)
    for f in c["files"]:
        shutil.copy(os.path.join(EX, f), w)
    ov = dict(c["ov"]); ov["outfile"] = "out.txt"
    write_ctl(os.path.join(w, "run.ctl"), os.path.join(EX, c["ctl"]), ov)
    fx = os.path.join(w, c["name"] + ".pbfx")
    env = dict(os.environ, PBX_BRANCH_DUMP=fx, PBX_BRANCH_CALL=str(c["call"]))
    r = subprocess.run([HOOK[c["prog"]], "run.ctl"], cwd=w, env=env, stdin=subprocess.DEVNULL,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    m = re.search(r"pbx: branch record (.*)", r.stdout)
    if r.returncode != 0 or not m:
        print(r.stdout[-3000:])
        raise SystemExit("branch case %s failed (rc=%d)" % (c["name"], r.returncode))
    os.makedirs(os.path.join(OUT, "branch"), exist_ok=True)
    shutil.copy(fx, os.path.join(OUT, "branch", c["name"] + ".pbfx"))
    print("%-24s %s  %d bytes" % (c["name"], m.group(1), os.path.getsize(fx)))


def add_ambiguity(path_in, path_out, kind):
    """Deterministically sprinkle ambiguity characters into a PHYLIP file whose names sit on their own lines."""
    lines = open(path_in).read().splitlines()
    out = [lines[0]]
    pos = 0      # nucleotide position within the current sequence
    iseq = 0
    for ln in lines[1:]:
        if not (len(ln.strip()) > 20 and re.fullmatch(r"[ACGTacgt\s]+", ln)):
            out.append(ln)
            if ln.strip():
                pos = 0; iseq += 1
            continue
        chars = list(ln)
        for i, ch in enumerate(chars):
            if ch not in "ACGTacgt":
                continue
            pos += 1
            key = pos * 7 + iseq * 13
            if kind == "nuc":
                if key % 17 == 0: chars[i] = "N"
                elif key % 23 == 0: chars[i] = "R" if ch in "AGag" else "Y"
                elif key % 41 == 0: chars[i] = "-"
            else:  # codon data: third positions only, so no stop codons can be implied
                if pos % 3 == 0 and ((pos // 3) * 5 + iseq * 3) % 19 == 0:
                    chars[i] = "N" if (pos // 3) % 2 else ("Y" if ch in "CTct" else "R")
        out.append("".join(chars))
    open(path_out, "w").write("\n".join(out) + "\n")


def run_case(c, workroot):
    w = os.path.join(workroot, c["name"])
    os.makedirs(w, exist_ok=True)
    for f in c["files"]:
        shutil.copy(os.path.join(EX, f), w)
    if c.get("ambig"):
        src = os.path.join(w, os.path.basename(c["files"][0]))
        add_ambiguity(src, os.path.join(w, c["ov"]["seqfile"]), c["ambig"])
    ctl = os.path.join(w, "run.ctl")
    ov = dict(c["ov"]); ov["outfile"] = "out.txt"
    write_ctl(ctl, os.path.join(EX, c["ctl"]), ov)
    fx = os.path.join(w, c["name"] + ".pbfx")
    env = dict(os.environ, PBX_DUMP=fx, PBX_DUMP_CALL=c.get("call", "final"))
    r = subprocess.run([HOOK[c["prog"]], "run.ctl"], cwd=w, env=env, stdin=subprocess.DEVNULL,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    open(os.path.join(w, "log.txt"), "w").write(r.stdout)
    m = re.search(r"pbx: dumped fixture .* -lnL=([-0-9.eE+]+)\)", r.stdout)
    if r.returncode != 0 or not m:
        print(r.stdout[-3000:])
        raise SystemExit("case %s failed (rc=%d)" % (c["name"], r.returncode))
    lnL = -float(m.group(1))
    if c["lnL"] is not None:
        assert abs(lnL - c["lnL"]) < 2e-5 * max(1, abs(c["lnL"])) * 1e-2 + 5e-4, (c["name"], lnL, c["lnL"])
    shutil.copy(fx, os.path.join(OUT, c["name"] + ".pbfx"))
    print("%-22s lnL = %.6f  (expected %s)  %d bytes" % (c["name"], lnL, c["lnL"], os.path.getsize(fx)))


if __name__ == "__main__":
    only = sys.argv[1:]
    workroot = tempfile.mkdtemp(prefix="pb200_golden_")
    print("work dir", workroot)
    for c in CASES:
        if only and c["name"] not in only:
            continue
        run_case(c, workroot)
    for c in BRANCH_CASES:
        if only and c["name"] not in only:
            continue
        run_branch_case(c, workroot)

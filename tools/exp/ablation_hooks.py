#!/usr/bin/env python3
"""Insert / remove the ablation switches of the line-marching apply kernel (development tool).

    python tools/exp/ablation_hooks.py apply     # edits portable-multigrid_b200/csrc/pmg_apply_sweep.h in place
    python tools/exp/ablation_hooks.py revert    # git checkout of that file

With the hooks in, tools/exp/build_exp.sh accepts -DPMG_EXP_NOSTAGE (no u / b / x_old copies), -DPMG_EXP_NOP12 (no y and x
sweeps), -DPMG_EXP_NOP3 (no z sweep, epilogue, stores); all three together leave the skeleton (barriers, mbarrier round trip,
address loops).  Results of round 1: profiles/r01_v6_ablation_experiment.txt.  The product source never carries the hooks."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HDR = os.path.join(ROOT, "portable-multigrid_b200", "csrc", "pmg_apply_sweep.h")
HOOKS = [  # (anchor, text inserted right after it)
    ("  static PMG_HD int issue_row(const double *vec, int64_t e, int nelem, int64_t n_local, double *dst, uint64_t *bar)\n  {\n",
     "#ifdef PMG_EXP_NOSTAGE\n    return 0;\n#endif\n"),
    ("  static PMG_HD void load_u_async(const PmgSweepParams<P> &p, const TileGeom &t, int lane, double *A, int gz0, int npl)\n  {\n",
     "#ifdef PMG_EXP_NOSTAGE\n    return;\n#endif\n"),
    ("  static PMG_HD void load_e_async(const PmgSweepParams<P> &p, const TileGeom &t, int lane, double *E, int gz0, int npl)\n  {\n",
     "#ifdef PMG_EXP_NOSTAGE\n    return;\n#endif\n"),
    ("                            int gz0, int npl)\n  {\n",
     "#ifdef PMG_EXP_NOP12\n    return;\n#endif\n"),
    ("  static PMG_HD void phase2(const PmgSweepParams<P> &p, const TileGeom &t, const ThreadState &st, double *Cb, double *Db, int npl,\n"
     "                            int alt = 0)\n  {\n",
     "#ifdef PMG_EXP_NOP12\n    return;\n#endif\n"),
    ("                            const double *Db, const double *E, int cz0, int nlay, int cz_write)\n  {\n    switch (mode_of(p)) {",
     None),  # handled below: the hook goes before the switch
]


def main():
    if len(sys.argv) != 2 or sys.argv[1] not in ("apply", "revert"):
        sys.exit(__doc__)
    if sys.argv[1] == "revert":
        subprocess.run(["git", "-C", ROOT, "checkout", "--", os.path.relpath(HDR, ROOT)], check=True)
        return
    s = open(HDR).read()
    if "PMG_EXP_NOSTAGE" in s:
        sys.exit("hooks already in")
    for anchor, text in HOOKS:
        if s.count(anchor) != 1:
            sys.exit("anchor not found exactly once:\n" + anchor)
        if text is None:
            s = s.replace(anchor, anchor.replace("    switch (mode_of(p)) {", "#ifdef PMG_EXP_NOP3\n    return;\n#endif\n    switch (mode_of(p)) {"))
        else:
            s = s.replace(anchor, anchor + text)
    open(HDR, "w").write(s)


if __name__ == "__main__":
    main()

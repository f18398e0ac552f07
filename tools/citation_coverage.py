"""Which executable lines of the reference's hot-path .m files does oracle/literal.py cite?

The literal oracle restates the reference line by line and cites the lines it follows: every function's docstring names
the file and the range it covers ("(iterEKF.m:30-210 ...)") and the body carries "# :a-b" / "# :a" comments relative to that
file.  This tool reads the reference sources (REF_ROOT, default /root/reference/src) and reports, per file, the executable
lines inside the restated ranges that no citation covers.  tests/test_literal_citations.py runs it where the reference is
present.  usage: citation_coverage.py [-v]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.environ.get("MEK_REF_ROOT", "/root/reference/src")
FILES = {"OB_step.m": "MPC-EKF4FastCharge", "iterEKF.m": "UTILITY", "initKF.m": "UTILITY", "EKFmatsHandler.m": "MPC-EKF4FastCharge",
         "predMat.m": "MPC-EKF4FastCharge", "constraintsMPC.m": "MPC-EKF4FastCharge", "hildreth.m": "MPC-EKF4FastCharge",
         "initMPC.m": "MPC-EKF4FastCharge", "iterMPC.m": "MPC-EKF4FastCharge", "runMPC.m": "MPC-EKF4FastCharge"}


def executable_lines(path):
    """1-based numbers of the lines that do something: not blank, not comment, not a bare block keyword / continuation tail."""
    out = set()
    cont = False
    for i, raw in enumerate(open(path, encoding="latin-1"), 1):
        line = raw.split("%", 1)[0].strip() if not raw.lstrip().startswith("%") else ""
        was_cont = cont
        cont = line.endswith("...")
        if not line or was_cont:
            continue
        if re.fullmatch(r"(end|else|otherwise|try|catch)\b;?", line) or line.startswith("function "):
            continue
        out.add(i)
    return out


def citations(literal_path):
    """{file: set(lines)} cited by comments in literal.py, and {file: [(lo, hi)]} the ranges its functions claim to restate.
    A claim is a docstring "(file.m:a-b ...)" or a "# :a-b" comment on a `def` line (nested helper functions); a citation is
    any other "# :a-b[,c-d]" comment, relative to the file of the enclosing top-level function."""
    cited, claimed, excluded = {}, {}, {}
    cur = None
    for ln in open(literal_path).read().splitlines():
        if re.match(r"^def \w+", ln):
            cur = None
        m = re.search(r'"""[^"]*?(\w+\.m):(\d+)-(\d+)', ln)
        if m and m.group(1) in FILES and (ln.strip().startswith('"""')):
            cur = m.group(1)
            claimed.setdefault(cur, []).append((int(m.group(2)), int(m.group(3))))
            continue
        if not cur:
            continue
        is_def = re.match(r"^\s+def \w+", ln) is not None
        for part, why in re.findall(r"#\s*not restated\s*(:\d[\d,\-:\s]*)\(([^)]*)\)", ln):
            for a, b in re.findall(r":?(\d+)(?:-(\d+))?", part):
                for i in range(int(a), int(b or a) + 1):
                    excluded.setdefault(cur, {})[i] = why
        if "# not restated" in ln:
            continue
        comment = ln.split("#", 1)[1] if "#" in ln else ""
        # ":a-b,c-d" lists anywhere in the comment (relative to the current file), and "file.m:a-b" naming that file
        for part in re.findall(r"(?:(?<![\w.])|%s):(\d[\d,\-\s]*)" % re.escape(cur), comment):
            for a, b in re.findall(r"(\d+)(?:-(\d+))?", part):
                lo, hi = int(a), int(b or a)
                if is_def:
                    claimed.setdefault(cur, []).append((lo, hi))
                else:
                    cited.setdefault(cur, set()).update(range(lo, hi + 1))
    return cited, claimed, excluded


def report(verbose=False):
    cited, claimed, excluded = citations(os.path.join(ROOT, "oracle", "literal.py"))
    res = {}
    for f, sub in FILES.items():
        path = os.path.join(REF_ROOT, sub, f)
        ex = executable_lines(path)
        inside = set()
        for lo, hi in claimed.get(f, []):
            inside |= {i for i in ex if lo <= i <= hi}
        outside = sorted(ex - inside)
        skipped = {i: w for i, w in excluded.get(f, {}).items() if i in inside}
        missing = sorted(inside - cited.get(f, set()) - set(skipped))
        res[f] = dict(executable=len(inside), cited=len(inside - set(skipped) - set(missing)), not_restated=skipped, missing=missing, unclaimed=outside)
        if verbose:
            lines = open(path, encoding="latin-1").read().splitlines()
            print("%-18s %4d executable lines in the restated ranges, %4d cited, %3d deliberately not restated, %3d uncited" %
                  (f, len(inside), res[f]["cited"], len(skipped), len(missing)))
            for i, w in sorted(skipped.items()):
                print("    %4d  [not restated: %s]  %s" % (i, w, lines[i - 1].strip()[:70]))
            for i in outside:
                print("    %4d  [outside every restated range]  %s" % (i, lines[i - 1].strip()[:80]))
            for i in missing:
                print("    %4d  %s" % (i, lines[i - 1].rstrip()[:110]))
    return res


if __name__ == "__main__":
    r = report(verbose=True)
    tot = sum(v["executable"] for v in r.values()); c = sum(v["cited"] for v in r.values())
    print("total: %d of %d executable lines cited (%.1f %%)" % (c, tot, 100.0 * c / max(tot, 1)))

#!/usr/bin/env python3
"""Runs the reference's UNMODIFIED World-API unit tests, linked through the World.cpp seam so that every World{} they
construct is a CudaWorld (playrho_b200/_build/UnitTests_cuda, built by playrho_b200/host/Makefile), one test per process
(a crash or hang of one test does not take the others down) and prints a summary: passed / failed / crashed per suite
and the failing test names. Needs a GPU.

  python tools/run_unittests_cuda.py [--filter 'World.*'] [--timeout 60] > profiles/r01_unittests_cuda.txt
"""
import argparse
import os
import subprocess
import sys
import collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "playrho_b200", "_build", "UnitTests_cuda")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filter", default="*")
    ap.add_argument("--timeout", type=float, default=60.0)
    args = ap.parse_args()
    out = subprocess.run([EXE, "--gtest_list_tests", "--gtest_filter=" + args.filter], capture_output=True, text=True).stdout
    tests, suite = [], None
    for line in out.splitlines():
        if not line.strip() or line.startswith("Running main"):
            continue
        if not line.startswith(" "):
            suite = line.split()[0]
        else:
            tests.append(suite + line.split()[0])
    res = collections.OrderedDict()
    per_suite = collections.defaultdict(lambda: collections.Counter())
    for t in tests:
        try:
            r = subprocess.run([EXE, "--gtest_filter=" + t, "--gtest_brief=1"], capture_output=True, text=True, timeout=args.timeout)
            if r.returncode == 0:
                st = "passed"
            elif "[  FAILED  ]" in r.stdout:
                st = "failed"
            else:
                st = "crashed(rc=%d)" % r.returncode
            detail = ""
            if st != "passed":
                lines = [l for l in (r.stdout + r.stderr).splitlines() if "Failure" in l or "exception" in l.lower() or "what()" in l or "rror" in l]
                detail = " | ".join(lines[:2])[:300]
        except subprocess.TimeoutExpired:
            st, detail = "timeout", ""
        res[t] = (st, detail)
        per_suite[t.split(".")[0]][st.split("(")[0]] += 1
    tot = collections.Counter(v[0].split("(")[0] for v in res.values())
    print("# the reference's unmodified UnitTests/{World,WorldBody,WorldContact,WorldShape,WorldJoint,<eleven joint types>,Joint,Dump,WorldManifold}.cpp on World{CudaWorld} (seam link), one process per test")
    print("# total %d: %s" % (len(res), ", ".join("%s %d" % kv for kv in sorted(tot.items()))))
    for s, c in per_suite.items():
        print("%-16s %s" % (s, ", ".join("%s %d" % kv for kv in sorted(c.items()))))
    print("# not passing:")
    for t, (st, detail) in res.items():
        if st != "passed":
            print("%-60s %s  %s" % (t, st, detail))
    return 0


if __name__ == "__main__":
    sys.exit(main())

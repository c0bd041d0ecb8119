"""Run the reference's OWN pytest files (/root/reference/tests, unmodified) against the reference
package executed over oracle/jaxshim, as a check of the stand-in itself: if the reference's tests
pass on it, the stand-in implements the jax / flax subset the reference relies on.

    python tools/run_reference_tests.py [-n WORKERS] [pytest args ...]

Build container only.  Writes profiles/r2_reference_tests_over_shim.txt: totals, and every failing
test with its cause.  Expected non-passes: tests that call jax.grad or jnp.bfloat16 (not provided —
the product path needs neither) and the one test whose own fp32 finite difference is noisier than
its tolerance.
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def main():
    args = sys.argv[1:]
    workers = "8"
    if args[:1] == ["-n"]:
        workers, args = args[1], args[2:]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "jaxshim"), os.path.join(REF, "src")]))
    with tempfile.TemporaryDirectory() as tmp:      # the reference checkout is read-only: no cache, no rootdir there
        cmd = [sys.executable, "-m", "pytest", os.path.join(REF, "tests"), "-q", "-p", "no:cacheprovider", f"--rootdir={tmp}",
               "--timeout", "1200", "-n", workers, "-rfEs", "--tb=line"] + args
        out = subprocess.run(cmd, env=env, cwd=tmp, capture_output=True, text=True).stdout
    summarize(out)


def summarize(out, dest=os.path.join(ROOT, "profiles", "r2_reference_tests_over_shim.txt")):
    tail = [l for l in out.splitlines() if re.search(r"\d+ (passed|failed)", l)]
    fails = [l for l in out.splitlines() if l.startswith(("FAILED", "ERROR"))]
    skips = [l for l in out.splitlines() if l.startswith("SKIPPED")]
    cause = lambda l: ("jax.grad (not in the stand-in)" if "jax.grad" in l else
                       "bfloat16 (not in the stand-in)" if "bfloat16" in l else
                       "other")
    lines = ["Reference test-suite (/root/reference/tests, unmodified) over oracle/jaxshim",
             "command: python tools/run_reference_tests.py", "", *tail, ""]
    by = {}
    for l in fails:
        by.setdefault(cause(l), []).append(l)
    for c, ls in sorted(by.items()):
        lines += [f"--- {len(ls)} x {c}"] + ["  " + l[:230] for l in ls] + [""]
    lines += [f"--- {len(skips)} skipped by the reference's own markers"] + ["  " + l[:200] for l in skips]
    open(dest, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--summarize":
        summarize(open(sys.argv[2]).read())
    else:
        main()

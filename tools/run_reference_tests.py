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


LIGHT = ["test_cosmology.py", "test_layers.py", "test_layers_vel.py", "test_style_layers.py",
         "test_style_layers_vel.py", "test_blocks.py", "test_blocks_vel.py", "test_style_blocks.py", "test_style_blocks_vel.py",
         "test_nbody_emulator.py"]
# whole 128^3 networks (or eight of them per process_box) in every test: hours of CPU over the stand-in
HEAVY = ["test_subbox.py", "test_style_nbody_emulator_vel_core.py", "test_style_nbody_emulator_core.py", "test_nbody_emulator_vel_core.py",
         "test_nbody_emulator_core.py"]


def main():
    args = sys.argv[1:]
    workers = "4"
    if args[:1] == ["-n"]:
        workers, args = args[1], args[2:]
    files = [a for a in args if a.endswith(".py")] or LIGHT + HEAVY
    args = [a for a in args if not a.endswith(".py")]
    threads = str(max(1, (os.cpu_count() or 8) // int(workers)))      # torch / MKL threads per xdist worker
    env = dict(os.environ, OMP_NUM_THREADS=threads, MKL_NUM_THREADS=threads, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "jaxshim"), os.path.join(REF, "src")]))
    out = ""
    with tempfile.TemporaryDirectory() as tmp:      # the reference checkout is read-only: no cache, no rootdir there
        for f in files:
            cmd = [sys.executable, "-m", "pytest", os.path.join(REF, "tests", f), "-q", "-p", "no:cacheprovider",
                   f"--rootdir={tmp}", "--timeout", "1500", "-n", workers, "-rfEs", "--tb=line"] + args
            r = subprocess.run(cmd, env=env, cwd=tmp, capture_output=True, text=True).stdout
            last = [l for l in r.splitlines() if re.search(r"\d+ (passed|failed|error)", l)]
            print(f, "|", last[-1] if last else "no summary", flush=True)
            out += f"##### {f}\n" + r
            open(os.path.join(ROOT, "gpurun_out", "reference_tests_over_shim.log"), "w").write(out)
            summarize(out, quiet=True)


def summarize(out, dest=os.path.join(ROOT, "profiles", "r2_reference_tests_over_shim.txt"), quiet=False):
    tail, cur = [], ""
    for l in out.splitlines():
        if l.startswith("##### "):
            cur = l[6:]
        elif re.search(r"\d+ (passed|failed|error)", l) and " in " in l:
            tail.append(f"{cur:42s} {l.strip('= ')}")
    tot = {k: sum(int(m) for l in tail for m in re.findall(r"(\d+) " + k, l)) for k in ("passed", "failed", "skipped", "error")}
    tail.append(f"{'TOTAL':42s} " + ", ".join(f"{v} {k}" for k, v in tot.items()))
    fails, skips, cur = [], [], ""
    for l in out.splitlines():
        if l.startswith("##### "):
            cur = l[6:]
        elif l.startswith(("FAILED ", "ERROR ")):
            fails.append(l.replace(" ::", f" {cur}::", 1))
        elif l.startswith("SKIPPED"):
            skips.append(f"{cur}: {l}")
    cause = lambda l: ("jax.grad -> NotImplementedError (reverse-mode AD is not in the stand-in; the forward path needs none)"
                       if "jax.grad" in l or "NotImplemen" in l or "gradient" in l else
                       "the test's own fp32 central difference (dz = 1e-4) of growth_factor: the stand-in's hyp2f1 is scipy's fp64 value "
                       "rounded to fp32, so neighbouring values carry uncorrelated rounding; the AD value it is compared with is correct"
                       if "finite_difference" in l else
                       "bfloat16 (not in the stand-in)" if "bfloat16" in l else
                       "float16 through numpy (overflow / precision of the stand-in's fp16, not the reference's)" if "float16" in l or "fp16" in l else
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
    if not quiet:
        print("\n".join(lines))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--summarize":
        summarize(open(sys.argv[2]).read())
    else:
        main()

"""Runs the reference's OWN executable (oracle/_ref/sparseBench-CRS-ref: its sources, CRS, no MPI, strict IEEE, built by
`make -C oracle ref` where /root/reference exists) and stores the solver lines of its stdout in ref_stdout.json --
the golden of tests/test_gpu_dropin_link.py. Run in the development container:  python tests/golden/make_ref_stdout.py"""
import json
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "sparseBench-CRS-ref")
KLEIN = os.path.join(HERE, "reference_fixtures", "matrix_band_klein.mtx")
CASES = {
    "gen8": ["-x", "8", "-y", "8", "-z", "8", "-i", "12"],
    "gen16_eps": ["-x", "16", "-y", "16", "-z", "16", "-i", "60", "-e", "1e-6"],
    "gen24x20x12": ["-x", "24", "-y", "20", "-z", "12", "-i", "40"],
    "klein": ["-m", KLEIN, "-i", "10"],
}
KEEP = re.compile(r"^(Initial Residual|Iteration =|Solution performed|Difference between)")


def solver_lines(text):
    out = []
    for line in text.splitlines():
        if KEEP.match(line):
            out.append(re.sub(r"and took .*", "and took", line))     # wall time differs, the count must not
    return out


if __name__ == "__main__":
    gold = {}
    for name, args in CASES.items():
        r = subprocess.run([EXE] + args, capture_output=True, text=True, check=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
        gold[name] = {"args": [a if a != KLEIN else "<klein>" for a in args], "lines": solver_lines(r.stdout)}
    json.dump(gold, open(os.path.join(HERE, "ref_stdout.json"), "w"), indent=1)
    for k, v in gold.items():
        print(k, len(v["lines"]), "lines;", v["lines"][-2:])

#!/usr/bin/env python
"""tests/golden/dp_driver_tree.json: the output tree of the UNMODIFIED reference program Percolation_FD/1storderDP_delta.cpp (BASELINE
config 1's program) for one stdin protocol with r = 4 replicates: the final table and the PRELIMINARY_* tables that
expanded_percolationDornic_2D dumps after the third replicate (P/:2101-2128) -- names, headers, line counts and the (a, c, L, t) columns.
The densities come from random_device-seeded streams and are compared elsewhere (ensemble tests).  The t of every a-block's row 0 is
recorded as 0 (what the runner sets, P/:1840), not as read from the reference's file: the
reference's surv_rep array is one element short (P/:1827) and the out-of-bounds `surv_rep[index] += 1` of the last bin lands, in this
build, on rho_rep_avg_var[0][0] -- the t of row 0 reads "number of survivors" (undefined behaviour, not reproduced).

Needs /root/reference (this container).  Run from the repo root:  python tests/golden/make_dp_prelim_golden.py"""
import json
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/simulations/Percolation_FD"
STDIN = ["0.1", "5", "16", "4", "-0.72", "-0.70", "2"]      # dt, t_max, L, r, a_start, a_end, div


def main():
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "bin"))
        os.makedirs(os.path.join(tmp, "Data", "SPDP"))
        exe = os.path.join(tmp, "bin", "dp_ref")
        subprocess.check_call(["g++", "-O1", "-w", "-std=c++17", "-fopenmp", "-I" + REF, os.path.join(REF, "1storderDP_delta.cpp"),
                               os.path.join(REF, "stochasticSPDP.cpp"), "-o", exe])
        subprocess.run([exe], input="\n".join(STDIN) + "\n", text=True, cwd=os.path.join(tmp, "bin"), stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, check=True)
        files = {}
        for f in sorted(os.listdir(os.path.join(tmp, "Data", "SPDP"))):
            lines = open(os.path.join(tmp, "Data", "SPDP", f)).read().splitlines()
            rows = [[float(x) for x in ln.split(",")] for ln in lines[1:]]
            files["Data/SPDP/" + f] = {"lines": len(lines), "header": lines[0], "a_c_L": [r[:3] for r in rows], "t": [0.0 if i % (len(rows) // (1 if f.startswith("PRELIMINARY") else int(STDIN[6]))) == 0 else r[3] for i, r in enumerate(rows)],
                                       "last_column": [r[-1] for r in rows] if f.startswith("PRELIMINARY") else None}
    out = {"stdin": STDIN, "files": files}
    with open(os.path.join(ROOT, "tests", "golden", "dp_driver_tree.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(len(files), "files:", *files)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""tests/golden/noise_off_driver_tree.json: the output tree of the UNMODIFIED reference program Rietkerk_FD/order_test_rietkerk.cpp
(the noise-off a-sweep) for two argument lists -- which files it writes (names carry L, t, dt, a, D, dx, alpha, gmax and
the replicate), their line counts and headers, and the (a, r, L, t) columns of the final table.  The fields themselves start from a
random_device-seeded speckle mask in the reference, so only what does not depend on the mask is recorded.

Needs /root/reference (this container): the reference's two sources are compiled where they lie with -DSPB=1 and the exprtk stand-in
of oracle/ref_shim, nothing is copied.  Run from the repo root:  python tests/golden/make_noise_off_driver_golden.py"""
import json
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/simulations/Rietkerk_FD"
CASES = {
    "small_kick": ["0.1", "3", "24", "2", "0.5", "0.6", "2", "5"],
    # dP = 10000: main() names the frame folder "..._dP_10000", the runner "..._dP_1e+04" (setprecision(3)): no frame is written
    "big_kick": ["0.1", "2", "16", "1", "1.5", "1.5", "1", "10000"],
    # the same program run twice into the same tree: a snapshot is never overwritten, the replicate suffix becomes (largest present) + 1
    # (save_frame R/:2845-2853, findMaxRepNo R/:1544)
    "small_kick_twice": ["0.1", "3", "24", "2", "0.5", "0.6", "2", "5"],
}


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "bin"))
        exe = os.path.join(tmp, "bin", "order_test_ref")
        subprocess.check_call(["g++", "-O1", "-w", "-std=c++23", "-DSPB=1", "-DINIT=1", "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-I" + REF,
                               os.path.join(REF, "order_test_rietkerk.cpp"), os.path.join(REF, "rietkerk_bjork_basic.cpp"), "-lgomp", "-o", exe])
        for name, args in CASES.items():
            data = os.path.join(tmp, "Data")
            subprocess.run(["rm", "-rf", data])
            for _ in range(2 if name.endswith("_twice") else 1):
                subprocess.run([exe] + args, cwd=os.path.join(tmp, "bin"), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
            files = {}
            for dirpath, _, fs in os.walk(data):
                for f in fs:
                    p = os.path.join(dirpath, f)
                    lines = open(p).read().splitlines()
                    rec = {"lines": len(lines), "header": lines[0]}
                    if "1stOrderCC" in f:
                        rec["a_r_L_t"] = [[float(x) for x in ln.split(",")[:4]] for ln in lines[1:]]
                    files[os.path.relpath(p, tmp)] = rec
            dirs = sorted(os.path.relpath(os.path.join(dp, d), tmp) for dp, ds, _ in os.walk(data) for d in ds)
            out[name] = {"argv": args, "files": files, "dirs": dirs}
    with open(os.path.join(ROOT, "tests", "golden", "noise_off_driver_tree.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, len(v["files"]), "files", len(v["dirs"]), "dirs")


if __name__ == "__main__":
    main()

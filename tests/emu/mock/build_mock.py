"""TEST TOOL: tests/emu/mock/libtm_lattice.so -- the product's C-ABI sources (csrc/tm_lattice.cu, tm_fke.cu, tm_consumer.cu,
tm_consumer3.cu) compiled by g++ against the stand-in CUDA runtime next to this file, with every `kernel<<<grid, block, smem,
stream>>>(args)` rewritten to emu_launch(kernel, args, grid, block, smem) (tests/emu/tm_cpu_emu.h: the kernels' own source on host
threads).  The product sources are not touched: the rewrite happens on copies under tests/emu/mock/_build/.

Use (see README.md here): dry-run `-m gpu` tests of code that has not yet been on a GPU, at small sizes:

    python tests/emu/mock/build_mock.py
    TM_TEST_MOCK_LIB=$PWD/tests/emu/mock/libtm_lattice.so LD_PRELOAD=$PWD/tests/emu/mock/libtm_lattice.so \
        python -m pytest tests/test_zz_consumer_step3.py -m gpu

Not a fallback: the library is never built by __graft_entry__.build(), lives under tests/, and nothing in trophic_maelstorm_b200/
refers to it."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
CSRC = os.path.join(ROOT, "trophic_maelstorm_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
KERNEL = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;]*>)?)\s*$")


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite(text):
    """kernel<<<grid, block[, smem[, stream]]>>>(args...)  ->  mock_launch(grid, block, smem, kernel, args...)"""
    out, pos, n = "", 0, 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            break
        m = KERNEL.search(text[pos:k])
        assert m, text[max(pos, k - 80):k + 40]
        k0 = pos + m.start(1)
        e = text.index(">>>", k)
        cfg = split_top(text[k + 3:e])
        assert len(cfg) in (2, 3, 4), text[k0:e + 3]
        a0 = e + 3
        while text[a0].isspace():
            a0 += 1
        assert text[a0] == "(", text[k0:a0 + 1]
        depth, a1 = 0, a0
        while True:
            if text[a1] == "(":
                depth += 1
            elif text[a1] == ")":
                depth -= 1
                if depth == 0:
                    break
            a1 += 1
        args = text[a0 + 1:a1].strip()
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += text[pos:k0] + "mock_launch((unsigned)(%s), (unsigned)(%s), (size_t)(%s), %s%s)" % (
            cfg[0], cfg[1], smem, m.group(1), ", " + args if args else "")
        pos = a1 + 1
        n += 1
    return out + text[pos:], n


def main(san=False):
    os.makedirs(BUILD, exist_ok=True)
    srcs = []
    for name in ("tm_lattice.cu", "tm_fke.cu", "tm_consumer.cu", "tm_consumer3.cu"):
        text, n = rewrite(open(os.path.join(CSRC, name)).read())
        # the sources include their headers relative to csrc/: keep that working from the copy
        text = text.replace('#include "../../include/', '#include "%s/include/' % ROOT)
        out = os.path.join(BUILD, name.replace(".cu", "_mock.cpp"))
        open(out, "w").write(text)
        srcs.append(out)
        print("%s: %d launch sites rewritten" % (name, n))
    lib = os.path.join(HERE, "libtm_lattice.so")
    flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"] if san else ["-O2"]
    cmd = ["/usr/bin/g++", "-std=c++20", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-Wno-unused-function"] + flags + \
          ["-include", "cmath", "-include", "math.h", "-I" + HERE, "-I" + os.path.dirname(HERE), "-I" + CSRC] + srcs + ["-Wl,-soname,libtm_lattice.so", "-Wl,--allow-multiple-definition", "-o", lib]
    subprocess.check_call(cmd)
    print(lib)


if __name__ == "__main__":
    main(san="--asan" in sys.argv)

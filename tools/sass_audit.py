"""Prove that a change left the shipped kernels alone: compile psnerf_b200/csrc/*.cu of a reference commit and of the working tree for
sm_100a and compare the SASS of every kernel the two have in common, instruction by instruction (addresses and encodings stripped).
New template parameters with a `false` default are matched to the old instantiation names.

    python tools/sass_audit.py <git-rev> [--jobs 8] [--allow 1]

Exit code 0 when every common kernel differs in at most --allow instruction lines (ptxas sometimes swaps the operands of a
commutative op when surrounding template code changes), 1 otherwise.  Needs nvcc and cuobjdump, no GPU."""
import argparse
import concurrent.futures
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    res, cur = {}, None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            line = re.sub(r"/\*[0-9a-f]{4}\*/", "", line)
            res[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
    return res


def compile_tree(src_root, out_dir, jobs):
    csrc = os.path.join(src_root, "psnerf_b200", "csrc")
    files = sorted(f for f in os.listdir(csrc) if f.endswith(".cu"))

    def one(f):
        o = os.path.join(out_dir, f[:-3] + ".o")
        r = subprocess.run(["nvcc"] + FLAGS + ["-c", os.path.join(csrc, f), "-o", o], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("nvcc failed on %s:\n%s" % (f, r.stderr[-2000:]))
        return f[:-3] + ".o"

    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs) as ex:
        return list(ex.map(one, files))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rev")
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--allow", type=int, default=1)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        old_src, old_obj, new_obj = (os.path.join(tmp, d) for d in ("src", "old", "new"))
        for d in (old_src, old_obj, new_obj):
            os.makedirs(d)
        tar = subprocess.run(["git", "-C", ROOT, "archive", a.rev, "psnerf_b200/csrc", "include"], capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", old_src], input=tar, check=True)
        old_files = compile_tree(old_src, old_obj, a.jobs)
        compile_tree(ROOT, new_obj, a.jobs)
        total = bad = 0
        for f in old_files:
            if not os.path.exists(os.path.join(new_obj, f)):
                print("object gone:", f)
                bad += 1
                continue
            o, n = kernels(os.path.join(old_obj, f)), kernels(os.path.join(new_obj, f))
            for name, body in o.items():
                cands = [name] + [name.replace("EEEvNS", "ELb0" * k + "EEEvNS", 1) for k in (1, 2, 3)]
                match = next((c for c in cands if c in n), None)
                total += 1
                if match is None:
                    print("no counterpart: %s %s" % (f, name))
                    bad += 1
                    continue
                d = sum(1 for x, y in zip(body, n[match]) if x != y) + abs(len(body) - len(n[match]))
                if d > a.allow:
                    print("DIFFERS (%d of %d lines): %s %s" % (d, len(body), f, name))
                    bad += 1
        print("%d kernels compared with %s, %d differ" % (total, a.rev, bad))
        sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

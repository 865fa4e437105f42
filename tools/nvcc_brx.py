#!/usr/bin/env python3
"""nvcc wrapper used for bt_tile.cu:  python tools/nvcc_brx.py <nvcc arguments of a -c compile>

Runs exactly the steps `nvcc --dryrun` prints (cpp, cudafe++, cicc, ptxas, fatbinary, host compile) and, between cicc and
ptxas, rewrites the micro-op switch of the register programs into an indirect branch (tools/ptx_brx.py).  If anything in the
rewrite fails the PTX is left as nvcc produced it, so the object is always built.
"""
import os
import re
import shlex
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, here)
import ptx_brx  # noqa: E402


def main():
    args = sys.argv[1:]
    out = args[args.index("-o") + 1]
    keep = os.path.join(os.path.dirname(out) or ".", "keep_" + os.path.splitext(os.path.basename(out))[0])
    os.makedirs(keep, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    dry = subprocess.run([nvcc, "--dryrun", "--keep", "--keep-dir", keep] + args, capture_output=True, text=True)
    if dry.returncode != 0:
        sys.stderr.write(dry.stderr)
        return dry.returncode
    env = dict(os.environ)
    for line in dry.stderr.splitlines():
        if not line.startswith("#$ "):
            continue
        cmd = line[3:].strip()
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)=(.*)$", cmd)
        if m:  # nvcc prints its environment as shell-style assignments
            val = m.group(2).strip()
            if len(val) >= 2 and val[0] == '"' and val[-1] == '"' and val.count('"') == 2:
                val = val[1:-1]
            env[m.group(1)] = re.sub(r"\$(\w+)", lambda mm: env.get(mm.group(1), ""), val)
            continue
        r = subprocess.run(cmd, shell=True, env=env, capture_output=True, text=True)
        sys.stderr.write(r.stderr)
        sys.stdout.write(r.stdout)
        if r.returncode != 0:
            return r.returncode
        if "cicc" in shlex.split(cmd)[0]:
            toks = shlex.split(cmd)
            ptx = toks[toks.index("-o") + 1]
            try:
                text = open(ptx).read()
                patched = ptx_brx.patch(text)
                if patched != text:
                    open(ptx + ".orig", "w").write(text)
                    open(ptx, "w").write(patched)
            except Exception as e:  # keep the build alive: the unpatched PTX is correct
                sys.stderr.write(f"nvcc_brx: rewrite skipped ({e})\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())

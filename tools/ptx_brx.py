#!/usr/bin/env python3
"""Replace the compare-and-branch trees that nvcc generates for the micro-op `switch` of the register programs
(bt_tile.cu, run_prog) by ONE indirect branch (`brx.idx` over a `.branchtargets` table).

nvcc (CUDA 12.9) never emits brx.idx: a 60-way switch becomes a 6-deep tree of setp / bra pairs, and at ~11 issue cycles
per dependent pair that decode costs more than the arithmetic of a micro-op.  The source marks the dispatch point and the
start of every case with PTX comments:

    // BT_DISPATCH %r<N>;        value switched on (a 32-bit register)
    // BT_CASE <site>;           first statement of `case <site>:`   (255 = default)

This script works on the PTX between cicc and ptxas (tools/nvcc_brx.py drives nvcc's own steps).  For every function
with exactly one dispatch marker it
  1. finds, for each case marker, the label of the basic block that contains it (nothing but straight-line code may sit
     between that label and the marker);
  2. checks that every block of the compare tree holds only setp / bra / cvt / and / mov (so that skipping it is safe);
  3. inserts  `brx.idx %r<N>, <table>;`  right after the dispatch marker; the tree becomes unreachable and ptxas drops it.
Functions that fail a check are left untouched (the build still works, just slower) and reported on stderr.
"""
import re
import sys

LABEL = re.compile(r"^(\$L__BB\d+_\d+):")
FUNC = re.compile(r"^\.(visible |weak )?\.?(entry|func)")
TREE_OK = re.compile(r"^\s*(@%p\d+\s+)?(setp|bra|cvt|and|mov)\b")


def patch_function(lines, start, end, fname, table_id):
    disp = [i for i in range(start, end) if "// BT_DISPATCH" in lines[i]]
    if not disp:
        return 0
    if len(disp) != 1:
        print(f"ptx_brx: {fname}: {len(disp)} dispatch markers (loop was duplicated) -- left as is", file=sys.stderr)
        return 0
    d = disp[0]
    m = re.search(r"BT_DISPATCH (%r\d+);", lines[d])
    if not m:
        print(f"ptx_brx: {fname}: cannot parse the dispatch marker -- left as is", file=sys.stderr)
        return 0
    reg = m.group(1)
    # case markers -> labels
    targets = {}
    for i in range(start, end):
        mc = re.search(r"// BT_CASE (\d+);", lines[i])
        if not mc:
            continue
        site = int(mc.group(1))
        if site in targets:
            print(f"ptx_brx: {fname}: case {site} appears twice -- left as is", file=sys.stderr)
            return 0
        j = i - 1
        label = None
        while j > d:
            s = lines[j].strip()
            ml = LABEL.match(s)
            if ml:
                label = ml.group(1)
                break
            if re.match(r"^(@%p\d+\s+)?bra\b", s) or s.startswith("brx") or s.startswith("ret") or s.startswith("bar."):
                break
            j -= 1
        if label is None:
            print(f"ptx_brx: {fname}: case {site} does not start a labelled block -- left as is", file=sys.stderr)
            return 0
        targets[site] = label
    if 255 not in targets:
        print(f"ptx_brx: {fname}: no default case marker -- left as is", file=sys.stderr)
        return 0
    # the compare tree: blocks reachable from the dispatch marker before a case label is reached
    case_labels = set(targets.values())
    label_line = {}
    for i in range(start, end):
        ml = LABEL.match(lines[i].strip())
        if ml:
            label_line[ml.group(1)] = i
    seen, todo = set(), [d + 1]
    while todo:
        i = todo.pop()
        while i < end:
            s = lines[i].strip()
            ml = LABEL.match(s)
            if ml:
                if ml.group(1) in case_labels:
                    break
                if ml.group(1) in seen:
                    break
                seen.add(ml.group(1))
                i += 1
                continue
            if not s or s.startswith("//") or s.startswith(".loc") or s.startswith(".pragma"):
                i += 1
                continue
            if not TREE_OK.match(s):
                print(f"ptx_brx: {fname}: compare tree holds `{s}` -- left as is", file=sys.stderr)
                return 0
            mb = re.match(r"^(@%p\d+\s+)?bra(\.uni)?\s+(\$L__BB\d+_\d+);", s)
            if mb:
                tgt = mb.group(3)
                if tgt not in case_labels and tgt not in seen:
                    if tgt not in label_line:
                        print(f"ptx_brx: {fname}: unknown label {tgt} -- left as is", file=sys.stderr)
                        return 0
                    seen.add(tgt)
                    todo.append(label_line[tgt] + 1)
                if not mb.group(1):  # unconditional: the block ends here
                    break
            i += 1
    n = max(k for k in targets if k != 255) + 1
    table = [targets.get(k, targets[255]) for k in range(n)]
    # indices >= n (the padding op 255 is never executed) must not reach brx.idx: clamp to the default entry
    table.append(targets[255])
    tname = f"$BT_TBL_{table_id}"
    ins = [
        "\t{",
        "\t.reg .u32 %bt_idx;",
        f"\tmin.u32 %bt_idx, {reg}, {n};",
        f"\t{tname}: .branchtargets " + ", ".join(table) + ";",
        f"\tbrx.idx %bt_idx, {tname};",
        "\t}",
    ]
    # after the "// end inline asm" line that closes the marker
    k = d + 1
    while k < end and "end inline asm" not in lines[k]:
        k += 1
    lines[k + 1:k + 1] = ins
    print(f"ptx_brx: {fname}: {len(targets) - 1} cases -> brx.idx ({len(seen)} tree blocks bypassed)", file=sys.stderr)
    return len(ins)


def patch(text):
    lines = text.split("\n")
    i = 0
    table_id = 0
    while i < len(lines):
        if FUNC.match(lines[i]):
            fname = lines[i].split("(")[0].split()[-1]
            # function body: from the first '{' to the matching closing brace at column 0
            j = i
            while j < len(lines) and not lines[j].startswith("{"):
                j += 1
            e = j
            while e < len(lines) and not lines[e].startswith("}"):
                e += 1
            added = patch_function(lines, j, e, fname, table_id)
            table_id += 1
            i = e + added + 1
        else:
            i += 1
    return "\n".join(lines)


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    out = patch(open(src).read())
    open(dst, "w").write(out)

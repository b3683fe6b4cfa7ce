#!/usr/bin/env python
"""Instruction counts of the hot loops of the SHIPPED kernels, read from the SASS of minorg_b200/libminorg_b200.so
(cuobjdump; needs no GPU).  bench.py takes the per-step instruction counts of its roofline from the JSON this writes
instead of literals, and the excerpts are committed next to it:

    python tools/sass_counts.py          -> profiles/sass_counts.json, profiles/r02_sass_<kernel>_inner_loop.txt

Hot path of screen_pairs_kernel<L,K,1024,0>: the (warp, 32-gRNA group) step that fails the warp vote - from the loop head
(target of the backward branch) through the pair-word loads, the threshold network and VOTE, then the branch-taken tail
back to the head.  The exact second stage lies on the fall-through side of the vote and is counted separately.
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "minorg_b200", "libminorg_b200.so")
INSTR = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*")


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, funcs = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = INSTR.match(line)
        if m and cur is not None:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return funcs


def mnemonic(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]


def branch_target(text):
    m = re.search(r"\b(BRA(?:\.\w+)*)\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", text)
    return int(m.group(2), 16) if m else None


def pairs_hot_path(ins):
    """-> (stage-1 hot path, stage-2 fall-through path) of a screen_pairs_kernel instantiation."""
    addr = {a: i for i, (a, _) in enumerate(ins)}
    # the run of consecutive shared loads = stage 1
    best, i = None, 0
    while i < len(ins):
        if mnemonic(ins[i][1]) == "LDS":
            j = i
            while j < len(ins) and mnemonic(ins[j][1]) == "LDS":
                j += 1
            if best is None or j - i > best[1] - best[0]:
                best = (i, j)
            i = j
        else:
            i += 1
    first_lds = best[0]
    # the vote and its forward branch
    v = next(k for k in range(best[1], len(ins)) if mnemonic(ins[k][1]) == "VOTE")
    b = next(k for k in range(v, len(ins)) if mnemonic(ins[k][1]) == "BRA")
    tail_at = addr[branch_target(ins[b][1])]
    # the tail runs to the backward branch whose target is at or before the first load = the loop head
    k = tail_at
    while True:
        t = branch_target(ins[k][1]) if mnemonic(ins[k][1]) == "BRA" else None
        if t is not None and t <= ins[first_lds][0]:
            head = addr[t]
            break
        k += 1
    hot = ins[head:b + 1] + ins[tail_at:k + 1]
    stage2 = ins[b + 1:tail_at]
    return hot, stage2


def tally(path):
    c = collections.Counter(mnemonic(t) for _, t in path)
    return dict(sorted(c.items()))


def rows_loop(ins):
    """Per-character body of gapped_rows_kernel: the innermost loop holding the 128-bit shared loads."""
    addr = {a: i for i, (a, _) in enumerate(ins)}
    lds = [i for i, (_, t) in enumerate(ins) if t.startswith("LDS.128") or " LDS.128" in t]
    if not lds:
        return None
    best = None
    for k, (a, t) in enumerate(ins):
        if mnemonic(t) == "BRA":
            tg = branch_target(t)
            if tg is not None and tg < a and tg in addr and addr[tg] <= lds[0] and k >= lds[-1]:
                if best is None or k - addr[tg] < best[1] - best[0]:
                    best = (addr[tg], k)
    return ins[best[0]:best[1] + 1] if best else None


def main():
    funcs = functions(LIB)
    prof = os.path.join(ROOT, "profiles")
    os.makedirs(prof, exist_ok=True)
    out = {"library": "minorg_b200/libminorg_b200.so", "how": "cuobjdump -sass, tools/sass_counts.py", "kernels": {}}
    for name, ins in funcs.items():
        m = re.search(r"screen_pairs_kernelILi(\d+)ELi(\d+)ELi1024ELb0EEE", name)
        if m and (int(m.group(1)), int(m.group(2))) in ((20, 4), (23, 4), (20, 0), (20, 2)):
            L, K = int(m.group(1)), int(m.group(2))
            hot, stage2 = pairs_hot_path(ins)
            key = "screen_pairs_kernel<%d,%d,1024,0>" % (L, K)
            out["kernels"][key] = {"mangled": name, "stage1_hot_path": tally(hot), "stage1_instructions": len(hot),
                                   "stage2_fall_through": tally(stage2), "stage2_instructions": len(stage2)}
            with open(os.path.join(prof, "r02_sass_pairs_L%d_K%d_inner_loop.txt" % (L, K)), "w") as f:
                f.write("# %s\n# stage-1 hot path (vote fails): loop head .. VOTE + branch, then the branch-taken tail\n" % name)
                for a, t in hot:
                    f.write("/*%04x*/  %s\n" % (a, t))
        m = re.search(r"gapped_rows_kernelILi20ELi20EEE", name)
        if m:
            body = rows_loop(ins)
            if body:
                out["kernels"]["gapped_rows_kernel<20,20>"] = {"mangled": name, "per_character": tally(body), "instructions": len(body)}
                with open(os.path.join(prof, "r02_sass_gapped_rows_L20_char_loop.txt"), "w") as f:
                    f.write("# %s\n# loop over the characters of one 16-character text block\n" % name)
                    for a, t in body:
                        f.write("/*%04x*/  %s\n" % (a, t))
    with open(os.path.join(prof, "sass_counts.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    json.dump(out["kernels"], sys.stdout, indent=1, sort_keys=True)
    print()


if __name__ == "__main__":
    main()

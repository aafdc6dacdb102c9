#!/usr/bin/env python
"""SASS evidence for the claims DESIGN.md makes about the kernels (no GPU needed: cuobjdump reads the built
library).  Writes profiles/r02_sass_*.txt:
  * the sector walk's inner loop of fused_kernel<256,8,...> with its instruction mix per block visit;
  * the blocked walk's loop of the same kernel (the round-1 default) for comparison;
  * knn_brute_kernel: the TMA bulk copy (UBLKCP) and mbarrier (SYNCS) instructions;
  * the evict-first 256-bit store of the leaf-id scratch.
python tools/sass_evidence.py"""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "rfsi_b200" / "librfsi_b200.so"
OUT = ROOT / "profiles"


def functions():
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    out, name, cur = {}, None, []
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                out[name] = cur
            name, cur = m.group(1), []
        elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            cur.append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln).rstrip())
    if name:
        out[name] = cur
    return out


def addr(ln):
    return int(re.match(r"\s+/\*([0-9a-f]+)\*/", ln).group(1), 16)


def opcode(ln):
    body = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", ln)
    body = re.sub(r"^@!?U?P\d\s+", "", body)
    return body.split()[0].rstrip(";")


def loops(lines):
    """backward branches: (start index, end index) of every loop body, innermost first by length"""
    at = {addr(l): i for i, l in enumerate(lines)}
    out = []
    for i, l in enumerate(lines):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", l)
        if m and int(m.group(1), 16) in at and at[int(m.group(1), 16)] <= i:
            out.append((at[int(m.group(1), 16)], i))
    return out


def mix(lines):
    c = collections.Counter(opcode(l) for l in lines)
    alu = sum(v for k, v in c.items() if re.match(r"(LOP3|IADD3|LEA|VIMNMX|VIADD|SEL|ISETP|SHF|PRMT|IMNMX)", k))
    fma = sum(v for k, v in c.items() if k.startswith("IMAD"))
    lsu = sum(v for k, v in c.items() if re.match(r"(LDG|LDS|STG|STS|LD\.|ST\.)", k))
    return c, alu, fma, lsu


def main():
    fn = functions()
    fused = next(k for k in fn if "fused_kernelILi256ELi8ELb0ELb0ELb1" in k)
    L = fn[fused]
    found = {}
    for a, b in loops(L):
        body = L[a:b + 1]
        ops = [opcode(l) for l in body]
        n64, n32, nlds = ops.count("LDG.E.64.CONSTANT"), ops.count("LDG.E.CONSTANT"), ops.count("LDS")
        if n64 == 8 and n32 == 16 and nlds == 24:
            found["sector"] = body            # 8 chains x (one 64-bit load + two 32-bit loads + three rank loads)
        if n64 == 0 and n32 == 40 and nlds == 32:
            found["blocked"] = body           # 8 chains x (four node loads + the exit word, four rank loads)
    for name, levels, what in (("sector", 3, "sector image (sector.cuh: traverse_s), 3 levels per block visit"),
                               ("blocked", 4, "blocked image (forest.cuh: traverse_b), 4 levels per block visit")):
        body = found.get(name)
        with open(OUT / f"r02_sass_{name}_loop.txt", "w") as f:
            f.write(f"# {fused}\n# inner loop of the {what}, J = 8 chains per thread\n")
            if body is None:
                f.write("# loop not identified in this build\n")
                continue
            c, alu, fma, lsu = mix(body)
            n = len(body)
            f.write(f"# {n} instructions per iteration = {n / 8:.1f} per chain and block visit = {n / 8 / levels:.2f} per tree level\n")
            f.write(f"# pipes: {alu} ALU (LOP3/IADD3/LEA/VIMNMX/...), {fma} FMA-pipe integer (IMAD*), {lsu} LSU; the ALU and FMA pipes\n"
                    f"# each take a warp instruction every 2 cycles per SM sub-partition -> floor {max(2 * alu, 2 * fma, n)} cycles per iteration\n")
            f.write("# mix: " + ", ".join(f"{k} {v}" for k, v in c.most_common()) + "\n#\n")
            f.write("\n".join(body) + "\n")
    knn = next(k for k in fn if "knn_brute_kernel" in k)
    with open(OUT / "r02_sass_knn_brute_tma.txt", "w") as f:
        f.write(f"# {knn}\n# TMA bulk copies of station tiles (cp.async.bulk -> UBLKCP) and their mbarrier (SYNCS)\n")
        for l in fn[knn]:
            if re.search(r"UBLKCP|SYNCS|DFMA|DADD|DMUL|DSETP", l):
                f.write(l + "\n")
    with open(OUT / "r02_sass_evict_first_store.txt", "w") as f:
        f.write(f"# {fused}\n# leaf-id / quantile-key scratch: 256-bit evict-first stores (st.global.L2::evict_first.v8.b32)\n")
        for l in L:
            if re.search(r"STG\.E\..*256|STG.*EF", l):
                f.write(l + "\n")
    qk = [k for k in fn if "qrf_direct_kernelILi16E" in k]
    with open(OUT / "r02_sass_qrf_direct.txt", "w") as f:
        f.write(f"# {qk[0] if qk else '?'}\n# DIRECT form of the quantile kernel with the counting-sort selection (qrf.cuh: qrf_row_fast): key rows arrive by\n"
                "# cp.async (LDGSTS), the histogram is 16 shared-memory atomics per lane (ATOMS.POPC.INC: lanes on the same bucket are\n"
                "# combined), the placement 16 returning atomics on the cursors (ATOMS.ADD) + 16 scattered stores, warp minimum / maximum /\n"
                "# counts and the peel ranking are REDUX / CREDUX, the only global loads left are the skip bytes and the two dictionary\n"
                "# look-ups per quantile.\n")
        if qk:
            Lq = fn[qk[0]]
            c = collections.Counter(opcode(l) for l in Lq)
            f.write("# whole kernel: " + ", ".join(f"{k} {v}" for k, v in c.most_common()
                                                     if re.match(r"LDG|LDGSTS|REDUX|CREDUX|ATOMS|LDS|STS|STG|VOTE|IMAD.HI|VIMNMX3", k)) + "\n#\n")
            # straight-line code of the fast path after the histogram: warp scan, cursor table, placement (ATOMS.ADD + STS)
            popc = [i for i, l in enumerate(Lq) if "ATOMS.POPC.INC" in l]
            last_atoms = popc[15] if len(popc) > 15 else popc[-1]              # the 16th: end of the fast path's histogram
            adds = [i for i in range(last_atoms, len(Lq)) if "ATOMS.ADD" in Lq[i]]
            stop = adds[15] if len(adds) > 15 else adds[-1]
            while stop + 1 < len(Lq) and not re.search(r"BRA|BSSY|BSYNC|WARPSYNC", Lq[stop + 1]) and stop - last_atoms < 200:
                stop += 1
            body = Lq[last_atoms + 1:stop + 1]
            cm = collections.Counter(opcode(l) for l in body)
            f.write(f"# after the histogram: warp scan of the 256 buckets, cursor table, and the placement of the row's 16 keys per lane\n"
                    f"# (a returning ATOMS.ADD on the bucket's cursor and a scattered STS each): {len(body)} instructions, straight line\n")
            f.write("# mix: " + ", ".join(f"{k} {v}" for k, v in cm.most_common(12)) + "\n")
            f.write("\n".join(body) + "\n#\n# the peel loop of a window (one distinct value per iteration):\n")
            for a2, b2 in loops(Lq):
                ops = [opcode(l) for l in Lq[a2:b2 + 1]]
                if any(o.startswith("CREDUX.MIN") for o in ops) and b2 - a2 < 24 and "VOTE.ANY" in ops:
                    f.write("\n".join(Lq[a2:b2 + 1]) + "\n")
                    break
    for p in sorted(OUT.glob("r02_sass_*.txt")):
        print(p.name, sum(1 for _ in open(p)), "lines")


if __name__ == "__main__":
    main()

"""Static cost model of a pair-kernel inner loop from its SASS (no GPU needed).

Model (measured on B200, profiles/r02_rf_probe.txt and r02_issue_model.txt): a sub-partition reads about two 32-bit
register operands per cycle, and operands served by the operand-reuse cache (`.reuse` on the PREVIOUS instruction, same
source slot, same register) are free.  An fp64 instruction occupies the fp64 pipe for 2 cycles (DADD/DMUL) or ~2.2
(DFMA) and costs max(pipe, fresh_reads / 2): a DFMA with three fresh 64-bit sources costs 3 cycles, with two 2.2.
Other instructions cost fresh_reads / 2 (IMAD with two register sources ~1 cycle, LOP3 / LDS address ~0.5).

usage: sass_cost.py FILE.sass [--pairs-per-shfl-pair R]
Finds the innermost loop(s) that contain SHFL.IDX (the ring walk), prints per loop: instruction histogram, fresh register
reads, and estimated cycles per pair = cost / (ring steps * R)."""
import re
import sys
from collections import Counter

INS = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")
FP64_PIPE = {"DFMA": 2.2, "DADD": 2.0, "DMUL": 2.0}


def parse(path):
    out = []
    for line in open(path):
        m = INS.match(line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        txt = m.group(2).strip()
        out.append((addr, txt))
    return out


def operands(txt):
    """-> (opcode, dest list, source operand strings)"""
    pred = ""
    if txt.startswith("@"):
        pred, txt = txt.split(None, 1)
    parts = txt.split(None, 1)
    op = parts[0]
    ops = [o.strip() for o in parts[1].split(",")] if len(parts) > 1 else []
    return op, ops


def src_regs(op, ops):
    """List of (slot, reg, width64, reuse_flag) for register SOURCE operands."""
    base = op.split(".")[0]
    if base in ("BRA", "BAR", "NOP", "EXIT", "BSSY", "BSYNC", "WARPSYNC", "DEPBAR", "CALL", "RET"):
        return []
    wide = base in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")
    srcs = ops[1:]
    if base in ("STS", "STG", "ST", "ATOMS", "ATOMG", "RED", "STL"):
        srcs = ops  # address + data
    if base in ("ISETP", "FSETP", "DSETP", "PLOP3"):
        srcs = ops[2:]
    res = []
    for slot, o in enumerate(srcs):
        reuse = ".reuse" in o
        o2 = o.replace(".reuse", "")
        m = re.search(r"(?<![A-Za-z_])R(\d+)", o2)
        if not m or "UR" in o2 and not re.search(r"(?<!U)R\d+", o2):
            continue
        m = re.search(r"(?<!U)R(\d+)", o2)
        if not m:
            continue
        reg = int(m.group(1))
        w64 = wide or (".64" in op and "[" in o2 and False)
        if base in ("STS", "STG", "ST", "STL") and slot == len(srcs) - 1:
            w64 = ".64" in op or ".128" in op  # data operand
            nreg = 4 if ".128" in op else (2 if ".64" in op else 1)
            res.append((slot, reg, nreg, reuse))
            continue
        if base == "I2F" and ".F64" in op:
            res.append((slot, reg, 1, reuse))
            continue
        if base == "SHFL":
            res.append((slot, reg, 1, reuse))
            continue
        res.append((slot, reg, 2 if w64 else 1, reuse))
    return res


def cost_block(instrs):
    hist = Counter()
    total = 0.0
    reads = 0
    prev_reuse = {}  # slot -> reg kept by the previous instruction
    fp64 = 0
    three = 0
    for _, txt in instrs:
        op, ops = operands(txt)
        base = op.split(".")[0]
        hist[base] += 1
        srcs = src_regs(op, ops)
        fresh = 0
        nfresh_ops = 0
        for slot, reg, n, _ in srcs:
            if prev_reuse.get(slot) == reg:
                continue
            fresh += n
            nfresh_ops += 1
        prev_reuse = {slot: reg for slot, reg, n, ru in srcs if ru}
        reads += fresh
        if base in FP64_PIPE:
            fp64 += 1
            if nfresh_ops >= 3:
                three += 1
            total += max(FP64_PIPE[base], fresh / 2.0)
        elif base == "SHFL":
            total += 1.9
        elif base == "I2F":
            total += 0.1
        else:
            total += fresh / 2.0
    return hist, total, reads, fp64, three


def main():
    path = sys.argv[1]
    R = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else None
    ins = parse(path)
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, txt) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", txt)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    seen = []
    for lo, hi in sorted(loops, key=lambda t: t[1] - t[0]):
        body = ins[lo:hi + 1]
        nsh = sum(1 for _, t in body if t.startswith("SHFL.IDX") or " SHFL.IDX" in t)
        if nsh == 0 or any(lo <= l2 and h2 <= hi and (l2, h2) != (lo, hi) for l2, h2 in seen):
            continue
        seen.append((lo, hi))
        hist, total, reads, fp64, three = cost_block(body)
        steps = nsh // 2
        print(f"loop 0x{ins[lo][0]:x}..0x{ins[hi][0]:x}: {len(body)} instr, {steps} ring steps, fp64 {fp64} (3-fresh-operand: {three}), "
              f"fresh 32-bit reads {reads}, est {total:.1f} cycles")
        if R:
            pairs = steps * R
            print(f"   per pair: {len(body) / pairs:.2f} instr, {fp64 / pairs:.2f} fp64, {reads / pairs:.1f} reads, est {total / pairs:.2f} cycles"
                  f" (fp64 pipe floor {fp64 / pairs * 2.1:.1f})")
        print("   " + ", ".join(f"{k}:{v}" for k, v in hist.most_common(14)))


if __name__ == "__main__":
    main()

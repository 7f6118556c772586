"""Instruction mix of one kernel from an ncu report's SASS page, split at the CTA barriers (= the kernel's stages).

    ncu -i REPORT.ncu-rep --page source --csv --kernel-name k_prefilter > /tmp/k.csv
    python tools/sass_mix.py /tmp/k.csv SAMPLES        # SAMPLES = samples (pixels) the profiled launch processed

Prints thread-instructions per sample per stage and per opcode (executed counts, not static counts)."""
import csv
import sys
from collections import Counter


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    samples = float(sys.argv[2])
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    h = rows[hi]
    I, S = h.index("Instructions Executed"), h.index("Source")
    data = []
    for r in rows[hi + 1:]:
        if len(r) > I and r[I].isdigit():
            data.append((r[S].strip(), int(r[I])))
    tot = sum(c for _, c in data)
    print(f"total warp instructions {tot}  thread instructions / sample {tot * 32 / samples:.1f}")
    reg, cur = [], []
    for s, c in data:
        cur.append((s, c))
        if "BAR.SYNC" in s or "BAR.RED" in s:
            reg.append(cur); cur = []
    reg.append(cur)
    allops = Counter()
    for i, r in enumerate(reg):
        t = sum(c for _, c in r)
        ops = Counter()
        for s, c in r:
            parts = s.split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            ops[op.split(".")[0]] += c
        allops.update(ops)
        print(f"stage {i}: {len(r)} static, {t * 32 / samples:6.1f} / sample  ",
              " ".join(f"{k}={v * 32 / samples:.1f}" for k, v in ops.most_common(12)))
    print("all:", " ".join(f"{k}={v * 32 / samples:.1f}" for k, v in allops.most_common(24)))


if __name__ == "__main__":
    main()

"""Joins an `ncu --page source --csv --print-source sass` export with `nvdisasm -gi` line info of the same cubin and
aggregates executed warp-instructions per (file:line) at a chosen inline depth.  Development aid.

    python tools/ncu_by_line.py <source.csv> <nvdisasm -gi output> <kernel> [--depth innermost|walker|N] [--top 40]
"""
import csv, re, sys, collections, argparse

def parse_dis(path, kernel):
    stacks = {}
    cur = []
    pending = []
    inside = False
    for line in open(path, errors="replace"):
        s = line.strip()
        if s.endswith(":") and not s.startswith("//") and not s.startswith(".L"):
            inside = s[:-1] in (kernel, ".text." + kernel)
        if not inside:
            continue
        m = re.match(r'//## File "([^"]+)", line (\d+)', s)
        if m:
            pending.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r'/\*([0-9a-f]{4,})\*/', s)
        if m:
            if pending:
                # nvdisasm prints innermost frame first, outermost last; a new group replaces the stack
                cur = pending
                pending = []
            stacks[int(m.group(1), 16)] = list(cur)
    return stacks

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv"); ap.add_argument("dis"); ap.add_argument("kernel")
    ap.add_argument("--depth", default="innermost")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    stacks = parse_dis(a.dis, a.kernel)
    rows = []
    lines = open(a.csv, errors="replace").read().splitlines()
    # find the block of this kernel
    start = None
    for i, l in enumerate(lines):
        if l.startswith('"Kernel Name"') and ('"%s"' % a.kernel) in l:
            start = i + 1
            break
    end = len(lines)
    for i in range(start + 1, len(lines)):
        if lines[i].startswith('"Kernel Name"'):
            end = i
            break
    rd = csv.DictReader(lines[start:end])
    tot = 0
    agg = collections.Counter(); samp = collections.Counter(); thr = collections.Counter()
    ops = collections.Counter()
    for r in rd:
        try:
            addr = int(r["Address"], 16) if not r["Address"].isdigit() else int(r["Address"])
        except Exception:
            continue
        n = int(r["Instructions Executed"] or 0)
        t = int(r["Thread Instructions Executed"] or 0)
        s = int(r["# Samples"] or 0)
        rows.append((addr, n))
        tot += n
        ops[r["Source"].split()[0] if not r["Source"].startswith("@") else r["Source"].split()[1]] += n
    base = min(x[0] for x in rows)
    rd = csv.DictReader(lines[start:end])
    for r in rd:
        try:
            addr = int(r["Address"], 16) if not r["Address"].isdigit() else int(r["Address"])
        except Exception:
            continue
        n = int(r["Instructions Executed"] or 0)
        st = stacks.get(addr - base, [("?", 0)])
        if a.depth == "innermost":
            key = st[0]
        elif a.depth == "walker":
            key = next((f for f in st if f[0] == "rv_walker.cu"), st[-1])
        elif a.depth == "chain":
            key = tuple(st)
        else:
            d = int(a.depth); key = st[min(d, len(st) - 1)]
        agg[key] += n
        samp[key] += int(r["# Samples"] or 0)
        thr[key] += int(r["Thread Instructions Executed"] or 0)
    print("kernel %s: %d warp-instructions executed" % (a.kernel, tot))
    for key, n in agg.most_common(a.top):
        print("%6.2f%%  %10d  lanes %.1f  samples %6d  %s" % (100.0 * n / tot, n, thr[key] / max(n, 1), samp[key], key))
    print("-- by opcode")
    for k, n in ops.most_common(25):
        print("%6.2f%%  %s" % (100.0 * n / tot, k))

main()

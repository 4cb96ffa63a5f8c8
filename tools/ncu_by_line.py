"""Join an `ncu --page source --csv` export with nvdisasm line info: stall samples and executed
instructions per CUDA source line (development helper).

    python tools/ncu_by_line.py <source.csv> <nvdisasm -g -c output> <kernel section substring> [top]
"""
import csv, re, sys, collections

src_csv, disasm, section, top = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1. instruction index -> (file line, inlined-at chain) from nvdisasm
lines = open(disasm).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text.") and section in l)
cur = None
per_instr = []
for l in lines[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3))
        continue
    m2 = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m2:
        per_instr.append((cur, m2.group(1).split()[0] if not m2.group(1).startswith("@") else m2.group(1).split()[1]))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
data = rows[2:]
isamp, iex, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
print(f"ncu rows {len(data)}, kernel instructions with line info {len(per_instr)} (the rest are callee functions)")
bad = 0
for r, (li, op) in zip(data, per_instr):
    t = r[isrc].split()
    o = t[1] if t[0].startswith("@") else t[0]
    bad += o.split(".")[0] != op.split(".")[0]
assert bad == 0, f"{bad} opcode mismatches: listings are not aligned"
per_instr = [li for li, _ in per_instr] + [("callee", 0, "")] * (len(data) - len(per_instr))
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot_s = tot_e = 0
for r, li in zip(data, per_instr):
    key = (li[0], li[1]) if li else ("?", 0)
    a = agg[key]
    a[0] += int(r[isamp]); a[1] += int(r[iex]); a[2] += 1
    for i, h in stall_cols:
        if r[i] not in ('', '0'): a[3][h[6:]] += int(r[i])
    tot_s += int(r[isamp]); tot_e += int(r[iex])
src_cache = {}
def text(f, n):
    if f not in src_cache:
        try:
            src_cache[f] = open("cellmapper_b200/csrc/" + f).read().splitlines()
        except Exception:
            src_cache[f] = []
    s = src_cache[f]
    return s[n - 1].strip()[:90] if 0 < n <= len(s) else ""
print(f"total samples {tot_s}, executed warp-instructions {tot_e}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = " ".join(f"{n}={c * 100 // max(a[0], 1)}%" for n, c in a[3].most_common(3))
    print(f"{a[0] / tot_s * 100:5.1f}% samples {a[1] / tot_e * 100:5.1f}% instr  {key[0]}:{key[1]:<5d} {text(*key)[:60]:60s} | {st}")

"""Join an ncu SASS source page (csv) with nvdisasm -g line info: samples and stall reasons per CUDA source line.

    ncu -i rep --page source --csv > src.csv ; nvdisasm -g cubin > all.sass (cut to the kernel) ;
    python tools/ncu_by_line.py src.csv kernel.sass [source.cu]
"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows[:6]) if "Address" in r)
hdr = rows[h]
ai, si = hdr.index("Address"), hdr.index("# Samples")
stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
ie = hdr.index("Instructions Executed")
line_of, cur = {}, None
for ln in open(sys.argv[2]):
    m = re.search(r'//## File ".*", line (\d+)', ln)
    if m:
        cur = int(m.group(1)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
src = open(sys.argv[3]).read().splitlines() if len(sys.argv) > 3 else None
agg, tot = {}, 0
base = None
for r in rows[h + 1:]:
    try:
        addr = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
        s = int(r[si] or 0)
    except ValueError:
        continue
    if base is None:
        base = addr
    L = line_of.get(addr - base)
    a = agg.setdefault(L, [0, 0, {}])
    a[0] += s; a[1] += int(r[ie] or 0); tot += s
    for i in stall:
        v = int(r[i] or 0)
        if v:
            a[2][hdr[i][6:]] = a[2].get(hdr[i][6:], 0) + v
print("total samples", tot)
for L, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    text = src[L - 1].strip()[:80] if (src and L) else ""
    print(f"{s:6d} {100 * s / max(tot, 1):5.1f}%  inst {n:8d}  line {L}: {text}   {top}")

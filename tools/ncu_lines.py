"""Map the warp-stall samples of an ncu report (SASS page) to CUDA source lines via nvdisasm -g.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; nvdisasm -g kernel.cubin > dis.txt ;
       python tools/ncu_lines.py src.csv dis.txt <mangled kernel name prefix> <source file> [top]"""
import collections, csv, re, sys

src_csv, dis_txt, kname, cu_file = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or "Stall" in h]
sass = [(r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)) for r in data if len(r) > isamp]
txt = open(dis_txt).read().split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + kname)][0]
line, seq = None, []
for l in txt[start + 1:]:
    if (l.startswith(".text.") or l.startswith("//-----")) and seq:
        break
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        seq.append((line, m.group(2)))
n = min(len(sass), len(seq))
assert abs(len(sass) - len(seq)) < 4, (len(sass), len(seq))
per, perex = collections.Counter(), collections.Counter()
for i in range(n):
    per[seq[i][0]] += sass[i][1]
    perex[seq[i][0]] += sass[i][2]
tot = sum(per.values())
src = open(cu_file).read().split("\n")
print("total samples", tot)
lo, hi = (int(sys.argv[6]), int(sys.argv[7])) if len(sys.argv) > 7 else (0, 10**9)
shown = 0
for ln, c in per.most_common():
    if ln is None or not (ln[0].endswith(cu_file.split("/")[-1]) and lo <= ln[1] <= hi) and len(sys.argv) > 7:
        continue
    t = src[ln[1] - 1].strip()[:105] if ln and ln[0] == cu_file.split("/")[-1] else ""
    print(f"{100*c/tot:5.1f}%  ex={perex[ln]:>11d}  {ln[0] if ln else '?'}:{ln[1] if ln else 0}: {t}")
    shown += 1
    if shown >= top:
        break

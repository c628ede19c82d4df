"""Executed instructions / stall samples of the lean rollout kernel per phase (development aid).
usage: python tools/ncu_phases.py src.csv dis.txt <mangled kernel> <source file> line:name line:name ...
A SASS instruction belongs to the phase of the last line of <source file> seen before it (inlined helpers inherit)."""
import collections, csv, re, sys
src_csv, dis_txt, kname, cu = sys.argv[1:5]
bounds = sorted((int(a.split(":")[0]), a.split(":")[1]) for a in sys.argv[5:])
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
sass = [(r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)) for r in data if len(r) > isamp]
txt = open(dis_txt).read().split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + kname)][0]
cur, seq = None, []
for l in txt[start + 1:]:
    if (l.startswith(".text.") or l.startswith("//-----")) and seq:
        break
    m = re.search(r'//## File "(.*)", line (\d+)( inlined at "(.*)", line (\d+))?', l)
    if m:
        if m.group(1).endswith(cu.split("/")[-1]):
            cur = int(m.group(2))
        elif m.group(4) and m.group(4).endswith(cu.split("/")[-1]):
            cur = int(m.group(5))
        continue
    if re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l):
        seq.append(cur)
n = min(len(sass), len(seq))
ex, sa = collections.Counter(), collections.Counter()
def phase(line):
    name = "prologue"
    for b, nm in bounds:
        if line is not None and line >= b:
            name = nm
    return name
for i in range(n):
    ph = phase(seq[i])
    ex[ph] += sass[i][2]
    sa[ph] += sass[i][1]
te, ts = sum(ex.values()), sum(sa.values())
for b, nm in [(0, "prologue")] + bounds:
    print(f"{nm:28s} executed {ex[nm]:>13d} {100*ex[nm]/te:5.1f}%   samples {100*sa[nm]/ts:5.1f}%")

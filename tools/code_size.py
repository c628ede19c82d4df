"""Code bytes per phase of one rollout-kernel instantiation (instruction-cache budget: L1.5 holds 32 KB).
usage: nvdisasm -g X.cubin > dis.txt ; python tools/code_size.py dis.txt <mangled kernel prefix> <source.cu>"""
import collections, re, sys

dis, kname, cu = sys.argv[1:4]
src = open(cu).read().split("\n")
stamps = [(i + 1, re.search(r"PHASE_STAMP\((\d+)\)", l).group(1)) for i, l in enumerate(src)
          if "PHASE_STAMP(" in l and "#define" not in l]
txt = open(dis).read().split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + kname)][0]
cur, line, n = 0, None, 0
cnt = collections.Counter()
fname = cu.split("/")[-1]
for l in txt[start + 1:]:
    if (l.startswith(".text.") or l.startswith("//-----")) and n:
        break
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        if m.group(1).endswith(fname) and " inlined at " not in l:
            cur = int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+", l):
        n += 1
        nxt = [s for s in stamps if s[0] >= cur]
        cnt[f"up to line {nxt[0][0]} (stamp {nxt[0][1]})" if nxt else "tail"] += 1
print("total", n, "instructions", n * 16 // 1024, "KB")
for k, v in sorted(cnt.items(), key=lambda kv: int(re.search(r"\d+", kv[0]).group()) if kv[0] != "tail" else 10**9):
    print(f"  {k:32s} {v:6d} instr {v * 16 / 1024:6.1f} KB")

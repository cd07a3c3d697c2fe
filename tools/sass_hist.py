"""Per-source-line SASS instruction histogram of one kernel (needs -lineinfo)."""
import re, collections, subprocess, sys, os, tempfile
so = sys.argv[1]; pat = sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if re.match(r"\s*\.section\s+\.text\." + pat, l))
end = next(i for i in range(start + 1, len(txt)) if re.match(r"\s*\.section\s+\.text\.", txt[i]))
cur = None; cnt = collections.Counter()
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+[A-Z@]", l) and cur:
        cnt[cur] += 1
print("total", sum(cnt.values()))
byfile = collections.Counter()
for (f, l), c in cnt.items(): byfile[f] += c
print(byfile.most_common(8))
for f, _ in byfile.most_common(5):
    items = sorted(((l, c) for (ff, l), c in cnt.items() if ff == f), key=lambda x: -x[1])[:22]
    print(f, items)

"""Shared-memory wavefronts (actual / ideal / excessive) per source line of one kernel of an ncu report:
python scripts/ncu_smem.py report.ncu-rep object.o kernel_substring [top]"""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(tmp + "/*.cubin"), capture_output=True, text=True).stdout.splitlines()
start = max(i for i, l in enumerate(dis) if l.startswith("//--------------------- .text.") and kname in l)
cur, seq = None, []
for l in dis[start + 1:]:
    if l.startswith("//--------------------- ") and seq:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[1], rows[2:]
iW, iI, iX = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("L1 Wavefronts Shared Excessive")
assert len(seq) == len(data), (len(seq), len(data))
w, ideal, exc = collections.Counter(), collections.Counter(), collections.Counter()
for (cur, txt), r in zip(seq, data):
    w[cur] += int(r[iW] or 0); ideal[cur] += int(r[iI] or 0); exc[cur] += int(r[iX] or 0)
tot = sum(w.values())
print("shared wavefronts", tot, "ideal", sum(ideal.values()), "excessive", sum(exc.values()))
for key, c in exc.most_common(top):
    print(f"{100 * c / max(1, tot):5.1f}% excessive  ({w[key]:>10} actual, {ideal[key]:>10} ideal)  {key}")

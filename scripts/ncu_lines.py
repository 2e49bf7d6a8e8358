"""Per-source-line executed-instruction breakdown of one kernel of an ncu report (needs -lineinfo and the matching .o):
python scripts/ncu_lines.py report.ncu-rep object.o kernel_substring [top]"""
import collections, csv, glob, os, re, subprocess, sys, tempfile

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
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
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[1], rows[2:]
iE, iSm = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(seq) == len(data), (len(seq), len(data))
byline, samp, ops = collections.Counter(), collections.Counter(), collections.Counter()
for (cur, txt), r in zip(seq, data):
    e = int(r[iE] or 0)
    byline[cur] += e
    samp[cur] += int(r[iSm] or 0)
    op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
    ops[op] += e
tot = sum(byline.values())
print("total warp instructions", tot, " total samples", sum(samp.values()))
cache = {}
for key, c in byline.most_common(top):
    if key is None:
        print(f"{100*c/tot:5.1f}%  (no line)")
        continue
    f, ln = key
    if f not in cache:
        try:
            cache[f] = open(f).read().splitlines()
        except OSError:
            cache[f] = []
    t = cache[f][ln - 1].strip()[:100] if ln - 1 < len(cache[f]) else ""
    print(f"{100*c/tot:5.1f}% smp {100*samp[key]/max(1,sum(samp.values())):4.1f}% {os.path.basename(f)}:{ln}  {t}")
print("--- opcodes")
for op, c in ops.most_common(16):
    print(f"{op:24s} {100*c/tot:5.1f}%")

"""Stall reasons per source-line range of a kernel from an ncu report (development aid).
usage: python tools/ncu_stalls.py report.ncu-rep mangled_substring lo-hi [lo-hi ...]   (line ranges of tile_ws_kernel.cu)"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, sub = sys.argv[1], sys.argv[2]
ranges = [tuple(map(int, r.split("-"))) for r in sys.argv[3:]]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {root}/bionumpy_b200/_lib/libbnpk.so > /dev/null", shell=True)
dis = None
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if sub in out:
            dis = out
            break
lines = dis.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and sub in l)
inl = re.compile(r'//## File "([^"]+)", line (\d+)')
ins = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);')
cur, seq = None, []
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = inl.search(l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = ins.match(l)
    if m:
        seq.append(cur)
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 5]
assert len(seq) == len(data), (len(seq), len(data))
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
isamp, iinst = hdr.index("# Samples"), hdr.index("Instructions Executed")
for lo, hi in ranges:
    agg = collections.Counter(); ns = ni = 0
    for loc, r in zip(seq, data):
        if loc and loc[0] == "tile_ws_kernel.cu" and lo <= loc[1] <= hi:
            ns += int(r[isamp]); ni += int(r[iinst])
            for i, h in stall_cols:
                agg[h] += int(r[i])
    tot = sum(agg.values()) or 1
    print(f"lines {lo}-{hi}: {ni} warp-inst, {ns} samples: " + ", ".join(f"{h[6:]} {100 * c / tot:.0f}%" for h, c in agg.most_common(7)))

"""Per-source-line instruction/sample shares of a kernel from an ncu report + nvdisasm line info.
usage: python tools/ncu_lines.py report.ncu-rep mangled_name_substring [top_n]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, sub = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {root}/bionumpy_b200/_lib/libbnpk.so > /dev/null", shell=True)
dis = None
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if ".text." + "" in out and sub in out:
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
        seq.append((cur, m.group(1).strip()))
src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src_csv.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 5]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(seq) == len(data), (len(seq), len(data))
agg = collections.defaultdict(lambda: [0, 0])
tot = tots = 0
for (loc, _), r in zip(seq, data):
    c, s = int(r[ia]), int(r[isamp])
    tot += c; tots += s
    agg[loc][0] += c; agg[loc][1] += s
print(f"kernel {sub}: {tot} warp-instructions, {tots} samples, {len(seq)} SASS instructions")
src = {}
for f in ("tile_kernels.cu", "tile_tma_kernel.cu", "tile_ws_kernel.cu", "tile_ws_kernel.inl", "tile_common.cuh", "bnpk_device.cuh", "row_kernels.cu", "misc_kernels.cu"):
    src[f] = open(os.path.join(root, "bionumpy_b200", "csrc", f)).read().split("\n")
for loc, (c, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:top_n]:
    f, ln = loc if loc else ("?", 0)
    text = src[f][ln - 1].strip()[:88] if f in src and 0 < ln <= len(src[f]) else ""
    print(f"{c / tot * 100:5.1f}% inst {s / max(tots, 1) * 100:5.1f}% smp  {f}:{ln}  {text}")

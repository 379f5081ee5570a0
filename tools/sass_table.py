"""SASS opcode table of the product kernels in lib/libroi3d_b200.so (cuobjdump -sass), written as markdown.
usage: sass_table.py <out.md>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "3d-multi-resolution-rcnn_b200", "lib", "libroi3d_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangled = {}
names = re.findall(r"Function : (\S+)", sass)
if names:
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    demangled = dict(zip(names, out))
WATCH = ["FFMA2", "FMUL2", "FFMA", "LDS", "STS", "LDG", "STG", "LDGSTS", "UTMALDG", "UBLKCP", "UTMASTG", "SYNCS", "REDG", "RED",
         "ATOMS", "ATOMG", "ATOM", "MATCH", "HMMA", "UTCHMMA", "BAR", "MUFU"]
rows = []
cur, cnt, tot = None, None, 0
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt, tot))
        cur, cnt, tot = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        cnt[op] += 1
        tot += 1
if cur:
    rows.append((cur, cnt, tot))
KEEP = ("roi_align3d_fwd_stream_kernel", "roi_align3d_fwd_stream_ncdhw_kernel", "roi_align3d_bwd_stream_kernel", "roi_align3d_bwd_planar_kernel", "roi_align3d_plan_kernel", "roi_align3d_fwd_planar_kernel", "roi_align3d_bwd2_kernel",
        "nms3d_", "topk_sample_kernel", "topk_sieve_kernel", "topk_first_kernel", "topk_second_kernel", "topk_split_keys_kernel", "topk_tail_kernel",
        "decode_proposals_batched_kernel", "assign_pass", "mask_paste_kernel", "transpose_r32c128")
with open(sys.argv[1], "w") as f:
    f.write("# SASS opcode counts of the product kernels (static instruction counts, `cuobjdump -sass lib/libroi3d_b200.so`)\n\n")
    f.write("No tensor-core opcodes (HMMA / UTCHMMA) anywhere: the path is gather / scatter with a 1e-5 fp32 budget "
            "(DESIGN.md section 4).  `UTMALDG` = TMA tensor loads, `UBLKCP` = bulk copies, `SYNCS` = mbarrier ops, "
            "`FFMA2` / `FMUL2` = packed fp32.\n\n")
    f.write("| kernel | total | " + " | ".join(WATCH) + " |\n|---|---|" + "---|" * len(WATCH) + "\n")
    for name, cnt, tot in rows:
        d = demangled.get(name, name)
        if not any(k in d for k in KEEP):
            continue
        d = re.sub(r"\((int|bool|unsigned int)\)", "", d)
        short = re.sub(r"\(.*", "", d).replace("roi3d::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        f.write("| `%s` | %d | %s |\n" % (short[:90], tot, " | ".join(str(cnt.get(w, 0)) for w in WATCH)))
    allc = collections.Counter()
    for _, cnt, _ in rows:
        allc.update(cnt)
    f.write("\nWhole library: %d kernels, %d instructions; " % (len(rows), sum(allc.values())))
    f.write(", ".join("%s %d" % (w, allc.get(w, 0)) for w in WATCH) + ".\n")
print("wrote", sys.argv[1])

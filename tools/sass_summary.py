"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (cuobjdump -sass of libgsttaco.so).
usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gst_tacotron_b200", "libgsttaco.so")
sass = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "MUFU.TANH", "UTCBAR", "REDG", "RED."]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("gstk::", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for k in MN:
            if op.startswith(k):
                counts[cur][k] += 1
print("# cuobjdump -sass gst_tacotron_b200/libgsttaco.so (sm_100a): instruction counts per kernel / out-of-line device function")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, HMMA = mma.sync")
print("| function | " + " | ".join(MN) + " |")
print("|---|" + "---|" * len(MN))
for k, c in counts.items():
    if any(c[m] for m in MN if m not in ("FFMA",)) or c["FFMA"] > 200:
        print("| `{}` | ".format(k[:70]) + " | ".join(str(c[m]) if c[m] else "" for m in MN) + " |")

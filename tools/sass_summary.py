"""Developer tool: SASS mnemonic counts per kernel of the built library (cuobjdump -sass) -- the evidence for which hardware paths a
kernel uses (UTCHMMA = tcgen05.mma, LDTM/STTM = TMEM access, UTMALDG/UTMASTG/UBLKCP = TMA, SYNCS = mbarrier, HMMA = mma.sync,
FFMA2/FMUL2/FADD2 = packed fp32 pairs).  usage: sass_summary.py [lib.so] > profiles/<round>_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "ffcnn_b200", "libffcnn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA2", "FMUL2", "FADD2", "FFMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR"]
fn = None; cnt = collections.OrderedDict()
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: fn = m.group(1); cnt[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m and fn: cnt[fn][m.group(1)] += 1; cnt[fn]["_total"] += 1
names = subprocess.run(["cu++filt"] + list(cnt), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(lib, REPO))
print("# tcgen05 = UTCHMMA (+ LDTM/STTM TMEM access, UTCBAR commit); TMA = UTMALDG / UTMASTG (tensor) and UBLKCP (bulk); mbarrier = SYNCS;")
print("# mma.sync = HMMA (tf32); packed fp32 pairs = FFMA2 / FMUL2 / FADD2\n")
for name, (f, c) in zip(names, cnt.items()):
    name = re.sub(r"\((int|bool)\)", "", name).replace("void ", "")
    name = re.sub(r"\((?!anonymous).*", "", name)
    print("%-66s %5d instr  %s" % (name[:66], c["_total"], "  ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))

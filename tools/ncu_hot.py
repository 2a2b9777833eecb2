"""Developer helper: stall / opcode / hottest-line summary of one .ncu-rep (source page, SASS + CUDA-C correlation)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_s = sum(f(r, '# Samples') for r in data) or 1; tot_i = sum(f(r, 'Instructions Executed') for r in data) or 1
print('kernel', rows[0][1][:100]); print('samples %d  warp instructions %d  sass lines %d' % (tot_s, tot_i, len(data)))
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for h, v in sorted(((h, sum(f(r, h) for r in data)) for h in st), key=lambda x: -x[1])[:8]: print('  %-24s %5.1f%%' % (h, 100 * v / tot_s))
op = collections.Counter(); ops = collections.Counter()
for r in data:
    s = r[ix['Source']].strip().split()
    if not s: continue
    o = (s[0] if not s[0].startswith('@') else s[1]).split('.')[0]
    op[o] += f(r, 'Instructions Executed'); ops[o] += f(r, '# Samples')
print('opcode mix:', ', '.join('%s %.1f%%/%.1f%%s' % (o, 100 * v / tot_i, 100 * ops[o] / tot_s) for o, v in op.most_common(14)))
wf = sum(f(r, 'L1 Wavefronts Shared') for r in data); wfi = sum(f(r, 'L1 Wavefronts Shared Ideal') for r in data)
print('shared wavefronts %d (ideal %d)' % (wf, wfi))
print('hottest SASS by samples:')
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    tops = sorted(((h, f(r, h)) for h in st), key=lambda x: -x[1])[:2]
    print('  %5.1f%%  %-70s %s' % (100 * f(r, '# Samples') / tot_s, r[ix['Source']].strip()[:70], ' '.join('%s=%d' % (h[6:], v) for h, v in tops if v)))
ex = sorted(data, key=lambda r: -(f(r, 'L1 Wavefronts Shared') - f(r, 'L1 Wavefronts Shared Ideal')))[:int(__import__("os").environ.get("NCU_HOT_CONFLICTS", 6))]
if wf > wfi:
    print('shared-memory bank conflicts by SASS line (wavefronts, ideal):')
    for r in ex:
        if f(r, 'L1 Wavefronts Shared') > f(r, 'L1 Wavefronts Shared Ideal'):
            print('  %9d %9d  %s' % (f(r, 'L1 Wavefronts Shared'), f(r, 'L1 Wavefronts Shared Ideal'), r[ix['Source']].strip()[:80]))
# CUDA-C view
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h2 = None
lines = []
for r in rows:
    if r and r[0] == 'Line No' or (r and r[0] == '#'):
        h2 = r; continue
    if h2 and len(r) == len(h2): lines.append(r)
if h2 and '# Samples' in h2:
    si = h2.index('# Samples'); src = h2.index('Source')
    def g(r):
        try: return float(r[si])
        except Exception: return 0.0
    t2 = sum(g(r) for r in lines) or 1
    print('hottest source lines:')
    for r in sorted(lines, key=lambda r: -g(r))[:16]: print('  %5.1f%%  L%-4s %s' % (100 * g(r) / t2, r[0], r[src].strip()[:110]))

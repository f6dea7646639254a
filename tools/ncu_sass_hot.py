"""Per-kernel instruction mix and hottest SASS lines from `ncu --page source --csv --print-source sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = None; hdr = None; data = []
def flush():
    if not data: return
    print("=====", kern[:110])
    iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[iE]) for r in data); tots = sum(int(r[iW]) for r in data)
    mix = collections.Counter(); smp = collections.Counter()
    for r in data:
        op = r[iS].split()[0] if not r[iS].strip().startswith('@') else r[iS].split()[1]
        op = op.split('.')[0]
        mix[op] += int(r[iE]); smp[op] += int(r[iW])
    print("total warp-inst %.3e samples %d" % (tot, tots))
    for op, c in mix.most_common(18): print("   %-10s %5.1f%% inst  %5.1f%% samples" % (op, 100.0 * c / tot, 100.0 * smp[op] / max(tots, 1)))
    print(" hottest lines by stall samples:")
    for idx in sorted(range(len(data)), key=lambda i: -int(data[i][iW]))[:top]:
        r = data[idx]
        print("   #%4d %5.2f%% smp  %4.2f%% inst  %s" % (idx, 100.0 * int(r[iW]) / max(tots, 1), 100.0 * int(r[iE]) / tot, r[iS].strip()[:100]))
for r in rows:
    if r and r[0] == "Kernel Name":
        flush(); kern = r[1]; data = []; hdr = None
    elif r and r[0] == "Address": hdr = r
    elif r and hdr: data.append(r)
flush()

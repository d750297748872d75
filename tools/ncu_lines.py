"""Per-source-line stall samples and instruction counts from an `ncu --page source --csv --print-source cuda,sass` export:
python tools/ncu_lines.py export.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            out.append((cur_file, int(r[0]), r[1].strip()[:110], int(d['# Samples']), int(d['Instructions Executed']),
                        {k: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '0', '-')}))
        except (ValueError, KeyError):
            pass
tot_s = sum(o[3] for o in out) or 1
tot_i = sum(o[4] for o in out) or 1
print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
print("---- by samples")
for f, ln, src, smp, ins, st in sorted(out, key=lambda o: -o[3])[:top]:
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s   %s" % (100.0 * smp / tot_s, 100.0 * ins / tot_i, f, ln, src, st))
print("---- by instructions")
for f, ln, src, smp, ins, st in sorted(out, key=lambda o: -o[4])[:top]:
    print("%5.1f%% ins %5.1f%% smp  %s:%d  %s" % (100.0 * ins / tot_i, 100.0 * smp / tot_s, f, ln, src))

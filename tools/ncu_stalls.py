"""Summarise an `ncu --page source --csv --print-source sass` export: per kernel section, the stall mix
and the hottest instructions with the instruction before them (stalls are charged to the waiter).
python tools/ncu_stalls.py file.csv [section] [min_samples]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
s, e = starts[want], starts[want + 1]
print(rows[s][1][:100])
hdr = rows[s + 1]; idx = {h: i for i, h in enumerate(hdr)}
sec = [r for r in rows[s + 2:e] if len(r) > 10]
tot = sum(int(r[idx['# Samples']]) for r in sec)
print('total samples', tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s_: sum(int(r[idx[s_]]) for r in sec) for s_ in stalls}
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
for i, r in enumerate(sec):
    if int(r[idx['# Samples']]) > thr:
        st = {s_: int(r[idx[s_]]) for s_ in stalls if int(r[idx[s_]]) > 0}
        print(i, r[idx['# Samples']], '| prev:', sec[i - 1][1].strip()[:64], '| cur:', r[1].strip()[:50],
              sorted(st.items(), key=lambda x: -x[1])[:2])

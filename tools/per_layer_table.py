"""Turn an `ncu --metrics ... --csv` log of one subbox forward into profiles/<name>.csv + traffic.json."""
import csv, json, sys
from collections import OrderedDict
src, out_csv, out_traffic = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else None)
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
iid, ik, im, iv, iu = [hdr.index(k) for k in ('ID', 'Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit')]
L = OrderedDict()
for r in rows[1:]:
    L.setdefault(r[iid], {'kernel': r[ik]})[r[im]] = (r[iv], r[iu])
names = ['pack_input', 'conv_l00.conv_0', 'conv_l00.conv_1+skip', 'conv_l01.conv_0', 'conv_l01.conv_1+skip', 'down_l0',
         'conv_l1.conv_0', 'conv_l1.conv_1+skip', 'down_l1', 'conv_l2.conv_0', 'conv_l2.conv_1+skip', 'down_l2',
         'conv_c.conv_0', 'conv_c.conv_1+skip', 'up_r2', 'conv_r2.conv_0', 'conv_r2.conv_1+skip', 'up_r1', 'conv_r1.conv_0',
         'conv_r1.conv_1+skip', 'up_r0', 'conv_r00.conv_0', 'conv_r00.conv_1+skip', 'conv_r01.conv_0', 'conv_r01.conv_1+skip']
f = lambda v: float(v[0].replace(',', ''))
sc = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
ids = list(L)
off = next(j for j, i in enumerate(ids) if 'pack_input' in L[i]['kernel'])
traffic = {}
with open(out_csv, 'w') as fo:
    fo.write('launch,kernel,duration_ms,tensor_pipe_active_pct,dram_GB,dram_GBps,tma_l2_to_sm_GB,tma_TBps,l2_hit_pct\n')
    for n, i in zip(names, ids[off:]):
        m = L[i]
        t, tu = f(m['gpu__time_duration.sum']), m['gpu__time_duration.sum'][1]
        t_ms = t / 1e6 if tu == 'ns' else (t / 1e3 if tu == 'us' else t)
        b = lambda k: f(m[k]) * sc[m[k][1]]
        dr = b('dram__bytes_read.sum') + b('dram__bytes_write.sum')
        tma = b('l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum')
        tp = f(m['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'])
        hit = f(m['lts__t_sector_hit_rate.pct'])
        kern = m['kernel'].split('(')[0][:48]
        fo.write('%s,"%s",%.3f,%.1f,%.3f,%.0f,%.2f,%.2f,%.1f\n' % (n, kern, t_ms, tp, dr / 1e9, dr / t_ms / 1e6, tma / 1e9, tma / t_ms / 1e9, hit))
        print('%-22s %8.3f ms tensor %5.1f%% dram %6.2f GB %5.0f GB/s tma %6.2f GB %5.2f TB/s L2hit %4.1f%%' % (n, t_ms, tp, dr / 1e9, dr / t_ms / 1e6, tma / 1e9, tma / t_ms / 1e9, hit))
        traffic[n] = dr
if out_traffic:
    json.dump(traffic, open(out_traffic, 'w'), indent=1)

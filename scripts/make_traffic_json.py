"""profiles/r2_traffic.json from the ncu --set full reports of this round: DRAM bytes (read + write) per launch
of the kernels bench.py reports a roofline for.  python scripts/make_traffic_json.py gpurun_out"""
import csv, io, json, os, subprocess, sys

d = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out'
REPORTS = {'zgemm3m_lower_T2': 'r2_prof_zgemm3m_T2', 'zgemm3m_pfield': 'r2_prof_zgemm3m_pfield', 'pvec_tiles': 'r2_prof_pvec',
           'hmv_tstack': 'r2_prof_hmv_tstack', 'hmv_ypass': 'r2_prof_hmv_ypass', 'fir_conv': 'r2_prof_firconv'}
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
out = {}
for key, rep in REPORTS.items():
    path = os.path.join(d, rep + '.ncu-rep')
    if not os.path.exists(path):
        continue
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(name)
        tot += float(vals[i].replace(',', '')) * UNIT[units[i]]
    out[key] = tot
if 'hmv_tstack' in out and 'hmv_ypass' in out:
    out['hmv_batched'] = out['hmv_tstack'] + out['hmv_ypass']
out['_source'] = 'ncu --set full --clock-control none, one launch each (scripts/profile_r2.sh), dram__bytes_read.sum + dram__bytes_write.sum'
json.dump(out, open(os.path.join('profiles', 'r2_traffic.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))

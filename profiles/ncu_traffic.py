"""Turn an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum` launch list of `python bench.py` into profiles/k2_traffic.json:
the DRAM traffic of ONE k_sample_stream<double> launch (the bench's K2 workload), stamped with the hash of the kernel sources it
was captured from -- bench.py only reports `roofline.traffic` when that hash matches the tree it runs from.

    python profiles/ncu_traffic.py gpurun_out/k2_dram.csv 192000000 > profiles/k2_traffic.json
"""
import csv
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import kernel_source_hash  # noqa: E402

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith('==')]
hdr = rows[0]
ik, im, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
# per launch (ncu "ID" column): read + written bytes of every k_sample_stream<double> launch; the bench's K2 launch is the largest
iid = hdr.index('ID')
per = {}
for r in rows[1:]:
    if 'k_sample_stream' in r[ik] and 'double' in r[ik]:
        v = float(r[iv].replace(',', '')) * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[r[iu]]
        d = per.setdefault(r[iid], {'name': r[ik]})
        d['rd' if r[im] == 'dram__bytes_read.sum' else 'wr'] = v
best = max(per.values(), key=lambda d: d.get('rd', 0.0) + d.get('wr', 0.0))
rd, wr, name = best.get('rd'), best.get('wr'), best['name']
npts = int(sys.argv[2])
print(json.dumps({'kernel': name, 'points_per_launch': npts, 'dram_bytes_read': rd, 'dram_bytes_write': wr, 'algorithmic_bytes': npts * 40,
                  'kernel_source_hash': kernel_source_hash(),
                  'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over `python bench.py`: the largest '
                            'k_sample_stream<double> launch (192 M points); profiles/runs/r02ai_split.sh'}, indent=1))

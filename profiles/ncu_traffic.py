"""profiles/traffic.json from `ncu --set full` reports: DRAM bytes (read + write) per launch and
per scale row of the W-writing kernels, keyed by the hash of the CUDA sources the reports were
captured from.  bench.py reports `roofline.traffic` from this file only while that hash matches
the sources it runs (otherwise null).

    python profiles/ncu_traffic.py rows_per_launch=<n> report1.ncu-rep [report2.ncu-rep ...]
"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def read(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, row = r[0], r[1], r[2]
    return {h: (row[i], units[i]) for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    f = float(v)
    return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def main():
    out = {"source_hash": bench.source_hash(), "kernels": {}, "how": "ncu --set full --clock-control none, one launch each"}
    for p in sys.argv[1:]:
        d = read(p)
        name = d['Kernel Name'][0]
        name = name[name.index('k_run<') + 6:name.rindex('>')] if 'k_run<' in name else name
        name = name.replace('cwtb::', '')
        rd = to_bytes(*d['dram__bytes_read.sum'])
        wr = to_bytes(*d['dram__bytes_write.sum'])
        grid = d.get('launch__grid_dim_y', d.get('launch__grid_size'))
        rows = int(float(d['launch__grid_dim_y'][0])) if 'launch__grid_dim_y' in d else None
        out["kernels"][name] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                                "rows_per_launch": rows,
                                "dram_bytes_per_row": (rd + wr) / rows if rows else None,
                                "time_us": float(d['gpu__time_duration.sum'][0]), "report": os.path.basename(p)}
    with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()

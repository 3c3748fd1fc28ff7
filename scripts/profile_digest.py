#!/usr/bin/env python
"""Digest the artefacts of scripts/gpu_launchlist.sh / gpu_ncu_full.sh into the tracked profiles/ directory:
    python scripts/profile_digest.py TAG
writes profiles/<round>_launches_TAG.csv (copy), <round>_launch_shares_TAG.txt, <round>_ncu_summary_TAG.txt and updates
profiles/traffic.json (dram bytes per launch of every kernel in the full capture)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
RND = os.environ.get('SMPLFIT_ROUND', 'r02')  # file-name prefix of the round
out = os.path.join(ROOT, 'profiles')
src = os.path.join(ROOT, 'gpurun_out')

lc = os.path.join(src, 'launches.csv')
if os.path.exists(lc):
    shutil.copy(lc, os.path.join(out, f'{RND}_launches_{tag}.csv'))
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = {}
    for r in rows[1:]:
        name = r[ki].split('(')[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', '')) / 1000
    tot = sum(v for _, v in agg.values())
    with open(os.path.join(out, f'{RND}_launch_shares_{tag}.txt'), 'w') as f:
        f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none), bench.py --steps 2 --warmup 3, B=4096 SMPL\n')
        f.write('# per-launch times are cold-cache and serialised: use the SHARES, not the absolutes\n')
        f.write('# kernel, launches, total_us, share\n')
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{k[:72]:72s} {n:5d} {v:10.1f} {100 * v / tot:5.1f}%\n')

rep = os.path.join(src, 'prof.ncu-rep')
if os.path.exists(rep):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_summary.py'), rep], capture_output=True,
                         text=True).stdout
    with open(os.path.join(out, f'{RND}_ncu_summary_{tag}.txt'), 'w') as f:
        f.write(f'# ncu --set full --clock-control none --import-source on, bench workload (B=4096 SMPL); digest by scripts/ncu_summary.py\n')
        f.write(txt)
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    units = rows[1]
    tj = os.path.join(out, 'traffic.json')
    traffic = json.load(open(tj)) if os.path.exists(tj) else {}
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    seen = {}
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0].split('<')[0].split('::')[-1].replace('void ', '').strip()
        rd = float(r[idx['dram__bytes_read.sum']]) * scale[units[idx['dram__bytes_read.sum']]]
        wr = float(r[idx['dram__bytes_write.sum']]) * scale[units[idx['dram__bytes_write.sum']]]
        seen.setdefault(name, []).append(rd + wr)
    for name, vals in seen.items():  # several instantiations of one template (k_fit_fused<2>, <3>): their mean
        traffic[name] = int(sum(vals) / len(vals))
    traffic['_comment'] = ('dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured instantiations of '
                           'a template), from ncu --set full captures (profiles/*_ncu_summary_*.txt), bench workload '
                           '(B=4096 SMPL); k_vposed_tc = the pair-term GEMM of the closed-form Gramian')
    json.dump(traffic, open(tj, 'w'), indent=2, sort_keys=True)
print(open(os.path.join(out, 'traffic.json')).read())

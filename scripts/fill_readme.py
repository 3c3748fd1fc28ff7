#!/usr/bin/env python
"""Fill the @PLACEHOLDER@ numbers of README.md / INTEGRATION.md from a bench line (profiles/r02_bench_line_*.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.loads([ln for ln in open(sys.argv[1]).read().splitlines() if ln.startswith('{')][-1])
pt = (d.get('reference_pt_b200') or {}).get('script_fits_per_s') or 33460.0
rep = {
    '@RES@': f"{d['value'] / 1e6:.3f} M", '@RESMS@': f"{d['ms_per_step']:.2f}", '@VSB@': f"{d['value'] / 9481.0:.0f}",
    '@VSPT@': f"{d['value'] / pt:.0f}", '@E2E@': f"{d['e2e']['value'] / 1e3:.0f} K", '@E2EMS@': f"{d['e2e']['ms_per_step']:.2f}",
    '@LAT@': f"{d.get('latency_batch32_ms') or 0:.2f}", '@FWD@': f"{(d.get('lbs_forward') or {}).get('forwards_per_s', 0) / 1e6:.1f}",
}
for name in ('README.md', 'INTEGRATION.md'):
    p = os.path.join(ROOT, name)
    s = open(p).read()
    for k, v in rep.items():
        s = s.replace(k, v)
    open(p, 'w').write(s)
print(rep)

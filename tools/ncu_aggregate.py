"""Aggregates an `ncu --csv` launch list (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum)
of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` into per-kernel-family totals of ONE step and writes
profiles/step_dram_traffic.json (read by bench.py for `roofline.traffic`) plus a markdown table.

    python tools/ncu_aggregate.py gpurun_out/traffic_r2.csv [--step-marker mask_draw_kernel] [--tag r2]

One step = the launches between the last two occurrences of the marker kernel on the main path (the mask generator of
the NEXT step's prefetch is launched once per step)."""
import csv
import io
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# algorithmic wgrad bytes per step at B=64 @512^2: every layer input and every dy read once (bf16), SURVEY App. A shapes
K2_ALGO_GB = 55.0


def family(name):
    name = re.sub(r'^(void\s+)?(cmu::)?', '', name)
    name = re.sub(r'\(.*$', '', name)
    m = re.match(r'([A-Za-z0-9_:]+(<[^>]*>)?)', name)
    base = m.group(1) if m else name
    if base.startswith('at::') or base.startswith('void at::'):
        base = base[:40]
    return base


def main():
    path = sys.argv[1]
    marker = 'mask_draw_kernel'
    tag = 'r2'
    for i, a in enumerate(sys.argv):
        if a == '--step-marker':
            marker = sys.argv[i + 1]
        if a == '--tag':
            tag = sys.argv[i + 1]
    text = open(path, errors='replace').read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    launches = {}
    order = []
    for r in rows:
        i = int(r['ID'])
        if i not in launches:
            launches[i] = {'name': r['Kernel Name'], 'ms': 0.0, 'rd': 0.0, 'wr': 0.0}
            order.append(i)
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        mn = r['Metric Name']
        if mn.startswith('gpu__time_duration'):
            launches[i]['ms'] = v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
        elif mn.startswith('dram__bytes_read'):
            launches[i]['rd'] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
        elif mn.startswith('dram__bytes_write'):
            launches[i]['wr'] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    marks = [i for i in order if marker in launches[i]['name']]
    # the prefetch launches the marker twice per step (online half, target half): take the last full step
    if len(marks) >= 4:
        lo, hi = marks[-4], marks[-2]
    elif len(marks) >= 2:
        lo, hi = marks[-2], marks[-1]
    else:
        lo, hi = order[0], order[-1] + 1
    step = [launches[i] for i in order if lo <= i < hi]
    fam = {}
    for l in step:
        f = fam.setdefault(family(l['name']), {'launches': 0, 'ms': 0.0, 'dram_read_GB': 0.0, 'dram_write_GB': 0.0})
        f['launches'] += 1
        f['ms'] += l['ms']
        f['dram_read_GB'] += l['rd'] / 1e9
        f['dram_write_GB'] += l['wr'] / 1e9
    for f in fam.values():
        for k in ('ms', 'dram_read_GB', 'dram_write_GB'):
            f[k] = round(f[k], 4)
    try:
        commit = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    except Exception:
        commit = '?'
    tot_rd = sum(f['dram_read_GB'] for f in fam.values())
    tot_wr = sum(f['dram_write_GB'] for f in fam.values())
    out = {'meta': {'source': os.path.basename(path), 'aggregated': time.strftime('%Y-%m-%d'), 'commit_at_aggregation': commit,
                    'command': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none '
                               'python bench.py --steps 1 --warmup 3 --no-cpu-baseline', 'launches_in_step': len(step)},
           'step_total_GB': round(tot_rd + tot_wr, 2), 'step_read_GB': round(tot_rd, 2), 'step_write_GB': round(tot_wr, 2),
           'step_ms_serialised': round(sum(f['ms'] for f in fam.values()), 2),
           'k2_algorithmic_GB_per_step': K2_ALGO_GB, 'families': dict(sorted(fam.items(), key=lambda kv: -kv[1]['ms']))}
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'step_dram_traffic.json'), 'w'), indent=1)
    json.dump(out, open(os.path.join(ROOT, 'profiles', f'{tag}_step_dram_traffic.json'), 'w'), indent=1)
    L = [f'# {tag} — per-kernel time and DRAM traffic of one step (B=64, S=512, 1xB200)', '',
         f'Source: `{out["meta"]["command"]}` (after the same command exited 0 without ncu); serialised, cold-cache: compare',
         f'shares.  {len(step)} launches, {out["step_ms_serialised"]} ms serialised, {out["step_read_GB"]} GB read + '
         f'{out["step_write_GB"]} GB written = {out["step_total_GB"]} GB.', '',
         '| kernel family | launches | ms | share | DRAM read GB | DRAM write GB | GB/s |', '|---|---|---|---|---|---|---|']
    tot = out['step_ms_serialised']
    for k, f in out['families'].items():
        gbs = (f['dram_read_GB'] + f['dram_write_GB']) / (f['ms'] / 1e3) if f['ms'] > 0 else 0
        L.append(f'| `{k}` | {f["launches"]} | {f["ms"]:.3f} | {100 * f["ms"] / tot:.1f} % | {f["dram_read_GB"]:.2f} | '
                 f'{f["dram_write_GB"]:.2f} | {gbs:.0f} |')
    open(os.path.join(ROOT, 'profiles', f'{tag}_step_breakdown.md'), 'w').write('\n'.join(L) + '\n')
    print(json.dumps({k: out[k] for k in ('step_total_GB', 'step_ms_serialised')}), len(step), 'launches')


if __name__ == '__main__':
    main()

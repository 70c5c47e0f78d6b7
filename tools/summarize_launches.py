"""dev: per-kernel summary of an ncu launch list (csv of --metrics gpu__time_duration.sum,dram__bytes_*,launch__registers_per_thread)."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
per = {}
for r in rows:
    kid, name, metric, unit, val = r[0], r[4], r[12], r[13], float(r[14].replace(',', ''))
    d = per.setdefault(kid, {'name': name})
    if metric == 'gpu__time_duration.sum':
        d['us'] = val / 1e3 if unit in ('nsecond', 'ns') else (val if unit in ('usecond', 'us') else val * 1e3)
    elif metric.startswith('dram__bytes_read'):
        d['rd'] = val * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    elif metric.startswith('dram__bytes_write'):
        d['wr'] = val * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    elif metric.startswith('launch__registers'):
        d['regs'] = int(val)
agg = collections.OrderedDict()
for d in per.values():
    n = re.sub(r'\(.*', '', d['name']); n = re.sub(r'^void ', '', n); n = n.replace('smpc::<unnamed>::', '').replace('smpc::', '')
    a = agg.setdefault(n, {'n': 0, 'us': 0.0, 'max': 0.0, 'rd': 0.0, 'wr': 0.0, 'regs': d.get('regs', 0)})
    a['n'] += 1; a['us'] += d.get('us', 0); a['max'] = max(a['max'], d.get('us', 0)); a['rd'] += d.get('rd', 0); a['wr'] += d.get('wr', 0)
tot = sum(a['us'] for a in agg.values())
print(f'total {tot / 1e3:.3f} ms over {sum(a["n"] for a in agg.values())} launches')
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    print(f'{n[:44]:44s} n={a["n"]:4d} tot={a["us"] / 1e3:8.2f}ms avg={a["us"] / a["n"]:8.1f}us max={a["max"]:8.1f} share={100 * a["us"] / tot:5.1f}% '
          f'rd={a["rd"] / 1e6:9.1f}MB wr={a["wr"] / 1e6:9.1f}MB regs={a["regs"]}')

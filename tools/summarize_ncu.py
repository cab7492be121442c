"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/."""
import csv, collections, subprocess, sys

def launches(src, dst):
    rows = list(csv.reader(open(src)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
        a = agg.setdefault(r[ki].split('(')[0][:90], [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    with open(dst, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n')
        f.write('# source: %s ; %d launches, %.1f us total\n' % (src, len(rows) - start, tot))
        f.write('%-92s %6s %12s %7s\n' % ('kernel', 'n', 'total_us', 'share'))
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('%-92s %6d %12.1f %6.1f%%\n' % (k, n, t, 100 * t / tot))

def raw(rep, dst, pattern=''):
    if rep.endswith('.csv'):                       # `ncu -i x.ncu-rep --page raw --csv` already exported on the GPU box
        out = open(rep).read()
    else:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(dst, 'w') as f:
        f.write('# ncu --set full --clock-control none, from %s\n' % rep)
        f.write(','.join('%s[%s]' % (w, units[i]) for w, i in idx) + '\n')
        for r in rows[2:]:
            if pattern in r[hdr.index('Kernel Name')]:
                f.write(','.join(r[i].split('(CUt')[0] if w == 'Kernel Name' else r[i] for w, i in idx) + '\n')

def traffic(rep, meta_json, dst, pattern='conv_igemm'):
    """Per-launch DRAM traffic of the dominant kernel over the launches of ONE step (an `ncu --set full` capture of
    tools/ncu_one_step.py) -> the JSON bench.py reads for `roofline.traffic`."""
    import json
    if rep.endswith('.csv'):                       # `ncu -i x.ncu-rep --page raw --csv` already exported on the GPU box
        out = open(rep).read()
    else:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    ki, ri, wi, ti = (hdr.index(k) for k in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum'))
    pi = hdr.index('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' in hdr else None
    units = rows[1]
    def scale(u):
        u = u.lower()
        return {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1.0)
    n, rd, wr, t_us, pipe_t = 0, 0.0, 0.0, 0.0, 0.0
    for r in rows[2:]:
        if pattern not in r[ki]:
            continue
        n += 1
        rd += float(r[ri].replace(',', '')) * scale(units[ri]); wr += float(r[wi].replace(',', '')) * scale(units[wi])
        dur = float(r[ti].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(units[ti], 1.0)
        t_us += dur
        if pi is not None:
            pipe_t += float(r[pi].replace(',', '')) * dur
    meta = json.load(open(meta_json))
    res = {'workload': meta['workload'], 'precision': meta['precision'], 'kernel': pattern, 'launches': n,
           'dram_bytes_per_step': rd + wr, 'dram_bytes_per_launch': (rd + wr) / max(n, 1), 'dram_read_bytes_per_step': rd,
           'dram_write_bytes_per_step': wr, 'algorithmic_bytes_per_step': meta['conv_algorithmic_bytes_per_step'],
           'ncu_duration_us_per_step': t_us, 'tensor_pipe_active_pct_time_weighted': pipe_t / t_us if t_us else None,
           'source': 'ncu --set full --clock-control none over the %d %s launches of one %s step (%s); per-launch = per-step / launches'
                     % (n, pattern, meta['workload'], rep.split('/')[-1])}
    json.dump(res, open(dst, 'w'), indent=1)
    print(json.dumps(res))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == 'traffic':
        traffic(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else 'conv_igemm')
    else:
        raw(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')

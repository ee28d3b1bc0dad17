"""Turns the raw ncu output in gpurun_out/ into the committed summaries under profiles/:
   launches_r1_batch.csv   per-kernel totals of ONE measured 107-fold batch (launch list)
   ncu_r1_<kernel>.txt     key counters of the `--set full` capture of each main kernel
Run here (no GPU): python profiles/summarize.py"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
SRC = os.path.join(ROOT, 'gpurun_out')

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers']


def launches():
    path = os.path.join(SRC, 'launches_r1.csv')
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    recs = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        v = float(d['Metric Value'].replace(',', ''))
        u = d['Metric Unit']
        v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v * 1e3 if u == 's' else v
        name = d['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        recs.append((name, d['Grid Size'], d['Block Size'], v))
    # the measured batch starts at the last class-mean launch over 107 folds
    idx = [i for i, r in enumerate(recs) if r[0] == 'k_class_mean' and ', 107)' in r[1]]
    batch = recs[idx[-1]:]
    agg = collections.OrderedDict()
    for name, grid, block, ms in batch:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, 'launches_r1_batch.csv'), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none, python '
                'profiles/profile_step.py: ONE measured batch of 107 folds (8 patients, MCCA), '
                'cold-cache serialised launch times -> compare shares, not absolutes\n')
        f.write('kernel,launches,total_ms,share_pct,example_grid,block\n')
        for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('%s,%d,%.3f,%.1f,"%s","%s"\n' % (name, a[0], a[1], 100 * a[1] / tot, a[2], a[3]))
        f.write('TOTAL,%d,%.3f,100.0,,\n' % (sum(a[0] for a in agg.values()), tot))
    print('batch total %.2f ms over %d launches' % (tot, sum(a[0] for a in agg.values())))


def full():
    for rep in sorted(glob.glob(os.path.join(SRC, 'prof_r1_*.ncu-rep'))):
        k = os.path.basename(rep)[len('prof_r1_'):-len('.ncu-rep')]
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        with open(os.path.join(OUT, 'ncu_r1_%s.txt' % k), 'w') as f:
            f.write('# ncu --set full --clock-control none --import-source on -k regex:%s  '
                    '(python profiles/profile_step.py, 107-fold batch)\n' % k)
            for vals in rows[2:]:
                d = dict(zip(hdr, vals))
                f.write('\n%s   grid %s block %s\n' % (d.get('Kernel Name', '?').split('(')[0],
                                                      d.get('Grid Size'), d.get('Block Size')))
                for w in WANT:
                    if w in d:
                        f.write('  %-66s %s %s\n' % (w, d[w], units[hdr.index(w)]))
        print('wrote ncu_r1_%s.txt' % k)


if __name__ == '__main__':
    launches()
    full()

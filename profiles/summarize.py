"""Turns the raw ncu output in gpurun_out/ into the committed summaries under profiles/:
   launches_<round>_batch.csv   per-kernel totals of ONE measured batch (launch list)
   ncu_<round>_<kernel>.txt     key counters of the `--set full` capture of each main kernel
   traffic.json                 dram read + write bytes per launch of the captured kernels and the
                                folds of that launch (bench.py's `roofline.traffic` reads it)
   sass_summary.txt             tcgen05 / TMA / TMEM instruction counts per object (cuobjdump -sass)
Run here (no GPU): python profiles/summarize.py [r2] [folds]"""
import json
import re
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


ROUND = sys.argv[1] if len(sys.argv) > 1 else 'r2'
FOLDS = int(sys.argv[2]) if len(sys.argv) > 2 else 138


def launches():
    path = os.path.join(SRC, 'launches_%s.csv' % ROUND)
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
        name = re.sub(r'\(anonymous namespace\)::', '', name)
        recs.append((name, d['Grid Size'], d['Block Size'], v))
    # the measured batch starts at the last class-mean launch over FOLDS folds
    idx = [i for i, r in enumerate(recs) if r[0] == 'k_class_mean' and ', %d)' % FOLDS in r[1]]
    batch = recs[idx[-1]:]
    agg = collections.OrderedDict()
    for name, grid, block, ms in batch:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, 'launches_%s_batch.csv' % ROUND), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none, python '
                'profiles/profile_step.py --folds %d: ONE measured batch of %d folds (8 patients, MCCA), '
                'cold-cache serialised launch times -> compare shares, not absolutes\n' % (FOLDS, FOLDS))
        f.write('kernel,launches,total_ms,share_pct,example_grid,block\n')
        for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('%s,%d,%.3f,%.1f,"%s","%s"\n' % (name, a[0], a[1], 100 * a[1] / tot, a[2], a[3]))
        f.write('TOTAL,%d,%.3f,100.0,,\n' % (sum(a[0] for a in agg.values()), tot))
    print('batch total %.2f ms over %d launches' % (tot, sum(a[0] for a in agg.values())))


def _num(v, unit):
    v = float(str(v).replace(',', ''))
    return v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1.0)


def full():
    traffic = {}
    pre = 'prof_%s_' % ROUND
    for rep in sorted(glob.glob(os.path.join(SRC, pre + '*.ncu-rep'))):
        k = os.path.basename(rep)[len(pre):-len('.ncu-rep')]
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        with open(os.path.join(OUT, 'ncu_%s_%s.txt' % (ROUND, k)), 'w') as f:
            f.write('# ncu --set full --clock-control none --import-source on --kernel-name-base demangled '
                    '-k regex:<%s>  (python profiles/profile_step.py --folds %d)\n' % (k, FOLDS))
            for vals in rows[2:]:
                d = dict(zip(hdr, vals))
                f.write('\n%s   grid %s block %s\n' % (d.get('Kernel Name', '?').split('(')[0],
                                                      d.get('Grid Size'), d.get('Block Size')))
                for w in WANT:
                    if w in d:
                        f.write('  %-66s %s %s\n' % (w, d[w], units[hdr.index(w)]))
                if 'dram__bytes_read.sum' in d:
                    rd = _num(d['dram__bytes_read.sum'], units[hdr.index('dram__bytes_read.sum')])
                    wr = _num(d['dram__bytes_write.sum'], units[hdr.index('dram__bytes_write.sum')])
                    dur = d.get('gpu__time_duration.sum')
                    traffic[d.get('Kernel Name', k).split('(')[0].split('::')[-1].split('<')[0] if False else
                            ('k_' + k if not k.startswith('k_') else k)] = {
                        'dram_bytes': rd + wr, 'dram_read': rd, 'dram_write': wr, 'folds': FOLDS,
                        'duration': '%s %s' % (dur, units[hdr.index('gpu__time_duration.sum')]) if dur else None,
                        'source': 'profiles/ncu_%s_%s.txt' % (ROUND, k)}
        print('wrote ncu_%s_%s.txt' % (ROUND, k))
    if traffic:
        json.dump(traffic, open(os.path.join(OUT, 'traffic.json'), 'w'), indent=1, sort_keys=True)
        print('wrote traffic.json (%d kernels)' % len(traffic))


def sass():
    """tcgen05 / TMA / TMEM instruction counts of every object the library is linked from."""
    csrc = os.path.join(ROOT, 'cross_patient_speech_decoding_b200', 'csrc')
    pats = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS', 'HMMA', 'DFMA', 'DMUL']
    with open(os.path.join(OUT, 'sass_summary.txt'), 'w') as f:
        f.write('# cuobjdump -sass <object> | grep -c <mnemonic>, per object of libcpsd_b200.so (sm_100a).\n'
                '# UTCHMMA = tcgen05.mma (kind::tf32 / f16), UTMALDG = TMA tensor load, LDTM = tcgen05.ld '
                '(TMEM -> registers), UTCBAR = tcgen05.commit, DFMA = fp64 FMA.\n')
        f.write('%-14s' % 'object' + ''.join('%11s' % p for p in pats) + '\n')
        for o in sorted(glob.glob(os.path.join(csrc, '*.o'))):
            txt = subprocess.run(['cuobjdump', '-sass', o], capture_output=True, text=True).stdout
            f.write('%-14s' % os.path.basename(o) + ''.join('%11d' % len(re.findall(r'\b%s' % p, txt)) for p in pats) + '\n')
            # per kernel for the tensor-core objects
            if 'UTCHMMA' in txt:
                for m in re.finditer(r'Function : (\S+)(.*?)(?=Function : |\Z)', txt, flags=re.S):
                    n_mma = len(re.findall(r'\bUTCHMMA', m.group(2)))
                    if n_mma:
                        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
                        name = name.replace('(anonymous namespace)::', '')
                        f.write('    %-70s UTCHMMA %3d  UTMALDG %3d  LDTM %3d\n' % (
                            name.split('(')[0][-70:], n_mma, len(re.findall(r'\bUTMALDG', m.group(2))),
                            len(re.findall(r'\bLDTM', m.group(2)))))
    print('wrote sass_summary.txt')


if __name__ == '__main__':
    if os.path.exists(os.path.join(SRC, 'launches_%s.csv' % ROUND)):
        launches()
    full()
    sass()

#!/usr/bin/env python
"""Writes / updates profiles/traffic.json: DRAM bytes per algorithmic unit of a kernel, from an `ncu --set full`
report and the bench line of the run that was captured (tools/gpu_ncu_one.sh keeps both: gpurun_out/<tag>.ncu-rep and
gpurun_out/ncu_<tag>.log).  bench.py multiplies the figure by the units of its own launch for `roofline.traffic`.

    python tools/ncu_traffic.py <kernel key> <report.ncu-rep> <bench log with the JSON line> <roofline kernel name>
    e.g. python tools/ncu_traffic.py viterbi_profile_q gpurun_out/vq.ncu-rep gpurun_out/ncu_vq.log viterbi_count
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    key, rep, log, rk = sys.argv[1:5]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def metric(name):
        i = hdr.index(name)
        v = float(vals[i].replace(',', ''))
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[units[i]]
        return v * scale

    traffic = metric('dram__bytes_read.sum') + metric('dram__bytes_write.sum')
    line = None
    for l in open(log):
        if l.startswith('{') and '"roofline_kernels"' in l:
            line = json.loads(l)
    per_launch = line['roofline_kernels'][rk]['units_per_launch']
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    data = json.load(open(path)) if os.path.exists(path) else {}
    commit = subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    data[key] = {'bytes_per_unit': traffic / per_launch, 'unit': 'DP cell' if rk == 'align_scan' else 'Viterbi column',
                 'dram_bytes_of_the_captured_launch': traffic, 'units_of_the_captured_launch': per_launch,
                 'kernel': vals[hdr.index('Kernel Name')][:80], 'source': 'profiles/ncu_' + os.path.basename(rep).replace('.ncu-rep', '.txt'),
                 'commit': commit}
    json.dump(data, open(path, 'w'), indent=1)
    print(key, data[key])


if __name__ == '__main__':
    main()

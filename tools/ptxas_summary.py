#!/usr/bin/env python
"""Registers / spills / shared memory per kernel from the -Xptxas -v build logs (build/obj/*.log),
kernel names demangled. Usage: python tools/ptxas_summary.py [regex]"""
import glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
rows = []
for log in sorted(glob.glob(os.path.join(ROOT, 'build', 'obj', '*.log'))):
    name = None
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
        m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
        if m and name:
            spill = m.groups()
        m = re.search(r'Used (\d+) registers', line)
        if m and name:
            rows.append((name, int(m.group(1)), spill))
            name = None
names = subprocess.run(['c++filt'], input='\n'.join(r[0] for r in rows), capture_output=True, text=True).stdout.split('\n')
for (mangled, regs, spill), nm in zip(rows, names):
    nm = re.sub(r'\(int\)', '', nm.split('(acq::DevPlan')[0]).replace('acq::', '').replace('void ', '')
    if pat and not pat.search(nm):
        continue
    print('%4d regs  stack %s spill st/ld %s/%s  %s' % (regs, spill[0], spill[1], spill[2], nm))

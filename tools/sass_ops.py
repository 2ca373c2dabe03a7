#!/usr/bin/env python
"""Static opcode histogram per kernel from `cuobjdump -sass` of an object / library.
Usage: python tools/sass_ops.py <file> [regex on the demangled kernel name]"""
import collections, re, subprocess, sys
src = subprocess.run(['cuobjdump', '-sass', sys.argv[1]], capture_output=True, text=True).stdout
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
kern, ops = None, collections.OrderedDict()
for line in src.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        kern = m.group(1); ops[kern] = collections.Counter(); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and kern:
        ops[kern][m.group(2)] += 1
names = subprocess.run(['c++filt'], input='\n'.join(ops), capture_output=True, text=True).stdout.split('\n')
for (k, c), nm in zip(ops.items(), names):
    nm = re.sub(r'\(int\)', '', nm.split('(acq::DevPlan')[0]).replace('acq::', '').replace('void ', '')
    if pat and not pat.search(nm):
        continue
    tot = sum(c.values())
    print('%6d  %s' % (tot, nm))
    print('        ' + ', '.join('%s %d' % kv for kv in c.most_common(24)))

"""Run the REFERENCE acquire-*.py scripts (unmodified, from /root/reference) on the seeded
recordings of tests/synth_files.py and store their stdout as tests/golden/cli_golden.json.
Build container only:  python tests/golden/make_cli_golden.py"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200'), os.path.join(ROOT, 'tests')]

import synth_files  # noqa: E402

REF = os.environ.get('GNSS_REFERENCE', '/root/reference')
out = {}
for case in synth_files.CLI_CASES:
    with tempfile.NamedTemporaryFile(suffix='.iq', delete=False) as f:
        f.write(synth_files.recording(case))
        path = f.name
    script, args = synth_files.command(case, path)
    r = subprocess.run([sys.executable, '-W', 'ignore', os.path.join(REF, 'acquire-%s.py' % script)] + args,
                       capture_output=True, text=True, cwd=REF)
    os.unlink(path)
    assert r.returncode == 0, r.stderr[-2000:]
    out[case] = r.stdout.splitlines()
    print(case, out[case], flush=True)
json.dump(out, open(os.path.join(HERE, 'cli_golden.json'), 'w'), indent=1)

# serial long-code acquisitions (acquire-gps-l2cl.py, acquire-glonass-l{1,2}-p.py)
out = {}
for case in synth_files.SERIAL_CASES:
    with tempfile.NamedTemporaryFile(suffix='.iq', delete=False) as f:
        f.write(synth_files.recording_serial(case))
        path = f.name
    script, args = synth_files.command_serial(case, path)
    r = subprocess.run([sys.executable, '-W', 'ignore', os.path.join(REF, 'acquire-%s.py' % script)] + args,
                       capture_output=True, text=True, cwd=REF)
    os.unlink(path)
    assert r.returncode == 0, r.stderr[-2000:]
    out[case] = r.stdout.splitlines()
    print(case, out[case], flush=True)
json.dump(out, open(os.path.join(HERE, 'cli_serial_golden.json'), 'w'), indent=1)

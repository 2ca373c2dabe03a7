"""acquire-*.py command lines end to end against the stdout of the reference scripts
(tests/golden/cli_golden.json, produced by tests/golden/make_cli_golden.py)."""
import io
import json
import os
import re
import subprocess
import sys

import pytest

import synth_files

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(HERE), 'gnss-dsp-tools_b200')
GOLD = json.load(open(os.path.join(HERE, 'golden', 'cli_golden.json')))
LINE = re.compile(r'^(prn|chan)\s*(-?\d+) doppler\s*(-?[\d.]+) metric\s*(-?[\d.]+) code_offset\s*(-?[\d.]+)$')


def compare(got_lines, want_lines):
    assert len(got_lines) == len(want_lines)
    for g, w in zip(got_lines, want_lines):
        mg, mw = LINE.match(g), LINE.match(w)
        assert mg and mw, (g, w)
        # key, Doppler and code offset are exact (integer lag / bin parity); same column layout
        assert mg.group(1, 2, 3, 5) == mw.group(1, 2, 3, 5), (g, w)
        assert len(g) == len(w)
        gm, wm = float(mg.group(4)), float(mw.group(4))
        digits = len(mw.group(4).split('.')[1])
        assert abs(gm - wm) <= 1e-4 * abs(wm) + 1.01 * 10 ** (-digits), (g, w)   # 1e-4 relative + print rounding


@pytest.fixture
def recording(tmp_path):
    def make(case):
        p = tmp_path / (case + '.iq')
        p.write_bytes(synth_files.recording(case))
        return str(p)
    return make


def test_config1_cli_on_emulated_kernels(recording):
    """BASELINE config 1 through the full script path, kernels compiled for the host (test harness)."""
    import emu_util
    from gnsstools import acquire_cli
    eng = emu_util.emu_engine()
    script, args = synth_files.command('config1-gps-l1', recording('config1-gps-l1'))
    out = io.StringIO()
    acquire_cli.main(script, args, out=out, engine=eng)
    compare(out.getvalue().splitlines(), GOLD['config1-gps-l1'])
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize('case', sorted(synth_files.CLI_CASES))
def test_scripts_match_reference_stdout(case, recording):
    script, args = synth_files.command(case, recording(case))
    r = subprocess.run([sys.executable, os.path.join(PKG, 'acquire-%s.py' % script)] + args,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    compare(r.stdout.splitlines(), GOLD[case])


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without the CUDA library the product path raises."""
    from gnsstools import _native
    monkeypatch.setattr(_native, 'LIB_PATH', str(tmp_path / 'nope.so'))
    monkeypatch.setattr(_native, '_lib', None)
    monkeypatch.setattr(_native, '_default_engine', None)
    with pytest.raises(_native.NativeError):
        _native.default_engine()


# --------------------------------------------------------------------------- serial long-code acquisitions
GOLD_SERIAL = json.load(open(os.path.join(HERE, 'golden', 'cli_serial_golden.json')))


def compare_serial(got_lines, want_lines):
    """'%f %f' % (code_phase, metric): the code phase (hypothesis index) is exact, the metric within 1e-4."""
    assert len(got_lines) == len(want_lines) == 1
    g, w = got_lines[0].split(), want_lines[0].split()
    assert g[0] == w[0], (got_lines, want_lines)
    assert abs(float(g[1]) - float(w[1])) <= 1e-4 * float(w[1]), (got_lines, want_lines)


def test_l2cl_cli_on_emulated_kernels(tmp_path, monkeypatch, capsys):
    """acquire-gps-l2cl.py end to end, correlator bank compiled for the host (test harness)."""
    import importlib.util
    import emu_util
    from gnsstools import _native
    eng = emu_util.emu_engine()
    monkeypatch.setattr(_native, '_default_engine', eng)
    p = tmp_path / 'l2cl.iq'
    p.write_bytes(synth_files.recording_serial('gps-l2cl'))
    script, args = synth_files.command_serial('gps-l2cl', str(p))
    spec = importlib.util.spec_from_file_location('acquire_gps_l2cl', os.path.join(PKG, 'acquire-%s.py' % script))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main(args)
    compare_serial(capsys.readouterr().out.splitlines(), GOLD_SERIAL['gps-l2cl'])
    monkeypatch.setattr(_native, '_default_engine', None)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize('case', sorted(synth_files.SERIAL_CASES))
def test_serial_scripts_match_reference_stdout(case, tmp_path):
    p = tmp_path / (case + '.iq')
    p.write_bytes(synth_files.recording_serial(case))
    script, args = synth_files.command_serial(case, str(p))
    r = subprocess.run([sys.executable, os.path.join(PKG, 'acquire-%s.py' % script)] + args,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    compare_serial(r.stdout.splitlines(), GOLD_SERIAL[case])

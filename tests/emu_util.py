"""TEST HARNESS ONLY: build/load the host-emulated kernel library (tests/cuda_emu)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, 'cuda_emu')
EMU_SO = os.path.join(EMU_DIR, 'libgnssacq_emu.so')
CSRC = os.path.join(os.path.dirname(HERE), 'gnss-dsp-tools_b200', 'csrc')


def _stale():
    if not os.path.isfile(EMU_SO):
        return True
    t = os.path.getmtime(EMU_SO)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(EMU_DIR, f) for f in ('cuda_emu.h', 'cuda_emu.cpp')]
    return any(os.path.getmtime(s) > t for s in srcs)


def emu_cdll():
    if _stale():
        subprocess.check_call([os.path.join(EMU_DIR, 'build_emu.sh')])
    return ctypes.CDLL(EMU_SO)


def emu_engine():
    from gnsstools import _native
    return _native.Engine(0, lib=emu_cdll())

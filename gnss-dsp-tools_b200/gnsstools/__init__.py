"""gnsstools — host-side mirror of the reference package surface.

Same module names and call signatures as pmonta/GNSS-DSP-tools' ``gnsstools``
(``nco``, ``io``, ``util`` and the per-constellation code generators); the
FFT acquisition search that each reference ``acquire-*.py`` carries as a local
``search()`` lives in :mod:`gnsstools.acquire` and runs on the GPU through the
C-ABI library ``libgnssacq.so`` (include/gnssacq.h).
"""

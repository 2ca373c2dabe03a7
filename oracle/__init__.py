"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference acquisition path.

Nothing under gnss-dsp-tools_b200/ may import this package. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
and only as the checker or the timed CPU baseline, never as the product path.
"""

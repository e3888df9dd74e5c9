"""Host-side helpers of bench.py that do not need a GPU: the static ncu traffic figure the `roofline` object quotes
and the algorithmic work table it is compared with (385.406 GFLOP, ~609 MB per image at 512 x 512)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_ncu_traffic_covers_every_launch_of_the_shipped_plan():
    """The newest committed `ncu --set full` capture must describe the 18-launch plan and every kernel family in it
    must be recognised (a renamed kernel once dropped two launches and the figure silently became null)."""
    import bench
    traffic, src = bench.ncu_traffic_bytes(18)
    assert traffic is not None and src.endswith("_ncu_raw.csv"), (traffic, src)
    # 17 non-stem launches of a batch-64 step: algorithmic 28.7 GB; a capture far off means mislabelled rows
    assert 24e9 < traffic < 34e9, traffic
    assert bench.ncu_traffic_bytes(22) == (None, None)          # another plan's launch count is refused


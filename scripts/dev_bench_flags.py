"""bench.py with knn_tc switches overridden (developer A/B on one box):
    python scripts/dev_bench_flags.py SYM_SAMPLE_FIRST=0 [SYM_WIDE=1] -- [bench.py arguments]"""
import sys
sys.path.insert(0, ".")
from reid_gan_b200 import knn_tc
args = sys.argv[1:]
rest = args[args.index("--") + 1:] if "--" in args else []
for a in (args[:args.index("--")] if "--" in args else args):
    name, val = a.split("=")
    assert hasattr(knn_tc, name), name
    setattr(knn_tc, name, bool(int(val)) if isinstance(getattr(knn_tc, name), bool) else int(val))
sys.argv = [sys.argv[0]] + rest
import bench
bench.main()

"""bench.py with the strip flavour of the symmetric candidate kernel switched on (developer A/B)."""
import sys
sys.path.insert(0, ".")
from reid_gan_b200 import knn_tc
knn_tc.SYM_WIDE = True
import bench
bench.main()

import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_scan import run
run(1_000_000, 384, 1, 5, iters=10)
run(10_000, 384, 1, 5, iters=10)

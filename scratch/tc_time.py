import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tc_check
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
os.environ.setdefault('NFK_ONLY_TC', '1')
tc_check.timeit((64, 64), 10, B)

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_probe_loss import run
run(4, 4, 64, 256, 256, iters=2)

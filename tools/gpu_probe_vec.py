import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_probe_loss import run
for cfg in [(4,4,64,256,256),(4,4,64,512,512),(4,4,8,200,200),(4,2,32,512,512),(6,4,32,512,512)]:
    run(*cfg)

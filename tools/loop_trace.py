"""Debug: per-phase timeline of the persistent mean-shift loop kernel (UOC_LOOP_TRACE).  usage: python tools/loop_trace.py [fields per launch]"""
import os, sys
os.environ["UOC_LOOP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unseenobjectclustering_b200 import mean_shift as MS
from unseenobjectclustering_b200 import synthetic

torch.manual_seed(0)
dev = torch.device("cuda:0")
n, D, M = 640 * 480, 64, 100
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
X = torch.nn.functional.normalize(torch.randn(B, D, 480, 640, device=dev), dim=1)
for rep in range(3):
    sys.stderr.write("--- rep %d\n" % rep)
    labels, sel = MS.cluster_fields(X, M, 10.0, 10, [123 + i for i in range(B)], epsilon=0.04)
    torch.cuda.synchronize()

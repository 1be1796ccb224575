"""one Float64 qrBlocked! (for ncu launch lists): python tools/one_qr.py n"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
src = torch.randn((n, n), device="cuda", dtype=torch.float64)
dtau = torch.zeros(n, device="cuda", dtype=torch.float64)
g.qr_blocked_dev(src.data_ptr(), n, n, n, dtau.data_ptr(), 0, torch.cuda.current_stream().cuda_stream, np.float64)
torch.cuda.synchronize()

import subprocess, sys
code = '''
import sys, torch
import kernel_checks as kc
from pe_b200 import native
L = native.lib(); P, S = kc.P, kc.S
st, no, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
L.pe_debug_pipeline(st, no)
x = torch.randn(B, 56, 56, 64, device="cuda"); w = torch.randn(64, 64, 1, 1, device="cuda")
tck, tkc = kc.pack(w); y = torch.empty(B, 56, 56, 64, device="cuda")
stats = torch.zeros(128, device="cuda", dtype=torch.float64)
use_stats = int(sys.argv[4])
for i in range(3):
    L.pe_conv2d_fwd(P(x), P(tck), P(y), B, 56, 56, 64, 64, 1, 1, 1, 0, None, None, None, 0, 0, P(stats) if use_stats else None, S())
    torch.cuda.synchronize()
print("ok", st, no, B, L.pe_device_error())
'''
for cfg in [(0, 0, 256, 1), (5, 3, 256, 1), (4, 2, 256, 1), (3, 2, 256, 1), (3, 2, 256, 0), (3, 2, 8, 1), (6, 1, 256, 1)]:
    r = subprocess.run([sys.executable, "-c", code] + [str(c) for c in cfg], capture_output=True, text=True)
    print(cfg, (r.stdout.strip() or r.stderr.strip()[-200:]))

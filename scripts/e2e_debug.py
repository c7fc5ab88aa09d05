"""Host-resident (streamed) PDGETRF against the device-resident one at a given N with the default options: where do they differ?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import scalapack_b200 as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
for kv in sys.argv[3:]:
    k, v = kv.split("="); S.set_option(k, int(v))
ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 1, 1)
desca, _ = S.descinit(n, n, nb, nb, 0, 0, ctx, n)
A = torch.empty(n * n, dtype=torch.float64, device="cuda")
S.matgen64(ctx, n, n, nb, nb, A, n, 20261017)
Ah = torch.empty(n * n, dtype=torch.float64, pin_memory=True); Ah.copy_(A)
ip_d = np.zeros(n + nb, np.int32); ip_h = np.zeros(n + nb, np.int32)
assert S.pdgetrf(n, n, A, 1, 1, desca, ip_d) == 0
assert S.pdgetrf(n, n, Ah.numpy(), 1, 1, desca, ip_h) == 0
print("ipiv equal:", np.array_equal(ip_d, ip_h), "first diff", (np.nonzero(ip_d != ip_h)[0][:5]).tolist())
D = (Ah.cuda() != A).view(n, n)              # [col, row]
nd = int(D.sum().item())
print("differing elements:", nd, "of", n * n)
if nd:
    cols = torch.nonzero(D.any(dim=1)).flatten(); rows = torch.nonzero(D.any(dim=0)).flatten()
    print("cols with diffs:", cols.numel(), "first", cols[:8].tolist(), "last", cols[-4:].tolist())
    print("rows with diffs:", rows.numel(), "first", rows[:8].tolist(), "last", rows[-4:].tolist())
    percol = D.sum(dim=1)
    blk = percol.view(-1, nb).sum(dim=1)
    print("diffs per block column:", blk.tolist()[:64])
    err = (Ah.cuda() - A).abs().max().item(); print("max abs diff", err)

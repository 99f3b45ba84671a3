"""CTA-pair tcgen05.mma self-test (cluster of 2, cta_group::2): K-major operands, N = 256 and N = 16."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reni_b200 import _lib

def image_kmajor(mat):
    R, K = mat.shape
    return np.ascontiguousarray(mat.reshape(R, K // 8, 8).transpose(1, 0, 2))

lib = _lib.load()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
def run(A, B, args, N, ksteps):
    a = np.concatenate([image_kmajor(A[:128]).reshape(-1), image_kmajor(A[128:]).reshape(-1)]).view(np.uint8)
    b = np.concatenate([image_kmajor(B[:N // 2]).reshape(-1), image_kmajor(B[N // 2:]).reshape(-1)]).view(np.uint8)
    ta, tb = torch.from_numpy(a.copy()).to(dev), torch.from_numpy(b.copy()).to(dev)
    d = torch.zeros(256, N, device=dev)
    rc = lib.reni_selftest_umma2(C.c_void_p(ta.data_ptr()), ta.numel() // 2, C.c_void_p(tb.data_ptr()), tb.numel() // 2,
                                 *args, N, ksteps, C.c_void_p(d.data_ptr()), None)
    _lib.check(rc, "selftest2")
    torch.cuda.synchronize()
    return d.cpu().numpy()
A = rng.uniform(-1, 1, (256, 64)).astype(np.float16)
B = rng.uniform(-1, 1, (256, 64)).astype(np.float16)
ref = A.astype(np.float32) @ B.astype(np.float32).T
got = run(A, B, (2048, 128, 2048, 128, 4096, 4096, 0, 0), 256, 4)
print("selftest2 N=256 max err", np.abs(got - ref).max(), "| rows 0..127", np.abs(got[:128] - ref[:128]).max(), "rows 128..255", np.abs(got[128:] - ref[128:]).max())
B16 = rng.uniform(-1, 1, (16, 64)).astype(np.float16)
ref = A.astype(np.float32) @ B16.astype(np.float32).T
got = run(A, B16, (2048, 128, 128, 128, 4096, 256, 0, 0), 16, 4)
print("selftest2 N=16 max err", np.abs(got - ref).max())
# probe: bulk copy completing on the peer CTA's mbarrier (plain bulk copy vs TMA tile load with .cta_group::2)
for use_tma in (0, 1):
    src = torch.randint(0, 255, (2 * 16384,), dtype=torch.uint8, device=dev)
    res = torch.zeros(4, dtype=torch.int32, device=dev)
    rc = lib.reni_probe_remote_tx(C.c_void_p(src.data_ptr()), 16384, C.c_void_p(res.data_ptr()), use_tma, None)
    try:
        _lib.check(rc, "probe"); torch.cuda.synchronize()
        r = res.cpu().numpy()
        print("probe remote complete_tx use_tma", use_tma, ": completed", r[0], "sums", r[1], r[2], "expected", int(src[:16384].sum()), int(src[16384:].sum()))
    except Exception as ex:
        print("probe failed:", ex)
# MN-major operands on a CTA pair, N = 16 (8 columns of B per CTA): the layer-0 reduction of the paired delta chain
At = [rng.uniform(-1, 1, (64, 128)).astype(np.float16) for _ in range(2)]
Bt = [rng.uniform(-1, 1, (64, 8)).astype(np.float16) for _ in range(2)]
a = np.concatenate([image_kmajor(x).reshape(-1) for x in At]).view(np.uint8)
b = np.concatenate([image_kmajor(x).reshape(-1) for x in Bt]).view(np.uint8)
ta, tb = torch.from_numpy(a.copy()).to(dev), torch.from_numpy(b.copy()).to(dev)
d = torch.zeros(256, 16, device=dev)
rc = lib.reni_selftest_umma2(C.c_void_p(ta.data_ptr()), ta.numel() // 2, C.c_void_p(tb.data_ptr()), tb.numel() // 2,
                             128, 1024, 128, 1024, 256, 256, 1, 1, 16, 4, C.c_void_p(d.data_ptr()), None)
_lib.check(rc, "selftest2 mn"); torch.cuda.synchronize()
got = d.cpu().numpy()
Bcat = np.concatenate([Bt[0], Bt[1]], axis=1).astype(np.float32)
ref = np.concatenate([At[0].astype(np.float32).T @ Bcat, At[1].astype(np.float32).T @ Bcat])
print("selftest2 MN-major N=16 max err", np.abs(got - ref).max(), "(own-columns only:",
      max(np.abs(got[:128, :8] - ref[:128, :8]).max(), np.abs(got[128:, 8:] - ref[128:, 8:]).max()), ")")

"""Timing of the environment-map shading kernels (FIT_INVERSE consumer): 128 x 128 render lit by a 64 x 128 map, B maps.
Beside it the reference formulation (materialised einsums, torch on the same GPU, one map at a time)."""
import os, sys, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.build()
from reni_b200 import EnvironmentMap, blinn_phong_shading_env_map, get_directions, get_sineweight
dev = torch.device("cuda:0")
torch.manual_seed(0)
H = W = 128
sidelen = 128
D, sw = get_directions(sidelen).to(dev), get_sineweight(sidelen).to(dev)
J = D.shape[1]
nrm = torch.nn.functional.normalize(torch.randn(H, W, 3, device=dev), dim=-1)
pos = torch.randn(H, W, 3, device=dev) * 0.3
cam = torch.tensor([[0.0, 0.0, 2.0]], device=dev)
def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)
for B in (1, 8, 32):
    env = torch.rand(B, J, 3, device=dev).requires_grad_(True)
    g = torch.randn(B, H, W, 3, device=dev)
    def fwd():
        return blinn_phong_shading_env_map(nrm, pos, cam, EnvironmentMap(env, D, sw), 500.0, 0.5, 0.5)[0]
    def fwdbwd():
        env.grad = None
        (fwd() * g).sum().backward()
    tf, tfb = timeit(fwd), timeit(fwdbwd)
    pairs = H * W * J
    print(f"B={B:3d}: forward {tf*1e3:8.1f} us, forward+adjoint {tfb*1e3:8.1f} us  ({pairs/1e6:.0f} M pixel-texel pairs, "
          f"{pairs*((B+7)//8)/tf/1e6:.1f} G weight evaluations/s forward)")
# reference formulation in torch on the GPU, one map (it materialises (H, W, J[, 3]) tensors: 1.6 GB for the half vectors)
def ref_one():
    n = nrm.reshape(-1, 3); l = D[0]
    diffuse = (n @ l.T).clamp(0, 1)
    v = torch.nn.functional.normalize(cam - pos.reshape(-1, 3), dim=-1)
    Hh = torch.nn.functional.normalize(v[:, None, :] + l[None, :, :], dim=-1)
    spec = (n[:, None, :] * Hh).sum(-1).clamp(0, 1) ** 500.0
    c = (500.0 + 2) / (4 * (2 - torch.exp(torch.tensor(-250.0))))
    return (0.5 * diffuse + c * 0.5 * spec) @ (torch.rand(J, 3, device=dev) * sw[0])
print(f"torch restatement of the reference einsums, 1 map: {timeit(ref_one, 5)*1e3:.1f} us")
